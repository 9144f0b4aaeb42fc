"""GPU parity, operator by operator: every per-operator C-ABI entry of libquokka_b200.so against the CPU
oracle (oracle/quokka_oracle.c, itself pinned bit-for-bit to the reference's templates) on the same seeded
inputs.  Bar: BIT-EXACT for every FP64 output (+0/-0 compare equal) and for the int redoFlag / counters.
All calls go through the C ABI with device pointers (torch is only the allocator).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import QK_HLLC, QK_LLF, QK_MC, QK_MINMOD, check, hydro_params, qk_box

pytestmark = pytest.mark.gpu

PROBLEMS = {0: (1.4, 0, 0, 0), 1: (1.4, 1, 0, 0), 2: (5.0 / 3.0, 1, 3, 2), 3: (1.0, 0, 1, 0), 4: (1.0, 1, 3, 2)}
# problem 3: the isothermal EOS (gamma = 1, cs_isothermal = 1.3: HydroSystem::is_eos_isothermal(), hydro_system.hpp:133); the oracle's
# isothermal branches are pinned to the reference's templates by tests/test_oracle_vs_ref.py (harness problem 3)
CS_ISO = {3: 1.3, 4: 0.7}
VALID = qk_box.make((3, -2, 5), (38, 17, 24))  # 36 x 20 x 20: not a multiple of the block size, non-zero origin
NG = 4


@pytest.fixture(scope="module")
def lib():
    return capi.load()


def params(problem):
    g, re, ns, nms = PROBLEMS[problem]
    return hydro_params(gamma=g, reconstruct_eint=re, nscalars=ns, nmscalars=nms, cs_isothermal=CS_ISO.get(problem, float("nan")))


def make_cons(problem, kind, seed=12345):
    g, re, ns, nms = PROBLEMS[problem]
    gb = VALID.grown(NG)
    rho, v, P, rng = ol.random_cons(gb, ns, seed, kind)
    cons = ol.HostFab(gb, 6 + ns)
    cons.a[...] = ol.cons_from_prim(rho, v, P, g if g != 1.0 else 1.4, rng, ns)  # the isothermal EOS never reads the energies
    return cons


def exact(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} mismatches, max abs diff {np.nanmax(np.abs(a - b))}"


def dev(hf):
    """HostFab -> DevFab with the same contents"""
    from quokka_b200.device import DevFab

    return DevFab(hf.box, hf.ncomp, dtype="f64" if hf.a.dtype == np.float64 else "i32", host=hf.a)


def one(x):
    return C.byref(x)


def oracle_prim(problem, kind):
    prm = params(problem)
    cons = make_cons(problem, kind)
    gb = VALID.grown(NG)
    po = ol.HostFab(gb, cons.ncomp)
    ol.oracle().orc_conserved_to_primitive(one(prm), one(cons.desc()), one(po.desc()), one(gb))
    return prm, cons, po


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_cons_to_prim(lib, problem, kind):
    prm, cons, po = oracle_prim(problem, kind)
    dc, dp = dev(cons), dev(ol.HostFab(po.box, po.ncomp))
    check(lib.qk_hydro_conserved_to_primitive(one(prm), 1, one(VALID), one(dc.desc()), one(dp.desc()), NG, None))
    exact(dp.numpy(), po.a, "prim")


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_flattening_coefficients(lib, problem, d):
    prm, cons, po = oracle_prim(problem, "shocked")
    g2 = VALID.grown(2)
    co = ol.HostFab(g2, 1)
    ol.oracle().orc_flattening_coefficients(one(prm), d, one(po.desc()), one(co.desc()), one(g2))
    dp, dc = dev(po), dev(ol.HostFab(g2, 1))
    check(lib.qk_hydro_flattening_coefficients(one(prm), d, 1, one(VALID), one(dp.desc()), one(dc.desc()), 2, None))
    exact(dc.numpy(), co.a, "chi")
    assert (co.a < 1.0).any() and (co.a == 1.0).any()


@pytest.mark.parametrize("order,limiter", [(1, 0), (2, QK_MINMOD), (2, QK_MC), (3, 0)])
@pytest.mark.parametrize("d", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_reconstruct(lib, order, limiter, d, kind):
    prm, cons, po = oracle_prim(2, kind)
    nv = po.ncomp
    g1 = VALID.grown(1)
    fb = ol.face_box(VALID, d, 1)
    L, R = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    ol.oracle().orc_reconstruct_states(order, limiter, d, one(po.desc()), one(L.desc()), one(R.desc()), one(g1), nv)
    dp, dl, dr = dev(po), dev(ol.HostFab(fb, nv)), dev(ol.HostFab(fb, nv))
    check(lib.qk_reconstruct_states(order, limiter, d, 1, one(VALID), one(dp.desc()), one(dl.desc()), one(dr.desc()), 1, nv, None))
    # the oracle and the GPU write the same face set; untouched faces stay 0 in both
    exact(dl.numpy(), L.a, "left")
    exact(dr.numpy(), R.a, "right")


def oracle_states(problem, d, kind, order=3, flatten=True):
    prm, cons, po = oracle_prim(problem, kind)
    nv = po.ncomp
    g1, g2 = VALID.grown(1), VALID.grown(2)
    chis = []
    for dd in range(3):
        c = ol.HostFab(g2, 1)
        ol.oracle().orc_flattening_coefficients(one(prm), dd, one(po.desc()), one(c.desc()), one(g2))
        chis.append(c)
    fb = ol.face_box(VALID, d, 1)
    L, R = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    ol.oracle().orc_reconstruct_states(order, QK_MINMOD, d, one(po.desc()), one(L.desc()), one(R.desc()), one(g1), nv)
    L0, R0 = L.a.copy(), R.a.copy()
    if flatten:
        ol.oracle().orc_flatten_shocks(d, one(po.desc()), one(chis[0].desc()), one(chis[1].desc()), one(chis[2].desc()), one(L.desc()),
                                       one(R.desc()), one(g1), nv)
    return prm, cons, po, chis, L, R, L0, R0


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_flatten_shocks(lib, problem, d):
    prm, cons, po, chis, L, R, L0, R0 = oracle_states(problem, d, "shocked")
    nv = po.ncomp
    fb = ol.face_box(VALID, d, 1)
    hl, hr = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    hl.a[...] = L0
    hr.a[...] = R0
    dp, dl, dr = dev(po), dev(hl), dev(hr)
    dch = [dev(c) for c in chis]
    check(lib.qk_hydro_flatten_shocks(d, 1, one(VALID), one(dp.desc()), one(dch[0].desc()), one(dch[1].desc()), one(dch[2].desc()),
                                      one(dl.desc()), one(dr.desc()), 1, nv, None))
    exact(dl.numpy(), L.a, "left")
    exact(dr.numpy(), R.a, "right")


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("solver", [QK_HLLC, QK_LLF])
@pytest.mark.parametrize("d", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_compute_fluxes(lib, problem, solver, d, kind):
    prm, cons, po, chis, L, R, _, _ = oracle_states(problem, d, kind, order=3 if solver == QK_HLLC else 1, flatten=(solver == QK_HLLC))
    nv = po.ncomp
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(one(prm), solver, d, one(Fo.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    dp, dl, dr = dev(po), dev(L), dev(R)
    dF, dV = dev(ol.HostFab(fb0, nv)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_compute_fluxes(one(prm), solver, d, 1, one(VALID), one(dF.desc()), one(dV.desc()), one(dl.desc()), one(dr.desc()),
                                      one(dp.desc()), None))
    exact(dF.numpy(), Fo.a, "flux")
    exact(dV.numpy(), Vo.a, "facevel")
    assert np.isfinite(Fo.a).all()


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_compute_fluxes_with_artificial_viscosity(lib, problem, d):
    """artificialViscosityK_ = 0.1 (hydro_system.hpp:1052-1076): operator entry and the one-kernel flux function, bit-exact vs the oracle
    (which tests/test_oracle_vs_ref.py pins to the reference with the same K_visc)"""
    prm, cons, po, chis, L, R, _, _ = oracle_states(problem, d, "shocked", order=3, flatten=True)
    prm.K_visc = 0.1
    nv = po.ncomp
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(one(prm), QK_HLLC, d, one(Fo.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    dp, dl, dr = dev(po), dev(L), dev(R)
    dF, dV = dev(ol.HostFab(fb0, nv)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_compute_fluxes(one(prm), QK_HLLC, d, 1, one(VALID), one(dF.desc()), one(dV.desc()), one(dl.desc()), one(dr.desc()),
                                      one(dp.desc()), None))
    exact(dF.numpy(), Fo.a, "flux with artificial viscosity")
    exact(dV.numpy(), Vo.a, "facevel")
    prm.K_visc = 0.0
    F0 = ol.HostFab(fb0, nv)
    ol.oracle().orc_compute_fluxes(one(prm), QK_HLLC, d, one(F0.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    assert (F0.a[0] != Fo.a[0]).any()


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_flux_function_fused(lib, problem, order, d):
    """hydroFluxFunction<DIR> in one kernel == Reconstruct -> FlattenShocks -> ComputeFluxes<HLLC> of the oracle."""
    prm, cons, po, chis, L, R, _, _ = oracle_states(problem, d, "shocked", order=order, flatten=True)
    prm.reconstruction_order = order
    nv = po.ncomp
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(one(prm), QK_HLLC, d, one(Fo.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    dp = dev(po)
    dch = [dev(c) for c in chis]
    dF, dV = dev(ol.HostFab(fb0, nv)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_flux_function(one(prm), 0, d, 1, one(VALID), one(dp.desc()), one(dch[0].desc()), one(dch[1].desc()), one(dch[2].desc()),
                                     one(dF.desc()), one(dV.desc()), None))
    exact(dF.numpy(), Fo.a, "flux")
    exact(dV.numpy(), Vo.a, "facevel")


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_fo_flux_function_fused(lib, problem, d):
    """hydroFOFluxFunction<DIR> (donor cell + LLF)."""
    prm, cons, po, chis, L, R, _, _ = oracle_states(problem, d, "shocked", order=1, flatten=False)
    nv = po.ncomp
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(one(prm), QK_LLF, d, one(Fo.desc()), one(Vo.desc()), one(L.desc()), one(R.desc()), one(po.desc()), one(fb0))
    dp = dev(po)
    dF, dV = dev(ol.HostFab(fb0, nv)), dev(ol.HostFab(fb0, 1))
    check(lib.qk_hydro_flux_function(one(prm), 1, d, 1, one(VALID), one(dp.desc()), None, None, None, one(dF.desc()), one(dV.desc()), None))
    exact(dF.numpy(), Fo.a, "flux")
    exact(dV.numpy(), Vo.a, "facevel")


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
def test_update_ops(lib, problem):
    prm = params(problem)
    cons = make_cons(problem, "shocked")
    nv = cons.ncomp
    rng = np.random.default_rng(7)
    dx = (C.c_double * 3)(0.01, 0.02, 0.03)
    F = [ol.HostFab(ol.face_box(VALID, d, 0), nv) for d in range(3)]
    V = [ol.HostFab(ol.face_box(VALID, d, 0), 1) for d in range(3)]
    for f in F + V:
        f.a[...] = rng.standard_normal(f.a.shape)
    dFl = [dev(f) for f in F]
    dVl = [dev(v) for v in V]
    dcons = dev(cons)
    # Saxpy
    sa = ol.HostFab(F[0].box, nv)
    sa.a[...] = rng.standard_normal(sa.a.shape)
    dsa = dev(sa)
    ol.oracle().orc_saxpy(one(sa.desc()), 0.5, one(F[0].desc()), one(F[0].box), nv)
    check(lib.qk_saxpy(1, one(F[0].box), one(dsa.desc()), 0.5, one(dFl[0].desc()), nv, None))
    exact(dsa.numpy(), sa.a, "saxpy")
    # ComputeRhsFromFluxes
    ro = ol.HostFab(VALID, nv)
    ol.oracle().orc_rhs_from_fluxes(one(ro.desc()), one(F[0].desc()), one(F[1].desc()), one(F[2].desc()), dx, one(VALID), nv)
    dr = dev(ol.HostFab(VALID, nv))
    check(lib.qk_hydro_rhs_from_fluxes(1, one(VALID), one(dr.desc()), one(dFl[0].desc()), one(dFl[1].desc()), one(dFl[2].desc()), dx, nv, None))
    exact(dr.numpy(), ro.a, "rhs")
    # AddInternalEnergyPdV with a few redo cells
    redo = ol.HostFab(VALID.grown(1), 1, dtype=np.int32, fill=0)
    redo.view(VALID)[0, 2, 3, 4] = 1
    redo.view(VALID)[0, 0, 0, 0] = 1
    dredo = dev(redo)
    ol.oracle().orc_add_internal_energy_pdv(one(prm), one(ro.desc()), one(cons.desc()), dx, one(V[0].desc()), one(V[1].desc()), one(V[2].desc()),
                                            one(redo.desc()), one(VALID))
    check(lib.qk_hydro_add_internal_energy_pdv(one(prm), 1, one(VALID), one(dr.desc()), one(dcons.desc()), dx, one(dVl[0].desc()),
                                               one(dVl[1].desc()), one(dVl[2].desc()), one(dredo.desc()), None))
    exact(dr.numpy(), ro.a, "rhs+pdv")
    # PredictStep (dt large enough that some densities go negative)
    dt = 0.05
    uo = ol.HostFab(VALID, nv)
    uo.a[...] = cons.view(VALID)
    no = ol.HostFab(VALID, nv)
    fo = ol.HostFab(VALID.grown(1), 1, np.int32)
    nbad_o = ol.oracle().orc_predict_step(one(prm), one(uo.desc()), one(no.desc()), one(ro.desc()), dt, nv, one(fo.desc()), one(VALID))
    duo, dno, dfo = dev(uo), dev(ol.HostFab(VALID, nv)), dev(ol.HostFab(VALID.grown(1), 1, np.int32))
    nbad = C.c_int64(-1)
    check(lib.qk_hydro_predict_step(one(prm), 1, one(VALID), one(duo.desc()), one(dno.desc()), one(dr.desc()), dt, nv, one(dfo.desc()),
                                    C.byref(nbad), None))
    exact(dno.numpy(), no.a, "predict")
    assert (dfo.numpy() == fo.a).all() and nbad.value == nbad_o and nbad_o > 0
    # EnforceLimits (with floors that bite) and SyncDualEnergy
    st = ol.HostFab(VALID, nv)
    st.a[...] = cons.view(VALID)
    st.a[0].flat[::7] *= 1e-3
    st.a[4].flat[::5] *= 1e-6
    dst = dev(st)
    prm.density_floor = 0.05
    prm.temp_floor = 1e14 if problem == 0 else 3e15
    ol.oracle().orc_enforce_limits(one(prm), one(st.desc()), one(VALID))
    check(lib.qk_hydro_enforce_limits(one(prm), 1, one(VALID), one(dst.desc()), None))
    exact(dst.numpy(), st.a, "enforce")
    nab_o = ol.oracle().orc_sync_dual_energy(one(prm), one(st.desc()), one(VALID))
    nab = C.c_int64(-1)
    check(lib.qk_hydro_sync_dual_energy(one(prm), 1, one(VALID), one(dst.desc()), C.byref(nab), None))
    exact(dst.numpy(), st.a, "sync")
    assert nab.value == nab_o
    # signal speeds
    cv = ol.HostFab(VALID, nv)
    cv.a[...] = cons.view(VALID)
    dcv = dev(cv)
    for which in (0, 1):
        m_o = ol.oracle().orc_max_signal_speed(one(prm), which, one(cv.desc()), one(VALID))
        m = C.c_double(0)
        check(lib.qk_hydro_max_signal_speed(one(prm), which, 1, one(VALID), one(dcv.desc()), C.byref(m), None))
        assert m.value == m_o


@pytest.mark.parametrize("d", [0, 1, 2])
def test_replace_fluxes(lib, d):
    nv = 6
    rng = np.random.default_rng(3)
    fb0 = ol.face_box(VALID, d, 0)
    F, FO = ol.HostFab(fb0, nv), ol.HostFab(fb0, nv)
    F.a[...] = rng.standard_normal(F.a.shape)
    FO.a[...] = rng.standard_normal(F.a.shape)
    redo = ol.HostFab(VALID.grown(1), 1, dtype=np.int32, fill=0)
    redo.a[...] = (rng.uniform(size=redo.a.shape) < 0.02).astype(np.int32)
    dF, dFO, dre = dev(F), dev(FO), dev(redo)
    ol.oracle().orc_replace_fluxes(d, one(F.desc()), one(FO.desc()), one(redo.desc()), one(VALID), nv)
    check(lib.qk_hydro_replace_fluxes(d, 1, one(VALID), one(dF.desc()), one(dFO.desc()), one(dre.desc()), nv, None))
    exact(dF.numpy(), F.a, "replace")


def test_eos_reset_edge(lib):
    """negative E - KE drives the EOS through eos_reset (T clamped to small_temp): SURVEY.md fact 5"""
    prm = params(0)
    cons = make_cons(0, "shocked")
    cons.a[4].flat[::11] = 1e-3  # E far below KE -> negative internal energy
    gb = VALID.grown(NG)
    po = ol.HostFab(gb, 6)
    ol.oracle().orc_conserved_to_primitive(one(prm), one(cons.desc()), one(po.desc()), one(gb))
    dc, dp = dev(cons), dev(ol.HostFab(gb, 6))
    check(lib.qk_hydro_conserved_to_primitive(one(prm), 1, one(VALID), one(dc.desc()), one(dp.desc()), NG, None))
    exact(dp.numpy(), po.a, "prim")


def test_unsupported_params(lib):
    prm = params(0)
    prm.nscalars = capi.QK_MAX_SCALARS + 1
    dc = dev(ol.HostFab(VALID.grown(NG), 6))
    assert lib.qk_hydro_conserved_to_primitive(one(prm), 1, one(VALID), one(dc.desc()), one(dc.desc()), NG, None) == capi.QK_ERR_UNSUPPORTED
