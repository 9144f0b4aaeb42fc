"""Pin the C oracle (oracle/quokka_oracle.c) against the REFERENCE's own operator templates compiled
from /root/reference (oracle/_ref/libquokka_ref.so; recipe oracle/ref_build/Makefile).

Bar: bit-exact (max |a-b| == 0; +0/-0 compare equal) on seeded smooth and shocked inputs, for the
three trait sets the harness instantiates and all three directions.  Skipped where oracle/_ref was
not built (the GPU box only has it if it travelled with the snapshot; /root/reference never does).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200.capi import QK_HLLC, QK_LLF, QK_MC, QK_MINMOD, hydro_params, qk_box

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libquokka_ref.so not built")

# harness problem id -> (gamma, reconstruct_eint, nscalars, nmscalars)   (oracle/ref_build/ref_harness.cpp)
PROBLEMS = {0: (1.4, 0, 0, 0), 1: (1.4, 1, 0, 0), 2: (5.0 / 3.0, 1, 3, 2), 3: (1.0, 0, 1, 0), 4: (1.0, 1, 3, 2)}
# problem 3: the isothermal EOS (gamma = 1, EOS_Traits::cs_isothermal = 1.3; every is_eos_isothermal() branch of hydro_system.hpp / HLLC.hpp)
CS_ISO = {3: 1.3, 4: 0.7}
VALID = qk_box.make((3, -2, 5), (14, 7, 12))
NG = 4


def params(problem):
    g, re, ns, nms = PROBLEMS[problem]
    return hydro_params(gamma=g, reconstruct_eint=re, nscalars=ns, nmscalars=nms, cs_isothermal=CS_ISO.get(problem, float("nan")))


def make_cons(problem, kind, seed=12345):
    g, re, ns, nms = PROBLEMS[problem]
    gb = VALID.grown(NG)
    rho, v, P, rng = ol.random_cons(gb, ns, seed, kind)
    cons = ol.HostFab(gb, 6 + ns)
    # the isothermal EOS never reads the energies: any finite values do (here those of a gamma = 1.4 gas)
    cons.a[...] = ol.cons_from_prim(rho, v, P, g if g != 1.0 else 1.4, rng, ns)
    return cons


def exact(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{bad.sum()} mismatches, max abs diff {np.nanmax(np.abs(a - b))}"


def both_prim(problem, kind):
    prm = params(problem)
    cons = make_cons(problem, kind)
    gb = VALID.grown(NG)
    nv = cons.ncomp
    po, pr = ol.HostFab(gb, nv), ol.HostFab(gb, nv)
    ol.oracle().orc_conserved_to_primitive(C.byref(prm), C.byref(cons.desc()), C.byref(po.desc()), C.byref(gb))
    ol.ref().ref_cons_to_prim(problem, C.byref(VALID), C.byref(cons.desc()), C.byref(pr.desc()), NG)
    return prm, cons, po, pr


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_cons_to_prim(problem, kind):
    _, _, po, pr = both_prim(problem, kind)
    exact(po.a, pr.a)


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_flattening_coefficients(problem, d):
    prm, cons, po, _ = both_prim(problem, "shocked")
    g2 = VALID.grown(2)
    co, cr = ol.HostFab(g2, 1), ol.HostFab(g2, 1)
    ol.oracle().orc_flattening_coefficients(C.byref(prm), d, C.byref(po.desc()), C.byref(co.desc()), C.byref(g2))
    ol.ref().ref_flattening_coefficients(problem, d, C.byref(VALID), C.byref(po.desc()), C.byref(cr.desc()), NG, 2)
    exact(co.a, cr.a)
    assert (co.a < 1.0).any() and (co.a == 1.0).any()  # both branches exercised


@pytest.mark.parametrize("order,limiter", [(1, 0), (2, QK_MINMOD), (2, QK_MC), (3, 0)])
@pytest.mark.parametrize("d", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_reconstruct(order, limiter, d, kind):
    problem = 2
    prm, cons, po, _ = both_prim(problem, kind)
    nv = po.ncomp
    g1 = VALID.grown(1)
    fb = ol.face_box(VALID, d, 1)
    lo_, ro_ = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    lr_, rr_ = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    ol.oracle().orc_reconstruct_states(order, limiter, d, C.byref(po.desc()), C.byref(lo_.desc()), C.byref(ro_.desc()), C.byref(g1), nv)
    ol.ref().ref_reconstruct(problem, order, limiter, d, C.byref(VALID), C.byref(po.desc()), C.byref(lr_.desc()), C.byref(rr_.desc()), NG, 1, nv)
    exact(lo_.a, lr_.a)
    exact(ro_.a, rr_.a)


def full_states(problem, d, kind, order=3):
    """prim, chi1..3, flattened L/R from the oracle (each step separately pinned above/below)."""
    prm, cons, po, _ = both_prim(problem, kind)
    nv = po.ncomp
    g1, g2 = VALID.grown(1), VALID.grown(2)
    chis = []
    for dd in range(3):
        c = ol.HostFab(g2, 1)
        ol.oracle().orc_flattening_coefficients(C.byref(prm), dd, C.byref(po.desc()), C.byref(c.desc()), C.byref(g2))
        chis.append(c)
    fb = ol.face_box(VALID, d, 1)
    L, R = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    ol.oracle().orc_reconstruct_states(order, QK_MINMOD, d, C.byref(po.desc()), C.byref(L.desc()), C.byref(R.desc()), C.byref(g1), nv)
    return prm, cons, po, chis, L, R


@pytest.mark.parametrize("problem", [0, 2])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_flatten_shocks(problem, d):
    prm, cons, po, chis, L, R = full_states(problem, d, "shocked")
    nv = po.ncomp
    g1 = VALID.grown(1)
    fb = ol.face_box(VALID, d, 1)
    Lr, Rr = ol.HostFab(fb, nv), ol.HostFab(fb, nv)
    Lr.a[...] = L.a
    Rr.a[...] = R.a
    ol.oracle().orc_flatten_shocks(d, C.byref(po.desc()), C.byref(chis[0].desc()), C.byref(chis[1].desc()), C.byref(chis[2].desc()),
                                   C.byref(L.desc()), C.byref(R.desc()), C.byref(g1), nv)
    ol.ref().ref_flatten_shocks(problem, d, C.byref(VALID), C.byref(po.desc()), C.byref(chis[0].desc()), C.byref(chis[1].desc()),
                                C.byref(chis[2].desc()), C.byref(Lr.desc()), C.byref(Rr.desc()), NG, 1, nv)
    exact(L.a, Lr.a)
    exact(R.a, Rr.a)


@pytest.mark.parametrize("problem", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("solver", [QK_HLLC, QK_LLF])
@pytest.mark.parametrize("d", [0, 1, 2])
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_compute_fluxes(problem, solver, d, kind):
    prm, cons, po, chis, L, R = full_states(problem, d, kind, order=3 if solver == QK_HLLC else 1)
    nv = po.ncomp
    g1 = VALID.grown(1)
    if solver == QK_HLLC:
        ol.oracle().orc_flatten_shocks(d, C.byref(po.desc()), C.byref(chis[0].desc()), C.byref(chis[1].desc()), C.byref(chis[2].desc()),
                                       C.byref(L.desc()), C.byref(R.desc()), C.byref(g1), nv)
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    Fr, Vr = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    ol.oracle().orc_compute_fluxes(C.byref(prm), solver, d, C.byref(Fo.desc()), C.byref(Vo.desc()), C.byref(L.desc()), C.byref(R.desc()),
                                   C.byref(po.desc()), C.byref(fb0))
    ol.ref().ref_compute_fluxes(problem, solver, d, C.byref(VALID), C.byref(Fr.desc()), C.byref(Vr.desc()), C.byref(L.desc()), C.byref(R.desc()),
                                C.byref(po.desc()), NG, 0.0)
    exact(Fo.a, Fr.a)
    exact(Vo.a, Vr.a)
    assert np.isfinite(Fo.a).all()
    if problem in CS_ISO:  # hydro_system.hpp:1083-1087: no energy fluxes
        assert (Fo.a[4] == 0).all() and (Fo.a[5] == 0).all() and (Fo.a[0] != 0).any()


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
@pytest.mark.parametrize("d", [0, 1, 2])
def test_compute_fluxes_with_artificial_viscosity(problem, d):
    """artificialViscosityK_ != 0 (Colella & Woodward 1984 eq. 4.2, hydro_system.hpp:1052-1076): K max(-div v, 0) (U_L - U_R) is added to
    the density, energy and scalar fluxes -- the momentum components are overwritten by the unmodified Riemann flux afterwards (:1078-1081)"""
    prm, cons, po, chis, L, R = full_states(problem, d, "shocked", order=3)
    prm.K_visc = 0.1
    nv = po.ncomp
    g1 = VALID.grown(1)
    ol.oracle().orc_flatten_shocks(d, C.byref(po.desc()), C.byref(chis[0].desc()), C.byref(chis[1].desc()), C.byref(chis[2].desc()),
                                   C.byref(L.desc()), C.byref(R.desc()), C.byref(g1), nv)
    fb0 = ol.face_box(VALID, d, 0)
    Fo, Vo = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    Fr, Vr = ol.HostFab(fb0, nv), ol.HostFab(fb0, 1)
    F0 = ol.HostFab(fb0, nv)
    ol.oracle().orc_compute_fluxes(C.byref(prm), QK_HLLC, d, C.byref(Fo.desc()), C.byref(Vo.desc()), C.byref(L.desc()), C.byref(R.desc()),
                                   C.byref(po.desc()), C.byref(fb0))
    ol.ref().ref_compute_fluxes(problem, QK_HLLC, d, C.byref(VALID), C.byref(Fr.desc()), C.byref(Vr.desc()), C.byref(L.desc()), C.byref(R.desc()),
                                C.byref(po.desc()), NG, 0.1)
    exact(Fo.a, Fr.a)
    exact(Vo.a, Vr.a)
    ol.ref().ref_compute_fluxes(problem, QK_HLLC, d, C.byref(VALID), C.byref(F0.desc()), C.byref(Vr.desc()), C.byref(L.desc()), C.byref(R.desc()),
                                C.byref(po.desc()), NG, 0.0)
    assert (Fr.a[0] != F0.a[0]).any(), "the viscosity term did nothing on shocked data"
    exact(Fr.a[1:4], F0.a[1:4])  # momentum fluxes carry no artificial viscosity


@pytest.mark.parametrize("problem", [0, 2, 3, 4])
def test_update_ops(problem):
    prm = params(problem)
    cons = make_cons(problem, "shocked")
    nv = cons.ncomp
    rng = np.random.default_rng(7)
    dx = (C.c_double * 3)(0.01, 0.02, 0.03)
    F = [ol.HostFab(ol.face_box(VALID, d, 0), nv) for d in range(3)]
    V = [ol.HostFab(ol.face_box(VALID, d, 0), 1) for d in range(3)]
    for f in F + V:
        f.a[...] = rng.standard_normal(f.a.shape)
    # ComputeRhsFromFluxes
    ro, rr = ol.HostFab(VALID, nv), ol.HostFab(VALID, nv)
    ol.oracle().orc_rhs_from_fluxes(C.byref(ro.desc()), C.byref(F[0].desc()), C.byref(F[1].desc()), C.byref(F[2].desc()), dx, C.byref(VALID), nv)
    ol.ref().ref_update_op(problem, 0, C.byref(VALID), C.byref(rr.desc()), None, None, C.byref(F[0].desc()), C.byref(F[1].desc()),
                           C.byref(F[2].desc()), None, dx, 0.0, 0.0, 0.0, None)
    exact(ro.a, rr.a)
    # AddInternalEnergyPdV with a few redo cells
    redo = ol.HostFab(VALID.grown(1), 1, dtype=np.int32, fill=0)
    redo.view(VALID)[0, 2, 3, 4] = 1
    redo.view(VALID)[0, 0, 0, 0] = 1
    ol.oracle().orc_add_internal_energy_pdv(C.byref(prm), C.byref(ro.desc()), C.byref(cons.desc()), dx, C.byref(V[0].desc()), C.byref(V[1].desc()),
                                            C.byref(V[2].desc()), C.byref(redo.desc()), C.byref(VALID))
    ol.ref().ref_update_op(problem, 1, C.byref(VALID), C.byref(rr.desc()), C.byref(cons.desc()), None, C.byref(V[0].desc()), C.byref(V[1].desc()),
                           C.byref(V[2].desc()), C.byref(redo.desc()), dx, 0.0, 0.0, 0.0, None)
    exact(ro.a, rr.a)
    # PredictStep (dt large enough that some densities go negative)
    dt = 0.05
    uo = ol.HostFab(VALID, nv)
    uo.a[...] = cons.view(VALID)
    no, nr = ol.HostFab(VALID, nv), ol.HostFab(VALID, nv)
    fo, fr = ol.HostFab(VALID.grown(1), 1, np.int32), ol.HostFab(VALID.grown(1), 1, np.int32)
    nbad_o = ol.oracle().orc_predict_step(C.byref(prm), C.byref(uo.desc()), C.byref(no.desc()), C.byref(ro.desc()), dt, nv, C.byref(fo.desc()),
                                          C.byref(VALID))
    nbad_r = C.c_double(0)
    ol.ref().ref_update_op(problem, 2, C.byref(VALID), C.byref(uo.desc()), C.byref(nr.desc()), C.byref(ro.desc()), None, None, None,
                           C.byref(fr.desc()), dx, dt, 0.0, 0.0, C.byref(nbad_r))
    exact(no.a, nr.a)
    assert (fo.a == fr.a).all() and nbad_o == int(nbad_r.value) and nbad_o > 0
    # EnforceLimits (with floors that bite) and SyncDualEnergy on a valid state
    st_o, st_r = ol.HostFab(VALID, nv), ol.HostFab(VALID, nv)
    st_o.a[...] = cons.view(VALID)
    st_o.a[0].flat[::7] *= 1e-3  # below the density floor
    st_o.a[4].flat[::5] *= 1e-6  # cold cells: temperature floor / dual energy branch
    st_r.a[...] = st_o.a
    prm.density_floor = 0.05
    prm.temp_floor = 1e14 if problem == 0 else 3e15
    ol.oracle().orc_enforce_limits(C.byref(prm), C.byref(st_o.desc()), C.byref(VALID))
    ol.ref().ref_update_op(problem, 3, C.byref(VALID), C.byref(st_r.desc()), None, None, None, None, None, None, dx, 0.0, prm.density_floor,
                           prm.temp_floor, None)
    exact(st_o.a, st_r.a)
    ol.oracle().orc_sync_dual_energy(C.byref(prm), C.byref(st_o.desc()), C.byref(VALID))
    ol.ref().ref_update_op(problem, 4, C.byref(VALID), C.byref(st_r.desc()), None, None, None, None, None, None, dx, 0.0, 0.0, 0.0, None)
    exact(st_o.a, st_r.a)
    # signal speeds
    for which, op in ((0, 5), (1, 6)):
        m_r = C.c_double(0)
        cv = ol.HostFab(VALID, nv)
        cv.a[...] = cons.view(VALID)
        m_o = ol.oracle().orc_max_signal_speed(C.byref(prm), which, C.byref(cv.desc()), C.byref(VALID))
        ol.ref().ref_update_op(problem, op, C.byref(VALID), C.byref(cv.desc()), None, None, None, None, None, None, dx, 0.0, 0.0, 0.0, C.byref(m_r))
        assert m_o == m_r.value
