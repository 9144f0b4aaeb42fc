"""The fused sweep kernels (csrc/qk_sweep.cu: shuffle x-sweep, marching y/z sweeps with the stage epilogue) against
the faithful per-operator path and the oracle.  Bar: BIT-EXACT new state after each RK stage, on ragged box sizes
(not multiples of the 30-cell x tile, the 32-lane warp or the 32-cell marching segment), several boxes of different
shapes in one launch, every instantiated trait set, both integrator orders.  Also checks that the production entry
really ran the fused kernels (kernel-class profiler) and that a flagged cell hands the stage to the faithful path."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check, qk_box
from test_gpu_level import GenericProblem, exact, level_desc, oracle_level, oracle_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load()


class RaggedProblem(GenericProblem):
    """boxes of unequal size: the domain is cut at arbitrary planes"""

    def __init__(self, ncell, cuts, periodic, bc_kind, nscalars=0, gamma=1.4):
        super().__init__(ncell, max(ncell), periodic, bc_kind, nscalars, gamma)
        edges = [[0] + list(c) + [n] for c, n in zip(cuts, ncell)]
        self.boxes = []
        for kz in range(len(edges[2]) - 1):
            for jy in range(len(edges[1]) - 1):
                for ix in range(len(edges[0]) - 1):
                    self.boxes.append(qk_box.make((edges[0][ix], edges[1][jy], edges[2][kz]),
                                                  (edges[0][ix + 1] - 1, edges[1][jy + 1] - 1, edges[2][kz + 1] - 1)))


def prof(lib):
    buf = (C.c_char * 8192)()
    lib.qk_prof_report(buf, 8192)
    return {ln.split()[0]: int(ln.split()[1]) for ln in buf.value.decode().splitlines()}


def run_pair(lib, p, prm, st, dt, entry, order=2):
    from quokka_b200.device import DevMultiFab

    desc, keep = level_desc(p)
    lev = C.c_void_p()
    check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
    U0 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost, host=st)
    U1 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost)
    U2 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost)
    b1, b2 = C.c_int64(-1), C.c_int64(-1)
    check(lib.qk_fill_boundary(lev, U0.descs, 0, p.ncomp, None))
    check(entry(lev, C.byref(prm), 1, U0.descs, U0.descs, U1.descs, dt, C.byref(b1), None))
    out1 = U1.numpy()
    out2 = None
    if order == 2:
        check(lib.qk_fill_boundary(lev, U1.descs, 0, p.ncomp, None))
        check(entry(lev, C.byref(prm), 2, U0.descs, U1.descs, U2.descs, dt, C.byref(b2), None))
        out2 = U2.numpy()
    lib.qk_level_destroy(lev)
    return out1, out2, b1.value, b2.value


CASES = {
    # name: (ncell, cuts, periodic, bc, nscalars, nmscalars, reint, gamma)
    "ragged_3boxes": ((71, 37, 45), ((33,), (), (20,)), (0, 0, 0), "reflect", 0, 0, 0, 1.4),
    "thin_boxes": ((64, 8, 40), ((), (4,), (33,)), (1, 1, 1), "periodic", 0, 0, 0, 1.4),
    "eint_outflow": ((35, 66, 12), ((), (31,), ()), (0, 0, 0), "outflow", 0, 0, 1, 5.0 / 3.0),
    "one_scalar": ((40, 33, 34), ((), (), ()), (1, 0, 1), "reflect", 1, 0, 0, 1.4),
    "mass_scalars_eint": ((34, 35, 36), ((16,), (), ()), (1, 1, 0), "reflect", 3, 2, 1, 5.0 / 3.0),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_fused_equals_faithful_and_oracle(lib, case, kind):
    ncell, cuts, periodic, bc, ns, nms, reint, gamma = CASES[case]
    p = RaggedProblem(ncell, cuts, periodic, bc, nscalars=ns, gamma=gamma)
    prm = p.params(nmscalars=nms, reconstruct_eint=reint)
    st = p.states(seed=11, kind=kind)
    dt = 1.0e-4 if kind == "shocked" else 3.0e-4
    if case == "mass_scalars_eint" and kind == "shocked":
        dt = 1.0e-5  # random mass scalars next to 50x pressure jumps go negative (and get flagged) at larger dt
    lib.qk_prof_enable(1)
    f1, f2, fb1, fb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert counts.get("sweep_x", 0) == 2 and counts.get("sweep_z", 0) == 2, f"fused kernels did not run: {counts}"
    assert counts.get("flux_function", 0) == 0, "the faithful path ran although no cell was flagged"
    g1, g2, gb1, gb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
    assert (fb1, fb2, gb1, gb2) == (0, 0, 0, 0)
    ng = p.nghost
    for b in range(len(p.boxes)):
        exact(f1[b][:, ng:-ng, ng:-ng, ng:-ng], g1[b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 1 box {b}")
        exact(f2[b][:, ng:-ng, ng:-ng, ng:-ng], g2[b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 2 box {b}")
    # and against the CPU oracle's level driver
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    ok = o.orc_advance_hydro_level(L, C.byref(prm), dt, 1.0e9, None, None)
    assert ok == 1
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)
        exact(f2[b][:, ng:-ng, ng:-ng, ng:-ng], ref[:, ng:-ng, ng:-ng, ng:-ng], f"oracle box {b}")
    o.orc_level_destroy(L)


def test_fused_forward_euler(lib):
    ncell, cuts, periodic, bc, ns, nms, reint, gamma = CASES["ragged_3boxes"]
    p = RaggedProblem(ncell, cuts, periodic, bc)
    prm = p.params()
    prm.integrator_order = 1
    st = p.states(seed=4, kind="smooth")
    f1, _, fb1, _ = run_pair(lib, p, prm, st, 2.0e-4, lib.qk_hydro_advance_stage, order=1)
    g1, _, gb1, _ = run_pair(lib, p, prm, st, 2.0e-4, lib.qk_hydro_advance_stage_faithful, order=1)
    ng = p.nghost
    for b in range(len(p.boxes)):
        exact(f1[b][:, ng:-ng, ng:-ng, ng:-ng], g1[b][:, ng:-ng, ng:-ng, ng:-ng], f"box {b}")


def test_flagged_stage_is_redone_by_the_faithful_path(lib):
    """large dt on a violent state: PredictStep flags cells, the fused stage must hand over to FOFC"""
    p = GenericProblem((32, 32, 32), 16, (1, 1, 1), "periodic")
    prm = p.params()
    prm.abort_on_fofc_failure = 0
    st = p.states(seed=9, kind="shocked")
    dt = 2.0e-3
    lib.qk_prof_enable(1)
    f1, f2, fb1, fb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert counts.get("flux_function", 0) > 0 and counts.get("replace_fluxes", 0) > 0
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    o.orc_advance_hydro_level(L, C.byref(prm), dt, 1.0e9, None, None)
    ng = p.nghost
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)
        exact(f2[b][:, ng:-ng, ng:-ng, ng:-ng], ref[:, ng:-ng, ng:-ng, ng:-ng], f"box {b}")
    o.orc_level_destroy(L)


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("arith", ["exact", "relaxed"])
def test_lower_order_reconstruction_through_the_stage_entry(lib, order, arith):
    """reconstructionOrder_ = 2 (PLM, minmod; config C4's hydro) runs through a compile-time PLM instantiation of the fused sweeps
    (scalar-free, reconstruct_eint = false trait set): bit-exact against the oracle in the exact mode, within 2e-14 of max|U| per
    component in the relaxed mode.  Order 1 (donor cell) takes the faithful per-operator path in either mode and is bit-exact."""
    ncell, cuts, periodic, bc, ns, nms, reint, gamma = CASES["thin_boxes"]
    p = RaggedProblem(ncell, cuts, periodic, bc, nscalars=ns, gamma=gamma)
    prm = p.params(nmscalars=nms, reconstruct_eint=reint, recon_order=order, arith=capi.QK_ARITH_FAST if arith == "relaxed" else capi.QK_ARITH_EXACT)
    st = p.states(seed=21, kind="shocked")
    dt = 1.0e-4
    lib.qk_prof_enable(1)
    f1, f2, fb1, fb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert (fb1, fb2) == (0, 0)
    fused = counts.get("sweep_x", 0) == 2 and counts.get("flux_function", 0) == 0
    assert fused == (order == 2), counts
    prm_exact = p.params(nmscalars=nms, reconstruct_eint=reint, recon_order=order)
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    assert o.orc_advance_hydro_level(L, C.byref(prm_exact), dt, 1.0e9, None, None) == 1
    ng = p.nghost
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)[:, ng:-ng, ng:-ng, ng:-ng]
        got = f2[b][:, ng:-ng, ng:-ng, ng:-ng]
        if arith == "exact" or order == 1:
            exact(got, ref, f"order {order} box {b}")
        else:
            scale = np.abs(ref).reshape(p.ncomp, -1).max(axis=1)
            assert (np.abs(got - ref).reshape(p.ncomp, -1).max(axis=1) / scale < 2e-14).all()
            assert not np.array_equal(got, ref)
    o.orc_level_destroy(L)


def test_artificial_viscosity_takes_the_operator_path(lib):
    """artificialViscosityK_ != 0: the production entry declines the fused sweeps (one stderr notice) and the one-kernel-per-operator path
    reproduces the oracle's level driver bit for bit"""
    p = GenericProblem((32, 32, 32), 16, (1, 1, 1), "periodic")
    prm = p.params()
    prm.K_visc = 0.1
    st = p.states(seed=21, kind="shocked")
    dt = 1.0e-4
    lib.qk_prof_enable(1)
    f1, f2, fb1, fb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert counts.get("sweep_x", 0) == 0 and counts.get("flux_function", 0) > 0
    L, keep = oracle_level(p, st)
    o = ol.oracle()
    ok = o.orc_advance_hydro_level(L, C.byref(prm), dt, 1.0e9, None, None)
    assert ok == 1
    ng = p.nghost
    for b in range(len(p.boxes)):
        ref = oracle_state(p, L, 0, b)
        exact(f2[b][:, ng:-ng, ng:-ng, ng:-ng], ref[:, ng:-ng, ng:-ng, ng:-ng], f"oracle box {b}")
    o.orc_level_destroy(L)
    # and it is not a no-op
    prm0 = p.params()
    g1, g2, _, _ = run_pair(lib, p, prm0, st, dt, lib.qk_hydro_advance_stage)
    assert any((a[:, ng:-ng, ng:-ng, ng:-ng] != b[:, ng:-ng, ng:-ng, ng:-ng]).any() for a, b in zip(f2, g2))


XCAT_CASES = {
    # name: (ncell, cuts, periodic, bc, nscalars, nmscalars, reint)
    "row128": ((128, 12, 10), ((), (), ()), (0, 0, 0), "reflect", 0, 0, 0),       # 130 slots per row: tiles run over row ends every 4.33 tiles
    "amr32": ((64, 32, 32), ((32,), (16,), ()), (1, 1, 1), "periodic", 0, 0, 0),  # 34 slots per row: almost every tile covers two rows
    "ragged": ((72, 37, 45), ((34,), (), (20,)), (0, 0, 0), "reflect", 1, 0, 1),  # nx = 34 and 38 (even: TMA-staged), a scalar, reconstruct_eint
    "nx30": ((30, 9, 7), ((), (), ()), (0, 0, 0), "outflow", 0, 0, 0),            # the narrowest box the concatenated kernel takes
}


@pytest.mark.parametrize("case", sorted(XCAT_CASES))
@pytest.mark.parametrize("arith", ["exact", "relaxed"])
def test_concatenated_x_sweep_equals_per_row_tiles(lib, case, arith, monkeypatch):
    """k_sweep_xc (rows of a box laid end to end, 30-slot tiles that may cover the end of one row and the start of the next; the x sweep
    every box >= 30 cells wide takes) against k_sweep_xt (QK_XCAT=0: 30-cell tiles per row): the same arithmetic per cell, so the new
    state is BIT-IDENTICAL after both stages in either arithmetic mode."""
    ncell, cuts, periodic, bc, ns, nms, reint = XCAT_CASES[case]
    p = RaggedProblem(ncell, cuts, periodic, bc, nscalars=ns)
    prm = p.params(nmscalars=nms, reconstruct_eint=reint, arith=capi.QK_ARITH_FAST if arith == "relaxed" else capi.QK_ARITH_EXACT)
    st = p.states(seed=31, kind="shocked")
    dt = 1.0e-4
    monkeypatch.delenv("QK_XCAT", raising=False)
    monkeypatch.setenv("QK_XCAT_MIN_TILES", "0")  # levels this small keep the per-row tiles by default (qk_sweep.cu): force the concatenated sweep
    lib.qk_prof_enable(1)
    a1, a2, ab1, ab2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert counts.get("sweep_x", 0) == 2 and counts.get("flux_function", 0) == 0, counts
    monkeypatch.setenv("QK_XCAT", "0")
    b1, b2, bb1, bb2 = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    assert (ab1, ab2, bb1, bb2) == (0, 0, 0, 0)
    ng = p.nghost
    for b in range(len(p.boxes)):
        exact(a1[b][:, ng:-ng, ng:-ng, ng:-ng], b1[b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 1 box {b}")
        exact(a2[b][:, ng:-ng, ng:-ng, ng:-ng], b2[b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 2 box {b}")
    if arith == "exact":  # and the reference's bits
        g1, g2, _, _ = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
        for b in range(len(p.boxes)):
            exact(a2[b][:, ng:-ng, ng:-ng, ng:-ng], g2[b][:, ng:-ng, ng:-ng, ng:-ng], f"faithful box {b}")
