"""GPU parity of the AMR transfer operators: qk_amr_interp_cons_lin_minmax and qk_amr_average_down through the C ABI
against the oracle (pinned bit-exactly to AMReX's mf_linear_slope_minmax_interp / average_down by tests/test_oracle_amr_transfer.py;
the kernels' arithmetic is checked on the CPU by tests/test_amr_host.py).  No libm call, no contraction: the bar is bit-exact.
(Written after the round's GPU budget was spent: this file sorts last and runs on a GPU for the first time at round end.)"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check, qk_array4, qk_box
from test_amr_host import interp_inputs, prepost_state
from test_oracle_amr_transfer import CASES

pytestmark = pytest.mark.gpu


def dev(fab: ol.HostFab):
    from quokka_b200.device import DevFab

    return DevFab(fab.box, fab.ncomp, host=fab.a)


@pytest.mark.parametrize("case", range(len(CASES)))
def test_interp_vs_oracle(case):
    lib = capi.load()
    ncomp, extra = 6, 2
    fine_region, cdomain, dest, fab_box, crse, r, lo, hi = interp_inputs(case, ncomp, extra)
    want = ol.HostFab(fab_box, ncomp + extra, fill=-7.0)
    ol.oracle().orc_interp_cons_lin_minmax(C.byref(crse.desc()), extra, C.byref(want.desc()), extra, ncomp, C.byref(fine_region), C.byref(dest),
                                           C.byref(cdomain), r, lo, hi)
    dc, df = dev(crse), dev(ol.HostFab(fab_box, ncomp + extra, fill=-7.0))
    cd, fd = (qk_array4 * 1)(dc.desc()), (qk_array4 * 1)(df.desc())
    reg = (qk_box * 1)(fine_region)
    n0 = lib.qk_launch_count()
    check(lib.qk_amr_interp_cons_lin_minmax(1, cd, extra, fd, extra, ncomp, reg, C.byref(dest), C.byref(cdomain), r, lo, hi, None))
    got = df.numpy()
    assert lib.qk_launch_count() - n0 == 1
    assert np.array_equal(got, want.a)


def test_interp_many_patches_one_call():
    """20 box pairs (more than one kernel-parameter table holds): the ghost shell of a fine level, each pair with its own FABs"""
    lib = capi.load()
    ncomp = 6
    cdomain = qk_box.make((0, 0, 0), (31, 31, 31))
    ratio = (2, 2, 2)
    r = (C.c_int * 3)(*ratio)
    lo = (C.c_int32 * (3 * ncomp))(*([capi.QK_BC_INT_DIR] * (3 * ncomp)))
    dest = qk_box.make((0, 0, 0), (63, 63, 63))
    rng = np.random.default_rng(9)
    crse_h, fine_h, want, regions = [], [], [], []
    for p in range(20):
        o = [int(x) for x in rng.integers(4, 40, 3)]
        n = [int(x) for x in rng.integers(1, 12, 3)]
        region = qk_box.make(tuple(o), tuple(o[d] + n[d] - 1 for d in range(3)))
        cb = qk_box.make(tuple(region.lo[d] // 2 - 1 for d in range(3)), tuple(region.hi[d] // 2 + 1 for d in range(3)))
        c = ol.HostFab(cb, ncomp)
        c.a[...] = rng.uniform(0.1, 10.0, c.a.shape)
        w = ol.HostFab(region.grown(1), ncomp, fill=-1.0)
        ol.oracle().orc_interp_cons_lin_minmax(C.byref(c.desc()), 0, C.byref(w.desc()), 0, ncomp, C.byref(region), C.byref(dest), C.byref(cdomain), r, lo, lo)
        crse_h.append(c)
        fine_h.append(ol.HostFab(region.grown(1), ncomp, fill=-1.0))
        want.append(w)
        regions.append(region)
    dcs, dfs = [dev(c) for c in crse_h], [dev(f) for f in fine_h]
    cd = (qk_array4 * 20)(*[d.desc() for d in dcs])
    fd = (qk_array4 * 20)(*[d.desc() for d in dfs])
    reg = (qk_box * 20)(*regions)
    n0 = lib.qk_launch_count()
    check(lib.qk_amr_interp_cons_lin_minmax(20, cd, 0, fd, 0, ncomp, reg, C.byref(dest), C.byref(cdomain), r, lo, lo, None))
    assert lib.qk_launch_count() - n0 == 2
    for d, w in zip(dfs, want):
        assert np.array_equal(d.numpy(), w.a)


@pytest.mark.parametrize("ratio", [(2, 2, 2), (4, 4, 4), (2, 1, 4)])
def test_average_down_vs_oracle(ratio):
    lib = capi.load()
    cbx = qk_box.make((2, -3, 1), (9, 4, 6))
    fb = qk_box.make(tuple(cbx.lo[d] * ratio[d] for d in range(3)), tuple((cbx.hi[d] + 1) * ratio[d] - 1 for d in range(3)))
    fine = ol.HostFab(fb.grown(1), 4)
    fine.a[...] = np.random.default_rng(6).uniform(-1.0, 10.0, fine.a.shape)
    want = ol.HostFab(cbx.grown(1), 4, fill=3.0)
    r = (C.c_int * 3)(*ratio)
    ol.oracle().orc_average_down(C.byref(want.desc()), 1, C.byref(fine.desc()), 1, 3, C.byref(cbx), r)
    dc, df = dev(ol.HostFab(cbx.grown(1), 4, fill=3.0)), dev(fine)
    check(lib.qk_amr_average_down(1, (qk_array4 * 1)(dc.desc()), 1, (qk_array4 * 1)(df.desc()), 1, 3, (qk_box * 1)(cbx), r, None))
    assert np.array_equal(dc.numpy(), want.a)


def test_interp_then_average_down_is_conservative_at_full_size():
    """size-independent property on a 128^3 coarse box refined by 2 (256^3 fine cells): average_down(interp(U)) == U to rounding, no new
    extrema, and components that are equal on the coarse level stay equal on the fine level"""
    import torch

    from quokka_b200.device import DevFab

    lib = capi.load()
    ncomp = 6
    cdomain = qk_box.make((0, 0, 0), (127, 127, 127))
    cb = cdomain.grown(1)
    fine_region = qk_box.make((0, 0, 0), (255, 255, 255))
    dc, df, back = DevFab(cb, ncomp), DevFab(fine_region, ncomp), DevFab(cdomain, ncomp)
    g = torch.Generator(device="cuda").manual_seed(4)
    dc.t.copy_(torch.rand(dc.t.shape, generator=g, device="cuda", dtype=torch.float64) * 9.9 + 0.1)
    dc.t[5] = dc.t[0]
    r = (C.c_int * 3)(2, 2, 2)
    lo = (C.c_int32 * (3 * ncomp))(*([capi.QK_BC_INT_DIR] * (3 * ncomp)))
    check(lib.qk_amr_interp_cons_lin_minmax(1, (qk_array4 * 1)(dc.desc()), 0, (qk_array4 * 1)(df.desc()), 0, ncomp, (qk_box * 1)(fine_region),
                                            C.byref(fine_region), C.byref(cdomain), r, lo, lo, None))
    check(lib.qk_amr_average_down(1, (qk_array4 * 1)(back.desc()), 0, (qk_array4 * 1)(df.desc()), 0, ncomp, (qk_box * 1)(cdomain), r, None))
    torch.cuda.synchronize()
    inner = dc.t[:, 1:-1, 1:-1, 1:-1]
    assert float((back.t - inner).abs().max()) <= 1e-14 * 10.0
    assert bool(torch.equal(df.t[0], df.t[5]))
    assert float(df.t.min()) >= float(dc.t.min()) - 1e-13 and float(df.t.max()) <= float(dc.t.max()) + 1e-13  # rounding of uc + offsets only


@pytest.mark.parametrize("post", [0, 1])
def test_pre_post_interp_state_vs_oracle(post):
    lib = capi.load()
    bx, want = prepost_state()
    _, start = prepost_state()
    (ol.oracle().orc_post_interp_state if post else ol.oracle().orc_pre_interp_state)(C.byref(want.desc()), C.byref(bx))
    d = dev(start)
    fn = lib.qk_amr_post_interp_state if post else lib.qk_amr_pre_interp_state
    check(fn(1, (qk_box * 1)(bx), (qk_array4 * 1)(d.desc()), None))
    assert np.array_equal(d.numpy(), want.a)
