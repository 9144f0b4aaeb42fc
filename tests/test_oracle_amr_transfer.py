"""Pin the AMR transfer operators of the oracle (orc_interp_cons_lin_minmax, orc_average_down; SURVEY 8(f)2, restated for the next
round's kernels) against AMReX's own code compiled from /root/reference/extern/amrex: amrex::mf_linear_slope_minmax_interp (the
interpolater Quokka selects, src/simulation.hpp:1389-1407) and amrex::average_down.  Bar: bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200.capi import QK_BC_EXT_DIR, QK_BC_FOEXTRAP, QK_BC_INT_DIR, QK_BC_REFLECT_EVEN, qk_box

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libquokka_ref.so not built")


def coarse_box(fine: qk_box, ratio):
    lo = [fine.lo[d] // ratio[d] - (1 if ratio[d] > 1 else 0) for d in range(3)]
    hi = [fine.hi[d] // ratio[d] + (1 if ratio[d] > 1 else 0) for d in range(3)]
    return qk_box.make(tuple(lo), tuple(hi))


def exact(a, b):
    bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{bad.sum()} mismatches of {a.size}"


CASES = [
    # fine region, coarse domain, ratio, bc (lo, hi per dim), kind
    (qk_box.make((8, 8, 8), (23, 19, 15)), qk_box.make((0, 0, 0), (15, 15, 15)), (2, 2, 2), (QK_BC_INT_DIR, QK_BC_INT_DIR), "smooth"),
    (qk_box.make((8, 8, 8), (23, 19, 15)), qk_box.make((0, 0, 0), (15, 15, 15)), (2, 2, 2), (QK_BC_INT_DIR, QK_BC_INT_DIR), "shocked"),
    (qk_box.make((-4, 0, 4), (11, 7, 19)), qk_box.make((0, 0, 0), (15, 15, 15)), (2, 2, 2), (QK_BC_REFLECT_EVEN, QK_BC_FOEXTRAP), "shocked"),  # ghost region left of the domain
    (qk_box.make((0, 0, 0), (15, 11, 7)), qk_box.make((0, 0, 0), (15, 15, 15)), (4, 4, 4), (QK_BC_INT_DIR, QK_BC_INT_DIR), "shocked"),
    (qk_box.make((0, 4, 8), (11, 15, 23)), qk_box.make((0, 0, 0), (7, 7, 11)), (2, 2, 2), (QK_BC_EXT_DIR, QK_BC_EXT_DIR), "shocked"),  # one-sided slopes at the walls
    (qk_box.make((0, 0, 0), (15, 7, 7)), qk_box.make((0, 0, 0), (7, 7, 7)), (2, 1, 1), (QK_BC_INT_DIR, QK_BC_INT_DIR), "shocked"),  # refinement in x only
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_interpolater_bit_exact(case):
    fine_region, cdomain, ratio, (bl, bh), kind = CASES[case]
    ncomp = 6
    rng = np.random.default_rng(100 + case)
    cb = coarse_box(fine_region, ratio)
    crse = ol.HostFab(cb, ncomp)
    nz, ny, nx = cb.shape()
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for n in range(ncomp):
        if kind == "smooth":
            crse.a[n] = 1.0 + 0.3 * np.sin(0.7 * x + 0.2 * n) * np.cos(0.5 * y) + 0.1 * z
        else:
            crse.a[n] = rng.uniform(0.1, 10.0, (nz, ny, nx)) * np.where((x + y + z) % 5 == 0, 100.0, 1.0)
    crse.a[5] = crse.a[0]  # identical components: their slopes must stay identical (one limiter per direction)
    fdom = qk_box.make(tuple(cdomain.lo[d] * ratio[d] for d in range(3)), tuple((cdomain.hi[d] + 1) * ratio[d] - 1 for d in range(3)))
    dest = qk_box.make(tuple(min(fdom.lo[d], fine_region.lo[d]) for d in range(3)), tuple(max(fdom.hi[d], fine_region.hi[d]) for d in range(3)))
    if case == 2:
        dest = fdom  # part of the fine region lies outside the destination domain and must not be written
    a, b = ol.HostFab(fine_region, ncomp, fill=-7.0), ol.HostFab(fine_region, ncomp, fill=-7.0)
    r = (C.c_int * 3)(*ratio)
    lo = (C.c_int32 * (3 * ncomp))(*([bl] * (3 * ncomp)))
    hi = (C.c_int32 * (3 * ncomp))(*([bh] * (3 * ncomp)))
    ol.oracle().orc_interp_cons_lin_minmax(C.byref(crse.desc()), 0, C.byref(a.desc()), 0, ncomp, C.byref(fine_region), C.byref(dest), C.byref(cdomain), r, lo, hi)
    assert ol.ref().ref_interp_cons_lin_minmax(C.byref(crse.desc()), C.byref(b.desc()), ncomp, C.byref(fine_region), C.byref(dest), C.byref(cdomain), r,
                                               lo, hi) == 0
    exact(a.a, b.a)
    assert np.array_equal(a.a[0], a.a[5])
    if case == 2:
        assert (a.a[:, :, :, :4] == -7.0).all() and not (a.a[:, :, :, 4:] == -7.0).any()
    # conservation: the mean of the fine cells of a coarse cell is the coarse value
    if case in (0, 1, 3):
        inner_c = qk_box.make(tuple(fine_region.lo[d] // ratio[d] for d in range(3)), tuple(fine_region.hi[d] // ratio[d] for d in range(3)))
        avg = ol.HostFab(inner_c, ncomp)
        ol.oracle().orc_average_down(C.byref(avg.desc()), 0, C.byref(a.desc()), 0, ncomp, C.byref(inner_c), r)
        sl = tuple(slice(inner_c.lo[d] - cb.lo[d], inner_c.hi[d] - cb.lo[d] + 1) for d in (2, 1, 0))
        ref_c = crse.a[(slice(None),) + sl]
        assert np.abs(avg.a - ref_c).max() <= 4e-16 * np.abs(ref_c).max() * 8


@pytest.mark.parametrize("ratio", [(2, 2, 2), (4, 4, 4), (2, 1, 4)])
def test_average_down_bit_exact(ratio):
    cbx = qk_box.make((2, -3, 1), (9, 4, 6))
    fb = qk_box.make(tuple(cbx.lo[d] * ratio[d] for d in range(3)), tuple((cbx.hi[d] + 1) * ratio[d] - 1 for d in range(3)))
    ncomp = 3
    fine = ol.HostFab(fb, ncomp)
    fine.a[...] = np.random.default_rng(5).uniform(-1.0, 10.0, fine.a.shape)
    a, b = ol.HostFab(cbx, ncomp), ol.HostFab(cbx, ncomp)
    r = (C.c_int * 3)(*ratio)
    ol.oracle().orc_average_down(C.byref(a.desc()), 0, C.byref(fine.desc()), 0, ncomp, C.byref(cbx), r)
    assert ol.ref().ref_average_down(C.byref(b.desc()), C.byref(fine.desc()), ncomp, C.byref(cbx), r) == 0
    exact(a.a, b.a)


@pytest.mark.parametrize("post", [0, 1])
def test_pre_post_interp_state_bit_exact(post):
    """QuokkaSimulation<P0>::PreInterpState / PostInterpState (src/QuokkaSimulation.hpp:804-841) called as static members"""
    bx = qk_box.make((3, -2, 5), (18, 9, 12))
    a, b = ol.HostFab(bx, 6), ol.HostFab(bx, 6)
    rng = np.random.default_rng(8)
    a.a[0] = rng.uniform(0.1, 10.0, a.a[0].shape)
    a.a[1:4] = rng.uniform(-3.0, 3.0, a.a[1:4].shape) * a.a[0]
    a.a[4] = rng.uniform(0.1, 10.0, a.a[0].shape) + 0.5 * (a.a[1:4] ** 2).sum(0) / a.a[0]
    a.a[5] = rng.uniform(0.1, 10.0, a.a[0].shape)
    if post:
        a.a[4] = rng.uniform(0.1, 10.0, a.a[0].shape)  # a specific internal energy
    b.a[...] = a.a
    before = a.a.copy()
    (ol.oracle().orc_post_interp_state if post else ol.oracle().orc_pre_interp_state)(C.byref(a.desc()), C.byref(bx))
    assert ol.ref().ref_pre_post_interp_state(post, C.byref(bx), C.byref(b.desc())) == 0
    exact(a.a, b.a)
    assert np.array_equal(a.a[[0, 1, 2, 3, 5]], before[[0, 1, 2, 3, 5]]) and not np.array_equal(a.a[4], before[4])


def test_interpolater_random_regions_bit_exact():
    """fine regions with odd bounds and negative indices, down to one cell, ratio 2 and 4, against AMReX"""
    rng = np.random.default_rng(12)
    ncomp = 5
    for trial in range(30):
        ratio = (2, 2, 2) if trial % 2 == 0 else (4, 4, 4)
        o = [int(x) for x in rng.integers(-9, 30, 3)]
        n = [int(x) for x in rng.integers(1, 11, 3)]
        region = qk_box.make(tuple(o), tuple(o[d] + n[d] - 1 for d in range(3)))
        cb = coarse_box(region, ratio)
        cdomain = qk_box.make((-8, -8, -8), (23, 23, 23))
        dest = qk_box.make((-8 * ratio[0],) * 3, (24 * ratio[0] - 1,) * 3)
        c = ol.HostFab(cb, ncomp)
        c.a[...] = rng.uniform(0.1, 10.0, c.a.shape)
        a, b = ol.HostFab(region, ncomp, fill=-1.0), ol.HostFab(region, ncomp, fill=-1.0)
        r = (C.c_int * 3)(*ratio)
        bc = (C.c_int32 * (3 * ncomp))()
        ol.oracle().orc_interp_cons_lin_minmax(C.byref(c.desc()), 0, C.byref(a.desc()), 0, ncomp, C.byref(region), C.byref(dest), C.byref(cdomain), r, bc, bc)
        assert ol.ref().ref_interp_cons_lin_minmax(C.byref(c.desc()), C.byref(b.desc()), ncomp, C.byref(region), C.byref(dest), C.byref(cdomain), r, bc, bc) == 0
        exact(a.a, b.a)
