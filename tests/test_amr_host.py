"""The arithmetic of the AMR transfer kernels without a GPU: quokka_b200/csrc/qk_amr.cuh (bodies of k_amr_interp / k_amr_avgdown:
one thread per coarse cell, limiter pass + write pass, no slope array) compiled for the host (tests/host_src/amr_host.cpp) against
the oracle (orc_interp_cons_lin_minmax / orc_average_down, pinned bit-exactly to AMReX by tests/test_oracle_amr_transfer.py).
No libm call is involved, so the bar is bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import qk_box
from test_oracle_amr_transfer import CASES, coarse_box

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_src", "amr_host.cpp")
HDR = os.path.join(HERE, "..", "quokka_b200", "csrc", "qk_amr.cuh")
SO = os.path.join(HERE, "host_src", "_build", "libamr_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC])
    lib = C.CDLL(SO)
    P = C.POINTER
    lib.host_amr_interp.argtypes = [P(capi.qk_array4), C.c_int, P(capi.qk_array4), C.c_int, C.c_int, P(qk_box), P(qk_box), P(qk_box), P(C.c_int),
                                    P(C.c_int32), P(C.c_int32)]
    lib.host_amr_prepost.argtypes = [C.c_int, P(capi.qk_array4), P(qk_box)]
    lib.host_amr_average_down.argtypes = [P(capi.qk_array4), C.c_int, P(capi.qk_array4), C.c_int, C.c_int, P(qk_box), P(C.c_int)]
    return lib


def interp_inputs(case, ncomp=6, extra_comps=2):
    """coarse FAB with `extra_comps` leading components that are not interpolated (ccomp = fcomp = extra_comps)"""
    fine_region, cdomain, ratio, (bl, bh), kind = CASES[case]
    rng = np.random.default_rng(300 + case)
    cb = coarse_box(fine_region, ratio)
    crse = ol.HostFab(cb, ncomp + extra_comps)
    crse.a[...] = rng.uniform(0.1, 10.0, crse.a.shape)
    if kind == "smooth":
        nz, ny, nx = cb.shape()
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        for n in range(ncomp + extra_comps):
            crse.a[n] = 1.0 + 0.3 * np.sin(0.7 * x + 0.2 * n) * np.cos(0.5 * y) + 0.1 * z
    fdom = qk_box.make(tuple(cdomain.lo[d] * ratio[d] for d in range(3)), tuple((cdomain.hi[d] + 1) * ratio[d] - 1 for d in range(3)))
    dest = fdom if case == 2 else qk_box.make(tuple(min(fdom.lo[d], fine_region.lo[d]) for d in range(3)),
                                              tuple(max(fdom.hi[d], fine_region.hi[d]) for d in range(3)))
    # the fine FAB is larger than the region that is filled (ghost cells of a fine patch)
    fab_box = fine_region.grown(2)
    r = (C.c_int * 3)(*ratio)
    lo = (C.c_int32 * (3 * ncomp))(*([bl] * (3 * ncomp)))
    hi = (C.c_int32 * (3 * ncomp))(*([bh] * (3 * ncomp)))
    return fine_region, cdomain, dest, fab_box, crse, r, lo, hi


@pytest.mark.parametrize("case", range(len(CASES)))
def test_interp_kernel_arithmetic_bit_exact(host, case):
    ncomp, extra = 6, 2
    fine_region, cdomain, dest, fab_box, crse, r, lo, hi = interp_inputs(case, ncomp, extra)
    a, b = ol.HostFab(fab_box, ncomp + extra, fill=-7.0), ol.HostFab(fab_box, ncomp + extra, fill=-7.0)
    ol.oracle().orc_interp_cons_lin_minmax(C.byref(crse.desc()), extra, C.byref(a.desc()), extra, ncomp, C.byref(fine_region), C.byref(dest), C.byref(cdomain),
                                           r, lo, hi)
    host.host_amr_interp(C.byref(crse.desc()), extra, C.byref(b.desc()), extra, ncomp, C.byref(fine_region), C.byref(dest), C.byref(cdomain), r, lo, hi)
    assert np.array_equal(a.a, b.a)
    assert (b.a[:extra] == -7.0).all()  # components outside [fcomp, fcomp + ncomp) untouched
    assert not (b.a[extra:, 2:-2, 2:-2, 2:-2] == -7.0).all()


@pytest.mark.parametrize("ratio", [(2, 2, 2), (4, 4, 4), (2, 1, 4)])
def test_avgdown_kernel_arithmetic_bit_exact(host, ratio):
    cbx = qk_box.make((2, -3, 1), (9, 4, 6))
    fb = qk_box.make(tuple(cbx.lo[d] * ratio[d] for d in range(3)), tuple((cbx.hi[d] + 1) * ratio[d] - 1 for d in range(3)))
    fine = ol.HostFab(fb.grown(1), 4)
    fine.a[...] = np.random.default_rng(6).uniform(-1.0, 10.0, fine.a.shape)
    a, b = ol.HostFab(cbx.grown(1), 4, fill=3.0), ol.HostFab(cbx.grown(1), 4, fill=3.0)
    r = (C.c_int * 3)(*ratio)
    ol.oracle().orc_average_down(C.byref(a.desc()), 1, C.byref(fine.desc()), 1, 3, C.byref(cbx), r)
    host.host_amr_average_down(C.byref(b.desc()), 1, C.byref(fine.desc()), 1, 3, C.byref(cbx), r)
    assert np.array_equal(a.a, b.a)
    assert (b.a[0] == 3.0).all() and not (b.a[1:, 1:-1, 1:-1, 1:-1] == 3.0).any()


def prepost_state(seed=8):
    bx = qk_box.make((3, -2, 5), (18, 9, 12))
    f = ol.HostFab(bx.grown(2), 10, fill=1.0)
    rng = np.random.default_rng(seed)
    f.a[0] = rng.uniform(0.1, 10.0, f.a[0].shape)
    f.a[1:4] = rng.uniform(-3.0, 3.0, f.a[1:4].shape) * f.a[0]
    f.a[4] = rng.uniform(0.1, 10.0, f.a[0].shape) + 0.5 * (f.a[1:4] ** 2).sum(0) / f.a[0]
    return bx, f


@pytest.mark.parametrize("post", [0, 1])
def test_prepost_kernel_arithmetic_bit_exact(host, post):
    bx, a = prepost_state()
    _, b = prepost_state()
    (ol.oracle().orc_post_interp_state if post else ol.oracle().orc_pre_interp_state)(C.byref(a.desc()), C.byref(bx))
    host.host_amr_prepost(post, C.byref(b.desc()), C.byref(bx))
    assert np.array_equal(a.a, b.a)


def test_pre_then_post_restores_the_energy(host):
    bx, a = prepost_state()
    before = a.a.copy()
    host.host_amr_prepost(0, C.byref(a.desc()), C.byref(bx))
    host.host_amr_prepost(1, C.byref(a.desc()), C.byref(bx))
    assert np.abs(a.a[4] / before[4] - 1).max() < 1e-14 and np.array_equal(np.delete(a.a, 4, 0), np.delete(before, 4, 0))


def test_interp_regions_that_cut_coarse_cells(host):
    """fine regions with odd bounds (a coarse cell contributes only some of its children), random sizes down to one cell, ratio 2 and 4"""
    rng = np.random.default_rng(12)
    ncomp = 5
    for trial in range(40):
        ratio = (2, 2, 2) if trial % 2 == 0 else (4, 4, 4)
        o = [int(x) for x in rng.integers(-9, 30, 3)]
        n = [int(x) for x in rng.integers(1, 11, 3)]
        region = qk_box.make(tuple(o), tuple(o[d] + n[d] - 1 for d in range(3)))
        cb = qk_box.make(tuple(region.lo[d] // ratio[d] - 1 for d in range(3)), tuple(region.hi[d] // ratio[d] + 1 for d in range(3)))
        cdomain = qk_box.make((-8, -8, -8), (23, 23, 23))
        dest = qk_box.make((-8 * ratio[0],) * 3, (24 * ratio[0] - 1,) * 3)
        c = ol.HostFab(cb, ncomp)
        c.a[...] = rng.uniform(0.1, 10.0, c.a.shape)
        a, b = ol.HostFab(region.grown(1), ncomp, fill=-1.0), ol.HostFab(region.grown(1), ncomp, fill=-1.0)
        r = (C.c_int * 3)(*ratio)
        bc = (C.c_int32 * (3 * ncomp))()
        ol.oracle().orc_interp_cons_lin_minmax(C.byref(c.desc()), 0, C.byref(a.desc()), 0, ncomp, C.byref(region), C.byref(dest), C.byref(cdomain), r, bc, bc)
        host.host_amr_interp(C.byref(c.desc()), 0, C.byref(b.desc()), 0, ncomp, C.byref(region), C.byref(dest), C.byref(cdomain), r, bc, bc)
        assert np.array_equal(a.a, b.a), (trial, o, n, ratio)
        assert not (b.a[:, 1:-1, 1:-1, 1:-1] == -1.0).any()
