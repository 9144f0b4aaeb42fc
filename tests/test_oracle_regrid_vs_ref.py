"""Pins the oracle's restatements of SURVEY 8(f)2's time interpolation and 8(f)4's regrid support to the reference itself (CPU, no GPU):
  * orc_time_interp           vs amrex::FillPatchSingleLevel (oracle/_ref/libquokka_ref.so: ref_time_interp) -- bit-exact;
  * orc_tag_pressure_gradient vs the reference's own QuokkaSimulation<SedovProblem>::ErrorEst, called on a QuokkaSimulation object built from
                              the reference's problem file (oracle/ref_build/errorest_harness.cpp -> libquokka_ref_sedov.so) -- identical tags;
  * orc_fixup_state           vs QuokkaSimulation<SedovProblem>::FixupState through the same probe -- bit-exact;
and checks the product's per-cell functions (quokka_b200/csrc/qk_amr.cuh, host build) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import hydro_params, qk_box

HERE = os.path.dirname(os.path.abspath(__file__))
SEDOV_SO = os.path.join(HERE, "..", "oracle", "_ref", "libquokka_ref_sedov.so")
P = C.POINTER


def sedov_like_state(n, ng, seed, kind):
    """6 conserved components on an n^3 box grown by ng: a blast-like pressure jump (tagged cells) on a quiet background"""
    bx = qk_box.make((0, 0, 0), (n - 1,) * 3)
    g = bx.grown(ng)
    rho, v, Pr, rng = ol.random_cons(g, 0, seed, kind)
    nz, ny, nx = g.shape()
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    r = np.sqrt((x - ng) ** 2 + (y - ng) ** 2 + (z - ng) ** 2)
    Pr = np.where(r < 0.4 * n, Pr * 30.0, 1.0e-3 * (1.0 + 0.05 * Pr))  # smooth ambient below/around P_min = 1e-3, steep edge at r = 0.4 n
    fab = ol.HostFab(g, 6)
    fab.a[...] = ol.cons_from_prim(rho, v, Pr, 1.4, rng)
    return bx, fab


@pytest.mark.parametrize("times", [(0.0, 1.0, 0.3), (2.0, 2.5, 2.125), (1.0, 3.0, 2.9999999), (0.0, 1.0, 1.0e-9)])
def test_time_interp_vs_amrex(times):
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = ol.ref()
    if not hasattr(ref, "ref_time_interp"):
        pytest.skip("libquokka_ref.so predates ref_time_interp (make -C oracle/ref_build harness)")
    t0, t1, t = times
    bx = qk_box.make((0, 0, 0), (11, 9, 7))
    rng = np.random.default_rng(5)
    s0, s1 = ol.HostFab(bx, 6), ol.HostFab(bx, 6)
    s0.a[...] = rng.uniform(-3, 3, s0.a.shape) * 10.0 ** rng.integers(-8, 8, s0.a.shape)
    s1.a[...] = rng.uniform(-3, 3, s1.a.shape) * 10.0 ** rng.integers(-8, 8, s1.a.shape)
    want, got = ol.HostFab(bx, 6), ol.HostFab(bx, 6)
    ref.ref_time_interp(C.byref(bx), C.byref(want.desc()), C.byref(s0.desc()), C.byref(s1.desc()), 6, t0, t1, t)
    which = ol.oracle().orc_time_interp(C.byref(got.desc()), 0, C.byref(s0.desc()), C.byref(s1.desc()), 0, 6, C.byref(bx), t0, t1, t)
    if which == 2:  # FillPatchSingleLevel interpolates whenever time != t0, t1; FillPatcher copies inside teps (checked below)
        assert np.array_equal(got.a, want.a)
    else:
        assert np.array_equal(got.a, s0.a if which == 0 else s1.a)


def test_time_interp_branches_follow_fillpatcher():
    o = ol.oracle()
    bx = qk_box.make((0, 0, 0), (3, 3, 3))
    s0, s1, d = ol.HostFab(bx, 1, fill=1.0), ol.HostFab(bx, 1, fill=2.0), ol.HostFab(bx, 1)
    call = lambda t: o.orc_time_interp(C.byref(d.desc()), 0, C.byref(s0.desc()), C.byref(s1.desc()), 0, 1, C.byref(bx), 10.0, 12.0, t)
    assert call(10.0) == 0 and call(10.0019) == 0 and call(9.9985) == 0  # teps = 2e-3
    assert call(12.0) == 1 and call(11.9985) == 1
    assert call(10.0021) == 2 and call(11.0) == 2
    assert d.a[0, 0, 0, 0] == 0.5 * 1.0 + 0.5 * 2.0
    assert o.orc_time_interp(C.byref(d.desc()), 0, C.byref(s0.desc()), None, 0, 1, C.byref(bx), 10.0, 12.0, 11.0) == 0


@pytest.fixture(scope="module")
def sedov_probe():
    if not os.path.exists(SEDOV_SO):
        pytest.skip("oracle/_ref/libquokka_ref_sedov.so not built (make -C oracle/ref_build errorest)")
    lib = C.CDLL(SEDOV_SO)
    lib.ref_sedov_error_est.argtypes = [C.c_int, P(capi.qk_array4), C.c_char_p]
    lib.ref_sedov_fixup_state.argtypes = [C.c_int, P(capi.qk_array4)]
    return lib


N_PROBE = 32  # one grid size per process (the probe's AmrCore geometry is fixed at first use)


@pytest.mark.parametrize("seed,kind", [(1, "smooth"), (2, "shocked"), (3, "smooth")])
def test_sedov_error_est_vs_reference(sedov_probe, seed, kind):
    bx, fab = sedov_like_state(N_PROBE, 1, seed, kind)
    want = C.create_string_buffer(N_PROBE ** 3)
    assert sedov_probe.ref_sedov_error_est(N_PROBE, C.byref(fab.desc()), want) == 0
    got = C.create_string_buffer(N_PROBE ** 3)
    prm = hydro_params(gamma=1.4, reconstruct_eint=0)
    ol.oracle().orc_tag_pressure_gradient(C.byref(prm), C.byref(fab.desc()), got, C.byref(bx), 0.1, 1.0e-3)
    w, g = np.frombuffer(want.raw, dtype=np.int8), np.frombuffer(got.raw, dtype=np.int8)
    assert 0 < (w != 0).sum() < w.size  # the input exercises both outcomes
    assert np.array_equal(w, g)


def test_fixup_state_vs_reference(sedov_probe):
    bx, fab = sedov_like_state(N_PROBE, 0, 11, "shocked")
    # cells where the dual-energy switch goes either way
    fab.a[5] *= np.where(np.arange(fab.a[5].size).reshape(fab.a[5].shape) % 3 == 0, 1.0e-6, 1.0)
    want, got = ol.HostFab(bx, 6), ol.HostFab(bx, 6)
    want.a[...] = fab.a
    got.a[...] = fab.a
    assert sedov_probe.ref_sedov_fixup_state(N_PROBE, C.byref(want.desc())) == 0
    prm = hydro_params(gamma=1.4, reconstruct_eint=0)
    ol.oracle().orc_fixup_state(C.byref(prm), C.byref(got.desc()), C.byref(bx))
    assert not np.array_equal(want.a, fab.a)
    assert np.array_equal(want.a, got.a)


# ---- the product's per-cell functions (host build of quokka_b200/csrc/qk_amr.cuh) against the oracle ----------------------------------
def test_kernel_cell_functions_on_host():
    from test_amr_host import host as host_fixture  # noqa: F401  (build recipe)
    import subprocess

    from test_amr_host import HDR, SO, SRC

    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.host_time_interp.argtypes = [P(capi.qk_array4)] * 3 + [C.c_int, P(qk_box), C.c_double, C.c_double, C.c_double]
    lib.host_tag_pressure.argtypes = [P(C.c_double), C.c_double, C.c_double]
    lib.host_tag_gradient_x.argtypes = [C.c_double] * 6
    o = ol.oracle()
    bx = qk_box.make((0, 0, 0), (9, 8, 7))
    rng = np.random.default_rng(8)
    s0, s1 = ol.HostFab(bx, 3), ol.HostFab(bx, 3)
    s0.a[...] = rng.standard_normal(s0.a.shape) * 1e3
    s1.a[...] = rng.standard_normal(s1.a.shape) * 1e-3
    for t in (0.0, 0.37, 0.999999, 1.0):
        a, b = ol.HostFab(bx, 3), ol.HostFab(bx, 3)
        wa = lib.host_time_interp(C.byref(a.desc()), C.byref(s0.desc()), C.byref(s1.desc()), 3, C.byref(bx), 0.0, 1.0, t)
        wb = o.orc_time_interp(C.byref(b.desc()), 0, C.byref(s0.desc()), C.byref(s1.desc()), 0, 3, C.byref(bx), 0.0, 1.0, t)
        assert wa == wb and np.array_equal(a.a, b.a)
    # tagging: the oracle on a 3^3 neighbourhood vs the per-cell function on the same seven pressures
    prm = hydro_params(gamma=1.4, reconstruct_eint=0)
    for seed in range(40):
        b1 = qk_box.make((0, 0, 0), (0, 0, 0))
        g = b1.grown(1)
        rho, v, Pr, r2 = ol.random_cons(g, 0, 100 + seed, "shocked" if seed % 2 else "smooth")
        Pr *= 10.0 ** r2.uniform(-4, 1)
        fab = ol.HostFab(g, 6)
        fab.a[...] = ol.cons_from_prim(rho, v, Pr, 1.4, r2)
        tag = C.create_string_buffer(1)
        o.orc_tag_pressure_gradient(C.byref(prm), C.byref(fab.desc()), tag, C.byref(b1), 0.1, 1.0e-3)
        pr = lambda i, j, k: o.orc_eos_pressure(C.byref(prm), fab.a[0, k, j, i],
                                                fab.a[4, k, j, i] - 0.5 * fab.a[0, k, j, i] * sum((fab.a[1 + m, k, j, i] / fab.a[0, k, j, i]) ** 2 for m in range(3)))
        P7 = (C.c_double * 7)(pr(1, 1, 1), pr(2, 1, 1), pr(0, 1, 1), pr(1, 2, 1), pr(1, 0, 1), pr(1, 1, 2), pr(1, 1, 0))
        assert 2 * lib.host_tag_pressure(P7, 0.1, 1.0e-3) == tag.raw[0]
        t2 = C.create_string_buffer(1)
        o.orc_tag_gradient_x(C.byref(fab.desc()), 0, t2, C.byref(b1), 0.37, 0.1, 0.01)
        assert lib.host_tag_gradient_x(fab.a[0, 1, 1, 0], fab.a[0, 1, 1, 1], fab.a[0, 1, 1, 2], 0.37, 0.1, 0.01) * 2 == t2.raw[0]
