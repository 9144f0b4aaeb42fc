"""GPU parity of SURVEY 8(f)2's time interpolation and 8(f)4's regrid support through the C ABI against the oracle (pinned to AMReX and to the
reference's own ErrorEst / FixupState by tests/test_oracle_regrid_vs_ref.py): qk_amr_time_interp bit-exact, qk_tag_pressure_gradient /
qk_tag_gradient_x identical tags and counts, qk_hydro_fixup_state bit-exact; plus a 256^3 property check (tags of a uniform state are empty, a
pressure step is tagged on exactly its two planes)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import check, hydro_params, qk_array4, qk_box, qk_carray4
from test_oracle_regrid_vs_ref import sedov_like_state

pytestmark = pytest.mark.gpu


def dev(fab):
    from quokka_b200.device import DevFab

    return DevFab(fab.box, fab.ncomp, host=fab.a)


def dev_tags(box):
    import torch

    nz, ny, nx = box.shape()
    t = torch.zeros((nz, ny, nx), dtype=torch.int8, device="cuda")
    d = qk_carray4()
    d.p = t.data_ptr()
    d.jstride, d.kstride, d.nstride = nx, nx * ny, nx * ny * nz
    d.begin[:] = list(box.lo)
    d.end[:] = [box.hi[i] + 1 for i in range(3)]
    d.ncomp = 1
    return t, d


@pytest.mark.parametrize("times", [(0.0, 1.0, 0.3), (2.0, 2.5, 2.125), (1.0, 3.0, 3.0), (1.0, 3.0, 1.0005)])
def test_time_interp_vs_oracle(times):
    lib = capi.load()
    t0, t1, t = times
    rng = np.random.default_rng(21)
    boxes = [qk_box.make((0, 0, 0), (15, 11, 9)), qk_box.make((16, 0, 0), (40, 6, 3)), qk_box.make((-4, -4, -4), (3, 3, 3))]
    s0 = [ol.HostFab(b.grown(2), 8) for b in boxes]
    s1 = [ol.HostFab(b.grown(2), 8) for b in boxes]
    for f in s0 + s1:
        f.a[...] = rng.uniform(-3, 3, f.a.shape) * 10.0 ** rng.integers(-6, 6, f.a.shape)
    want = [ol.HostFab(b.grown(1), 7, fill=-9.0) for b in boxes]
    regions = [b.grown(1) for b in boxes]
    whiches = {ol.oracle().orc_time_interp(C.byref(w.desc()), 1, C.byref(a.desc()), C.byref(b.desc()), 2, 6, C.byref(r), t0, t1, t)
               for w, a, b, r in zip(want, s0, s1, regions)}
    d0, d1 = [dev(f) for f in s0], [dev(f) for f in s1]
    dd = [dev(ol.HostFab(b.grown(1), 7, fill=-9.0)) for b in boxes]
    n = len(boxes)
    which = C.c_int(-1)
    n0 = lib.qk_launch_count()
    check(lib.qk_amr_time_interp(n, (qk_array4 * n)(*[x.desc() for x in dd]), 1, (qk_array4 * n)(*[x.desc() for x in d0]),
                                 (qk_array4 * n)(*[x.desc() for x in d1]), 2, 6, (qk_box * n)(*regions), t0, t1, t, C.byref(which), None))
    assert lib.qk_launch_count() - n0 == 1
    assert whiches == {which.value}
    for g, w in zip(dd, want):
        assert np.array_equal(g.numpy(), w.a)


@pytest.mark.parametrize("seed,kind", [(1, "smooth"), (2, "shocked")])
def test_tags_vs_oracle(seed, kind):
    lib = capi.load()
    prm = hydro_params(gamma=1.4, reconstruct_eint=0)
    cases = [sedov_like_state(20, 4, seed, kind), sedov_like_state(12, 4, seed + 10, kind)]
    n = len(cases)
    devs = [dev(f) for _, f in cases]
    tags = [dev_tags(b) for b, _ in cases]
    cnt = C.c_int64(-1)
    check(lib.qk_tag_pressure_gradient(C.byref(prm), n, (qk_box * n)(*[b for b, _ in cases]), (qk_array4 * n)(*[d.desc() for d in devs]),
                                       (qk_carray4 * n)(*[d for _, d in tags]), 0.1, 1.0e-3, C.byref(cnt), None))
    total = 0
    for (bx, fab), (t, _) in zip(cases, tags):
        want = C.create_string_buffer(bx.ncells())
        ol.oracle().orc_tag_pressure_gradient(C.byref(prm), C.byref(fab.desc()), want, C.byref(bx), 0.1, 1.0e-3)
        w = np.frombuffer(want.raw, dtype=np.int8).reshape(t.shape)
        assert 0 < (w != 0).sum() < w.size
        assert np.array_equal(t.cpu().numpy(), w)
        total += int((w != 0).sum())
    assert cnt.value == total
    # density-gradient criterion of the shock tube
    tags2 = [dev_tags(b) for b, _ in cases]
    check(lib.qk_tag_gradient_x(n, (qk_box * n)(*[b for b, _ in cases]), (qk_array4 * n)(*[d.desc() for d in devs]), 0,
                                (qk_carray4 * n)(*[d for _, d in tags2]), 0.37, 0.1, 0.01, C.byref(cnt), None))
    total = 0
    for (bx, fab), (t, _) in zip(cases, tags2):
        want = C.create_string_buffer(bx.ncells())
        ol.oracle().orc_tag_gradient_x(C.byref(fab.desc()), 0, want, C.byref(bx), 0.37, 0.1, 0.01)
        w = np.frombuffer(want.raw, dtype=np.int8).reshape(t.shape)
        assert np.array_equal(t.cpu().numpy(), w)
        total += int((w != 0).sum())
    assert cnt.value == total and total > 0


def test_fixup_state_vs_oracle():
    lib = capi.load()
    prm = hydro_params(gamma=1.4, reconstruct_eint=0, density_floor=0.5, temp_floor=0.0)
    bx, fab = sedov_like_state(16, 0, 5, "shocked")
    fab.a[5] *= np.where(np.arange(fab.a[5].size).reshape(fab.a[5].shape) % 3 == 0, 1.0e-6, 1.0)
    want = ol.HostFab(bx, 6)
    want.a[...] = fab.a
    ol.oracle().orc_fixup_state(C.byref(prm), C.byref(want.desc()), C.byref(bx))
    d = dev(fab)
    check(lib.qk_hydro_fixup_state(C.byref(prm), 1, (qk_box * 1)(bx), (qk_array4 * 1)(d.desc()), None))
    assert not np.array_equal(want.a, fab.a)
    assert np.array_equal(d.numpy(), want.a)


def test_tags_full_size_property():
    """256^3 (configs[1] grid): a uniform state tags nothing; a pressure step between planes i = 99 | 100 tags exactly those two planes"""
    import torch

    lib = capi.load()
    prm = hydro_params(gamma=1.4, reconstruct_eint=0)
    n = 256
    bx = qk_box.make((0, 0, 0), (n - 1,) * 3)
    g = bx.grown(1)
    from quokka_b200.device import DevFab

    st = DevFab(g, 6, fill=0.0)
    st.t[0] = 1.0
    st.t[4] = 2.5
    st.t[5] = 2.5
    t, d = dev_tags(bx)
    cnt = C.c_int64(-1)
    check(lib.qk_tag_pressure_gradient(C.byref(prm), 1, (qk_box * 1)(bx), (qk_array4 * 1)(st.desc()), (qk_carray4 * 1)(d), 0.1, 1.0e-3, C.byref(cnt), None))
    assert cnt.value == 0 and int(t.sum().item()) == 0
    st.t[4, :, :, 101:] = 25.0  # cells i >= 100 (FAB index = i + 1)
    check(lib.qk_tag_pressure_gradient(C.byref(prm), 1, (qk_box * 1)(bx), (qk_array4 * 1)(st.desc()), (qk_carray4 * 1)(d), 0.1, 1.0e-3, C.byref(cnt), None))
    assert cnt.value == 2 * n * n
    assert bool((t[:, :, 99] == 2).all()) and bool((t[:, :, 100] == 2).all()) and int(t.sum().item()) == 2 * 2 * n * n
    del st, t
    torch.cuda.empty_cache()
