"""The fused stage that also keeps its face fluxes (qk_hydro_advance_stage_keep_fluxes, csrc/qk_sweep_keepf.cu): what a level with flux
registers hands to incrementFluxRegisters (src/QuokkaSimulation.hpp:1195-1198,1280-1283 -> src/simulation.hpp:1345-1387).  Bar, exact
arithmetic: the new state AND the kept fluxes of both RK stages are BIT-IDENTICAL to the faithful one-kernel-per-operator path (which is
bit-identical to the oracle, tests/test_gpu_level.py) on ragged multi-box levels and every instantiated trait set; the flux arrays are
tight (an alias FArrayBox views them).  Relaxed arithmetic: same state bits as the relaxed stage without flux keeping, fluxes within
1e-12 of the faithful ones.  A flagged stage falls back to the faithful path and returns ITS (FOFC-replaced) fluxes."""
import ctypes as C

import numpy as np
import pytest

from quokka_b200 import capi
from quokka_b200.capi import check, qk_array4
from test_gpu_level import GenericProblem, exact, level_desc
from test_gpu_sweeps import CASES, RaggedProblem, prof

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load()


_cudart = None


def device_doubles(ptr, shape):
    """host copy of a device array the library owns (raw pointer from a qk_array4)"""
    global _cudart
    import torch  # noqa: F401  (loads the CUDA runtime the process uses)

    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        assert _cudart is not None, "CUDA runtime library not found"
        _cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _cudart.cudaMemcpy.restype = C.c_int
    host = np.empty(shape)
    rc = _cudart.cudaMemcpy(host.ctypes.data, ptr, host.nbytes, 2)  # cudaMemcpyDeviceToHost
    assert rc == 0, f"cudaMemcpy failed: {rc}"
    return host


def read_fluxes(lib, lev, p, nv):
    """the level's stage fluxes as host arrays [dir][box] -> (nv, nz(+1), ny(+1), nx(+1)); asserts the arrays are tight"""
    out = []
    nb = len(p.boxes)
    for d in range(3):
        descs = (qk_array4 * nb)()
        check(lib.qk_level_stage_fluxes(lev, d, descs))
        per = []
        for b, bx in enumerate(p.boxes):
            n = [bx.hi[k] - bx.lo[k] + 1 + (1 if k == d else 0) for k in range(3)]
            a = descs[b]
            assert a.jstride == n[0] and a.kstride == n[0] * n[1] and a.nstride == n[0] * n[1] * n[2], "kept flux arrays must be tight"
            assert a.ncomp >= nv and [a.begin[k] for k in range(3)] == [bx.lo[k] for k in range(3)]
            per.append(device_doubles(a.p, (nv, n[2], n[1], n[0])))
        out.append(per)
    return out


def run_stages(lib, p, prm, st, dt, entry):
    from quokka_b200.device import DevMultiFab

    desc, keep = level_desc(p)
    lev = C.c_void_p()
    check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
    U0 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost, host=st)
    U1 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost)
    U2 = DevMultiFab(p.boxes, p.ncomp, ngrow=p.nghost)
    nv = 6 + prm.nscalars
    b1, b2 = C.c_int64(-1), C.c_int64(-1)
    check(lib.qk_fill_boundary(lev, U0.descs, 0, p.ncomp, None))
    check(entry(lev, C.byref(prm), 1, U0.descs, U0.descs, U1.descs, dt, C.byref(b1), None))
    import torch

    torch.cuda.synchronize()
    fl1 = read_fluxes(lib, lev, p, nv)
    s1 = U1.numpy()
    check(lib.qk_fill_boundary(lev, U1.descs, 0, p.ncomp, None))
    check(entry(lev, C.byref(prm), 2, U0.descs, U1.descs, U2.descs, dt, C.byref(b2), None))
    torch.cuda.synchronize()
    fl2 = read_fluxes(lib, lev, p, nv)
    s2 = U2.numpy()
    lib.qk_level_destroy(lev)
    return s1, s2, fl1, fl2, b1.value, b2.value


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("kind", ["smooth", "shocked"])
def test_kept_fluxes_equal_the_faithful_path(lib, case, kind):
    ncell, cuts, periodic, bc, ns, nms, reint, gamma = CASES[case]
    p = RaggedProblem(ncell, cuts, periodic, bc, nscalars=ns, gamma=gamma)
    prm = p.params(nmscalars=nms, reconstruct_eint=reint)
    st = p.states(seed=11, kind=kind)
    dt = 1.0e-4 if kind == "shocked" else 3.0e-4
    if case == "mass_scalars_eint" and kind == "shocked":
        dt = 1.0e-5
    lib.qk_prof_enable(1)
    f = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_keep_fluxes)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    even_x = all((bx.hi[0] - bx.lo[0] + 1) % 2 == 0 for bx in p.boxes)
    if even_x:  # odd row lengths cannot be bulk-copied: those levels take the faithful path (still the right fluxes)
        assert counts.get("sweep_x", 0) == 2 and counts.get("flux_function", 0) == 0, f"the flux-keeping fused kernels did not run: {counts}"
    g = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
    assert (f[4], f[5], g[4], g[5]) == (0, 0, 0, 0)
    ng = p.nghost
    for b in range(len(p.boxes)):
        exact(f[0][b][:, ng:-ng, ng:-ng, ng:-ng], g[0][b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 1 state box {b}")
        exact(f[1][b][:, ng:-ng, ng:-ng, ng:-ng], g[1][b][:, ng:-ng, ng:-ng, ng:-ng], f"stage 2 state box {b}")
        for d in range(3):
            exact(f[2][d][b], g[2][d][b], f"stage 1 flux dir {d} box {b}")
            exact(f[3][d][b], g[3][d][b], f"stage 2 flux dir {d} box {b}")


def test_kept_fluxes_plm_and_relaxed(lib):
    """PLM instantiation (config C4's hydro) bit-exact; relaxed arithmetic: state bits of the plain relaxed stage, fluxes near the faithful ones"""
    p = GenericProblem((64, 32, 32), 32, (1, 1, 1), "periodic")
    st = p.states(seed=5, kind="smooth")
    dt = 3.0e-4
    prm = p.params()
    prm.reconstruction_order = 2
    f = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_keep_fluxes)
    g = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
    for b in range(len(p.boxes)):
        exact(f[1][b][:, 4:-4, 4:-4, 4:-4], g[1][b][:, 4:-4, 4:-4, 4:-4], f"PLM state box {b}")
        for d in range(3):
            exact(f[2][d][b], g[2][d][b], f"PLM stage 1 flux dir {d} box {b}")
            exact(f[3][d][b], g[3][d][b], f"PLM stage 2 flux dir {d} box {b}")
    prm = p.params(arith=capi.QK_ARITH_FAST)
    r = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_keep_fluxes)
    from test_gpu_sweeps import run_pair

    _, plain2, _, _ = run_pair(lib, p, prm, st, dt, lib.qk_hydro_advance_stage)
    g = run_stages(lib, p, p.params(), st, dt, lib.qk_hydro_advance_stage_faithful)
    worst = 0.0
    for b in range(len(p.boxes)):
        exact(r[1][b][:, 4:-4, 4:-4, 4:-4], plain2[b][:, 4:-4, 4:-4, 4:-4], f"relaxed state box {b}")
        for d in range(3):
            for k in (2, 3):
                scale = np.abs(g[k][d][b]).reshape(6, -1).max(axis=1)
                scale[scale == 0] = 1.0
                worst = max(worst, (np.abs(r[k][d][b] - g[k][d][b]).reshape(6, -1).max(axis=1) / scale).max())
    print(f"relaxed kept fluxes vs faithful: worst L-inf / max = {worst:.3e}")
    assert worst < 1e-12


def test_flagged_flux_keeping_stage_returns_the_fofc_fluxes(lib):
    p = GenericProblem((32, 32, 32), 16, (1, 1, 1), "periodic")
    prm = p.params()
    prm.abort_on_fofc_failure = 0
    st = p.states(seed=9, kind="shocked")
    dt = 2.0e-3
    lib.qk_prof_enable(1)
    f = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_keep_fluxes)
    counts = prof(lib)
    lib.qk_prof_enable(0)
    assert counts.get("replace_fluxes", 0) > 0
    g = run_stages(lib, p, prm, st, dt, lib.qk_hydro_advance_stage_faithful)
    for b in range(len(p.boxes)):
        exact(f[1][b][:, 4:-4, 4:-4, 4:-4], g[1][b][:, 4:-4, 4:-4, 4:-4], f"state box {b}")
        for d in range(3):
            exact(f[2][d][b], g[2][d][b], f"stage 1 flux dir {d} box {b}")
            exact(f[3][d][b], g[3][d][b], f"stage 2 flux dir {d} box {b}")
