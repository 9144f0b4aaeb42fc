#!/usr/bin/env python3
"""First GPU call of the next round: everything that was built after round 1's GPU budget ran out, in one process.
    gpurun --timeout 600 -- python scripts/gpu_next.py          -> gpurun_out/next_*.{log,json}
 1. the GPU parity tests that have not run on a GPU yet (tests/test_zzzgpu_amr_transfer.py) and the newest verified ones;
 2. CUDA-event timings of k_rad_source in the exact and the relaxed arithmetic mode (8 x 128^3, RadhydroShell traits);
 3. CUDA-event timings of the AMR transfer kernels (128^3 coarse -> 256^3 fine, 6 components) with their HBM fractions;
 4. bench.py --workload radhydro in both arithmetic modes (run as subprocesses);
 5. bench.py --ncell 512 (Sedov 512^3 on one GPU, BASELINE.json's north-star size) in both arithmetic modes."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

import pytest
import torch

rc = pytest.main(["-q", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_zzzgpu_amr_transfer.py"),
                  os.path.join(ROOT, "tests", "test_zzgpu_relaxed_radhydro.py"), os.path.join(ROOT, "tests", "test_zgpu_shell.py"),
                  f"--junitxml={OUT}/next_tests.xml"])

from quokka_b200 import capi
from quokka_b200.capi import check, qk_array4, qk_box
from quokka_b200.device import DevFab, DevMultiFab
from test_rad_source_host import trait_set

lib = capi.load()
PEAK = 6455.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = {"pytest_rc": int(rc), "gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAK}


def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])


# ---- 2. source terms, exact vs relaxed --------------------------------------------------------------------------
hp, rp, sp, gen = trait_set("shell")
boxes = [qk_box.make((128 * i, 128 * j, 128 * k), (128 * i + 127, 128 * j + 127, 128 * k + 127)) for k in range(2) for j in range(2) for i in range(2)]
U = DevMultiFab(boxes, 10, ngrow=4, fill=1.0)
init = []
for f in U.fabs:
    shp = f.t.shape[1:]
    z, y, x = torch.meshgrid(*[torch.linspace(0, 3.14159, n, device="cuda", dtype=torch.float64) for n in shp], indexing="ij")
    s = torch.sin(x) * torch.sin(y) * torch.sin(z)
    rho = 1e-19 * 10.0 ** (2 * s - 1)
    Tg, Tr = 100.0 * 10.0 ** (s - 0.5), 100.0 * 10.0 ** (0.5 - s)
    v = 3e5 * torch.stack([torch.cos(x), torch.cos(y), torch.cos(z)])
    eint = rho * hp.boltzmann_constant * Tg / (hp.mean_molecular_weight * (hp.gamma - 1.0))
    E = sp.radiation_constant * Tr ** 4
    f.t[0] = rho
    f.t[1:4] = rho * v
    f.t[4] = eint + 0.5 * rho * (v ** 2).sum(0)
    f.t[5] = eint
    f.t[6] = E
    f.t[7:10] = 0.3 * torch.stack([torch.cos(x), torch.cos(y), torch.cos(z)]) * rp.c_light * E
    init.append(f.t.clone())
ncell = 8 * 128 ** 3
for mode, arith in (("exact", capi.QK_ARITH_EXACT), ("relaxed", capi.QK_ARITH_FAST)):
    hp.arith = arith

    def run():
        check(lib.qk_rad_add_source_terms(C.byref(hp), C.byref(rp), C.byref(sp), 1, len(boxes), U.boxes_c, U.descs, None, gen["dts"][1], None, None))

    ts = []
    for _ in range(4):
        for f, t0 in zip(U.fabs, init):
            f.t.copy_(t0)
        ts.append(timed(run, reps=2))
    ms = min(ts)
    res[f"rad_source_{mode}"] = {"ms": ms, "Mcell_per_s": ncell / ms / 1e3, "GBps_algorithmic_152B": ncell * 152 / ms / 1e6, "frac_hbm": ncell * 152 / ms / 1e6 / PEAK}
hp.arith = capi.QK_ARITH_EXACT
del U, init

# ---- 3. AMR transfer kernels --------------------------------------------------------------------------------------
ncomp = 6
cdomain = qk_box.make((0, 0, 0), (127, 127, 127))
fine_region = qk_box.make((0, 0, 0), (255, 255, 255))
dc, df, back = DevFab(cdomain.grown(1), ncomp), DevFab(fine_region, ncomp), DevFab(cdomain, ncomp)
dc.t.copy_(torch.rand(dc.t.shape, device="cuda", dtype=torch.float64) + 0.5)
r = (C.c_int * 3)(2, 2, 2)
bc = (C.c_int32 * (3 * ncomp))()
cd, fd, bd = (qk_array4 * 1)(dc.desc()), (qk_array4 * 1)(df.desc()), (qk_array4 * 1)(back.desc())
fr, cb = (qk_box * 1)(fine_region), (qk_box * 1)(cdomain)
ms = timed(lambda: check(lib.qk_amr_interp_cons_lin_minmax(1, cd, 0, fd, 0, ncomp, fr, C.byref(fine_region), C.byref(cdomain), r, bc, bc, None)))
nb = ncomp * 8 * (128 ** 3 + 256 ** 3)  # coarse read + fine write
res["amr_interp_128_to_256"] = {"ms": ms, "GBps_algorithmic": nb / ms / 1e6, "frac_hbm": nb / ms / 1e6 / PEAK, "bytes": nb}
ms = timed(lambda: check(lib.qk_amr_average_down(1, bd, 0, fd, 0, ncomp, cb, r, None)))
res["amr_average_down_256_to_128"] = {"ms": ms, "GBps_algorithmic": nb / ms / 1e6, "frac_hbm": nb / ms / 1e6 / PEAK, "bytes": nb}
ms = timed(lambda: check(lib.qk_amr_pre_interp_state(1, fr, fd, None)))
nb2 = 8 * 256 ** 3 * 6  # 5 components read, 1 written
res["amr_pre_interp_256"] = {"ms": ms, "GBps_algorithmic": nb2 / ms / 1e6, "frac_hbm": nb2 / ms / 1e6 / PEAK, "bytes": nb2}
del dc, df, back
torch.cuda.empty_cache()

# ---- 4. config C4 bench, both modes ----------------------------------------------------------------------------------
for mode in ("exact", "relaxed"):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "radhydro", "--arith", mode, "--steps", "3", "--warmup", "1", "--no-extras"],
                       capture_output=True, text=True)
    try:
        res[f"bench_radhydro_{mode}"] = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        res[f"bench_radhydro_{mode}"] = {"error": (p.stdout + p.stderr)[-1500:]}
# ---- 5. the north-star size on one GPU: Sedov 512^3 (64 boxes of 128^3), both arithmetic modes ------------------------------
for mode in ("relaxed", "exact"):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--ncell", "512", "--arith", mode, "--steps", "5", "--warmup", "3", "--no-extras"],
                       capture_output=True, text=True)
    try:
        res[f"bench_sedov512_{mode}"] = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        res[f"bench_sedov512_{mode}"] = {"error": (p.stdout + p.stderr)[-1500:]}
json.dump(res, open(os.path.join(OUT, "next_measurements.json"), "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("ms", "frac_hbm", "value", "ms_per_step", "error")}) for k, v in res.items()}, indent=1))
