#!/usr/bin/env python3
"""One GPU call for the matter-radiation source terms: the parity tests of tests/test_zgpu_rad_source.py, then a device timing
of qk_rad_add_source_terms on 8 x 128^3 (RadhydroShell traits, config C4's box layout) with CUDA events.
    gpurun -- python scripts/gpu_radsrc.py      -> gpurun_out/radsrc_tests.log, gpurun_out/radsrc_timing.json"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

import pytest
import torch

TIME_ONLY = "--time-only" in sys.argv  # (for the ncu capture: no tests, one repetition)
rc = -1
if not TIME_ONLY:
    rc = pytest.main(["-q", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_zgpu_rad_source.py"),
                      f"--junitxml={ROOT}/gpurun_out/radsrc_tests.xml"])
    open(os.path.join(ROOT, "gpurun_out", "radsrc_tests.log"), "w").write(f"pytest exit code {int(rc)}\n")

from quokka_b200 import capi
from quokka_b200.capi import check, qk_box
from quokka_b200.device import DevMultiFab
from test_rad_source_host import trait_set

lib = capi.load()
hp, rp, sp, gen = trait_set("shell")
boxes = [qk_box.make((128 * i, 128 * j, 128 * k), (128 * i + 127, 128 * j + 127, 128 * k + 127)) for k in range(2) for j in range(2) for i in range(2)]
U = DevMultiFab(boxes, 10, ngrow=4, fill=1.0)
g = torch.Generator(device="cuda").manual_seed(1)
init = []
for f in U.fabs:
    shp = f.t.shape[1:]
    # smooth-ish fields (a real shell has neighbouring cells with similar iteration counts): large-scale modes + 10 % noise
    z, y, x = torch.meshgrid(*[torch.linspace(0, 3.14159, n, device="cuda", dtype=torch.float64) for n in shp], indexing="ij")
    s = torch.sin(x) * torch.sin(y) * torch.sin(z)
    noise = lambda: 1 + 0.1 * (torch.rand(shp, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    rho = 1e-19 * 10.0 ** (2 * s - 1) * noise()
    Tg = 100.0 * 10.0 ** (s - 0.5) * noise()
    Tr = 100.0 * 10.0 ** (0.5 - s) * noise()
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
res = {}
VARIANTS = [("default", None, None)] if TIME_ONLY else [("default", None, None)] + [(f"{d}{m}", d, m) for d, m in
                                                                                  [("plain", 4), ("plain", 5), ("plain", 6), ("plain", 7), ("plain", 8),
                                                                                   ("plain", 10), ("shared", 5)]]
for variant, div, minb in VARIANTS:
  for dt in (gen["dts"][1:2] if (TIME_ONLY or div) else gen["dts"]):
    os.environ.pop("QK_RADSRC_PLAIN_DIV", None)
    os.environ.pop("QK_RADSRC_MINB", None)
    if div:
        os.environ["QK_RADSRC_PLAIN_DIV"] = "1" if div == "plain" else "0"
        os.environ["QK_RADSRC_MINB"] = str(minb)
    times = []
    cnt = (C.c_int64 * 7)()
    for rep in range(2 if TIME_ONLY else 4):
        for f, t0 in zip(U.fabs, init):
            f.t.copy_(t0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.qk_rad_add_source_terms(C.byref(hp), C.byref(rp), C.byref(sp), 1, len(boxes), U.boxes_c, U.descs, None, dt, None, None))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    for f, t0 in zip(U.fabs, init):
        f.t.copy_(t0)
    check(lib.qk_rad_add_source_terms(C.byref(hp), C.byref(rp), C.byref(sp), 1, len(boxes), U.boxes_c, U.descs, None, dt, cnt, None))
    ms = min(times[1:])
    res[f"{variant}:{dt}"] = {"ms": ms, "Mcell_per_s": ncell / ms / 1e3, "GBps_algorithmic_152B": ncell * 152 / ms / 1e6,
                    "newton_iters_per_cell": cnt[1] / ncell, "solves_per_cell": cnt[0] / ncell, "max_newton": cnt[2], "fail": [cnt[4], cnt[6]]}
out = {"kernel": "k_rad_source", "workload": "8 x 128^3, RadhydroShell traits, stage 1", "pytest_rc": int(rc), "by_dt_radiation": res,
       "gpu": torch.cuda.get_device_name(0)}
if not TIME_ONLY:
    open(os.path.join(ROOT, "gpurun_out", "radsrc_timing.json"), "w").write(json.dumps(out, indent=1))
print(json.dumps(out))
