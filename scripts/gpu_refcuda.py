#!/usr/bin/env python3
"""Run the reference's own executables on the B200: the STOCK CUDA build (oracle/_ref/cuda/*_cuda, the unmodified reference
compiled by oracle/ref_build/Makefile.cuda) and the PATCHED build (*_b200: the same problem files with advanceHydroAtLevel /
subcycleRadiationAtLevel routed through libquokka_b200 by oracle/ref_build/apply_b200_patch.py), on the same inputs.

    python scripts/gpu_refcuda.py [--quick] [--out gpurun_out/r02_ref_cuda.json]

Records, per case, the reference's own figure-of-merit line (src/simulation.hpp:972-977), and compares the plotfiles of the
stock and the patched run FAB by FAB (all AMR levels).  Input parameters are those of the reference's tests/blast_unigrid_256.in,
tests/blast_amr_maxlev2.in and tests/radhydro_shell_256.in (BASELINE.json configs C2, C5, C4)."""
import argparse
import glob
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.path.join(ROOT, "oracle", "_ref", "cuda")

SEDOV = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  1.2  1.2  1.2
geometry.is_periodic =  0    0    0
amr.v = {v}
amr.n_cell = {n} {n} {n}
amr.max_level = {maxlev}
amr.max_grid_size = {grid}
amr.blocking_factor = {bf}
amr.n_error_buf = 3
amr.grid_eff = 0.7
do_reflux = {amr}
do_subcycle = {amr}
max_timesteps = {steps}
plotfile_interval = {plot}
checkpoint_interval = -1
"""
SHELL = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  6.172e19  6.172e19  6.172e19
geometry.is_periodic =  1    1    1
amr.v = 0
amr.n_cell = {n} {n} {n}
amr.max_level = 0
amr.max_grid_size = {grid}
amr.blocking_factor = {grid}
amr.n_error_buf = 3
amr.grid_eff = 0.7
do_reflux = 0
do_subcycle = 0
max_timesteps = {steps}
plotfile_interval = {plot}
checkpoint_interval = -1
"""


def read_plotfile_levels(path):
    """{(level, lo): array[ncomp, nz, ny, nx]} for every FAB of every level of an AMReX plotfile (native FP64)."""
    out = {}
    for lev_dir in sorted(glob.glob(os.path.join(path, "Level_*"))):
        lev = int(lev_dir.rsplit("_", 1)[1])
        with open(os.path.join(lev_dir, "Cell_H")) as f:
            cellh = f.read()
        for fname, off in re.findall(r"FabOnDisk: (\S+) (\d+)", cellh):
            with open(os.path.join(lev_dir, fname), "rb") as f:
                f.seek(int(off))
                hdr = f.readline().decode()
                m = re.search(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \(\d+,\d+,\d+\)\) (\d+)", hdr)
                blo = tuple(int(m.group(i)) for i in (1, 2, 3))
                bhi = [int(m.group(i)) for i in (4, 5, 6)]
                nc = int(m.group(7))
                bn = [bhi[d] - blo[d] + 1 for d in range(3)]
                out[(lev, blo)] = np.frombuffer(f.read(8 * nc * bn[0] * bn[1] * bn[2]), dtype="<f8").reshape(nc, bn[2], bn[1], bn[0])
    return out


def compare_plotfiles(a, b, ncomp):
    A, B = read_plotfile_levels(a), read_plotfile_levels(b)
    res = {"same_grids": sorted(A.keys()) == sorted(B.keys()), "nfabs": len(A), "levels": sorted({k[0] for k in A})}
    if not res["same_grids"]:
        res["grids_a"], res["grids_b"] = len(A), len(B)
        return res
    scale = np.zeros(ncomp)
    for k in A:
        scale = np.maximum(scale, np.abs(A[k][:ncomp]).reshape(ncomp, -1).max(axis=1))
    scale[scale == 0] = 1.0
    err = np.zeros(ncomp)
    ndiff = 0
    ncell = 0
    for k in A:
        d = np.abs(A[k][:ncomp] - B[k][:ncomp]).reshape(ncomp, -1)
        err = np.maximum(err, d.max(axis=1) / scale)
        ndiff += int((A[k][:ncomp] != B[k][:ncomp]).sum())
        ncell += A[k][0].size
    res.update(bit_identical=(ndiff == 0), values_differing=ndiff, cells=ncell, linf_rel_per_comp=[float(x) for x in err], linf_rel=float(err.max()))
    h = hashlib.sha256()
    for k in sorted(A):
        h.update(np.ascontiguousarray(A[k][:ncomp]).tobytes())
    res["sha256_a"] = h.hexdigest()
    return res


def run(exe, inputs, extra, workdir, table=False, timeout=1500):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "in"), "w") as f:
        f.write(inputs)
    if table:
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "dust_shell_initial_conditions.txt"), os.path.join(workdir, "initial_conditions.txt"))
    t0 = time.time()
    p = subprocess.run([os.path.join(CUDA, exe), "in"] + extra, cwd=workdir, capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    log = p.stdout + p.stderr
    with open(os.path.join(workdir, "log.txt"), "w") as f:
        f.write(log)
    fom = re.search(r"Performance figure-of-merit: (\S+) .s/zone-update \[(\S+) Mupdates/s\]", log)
    el = re.search(r"elapsed time: (\S+) seconds", log)
    upd = re.search(r"zone-updates? on level|Zone-updates", log)
    steps = len(re.findall(r"ADVANCE with time", log))
    r = {"exe": exe, "args": extra, "rc": p.returncode, "wall_s": round(wall, 2), "coarse_steps_logged": steps,
         "fom_Mupdates_s": float(fom.group(2)) if fom else None, "elapsed_s": float(el.group(1)) if el else None}
    rad = re.search(r"radiation.*?\[(\S+) Mupdates/s\]", log)
    if rad:
        r["rad_fom_Mupdates_s"] = float(rad.group(1))
    m = re.search(r"\[b200\].*", log)
    if m:
        r["b200_banner"] = m.group(0)
    if p.returncode != 0 or not fom:
        r["log_tail"] = log[-3000:]
    # TinyProfiler top entries (exclusive time table)
    tp = re.search(r"Name\s+NCalls\s+Excl\. Min.*?\n-+\n(.*?)\n-+\n", log, re.S)
    if tp:
        r["tinyprofiler_excl_top"] = tp.group(1).split("\n")[:12]
    return r


def last_plt(workdir):
    c = sorted(glob.glob(os.path.join(workdir, "plt*")))
    c = [x for x in c if os.path.isdir(x) and not x.endswith(".old")]
    return c[-1] if c else None


def case(name, inputs, stock, patched, ncomp, out, table=False, modes=("exact", "relaxed"), keep=False, arena=None):
    base = tempfile.mkdtemp(prefix=f"qk_{name}_")
    extra0 = [f"amrex.the_arena_init_size={arena}"] if arena else []
    rec = {"stock": run(stock, inputs, extra0 + [], os.path.join(base, "stock"), table)}
    print(name, "stock", rec["stock"].get("fom_Mupdates_s"), "rc", rec["stock"]["rc"], flush=True)
    ps = last_plt(os.path.join(base, "stock"))
    for mode in modes:
        wd = os.path.join(base, mode)
        rec[mode] = run(patched, inputs, extra0 + ["b200.enabled=1", f"b200.arith={mode}"], wd, table)
        print(name, mode, rec[mode].get("fom_Mupdates_s"), "rc", rec[mode]["rc"], flush=True)
        pb = last_plt(wd)
        if ps and pb:
            rec[mode]["vs_stock"] = compare_plotfiles(ps, pb, ncomp)
            rec[mode]["vs_stock"]["plotfile"] = os.path.basename(pb)
            print(name, mode, "vs stock:", {k: rec[mode]["vs_stock"].get(k) for k in ("same_grids", "bit_identical", "linf_rel", "values_differing")}, flush=True)
        if rec["stock"].get("fom_Mupdates_s") and rec[mode].get("fom_Mupdates_s"):
            rec[mode]["speedup_vs_stock"] = rec[mode]["fom_Mupdates_s"] / rec["stock"]["fom_Mupdates_s"]
    # the patched executable with the driver switched off must reproduce the stock executable (same code path)
    out[name] = rec
    if not keep:
        shutil.rmtree(base, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_ref_cuda.json"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    out = {"what": "stock reference CUDA build vs the same executable routed through libquokka_b200, on one B200",
           "nvidia_smi": subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()}
    only = set(a.only.split(",")) if a.only else None

    def want(n):
        return only is None or n in only

    B, C = "test_hydro3d_blast_cuda", "test_hydro3d_blast_b200"
    # parity first, small: uniform 64^3 in 32^3 boxes, 20 steps
    if want("sedov64"):
        case("sedov64_parity", SEDOV.format(v=0, n=64, maxlev=0, grid=32, bf=32, amr=0, steps=20, plot=1000), B, C, 6, out)
    # C5 in small: 64^3 base + 2 levels, subcycling + reflux, 10 coarse steps
    if want("amr64"):
        case("sedov_amr64_maxlev2", SEDOV.format(v=1, n=64, maxlev=2, grid=32, bf=16, amr=1, steps=10, plot=1000), B, C, 6, out)
    if not a.quick:
        # C2: tests/blast_unigrid_256.in, 100 steps (the paper's protocol)
        if want("sedov256"):
            case("sedov256_c2", SEDOV.format(v=0, n=256, maxlev=0, grid=128, bf=128, amr=0, steps=100, plot=1000), B, C, 6, out, arena=60000000000)
        # C5: tests/blast_amr_maxlev2.in, 20 coarse steps
        if want("amr256"):
            case("sedov_amr256_c5", SEDOV.format(v=1, n=256, maxlev=2, grid=128, bf=32, amr=1, steps=20, plot=1000), B, C, 6, out, arena=60000000000)
    if want("shell64"):
        case("shell64_parity", SHELL.format(n=64, grid=32, steps=3, plot=1000), "shell_cuda", "shell_b200", 10, out, table=True)
    if not a.quick and want("shell256"):
        case("shell256_c4", SHELL.format(n=256, grid=128, steps=10, plot=1000), "shell_cuda", "shell_b200", 10, out, table=True, arena=60000000000)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
