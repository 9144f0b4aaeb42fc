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
SOD = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  5.0  1.0  1.0
geometry.is_periodic =  0    1    1
amr.v = 1
amr.n_cell = 1024 16 16
amr.max_level = {maxlev}
amr.blocking_factor = 16
do_reflux = 1
do_subcycle = 1
cfl = 0.6
hydro.reconstruction_order = 3
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
    res["sha256_a"], res["sha256_b"] = plotfile_sha(A, ncomp), plotfile_sha(B, ncomp)
    return res


def plotfile_sha(fabs, ncomp):
    """SHA-256 over the first ncomp components of every FAB, FABs ordered by (level, lo)"""
    h = hashlib.sha256()
    for k in sorted(fabs):
        h.update(np.ascontiguousarray(fabs[k][:ncomp]).tobytes())
    return h.hexdigest()


_BIN = None


def exe_path(exe):
    """the executables travel xz-compressed (oracle/ref_build/Makefile.cuda `pack`); unpack once per process tree into a scratch dir"""
    global _BIN
    plain = os.path.join(CUDA, exe)
    packed = plain + ".xz"
    if not os.path.exists(packed):
        return plain
    if _BIN is None:
        _BIN = os.environ.get("QK_REFCUDA_BIN") or os.path.join(tempfile.gettempdir(), "qk_refcuda_bin")
        os.makedirs(_BIN, exist_ok=True)
    out = os.path.join(_BIN, exe)
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(packed):
        with open(out + ".part", "wb") as f:
            subprocess.run(["xz", "-dc", packed], stdout=f, check=True)
        os.chmod(out + ".part", 0o755)
        os.replace(out + ".part", out)
    return out


def run(exe, inputs, extra, workdir, table=False, timeout=1500, exe_dir=None):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "in"), "w") as f:
        f.write(inputs)
    if table:  # RadhydroShell reads ./initial_conditions.txt
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "dust_shell_initial_conditions.txt"), os.path.join(workdir, "initial_conditions.txt"))
    if "shocktube" in exe:  # HydroShocktube reads ../extern/ppm1d/output (exact solution) relative to its cwd
        ext = os.path.join(os.path.dirname(workdir), "extern", "ppm1d")
        os.makedirs(ext, exist_ok=True)
        shutil.copy(os.path.join(ROOT, "oracle", "_ref", "cuda", "ppm1d_output.txt"), os.path.join(ext, "output"))
    t0 = time.time()
    env = dict(os.environ)  # the patched executables resolve libquokka_b200.so through their RUNPATH only from oracle/_ref/cuda
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "quokka_b200", "csrc") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    p = subprocess.run([os.path.join(exe_dir, exe) if exe_dir else exe_path(exe), "in"] + extra, cwd=workdir, capture_output=True, text=True, timeout=timeout,
                       env=env)
    wall = time.time() - t0
    log = p.stdout + p.stderr
    with open(os.path.join(workdir, "log.txt"), "w") as f:
        f.write(log)
    fom = re.search(r"Performance figure-of-merit: (\S+) .s/zone-update \[(\S+) Mupdates/s\]", log)
    steps = len(re.findall(r"\nSTEP \d+|Coarse STEP", log))
    r = {"exe": exe, "args": extra, "rc": p.returncode, "wall_s": round(wall, 2),
         "fom_Mupdates_s": float(fom.group(2)) if fom else None, "fom_us_per_update": float(fom.group(1)) if fom else None}
    zu = re.findall(r"Zone-updates on level (\d+): (\d+)", log)
    if zu:
        r["zone_updates_per_level"] = [int(b) for _, b in zu]
    m = re.search(r"\[b200\].*", log)
    if m:
        r["b200_banner"] = m.group(0)
    m = re.search(r"[Rr]elative (?:rms )?L1 (?:error )?norm = (\S+)", log)
    if m:
        r["l1_error_norm"] = m.group(1)
    if not fom:
        r["log_tail"] = log[-3000:]
    tp = re.search(r"Name\s+NCalls\s+Excl\. Min.*?\n-+\n(.*?)\n-+\n", log, re.S)
    if tp:
        r["tinyprofiler_excl_top"] = [" ".join(x.split()) for x in tp.group(1).split("\n")[:10]]
    return r


def last_plt(workdir):
    c = sorted(glob.glob(os.path.join(workdir, "plt*")))
    c = [x for x in c if os.path.isdir(x) and not x.endswith(".old")]
    return c[-1] if c else None


def case(name, inputs, stock, patched, ncomp, out, table=False, modes=("exact", "relaxed"), arena=None, golden=None, extra=()):
    """stock executable, then the patched one in each arithmetic mode, on the same inputs; with a final plotfile the states are compared
    FAB by FAB (all levels) with each other and with the CPU reference's SHA-256 (tests/golden/refcuda_hashes.json)"""
    base = tempfile.mkdtemp(prefix=f"qk_{name}_")
    extra0 = ([f"amrex.the_arena_init_size={arena}"] if arena else []) + list(extra)
    rec = {"stock": run(stock, inputs, extra0 + [], os.path.join(base, "stock", "run"), table)}
    print(name, "stock", rec["stock"].get("fom_Mupdates_s"), "rc", rec["stock"]["rc"], flush=True)
    ps = last_plt(os.path.join(base, "stock", "run"))
    if ps:
        rec["stock"]["sha256"] = plotfile_sha(read_plotfile_levels(ps), ncomp)
        if golden:
            rec["stock"]["equals_cpu_reference"] = (rec["stock"]["sha256"] == golden.get("sha256"))
    for mode in modes:
        wd = os.path.join(base, mode, "run")
        rec[mode] = run(patched, inputs, extra0 + ["b200.enabled=1", f"b200.arith={mode}"], wd, table)
        print(name, mode, rec[mode].get("fom_Mupdates_s"), "rc", rec[mode]["rc"], flush=True)
        pb = last_plt(wd)
        if ps and pb:
            rec[mode]["vs_stock"] = compare_plotfiles(ps, pb, ncomp)
            rec[mode]["vs_stock"]["plotfile"] = os.path.basename(pb)
            if golden:
                rec[mode]["equals_cpu_reference"] = (rec[mode]["vs_stock"].get("sha256_b") == golden.get("sha256"))
            print(name, mode, "vs stock CUDA:", {k: rec[mode]["vs_stock"].get(k) for k in ("same_grids", "bit_identical", "linf_rel", "values_differing")},
                  "== CPU reference bits:", rec[mode].get("equals_cpu_reference"), flush=True)
        if rec["stock"].get("fom_Mupdates_s") and rec[mode].get("fom_Mupdates_s"):
            rec[mode]["speedup_vs_stock"] = round(rec[mode]["fom_Mupdates_s"] / rec["stock"]["fom_Mupdates_s"], 3)
    if golden:
        rec["cpu_reference_golden"] = golden
    out[name] = rec
    shutil.rmtree(base, ignore_errors=True)


def pow_selftest(out):
    """does CUDA's pow(x, 2) equal x*x?  (hydro_system.hpp:602 forms K_S = std::pow(c_s, 2) * rho; gcc folds it to c_s*c_s)"""
    src = r"""
#include <cstdio>
#include <cmath>
__global__ void k(unsigned long long *n, unsigned long long *bad) {
  unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  unsigned long long b = 0;
  for (int i = 0; i < 4096; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    double x = 1e-3 + (double)(s >> 11) * (1.0 / 9007199254740992.0) * 1e3;
    if (std::pow(x, 2) != x * x) ++b;
  }
  atomicAdd(n, 4096ull); atomicAdd(bad, b);
}
int main() { unsigned long long *d, h[2] = {0, 0}; cudaMalloc(&d, 16); cudaMemset(d, 0, 16); k<<<256, 256>>>(d, d + 1);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%llu %llu\n", h[0], h[1]); return 0; }
"""
    tmp = tempfile.mkdtemp(prefix="qk_pow_")
    try:
        with open(os.path.join(tmp, "p.cu"), "w") as f:
            f.write(src)
        subprocess.run(["nvcc", "-gencode", "arch=compute_100,code=sm_100", "--fmad=false", "-o", "p", "p.cu"], cwd=tmp, check=True, capture_output=True)
        n, bad = subprocess.run(["./p"], cwd=tmp, capture_output=True, text=True).stdout.split()
        out["cuda_pow_x_2_vs_x_times_x"] = {"samples": int(n), "different": int(bad)}
    except Exception as e:
        out["cuda_pow_x_2_vs_x_times_x"] = {"error": str(e)}
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_ref_cuda.json"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    out = {"what": "stock reference CUDA build vs the same executable routed through libquokka_b200, on one B200",
           "nvidia_smi": subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()}
    only = set(a.only.split(",")) if a.only else None
    gold = {}
    try:
        with open(os.path.join(ROOT, "tests", "golden", "refcuda_hashes.json")) as f:
            gold = json.load(f)
    except Exception:
        pass

    def want(n):
        return only is None or n in only

    B, C = "test_hydro3d_blast_cuda", "test_hydro3d_blast_b200"
    if want("pow"):
        pow_selftest(out)
    # ---- parity (final plotfile, small step counts) ----
    for name in CASES:
        if not want(name):
            continue
        kind, inputs, ncomp = CASES[name]
        if kind == "sedov":
            case(name, inputs, B, C, ncomp, out, golden=gold.get(name))
        elif kind == "sod":
            case(name, inputs, "test_hydro_shocktube_cuda", "test_hydro_shocktube_b200", ncomp, out, golden=gold.get(name))
        else:
            case(name, inputs, "shell_cuda", "shell_b200", ncomp, out, table=True, golden=gold.get(name))
    # ---- timing (no plotfiles: the reference's protocol, plotfile_interval = -1) ----
    big = 60000000000
    if want("sedov256_c2"):  # C2: tests/blast_unigrid_256.in, 100 steps
        case("sedov256_c2", SEDOV.format(v=0, n=256, maxlev=0, grid=128, bf=128, amr=0, steps=100, plot=-1), B, C, 6, out, arena=big)
    if want("sedov512_c3"):  # C3 on one GPU: tests/blast_unigrid_512.in, 30 steps
        case("sedov512_c3_1gpu", SEDOV.format(v=0, n=512, maxlev=0, grid=128, bf=128, amr=0, steps=30, plot=-1), B, C, 6, out, arena=big)
    if want("sedov_amr256_c5"):  # C5: tests/blast_amr_maxlev2.in, 20 coarse steps
        case("sedov_amr256_c5", SEDOV.format(v=0, n=256, maxlev=2, grid=128, bf=32, amr=1, steps=20, plot=-1), B, C, 6, out, arena=big)
    if want("shell256_c4"):  # C4: tests/radhydro_shell_256.in, 10 coarse steps
        case("shell256_c4", SHELL.format(n=256, grid=128, steps=10, plot=-1), "shell_cuda", "shell_b200", 10, out, table=True, arena=big)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


# parity cases shared with tests/golden/make_golden_refcuda.py (which runs the CPU build of the same problem files): name -> (kind, inputs, ncomp)
CASES = {
    "sedov64_parity": ("sedov", SEDOV.format(v=0, n=64, maxlev=0, grid=32, bf=32, amr=0, steps=20, plot=1000), 6),
    "sedov_amr64_maxlev2": ("sedov", SEDOV.format(v=0, n=64, maxlev=2, grid=32, bf=16, amr=1, steps=10, plot=1000), 6),
    "sedov128_b64_s100": ("sedov", SEDOV.format(v=0, n=128, maxlev=0, grid=64, bf=64, amr=0, steps=100, plot=1000), 6),
    "sod_c1_amr": ("sod", SOD.format(maxlev=1, steps=8000, plot=100000), 6),
    "shell64_parity": ("shell", SHELL.format(n=64, grid=32, steps=3, plot=1000), 10),
}

if __name__ == "__main__":
    main()
