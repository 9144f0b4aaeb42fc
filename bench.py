#!/usr/bin/env python3
"""bench.py -- Mcell-updates/s of the hydro hot path on the Sedov blast (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one coarse time step of the whole level: a full RK2-SSP update of every cell (2 ghost fills, 2 x
(prim + flattening + PPM + HLLC in x,y,z + update), dt and CFL reductions) -- exactly what the reference counts in
its figure of merit (src/simulation.hpp:972-977, cellUpdates_ += CountCells(lev) :1285).
N = 1: configs[1], Sedov 256^3 in eight 128^3 boxes.  N > 1: weak scaling, 256^3 cells (8 boxes) per GPU
(N = 8 is configs[2], Sedov 512^3), boxes mapped to ranks as AMReX's SFC DistributionMapping does, ghost exchange
over the library's NCCL transport.  `value` is timed with CUDA events around the C++ driver's loop with the state
resident in HBM; `e2e` drives the same steps through the C ABI from HOST buffers (pinned): state upload, step,
state download inside the timed region.  The state (0.8 GB per GPU) is far larger than L2, so no explicit flush.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcell-updates/s (Sedov 3D PPM+HLLC)"
UNIT = "Mcell-updates/s"
# the reference's own CPU executable: the -O3 -march=x86-64-v3 build when present (oracle/ref_build/Makefile OPT=...), else the -O2 one the golden
# vectors were made with (both -ffp-contract=off; same bits)
REF_EXE = next((p for p in (os.path.join(ROOT, "oracle", "_ref", "o3", "test_hydro3d_blast"), os.path.join(ROOT, "oracle", "_ref", "test_hydro3d_blast"))
                if os.path.exists(p)), os.path.join(ROOT, "oracle", "_ref", "test_hydro3d_blast"))
REF_OPT = "-O3 -march=x86-64-v3" if os.sep + "o3" + os.sep in REF_EXE else "-O2"

REF_INPUT = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  1.2  1.2  1.2
geometry.is_periodic =  0    0    0
amr.v = 0
amr.max_level = 0
amr.n_error_buf = 3
amr.grid_eff = 0.7
do_reflux = 0
do_subcycle = 0
plotfile_interval = -1
checkpoint_interval = -1
"""


def ncell_for(ngpus):
    """256^3 cells per GPU: 1 -> 256^3, 2 -> 512x256x256, 4 -> 512x512x256, 8 -> 512^3"""
    n = [256, 256, 256]
    d = 0
    g = ngpus
    while g > 1:
        n[d] *= 2
        d = (d + 1) % 3
        g //= 2
    return n


def make_config(world, ncell):
    """identical for both arms (the driver compares them); the arithmetic mode of our arm is the top-level key `arith`"""
    which = "configs[1]" if (world == 1 and ncell[0] == 256) else "configs[2]" if ncell[0] * ncell[1] * ncell[2] == 512 ** 3 else "weak-scaled configs[1]"
    return {"workload": f"Sedov blast {ncell[0]}x{ncell[1]}x{ncell[2]} uniform ({which}), 128^3 boxes, PPM+HLLC RK2, gamma=1.4, reflecting BCs, cfl 0.3",
            "cells": ncell[0] * ncell[1] * ncell[2], "l2": "state per GPU (0.8 GB) >> 126 MB L2, no flush needed",
            "parallelism": f"dp{world} (boxes over ranks, NCCL ghost exchange)" if world > 1 else "1 GPU"}


ARITH_TEXT = {"exact": "exact (IEEE order, no FMA contraction; bit-identical to the reference's CPU build)",
              "relaxed": "relaxed (closed-form gamma-law EOS, ~1-ulp reciprocals, FMA; rel L_inf vs reference < 1e-12 after 100 steps, tests/test_gpu_relaxed.py)"}


# ---- clocks --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi -lms 20 in the background (the recipe's clocks line).  Started BEFORE the problem is set up -- nvidia-smi needs ~1 s to
    deliver its first row, longer than a K-step timed region -- and every row is time-stamped on arrival, so that stop() can report the
    median over the window [t_load_start, t_load_end] in which this process kept the GPU busy with the benchmark's own kernels."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def count(self, t0):
        return sum(1 for t, _ in self.rows if t >= t0)

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": "warm-up + timed region + identical untimed steps until >= 10 samples"}


# ---- the reference's own CPU implementation ---------------------------------------------------------------------
def run_reference_cpu(ncell, box, nsteps, threads):
    """the UNMODIFIED reference executable (oracle/_ref/test_hydro3d_blast, built from /root/reference by
    oracle/ref_build/Makefile) on the host cores; returns (Mupdates/s from its own FOM line, elapsed s)"""
    tmp = tempfile.mkdtemp(prefix="qkref_")
    try:
        with open(os.path.join(tmp, "in"), "w") as f:
            f.write(REF_INPUT)
            f.write(f"amr.n_cell = {ncell[0]} {ncell[1]} {ncell[2]}\namr.max_grid_size = {box}\namr.blocking_factor = {box}\nmax_timesteps = {nsteps}\n")
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        t0 = time.time()
        out = subprocess.run([REF_EXE, "in"], cwd=tmp, env=env, capture_output=True, text=True).stdout
        el = time.time() - t0
        m = re.search(r"figure-of-merit:\s*([0-9.eE+-]+)\s*\S+/zone-update\s*\[([0-9.eE+-]+)\s*Mupdates/s\]", out)
        me = re.search(r"elapsed time:\s*([0-9.eE+-]+)", out)
        if not m:
            raise RuntimeError("reference run printed no figure of merit:\n" + out[-2000:])
        return float(m.group(2)), (float(me.group(1)) if me else el)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference_shell_cpu(ncell, box, nsteps, threads):
    """the reference's own RadhydroShell problem (oracle/_ref/test_radhydro_shell is hard-wired to 50 steps, so the same problem
    file behind oracle/ref_build/shell_golden.cpp, which leaves the step count to the inputs) on the host cores; its
    interpolation table travels as oracle/_ref/dust_shell_initial_conditions.txt.  Returns (Mupdates/s, elapsed s)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "shell_golden")
    table = os.path.join(ROOT, "oracle", "_ref", "dust_shell_initial_conditions.txt")
    tmp = tempfile.mkdtemp(prefix="qkref_")
    try:
        shutil.copy(table, os.path.join(tmp, "initial_conditions.txt"))
        with open(os.path.join(tmp, "in"), "w") as f:
            f.write("geometry.prob_lo = 0.0 0.0 0.0\ngeometry.prob_hi = 3.086e19 3.086e19 3.086e19\ngeometry.is_periodic = 1 1 1\namr.v = 0\n"
                    "amr.max_level = 0\ndo_reflux = 0\ndo_subcycle = 0\nplotfile_interval = -1\ncheckpoint_interval = -1\n")
            f.write(f"amr.n_cell = {ncell} {ncell} {ncell}\namr.max_grid_size = {box}\namr.blocking_factor = {box}\nmax_timesteps = {nsteps}\n")
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        t0 = time.time()
        out = subprocess.run([exe, "in"], cwd=tmp, env=env, capture_output=True, text=True).stdout
        el = time.time() - t0
        m = re.search(r"figure-of-merit:\s*([0-9.eE+-]+)\s*\S+/zone-update\s*\[([0-9.eE+-]+)\s*Mupdates/s\]", out)
        me = re.search(r"elapsed time:\s*([0-9.eE+-]+)", out)
        if not m:
            raise RuntimeError("reference run printed no figure of merit:\n" + out[-2000:])
        return float(m.group(2)), (float(me.group(1)) if me else el)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_oracle_port(ncell, box, nsteps):
    """fallback CPU baseline when oracle/_ref did not travel: the single-core C restatement (oracle/liboracle.so)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_oracle_golden import run_oracle_sedov

    t0 = time.time()
    run_oracle_sedov(ncell, box, nsteps)
    el = time.time() - t0
    return ncell ** 3 * nsteps / el / 1e6, el


def cpu_baseline(steps):
    cores = os.cpu_count() or 1
    if os.path.exists(REF_EXE):
        v, el = run_reference_cpu([128, 128, 128], 64, steps, cores)
        return {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"BOUNDED SAMPLE of the workload: reference executable ({REF_OPT}, OpenMP, {cores} threads), Sedov 128^3 in 64^3 boxes, {steps} steps, "
                          f"{el:.1f} s (the full 256^3 configuration is what `bench.py --impl reference` runs)"}
    v, el = run_oracle_port(48, 48, steps)
    return {"value": round(v, 4), "unit": UNIT, "cores": 1, "kind": "port", "sample": f"C oracle, 1 thread, Sedov 48^3, {steps} steps, {el:.1f} s"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = args.steps + args.warmup
    cores = os.cpu_count() or 1
    if getattr(args, "workload", "hydro") == "radhydro":  # config C4: the reference's own RadhydroShell problem on the host cores
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "shell_golden")):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/shell_golden not built (make -C oracle/ref_build shell_golden)"}))
            return 0
        nst = max(1, min(steps, 2))
        v, el = run_reference_shell_cpu(64, 32, nst, cores)
        txt = f"reference RadhydroShell problem (OpenMP, {cores} threads), 64^3 in 32^3 boxes, {nst} coarse steps of 10 radiation substeps, {el:.1f} s"
        print(json.dumps({"impl": "reference", "metric": "Mcell-updates/s (radiation hydrodynamics coarse step: hydro PLM+HLLC RK2 + 10 two-moment IMEX "
                          "substeps with matter-radiation coupling)", "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": round(el * 1e3 / nst, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "the reference's own initial condition",
                          "config": {"workload": "RadhydroShell (configs[3]); bounded sample 64^3", "cells": 64 ** 3},
                          "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference", "sample": txt},
                          "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    ncell = ncell_for(args.gpus)
    if os.path.exists(REF_EXE):
        # the REAL configuration of our arm (same grid, same 128^3 boxes).  Bounded in wall time, not in problem size: the step count is
        # cut so that the run stays near 150 s on the box's cores (a 256^3 step takes ~4.5 s on 16 cores); Mcell-updates/s is a rate.
        cells = ncell[0] * ncell[1] * ncell[2]
        nrun = max(2, min(steps, int(150.0 * 0.23e6 * cores / cells)))
        v, el = run_reference_cpu(ncell, 128, nrun, cores)
        kind = "reference"
        sample_txt = (f"reference executable ({REF_OPT} -ffp-contract=off, OpenMP, {cores} threads), the full configuration: Sedov {ncell[0]}x{ncell[1]}x{ncell[2]} in "
                      f"128^3 boxes, {nrun} of the {steps} requested steps, {el:.1f} s")
        ms = el * 1e3 / nrun
    else:
        v, el = run_oracle_port(48, 48, steps)
        cores, kind, sample_txt = 1, "port", f"C oracle, 1 thread, Sedov 48^3, {steps} steps, {el:.1f} s"
        ms = el * 1e3 / steps
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.gpus, ncell),
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_txt},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---- our arm ---------------------------------------------------------------------------------------------------
def read_prof(lib, capi):
    buf = (capi.C.c_char * 8192)()
    lib.qk_prof_report(buf, 8192)
    prof = {}
    for ln in buf.value.decode().splitlines():
        nm, cnt, ms = ln.split()
        prof[nm] = (int(cnt), float(ms))
    return prof


def sweep_fracs(prof, ncell_local, steps):
    """per direction sweep: average launch ms and fraction of the measured HBM copy bandwidth at SURVEY 8(d)'s 96 B per cell-sweep"""
    peak, _ = hbm_peak()
    out = {}
    for k in ("sweep_x", "sweep_y", "sweep_z"):
        if k in prof:
            ms = prof[k][1] / (2 * steps)
            out[k] = {"avg_launch_ms": round(ms, 4), "achieved_gbs": round(96 * ncell_local / (ms * 1e-3) / 1e9, 1),
                      "frac": round(96 * ncell_local / (ms * 1e-3) / 1e9 / peak, 4)}
    return out


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def gpu_reference_record(steps):
    """the reference's own CUDA build (oracle/_ref/cuda, made from /root/reference by oracle/ref_build/Makefile.cuda) on this GPU: the stock
    executable, and the same problem file with advanceHydroAtLevel routed through libquokka_b200, both timed by the
    reference's own figure-of-merit line (src/simulation.hpp:972-977) on configs[1] with plotfiles off"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import gpu_refcuda as g

    if not (os.path.exists(os.path.join(g.CUDA, "test_hydro3d_blast_cuda.xz")) or os.path.exists(os.path.join(g.CUDA, "test_hydro3d_blast_cuda"))):
        return {"unavailable": "oracle/_ref/cuda not built (make -f Makefile.cuda -C oracle/ref_build stock patched pack)"}
    inputs = g.SEDOV.format(v=0, n=256, maxlev=0, grid=128, bf=128, amr=0, steps=steps, plot=-1)
    base = tempfile.mkdtemp(prefix="qk_gpuref_")
    arena = ["amrex.the_arena_init_size=40000000000"]
    rec = {"workload": f"Sedov 256^3, 128^3 boxes, {steps} steps from t = 0, plotfile_interval = -1; the reference's own FOM line (wall clock of evolve())",
           "build": "nvcc 12.9 -gencode arch=compute_100,code=sm_100 --fmad=false -maxrregcount=255, AMReX 24.09 CUDA, no MPI"}
    try:
        r = g.run("test_hydro3d_blast_cuda", inputs, arena, os.path.join(base, "stock", "run"))
        rec["stock_cuda"] = {"value": r.get("fom_Mupdates_s"), "unit": UNIT}
        for mode in ("exact", "relaxed"):
            r = g.run("test_hydro3d_blast_b200", inputs, arena + ["b200.enabled=1", f"b200.arith={mode}"], os.path.join(base, mode, "run"))
            rec[f"through_libquokka_b200_{mode}"] = {"value": r.get("fom_Mupdates_s"), "unit": UNIT}
            if rec["stock_cuda"]["value"] and r.get("fom_Mupdates_s"):
                rec[f"through_libquokka_b200_{mode}"]["ratio_vs_stock_cuda"] = round(r["fom_Mupdates_s"] / rec["stock_cuda"]["value"], 3)
    finally:
        shutil.rmtree(base, ignore_errors=True)
    return rec


def nrank_parity(torch, dist, world, rank, comm, capi):
    """N-rank parity inside the scaling run: 5 exact-mode steps of Sedov on a (64^3 x N) domain in 32^3 boxes across the N ranks, and the
    same problem on rank 0 alone; SHA-256 of the assembled global state must agree (bit-exact, exact arithmetic)."""
    import hashlib

    import numpy as np

    from quokka_b200.problems import SedovProblem
    from quokka_b200.simulation import HydroSimulation

    ncell = [c // 4 for c in ncell_for(world)]
    prob = SedovProblem(ncell, 32)
    prm = prob.params(arith=capi.QK_ARITH_EXACT)
    nsteps = 5

    def run(nranks, r, cm):
        sim = HydroSimulation(prob, nranks=nranks, rank=r, comm=cm, params=prm)
        sim.setInitialConditions()
        nd, _, _ = sim.evolve(nsteps)
        st = sim.state_valid()
        t = sim.time
        ids = list(sim.local_ids)
        sim.close()
        return nd, t, ids, st

    nd, t, ids, st = run(world, rank, comm)
    mine = torch.from_numpy(np.stack([st[i] for i in ids])).cuda()
    allt = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device="cuda")
    dist.all_gather_into_tensor(allt, mine)
    idt = torch.tensor(ids, dtype=torch.int64, device="cuda")
    allid = torch.empty((world, len(ids)), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allid, idt)
    rec = None
    if rank == 0:
        def assemble(pieces):
            out = np.zeros((6,) + tuple(reversed(ncell)))
            for gid, a in pieces.items():
                bx = prob.boxes[gid]
                out[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1] = a
            return out
        allt_h, allid_h = allt.cpu().numpy(), allid.cpu().numpy()
        pieces = {int(allid_h[r][b]): allt_h[r][b] for r in range(world) for b in range(allid_h.shape[1])}
        g_n = assemble(pieces)
        nd1, t1, ids1, st1 = run(1, 0, None)
        g_1 = assemble(st1)
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
        rec = {"what": f"Sedov {ncell[0]}x{ncell[1]}x{ncell[2]} in 32^3 boxes, {nsteps} exact-arithmetic steps: {world} ranks vs rank 0 alone",
               "sha_nrank": sha(g_n), "sha_1rank": sha(g_1), "equal": bool(sha(g_n) == sha(g_1)), "time_equal": bool(t == t1),
               "max_abs_diff": float(np.abs(g_n - g_1).max())}
    dist.barrier()
    return rec


def main_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libquokka_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # nvidia-smi needs ~1 s for its first row: start it before the problem is built
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from quokka_b200 import capi
    from quokka_b200.problems import SedovProblem
    from quokka_b200.simulation import Communicator, HydroSimulation

    lib = capi.load()
    ncell = [args.ncell] * 3 if getattr(args, "ncell", 0) else ncell_for(world)
    prob = SedovProblem(ncell, 128)
    comm = None
    if world > 1:
        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        comm = Communicator(rank, world, bcast)
    prm = prob.params(arith=capi.QK_ARITH_FAST if args.arith == "relaxed" else capi.QK_ARITH_EXACT)
    sim = HydroSimulation(prob, nranks=world, rank=rank, comm=comm, params=prm)
    sim.setInitialConditions()
    ncells_total = ncell[0] * ncell[1] * ncell[2]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up (also builds the scratch pools)
    t_load0 = time.perf_counter()
    sim.evolve(args.warmup)
    barrier()
    lib.qk_prof_enable(1)
    l0 = lib.qk_launch_count()
    barrier()
    nd, elapsed, dev_ms = sim.evolve(args.steps)
    barrier()
    launches = lib.qk_launch_count() - l0
    lib.qk_prof_enable(0)
    assert nd == args.steps, (nd, args.steps)
    ms_total = max_over_ranks(dev_ms)
    value = ncells_total * args.steps / (ms_total * 1e-3) / 1e6
    prof = read_prof(lib, capi)
    # clocks under load: the timed region (K steps of ~8 ms) is shorter than nvidia-smi's sampling can resolve, so the same kernels keep
    # running, untimed, until the sampler has >= 10 rows inside the load window (at most 4 s)
    clk = None
    extra_steps = 0
    while True:
        enough = (clocks.count(t_load0) >= 10) if rank == 0 else True
        if dist is not None:
            flag = torch.tensor([1 if enough else 0], device="cuda")
            dist.broadcast(flag, src=0)
            enough = bool(flag.item())
        if enough or time.perf_counter() - t_load0 > 4.0 + 1e-3 * ms_total or args.no_extras:
            break
        sim.evolve(args.steps)
        extra_steps += args.steps
    barrier()
    if rank == 0:
        clk = clocks.stop(t_load0, time.perf_counter())
        clk["untimed_steps_for_sampling"] = extra_steps

    ncell_local = sum(b.ncells() for b in sim.local_boxes)
    roof = roofline(prof, ncell_local, args.steps, ms_total, args.arith)

    if args.no_extras:  # profiling runs (ncu): kernels only
        if rank == 0:
            print(json.dumps({"value": round(value, 2), "ms_per_step": round(ms_total / args.steps, 4), "gpu_launches": int(launches), "roofline": roof,
                              "sweeps": sweep_fracs(prof, ncell_local, args.steps),
                              "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}))
        sim.close()
        return 0
    # end to end through the C ABI from host buffers: every step uploads the state (pinned host memory, the VALID cells: ghost cells are the
    # step's own business), advances it and downloads the result; per box the copy runs on its own stream and overlaps the (un)packing
    # kernel of the previous box (qk_sim_set_state_valid / qk_sim_get_state_valid)
    e2e_steps = max(1, min(args.steps, 5))
    sim.download_valid()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.upload_valid()
        dt = sim.computeTimestep()
        r = sim.advanceSingleTimestepAtLevel(dt)
        assert r >= 0
        sim.download_valid()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = ncells_total * e2e_steps / e2e_s / 1e6
    hb = sim.valid_bytes()
    sim.close()
    del sim
    torch.cuda.empty_cache()
    # the same workload with the bit-exact arithmetic mode, for the record (short: 5 steps after the warm-up)
    other = None
    if args.arith == "relaxed":
        sim2 = HydroSimulation(prob, nranks=world, rank=rank, comm=comm, params=prob.params(arith=capi.QK_ARITH_EXACT))
        sim2.setInitialConditions()
        sim2.evolve(args.warmup)
        barrier()
        lib.qk_prof_enable(1)
        nd2, _, ms2 = sim2.evolve(5)
        barrier()
        lib.qk_prof_enable(0)
        ms2 = max_over_ranks(ms2)
        other = {"arith": ARITH_TEXT["exact"], "value": round(ncells_total * nd2 / (ms2 * 1e-3) / 1e6, 2), "ms_per_step": round(ms2 / nd2, 4),
                 "sweeps": sweep_fracs(read_prof(lib, capi), ncell_local, nd2)}
        sim2.close()
        del sim2
        torch.cuda.empty_cache()

    parity = None
    if world > 1:
        try:
            parity = nrank_parity(torch, dist, world, rank, comm, capi)
        except Exception as e:  # never lose the scaling line
            parity = {"error": repr(e)}

    # ---- sub-records of the N = 1 line: the north-star size on one GPU, config C4, the reference's own CUDA build on this GPU -------------
    north = radhydro = gpu_ref = None
    if world == 1 and not getattr(args, "ncell", 0) and not args.no_subrecords:
        try:
            free, _ = torch.cuda.mem_get_info()
            if free > 90e9:
                p512 = SedovProblem([512] * 3, 128)
                s5 = HydroSimulation(p512, params=p512.params(arith=capi.QK_ARITH_FAST if args.arith == "relaxed" else capi.QK_ARITH_EXACT))
                s5.setInitialConditions()
                s5.evolve(2)
                torch.cuda.synchronize()
                lib.qk_prof_enable(1)
                n5, _, ms5 = s5.evolve(4)
                lib.qk_prof_enable(0)
                pr5 = read_prof(lib, capi)
                north = {"workload": "Sedov blast 512^3 uniform on ONE GPU (configs[2]'s grid, tests/blast_unigrid_512.in; BASELINE.json's north-star size), 64 boxes of 128^3",
                         "arith": args.arith, "steps": n5, "warmup": 2, "value": round(512 ** 3 * n5 / (ms5 * 1e-3) / 1e6, 2), "unit": UNIT,
                         "ms_per_step": round(ms5 / n5, 3), "sweeps": sweep_fracs(pr5, 512 ** 3, n5),
                         "kernel_ms_per_step": {k: round(v[1] / n5, 3) for k, v in sorted(pr5.items(), key=lambda kv: -kv[1][1])},
                         "target": "north_star: >= 0.70 of the HBM roofline on the x sweep at 96 B/cell"}
                s5.close()
                del s5
                torch.cuda.empty_cache()
            else:
                north = {"skipped": f"only {free / 1e9:.0f} GB of device memory free"}
        except Exception as e:
            north = {"error": repr(e)}
        try:
            radhydro = radhydro_record(args.arith, 2, 1, extras=False)
            radhydro = {k: radhydro[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "config", "gpu_launches", "roofline",
                                                 "kernel_ms_per_step")}
        except Exception as e:
            radhydro = {"error": repr(e)}
        try:
            gpu_ref = gpu_reference_record(args.steps + args.warmup)
        except Exception as e:
            gpu_ref = {"error": repr(e)}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": make_config(world, ncell), "arith": ARITH_TEXT[args.arith],
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": hb, "d2h_bytes_per_step": hb, "steps": e2e_steps,
                        "note": "host-buffer plugin call: pinned upload of the valid cells + step + download per step (PCIe-bound)"},
                "roofline": roof, "sweeps": sweep_fracs(prof, ncell_local, args.steps),
                "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
        if other:
            line["exact_arith"] = other
        if parity is not None:
            line["parity"] = parity
        if north is not None:
            line["north_star_512"] = north
        if radhydro is not None:
            line["radhydro"] = radhydro
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(3)
            except Exception as e:  # never lose the GPU line to a CPU-side problem
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if comm:
        comm.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


# ---- radiation transport sweep (SURVEY.md section 8(a) RadSystem rows; not the headline metric) ----------------------------------
def main_radiation(args):
    """python bench.py --workload radiation: one step = one two-moment transport substep (ghost fill, stage 1, ghost fill,
    stage 2 through qk_rad_advance_stage) of a free-streaming pulse on 256^3 in eight 128^3 boxes, one photon group, PPM."""
    import ctypes as C

    import numpy as np
    import torch

    from quokka_b200 import capi
    from quokka_b200.capi import check, make_level_desc, qk_box, rad_params
    from quokka_b200.device import DevMultiFab
    from quokka_b200.problems import chop_domain

    lib = capi.load()
    n, box, ng, ncomp = 256, 128, 4, 10
    order = int(os.environ.get("QK_BENCH_RAD_ORDER", "3"))  # 3 = PPM (default), 2 = PLM(MC) as config C4 uses
    prm = rad_params(c_light=1.0, c_hat=1.0, recon_order=order, arith=capi.QK_ARITH_FAST if args.arith == "relaxed" else capi.QK_ARITH_EXACT)
    boxes = chop_domain(n, box)
    dom = qk_box.make((0, 0, 0), (n - 1,) * 3)
    dx = [1.0 / n] * 3
    bc = [capi.QK_BC_INT_DIR] * (3 * ncomp)
    desc, keep = make_level_desc(dom, (1, 1, 1), dx, ng, ncomp, boxes, [0] * len(boxes), 0, bc, bc)
    lev = C.c_void_p()
    check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
    host = []
    for bx in boxes:
        g = bx.grown(ng)
        nz, ny, nx = g.shape()
        z, y, x = np.meshgrid((np.arange(g.lo[2], g.hi[2] + 1) + 0.5) / n, (np.arange(g.lo[1], g.hi[1] + 1) + 0.5) / n,
                              (np.arange(g.lo[0], g.hi[0] + 1) + 0.5) / n, indexing="ij", sparse=True)
        a = np.ones((ncomp, nz, ny, nx))
        E = 1.0e-3 + np.exp(-((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2) / (2 * 0.05 ** 2))
        a[6] = E
        a[7] = 0.9 * E
        a[8] = 0.0
        a[9] = 0.0
        host.append(a)
    U = [DevMultiFab(boxes, ncomp, ngrow=ng, host=host) for _ in range(3)]
    dt = 0.3 * dx[0]

    def step():
        check(lib.qk_fill_boundary(lev, U[0].descs, 0, ncomp, None))
        check(lib.qk_rad_advance_stage(lev, C.byref(prm), 1, U[0].descs, U[0].descs, U[1].descs, dt, None))
        check(lib.qk_fill_boundary(lev, U[1].descs, 0, ncomp, None))
        check(lib.qk_rad_advance_stage(lev, C.byref(prm), 2, U[0].descs, U[1].descs, U[2].descs, dt, None))
        U[0], U[2] = U[2], U[0]

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    lib.qk_prof_enable(1)
    l0 = lib.qk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.qk_launch_count() - l0
    lib.qk_prof_enable(0)
    buf = (capi.C.c_char * 8192)()
    lib.qk_prof_report(buf, 8192)
    prof = {ln.split()[0]: (int(ln.split()[1]), float(ln.split()[2])) for ln in buf.value.decode().splitlines()}
    ncell = n ** 3
    value = ncell * args.steps / (ms * 1e-3) / 1e6
    peak, src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    st_ms = prof.get("rad_stage", (0, 0.0))[1] / (2 * args.steps)
    ach = 64 * ncell / (st_ms * 1e-3) / 1e9 if st_ms > 0 else None
    line = {"metric": "Mcell-updates/s (two-moment radiation transport substep, PPM + HLL, RK2)", "value": round(value, 2), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "free-streaming Gaussian pulse 256^3 periodic, eight 128^3 boxes, 1 photon group, c_hat = c, cfl 0.3 (radiation rows of SURVEY 8a; "
                                   "config C4's source terms are not part of this path)", "cells": ncell, "reconstruction_order": order, "arith": ARITH_TEXT[args.arith]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_rad_stage", "achieved": round(ach, 1) if ach else None, "peak": peak, "peak_source": src, "unit": "GB/s",
                         "frac": round(ach / peak, 4) if ach else None, "traffic": None, "algorithmic_bytes_per_cell": 64, "design_bytes_per_cell": 160,
                         "avg_launch_ms": round(st_ms, 4)},
            "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
    if not args.no_extras:
        # CPU baseline: the C oracle's transport substep (1 thread) on 48^3
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol

        m = 48
        bx = qk_box.make((0, 0, 0), (m - 1,) * 3)
        d2, k2 = make_level_desc(bx, (1, 1, 1), [1.0 / m] * 3, ng, ncomp, [bx], [0], 0, bc, bc)
        o = ol.oracle()
        L = o.orc_level_create(C.byref(d2))
        for which in (0, 1):
            d = o.orc_level_state(L, which, 0)
            v = np.ctypeslib.as_array(C.cast(d.p, C.POINTER(C.c_double)), shape=(ncomp, m + 8, m + 8, m + 8))
            v[...] = 1.0
            v[7:] = 0.1
        t0 = time.time()
        reps = 3
        for _ in range(reps):
            o.orc_level_swap(L)
            o.orc_rad_advance_level(L, C.byref(prm), 0.3 / m)
        el = time.time() - t0
        o.orc_level_destroy(L)
        line["cpu_baseline"] = {"value": round(m ** 3 * reps / el / 1e6, 4), "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"C oracle, 1 thread, {m}^3, {reps} substeps, {el:.1f} s"}
    print(json.dumps(line))
    lib.qk_level_destroy(lev)
    return 0


def radhydro_record(arith_name, steps, warmup, extras=True, dist_ok=False, ncell_override=0):
    """python bench.py --workload radhydro: config C4's coarse step on one GPU -- hydro PLM(minmod)+HLLC RK2 advance, then
    subcycleRadiationAtLevel: 10 IMEX substeps of (ghost fill, transport stage 1, matter-radiation source terms, ghost fill,
    transport stage 2, source terms) -- RadhydroShell traits on 256^3 periodic in eight 128^3 boxes, through the C++ driver
    (qk_sim_evolve).  The initial condition is SYNTHETIC: the reference's Gaussian shell density with an analytic radiation field
    (its interpolation table lives in /root/reference and does not travel)."""
    import ctypes as C

    import numpy as np

    from quokka_b200 import capi
    from quokka_b200.device import DevMultiFab
    from quokka_b200.problems import ShellProblem
    from quokka_b200.simulation import HydroSimulation

    lib = capi.load()
    world = int(os.environ.get("WORLD_SIZE", "1")) if dist_ok else 1
    rank = int(os.environ.get("RANK", "0")) if dist_ok else 0
    # N = 1: config C4's grid on one GPU (256^3 in eight 128^3 boxes).  N > 1: weak scaling with 256^3 cells per GPU unless --ncell fixes the grid
    # (--ncell 256 --gpus 8 is configs[3] literally: one 128^3 box per GPU)
    nc3 = [ncell_override] * 3 if ncell_override else ncell_for(world)
    box = 128
    P = ShellProblem
    h = P.prob_hi / min(nc3)  # cubic cells; the periodic domain grows with the rank count (weak scaling at the 256^3 problem's cell size)

    class SyntheticShell(ShellProblem):
        """the reference's Gaussian shell density with an analytic radiation field, evaluated box by box (nothing of size n^3 on the host)"""

        def fields(self, bx, ng):
            g = bx.grown(ng)
            ax = [((np.arange(g.lo[d], g.hi[d] + 1) % nc3[d]) + 0.5) * h - 0.5 * h * nc3[d] for d in range(3)]
            z, y, x = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij", sparse=True)
            return np.sqrt(x * x + y * y + z * z)

        def initial_state(self, bx, ng=None):
            ng = self.nghost if ng is None else ng
            r = self.fields(bx, ng)
            sigma_sh = 0.3 * P.r_0 / (2.0 * np.sqrt(2.0 * np.log(2.0)))
            M_shell = 0.5 * 1.0e6 * 2.0e33
            rho = np.maximum(M_shell / (4.0 * np.pi * r * r * np.sqrt(2.0 * np.pi * sigma_sh * sigma_sh)) * np.exp(-(r - P.r_0) ** 2 / (2.0 * sigma_sh * sigma_sh)),
                             1.0e-8 * P.rho_0)
            T = 300.0 * (1.0 + (r / P.r_0) ** 2) ** -0.25  # K, gas and radiation in equilibrium
            Er = P.a_rad * T ** 4
            c_v = capi.K_B / ((2.2 * capi.M_U) * (P.gamma - 1.0))
            a = np.zeros((10,) + r.shape)
            a[0] = rho
            a[4] = rho * c_v * T
            a[5] = a[4]
            a[6] = Er
            a[7] = a[8] = a[9] = 0.1 * P.c_light * Er / np.sqrt(3.0)
            return a

        def source(self, bx):
            r = self.fields(bx, 0)
            sigma_star = 0.3 * P.r_0
            return np.ascontiguousarray(((1.0 / P.c_light) * (0.5 * 1.0e6 * 2.0e33 * 2000.0) / (2.0 * np.pi * sigma_star * sigma_star) ** 1.5
                                         * np.exp(-(r * r) / (2.0 * sigma_star * sigma_star)))[None])

    prob = SyntheticShell(nc3, box)
    prob.dx = [h] * 3
    # --arith relaxed (the parser's default) selects the relaxed fused PLM sweeps, the relaxed transport sweeps AND the relaxed source-term solve
    arith = capi.QK_ARITH_FAST if arith_name == "relaxed" else capi.QK_ARITH_EXACT
    comm = dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        from quokka_b200.simulation import Communicator

        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        comm = Communicator(rank, world, bcast)
    sim = HydroSimulation(prob, nranks=world, rank=rank, comm=comm, params=prob.params(arith=arith))
    esrc = DevMultiFab(sim.local_boxes, 1, ngrow=0, host=[prob.source(b) for b in sim.local_boxes])
    sim.enableRadiation(prob.rad_params(), prob.rad_source_params(), esrc, rad_cfl=prob.rad_cfl, max_substeps=prob.max_substeps)
    sim.setInitialConditions()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")) if dist_ok else 0)
    clocks.start()
    nd, _, _ = sim.evolve(warmup)
    assert nd == warmup
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
    lib.qk_prof_enable(1)
    l0 = lib.qk_launch_count()
    nd, elapsed, ms = sim.evolve(steps)
    launches = lib.qk_launch_count() - l0
    lib.qk_prof_enable(0)
    if dist is not None:
        torch.cuda.synchronize()
        dist.barrier()
    ms = max_over_ranks(ms)
    clk = clocks.stop()
    assert nd == steps, (nd, steps)
    buf = (capi.C.c_char * 8192)()
    lib.qk_prof_report(buf, 8192)
    prof = {ln.split()[0]: (int(ln.split()[1]), float(ln.split()[2])) for ln in buf.value.decode().splitlines()}
    ncell = nc3[0] * nc3[1] * nc3[2]
    nsub = sim.radiationSubsteps
    value = ncell * steps / (ms * 1e-3) / 1e6
    peak, psrc = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, psrc = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    # dominant kernel class of the step: the source-term solve (2 launches per substep); 152 algorithmic bytes per cell
    src_cnt, src_ms = prof.get("rad_source_terms", (0, 0.0))
    per = src_ms / max(1, 2 * nsub * steps)
    ach = 152 * ncell / (per * 1e-3) / 1e9 if per > 0 else None
    ncell_local = sum(b.ncells() for b in sim.local_boxes)
    ach = 152 * ncell_local / (per * 1e-3) / 1e9 if per > 0 else None
    st = sim.gather_global() if (extras and world == 1) else None
    line = {"metric": "Mcell-updates/s (radiation hydrodynamics coarse step: hydro PLM+HLLC RK2 + 10 two-moment IMEX substeps with matter-radiation coupling)",
            "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"RadhydroShell {nc3[0]}x{nc3[1]}x{nc3[2]} periodic (configs[3]'s problem; {world} GPU(s)), {len(prob.boxes)} boxes of 128^3, 1 photon group, kappa = 20, "
                                   "beta_order 1, PLM hydro + PLM radiation, cfl 0.3; state per GPU (>= 0.2 GB) >> L2, no flush", "cells": ncell, "radiation_substeps_per_step": nsub,
                       "radiation_Mcell_updates_per_s": round(value * nsub, 1), "arith": arith_name},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_rad_source", "achieved": round(ach, 1) if ach else None, "peak": peak, "peak_source": psrc, "unit": "GB/s",
                         "frac": round(ach / peak, 4) if ach else None, "traffic": None, "algorithmic_bytes_per_cell": 152, "avg_launch_ms": round(per, 4),
                         "note": "FP64-pipe / latency bound implicit solve (DESIGN.md section 3); ncu: profiles/r01_ncu_radsrc.txt"},
            "kernel_ms_per_step": {k: round(v[1] / steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
    if extras and world == 1 and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "shell_golden")):
        cores = os.cpu_count() or 1
        v, el = run_reference_shell_cpu(64, 32, 1, cores)
        line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": f"reference RadhydroShell problem (OpenMP, {cores} threads), 64^3 in 32^3 boxes, 1 coarse step = 10 radiation substeps, {el:.1f} s"}
    if st is not None:
        line["sanity"] = {"finite": bool(np.isfinite(st).all()), "min_rho": float(st[0].min()), "min_Erad": float(st[6].min()),
                          "sim_time": sim.time}
    sim.close()
    del sim, esrc
    if comm:
        comm.close()
    import torch

    torch.cuda.empty_cache()
    return line


def main_radhydro(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = radhydro_record(args.arith, args.steps, args.warmup, extras=not args.no_extras, dist_ok=True, ncell_override=getattr(args, "ncell", 0))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def roofline(prof, ncell_local, steps, ms_total, arith="relaxed"):
    """dominant kernel class vs the measured HBM copy bandwidth (MEASURED_PEAKS.json, else the recipe's fallback).
    Algorithmic and design bytes per cell are documented in DESIGN.md section 3."""
    peak, src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    # the unit is one cell of one direction sweep; ALGORITHMIC bytes = 96 (read 6 state variables, write 6).
    # DESIGN bytes = what this implementation moves per unit (equals the ncu DRAM traffic within a few %).
    ALG = 96
    if arith == "exact":  # keeps 0.5*F(U0) on the faces between the RK stages
        DESIGN = {"flux_function": 48 + 24 + 56, "sweep_x": 56 + 56 + 56, "sweep_y": 56 + 56 + 112, "sweep_z": 56 + 56 + 56 + 48 + 48}
        fp64_per_cell, opmix = 784, "profiles/r02_ncu_opmix_exact_2.txt (y sweep: 410.9 M FP64-pipe warp instructions over 524 288 rows of 32 cells)"
    else:  # keeps R(U0) per cell instead
        DESIGN = {"sweep_x": 56 + 56, "sweep_y": 56 + 112, "sweep_z": 56 + 56 + 48 + 48 + 48}
        fp64_per_cell, opmix = 431, "profiles/r02_ncu_opmix_relaxed_2.txt (y sweep: 226.07 M FP64-pipe warp instructions over 524 288 rows of 32 cells)"
    # measured DRAM traffic per launch (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, stage-1/stage-2 launches averaged) from the
    # committed capture profiles/ncu_traffic.json -- reported only while the kernel sources are the ones that were profiled
    TRAFFIC, traffic_note = {}, "profiles/ncu_traffic.json missing"
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from make_ncu_traffic import sources_sha

        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("sources_sha256") == sources_sha() and arith in tj:
            TRAFFIC, traffic_note = tj[arith]["bytes_per_launch"], tj[arith]["from"]
        else:
            traffic_note = "stale: the kernel sources changed since profiles/ncu_traffic.json was captured"
    except Exception as e:
        traffic_note = f"unavailable: {e}"
    cand = [(v[1], k) for k, v in prof.items() if k in DESIGN]
    if not cand:
        return None
    ms, name = max(cand)
    nl = prof[name][0]
    # every launch of a sweep class covers all local cells once; two launches (one per RK stage) per step
    passes = {"flux_function": 6, "sweep_x": 2, "sweep_y": 2, "sweep_z": 2}[name] * steps
    t = ms * 1e-3 / passes
    achieved = ALG * ncell_local / t / 1e9
    # second roof: the FP64 pipe (64 lanes/SM/clk at 1965 MHz)
    fp64_peak = 148 * 64 * 1.965e9
    return {"bound": "hbm", "kernel": name, "achieved": round(achieved, 1), "peak": peak, "peak_source": src, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": TRAFFIC.get(name) if ncell_local == 256 ** 3 else None, "traffic_source": traffic_note, "algorithmic_bytes_per_cell": ALG, "algorithmic_bytes_per_launch": ALG * ncell_local,
            "design_bytes_per_cell": DESIGN[name], "achieved_design_gbs": round(DESIGN[name] * ncell_local / t / 1e9, 1),
            "frac_design": round(DESIGN[name] * ncell_local / t / 1e9 / peak, 4),
            "gcell_sweeps_per_s": round(ncell_local / t / 1e9, 3), "avg_launch_ms": round(ms / passes, 4), "launches": nl, "share_of_step": round(ms / ms_total, 4),
            "fp64_roof": {"fp64_instr_per_cell": fp64_per_cell, "peak_instr_per_s": fp64_peak, "frac": round(fp64_per_cell * ncell_local / t / fp64_peak, 4),
                          "note": f"FP64-pipe instructions per cell-sweep from {opmix}; the sweep is issue/FP64-pipe bound, not HBM bound (DESIGN.md section 3)"}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arith", default="relaxed", choices=["exact", "relaxed"], help="arithmetic mode of the fused sweeps (DESIGN.md section 3)")
    ap.add_argument("--workload", default="hydro", choices=["hydro", "radiation", "radhydro"],
                    help="hydro = the BASELINE.json metric (default); radiation = the transport sweep; radhydro = config C4's coarse step")
    ap.add_argument("--ncell", type=int, default=0, help="override the grid: N^3 cells in 128^3 boxes over all ranks (e.g. 512 on one GPU, BASELINE.json's "
                    "north-star size: use with --no-extras, the e2e leg would pin 7.7 GB of host memory)")
    ap.add_argument("--no-extras", action="store_true", help="skip the e2e and cpu_baseline legs (profiling runs)")
    ap.add_argument("--no-subrecords", action="store_true", help="N = 1 only: skip the north_star_512 / radhydro / gpu_reference sub-records")
    a = ap.parse_args()
    if a.workload == "radiation" and a.impl == "ours":
        sys.exit(main_radiation(a))
    if a.workload == "radhydro" and a.impl == "ours":
        sys.exit(main_radhydro(a))
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
