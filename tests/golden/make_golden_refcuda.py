#!/usr/bin/env python3
"""Generate tests/golden/refcuda_hashes.json: SHA-256 digests of the final state of the reference's own CPU executables
(oracle/_ref, built from /root/reference by oracle/ref_build/Makefile) on the parity cases of scripts/gpu_refcuda.py -- the
same inputs the B200 runs of the stock CUDA build and of the build routed through libquokka_b200 use.  A digest covers the
conserved components of every FAB of every AMR level (FABs ordered by (level, lo)), i.e. it also pins the grids the
reference's regridding chose.  Run in the build container only:   python tests/golden/make_golden_refcuda.py [case ...]"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gpu_refcuda as g  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
EXE = {"sedov": "test_hydro3d_blast", "sod": "test_hydro_shocktube_3d", "shell": "shell_golden"}
OUT = os.path.join(HERE, "refcuda_hashes.json")


def run_case(name):
    kind, inputs, ncomp = g.CASES[name]
    base = tempfile.mkdtemp(prefix=f"qkgold_{name}_")
    wd = os.path.join(base, "run")
    os.makedirs(wd)
    try:
        with open(os.path.join(wd, "in"), "w") as f:
            f.write(inputs)
        if kind == "shell":
            shutil.copy("/root/reference/extern/dust_shell/initial_conditions.txt", os.path.join(wd, "initial_conditions.txt"))
        if kind == "sod":
            os.makedirs(os.path.join(base, "extern", "ppm1d"))
            shutil.copy("/root/reference/extern/ppm1d/output", os.path.join(base, "extern", "ppm1d", "output"))
        env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
        log = subprocess.run([os.path.join(REF, EXE[kind]), "in"], cwd=wd, env=env, capture_output=True, text=True).stdout
        plt = g.last_plt(wd)
        fabs = g.read_plotfile_levels(plt)
        rec = {"sha256": g.plotfile_sha(fabs, ncomp), "plotfile": os.path.basename(plt), "nfabs": len(fabs), "levels": sorted({k[0] for k in fabs}),
               "cells": int(sum(v[0].size for v in fabs.values())), "exe": f"oracle/_ref/{EXE[kind]}", "ncomp": ncomp}
        zu = re.findall(r"Zone-updates on level (\d+): (\d+)", log)
        if zu:
            rec["zone_updates_per_level"] = [int(b) for _, b in zu]
        m = re.search(r"[Rr]elative (?:rms )?L1 (?:error )?norm = (\S+)", log)
        if m:
            rec["l1_error_norm"] = m.group(1)
        fom = re.search(r"\[(\S+) Mupdates/s\]", log)
        if fom:
            rec["cpu_fom_Mupdates_s"] = float(fom.group(1))
        return rec
    finally:
        shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(g.CASES)
    out = {}
    if os.path.exists(OUT):
        out = json.load(open(OUT))
    for n in names:
        out[n] = run_case(n)
        print(n, out[n], flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
