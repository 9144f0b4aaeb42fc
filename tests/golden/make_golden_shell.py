#!/usr/bin/env python3
"""Generate tests/golden/shell*.npz / shell_hashes.json (config C4: hydro PLM + radiation subcycles + matter-radiation source
terms) by running the reference's own RadhydroShell problem file behind oracle/ref_build/shell_golden.cpp (the stock
problem_main disables plotfiles after reading the inputs; ours sets the same parameters and leaves them to the inputs file)
and reading its plotfiles.  Run in the build container only:   python tests/golden/make_golden_shell.py

Each .npz holds the reference's state_new_cc_ (10 components: 6 gas + E_r, F_r; valid cells) at step 0 (the initial condition,
which comes from an interpolation table and libm pow calls and is therefore taken from the reference, not re-derived) and after
each coarse step, the dt and time the reference prints, the number of radiation substeps, and the Newton-Raphson statistics it
prints per substep (radiation.print_iteration_counts = 1)."""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import ROOT, read_plotfile

EXE = os.path.join(ROOT, "oracle", "_ref", "shell_golden")
TABLE = "/root/reference/extern/dust_shell/initial_conditions.txt"
OUT = os.path.dirname(os.path.abspath(__file__))
PROB_HI = 3.086e19

INPUT = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  {hi}  {hi}  {hi}
geometry.is_periodic =  1    1    1
amr.v = 1
amr.max_level = 0
amr.n_error_buf = 3
amr.grid_eff = 0.7
do_reflux = 0
do_subcycle = 0
checkpoint_interval = -1
radiation.print_iteration_counts = 1
amr.n_cell = {n} {n} {n}
amr.max_grid_size = {box}
amr.blocking_factor = {box}
max_timesteps = {steps}
plotfile_interval = 1
"""


def run_reference(ncell, box, nsteps, threads=8):
    tmp = tempfile.mkdtemp(prefix="qkshell_")
    try:
        shutil.copy(TABLE, os.path.join(tmp, "initial_conditions.txt"))
        with open(os.path.join(tmp, "in"), "w") as f:
            f.write(INPUT.format(hi=repr(PROB_HI), n=ncell, box=box, steps=nsteps))
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        log = subprocess.run([EXE, "in"], cwd=tmp, env=env, capture_output=True, text=True).stdout
        dts = [float(x) for x in re.findall(r"ADVANCE with time = \S+ dt = (\S+)", log)]
        nsub = [int(x) for x in re.findall(r"Radiation substeps: (\d+)", log)]
        stats = [[float(a), float(b), float(c)] for a, b, c in re.findall(
            r"Newton-Raphson solvings per IMEX stage is (\S+), \(mean, max\) number of Newton-Raphson iterations are (\S+), (\d+)\.", log)]
        states, times = [], []
        for n in range(nsteps + 1):
            st, t = read_plotfile(os.path.join(tmp, f"plt{n:05d}"), ncomp_keep=10)
            states.append(st)
            times.append(t)
        return np.stack(states), np.array(times), np.array(dts), np.array(nsub), np.array(stats)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    if not os.path.exists(EXE):
        sys.exit(f"{EXE} missing: make -C oracle/ref_build shell_golden")
    states, times, dts, nsub, stats = run_reference(16, 8, 3)
    np.savez_compressed(os.path.join(OUT, "shell16_b8_s3.npz"), states=states, times=times, dts_printed=dts, nsub=nsub, nr_stats=stats, ncell=16, box=8,
                        prob_hi=PROB_HI)
    print("shell16_b8_s3", states.shape, "t =", times, "nsub", nsub, "sum(Er) =", repr(states[-1][6].sum()))
    hashes = {}
    for ncell, box, nsteps in [(32, 16, 2)]:
        states, times, dts, nsub, stats = run_reference(ncell, box, nsteps)
        hashes[f"shell{ncell}_b{box}_s{nsteps}"] = {
            "ncell": ncell, "box": box, "nsteps": nsteps, "time": repr(float(times[-1])), "nsub": [int(x) for x in nsub],
            "sha256_initial": hashlib.sha256(np.ascontiguousarray(states[0]).tobytes()).hexdigest(),
            "sha256_final": hashlib.sha256(np.ascontiguousarray(states[-1]).tobytes()).hexdigest(),
            "sum_Er_final": repr(float(states[-1][6].sum())), "nr_stats_last": [float(x) for x in stats[-1]]}
        np.savez_compressed(os.path.join(OUT, f"shell{ncell}_b{box}_initial.npz"), state=states[0].astype(np.float64))
        print(ncell, box, nsteps, hashes)
    json.dump(hashes, open(os.path.join(OUT, "shell_hashes.json"), "w"), indent=1)
