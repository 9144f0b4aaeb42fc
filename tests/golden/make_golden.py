#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/test_hydro3d_blast,
built from /root/reference by oracle/ref_build/Makefile) and reading its own plotfiles.

Run in the build container only (needs oracle/_ref):   python tests/golden/make_golden.py
Each fixture holds the reference's state_new_cc_ (6 conserved components, valid cells) after N
coarse steps of the Sedov blast (src/problems/HydroBlast3D/test_hydro3d_blast.cpp) plus the run
configuration and the simulation time the reference reports.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(ROOT, "oracle", "_ref", "test_hydro3d_blast")
OUT = os.path.dirname(os.path.abspath(__file__))

BASE_INPUT = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  1.2  1.2  1.2
geometry.is_periodic =  0    0    0
amr.v = 1
amr.max_level = 0
amr.n_error_buf = 3
amr.grid_eff = 0.7
do_reflux = 0
do_subcycle = 0
"""


def read_plotfile(path, ncomp_keep=6):
    """Minimal AMReX plotfile reader (single level, native FP64 FABs)."""
    with open(os.path.join(path, "Header")) as f:
        lines = f.read().split("\n")
    ncomp = int(lines[1])
    time = float(lines[2 + ncomp + 1])
    dom = re.findall(r"\((-?\d+),(-?\d+),(-?\d+)\)", lines[2 + ncomp + 1 + 5])
    lo = [int(x) for x in dom[0]]
    hi = [int(x) for x in dom[1]]
    n = [hi[d] - lo[d] + 1 for d in range(3)]
    out = np.zeros((ncomp, n[2], n[1], n[0]))
    with open(os.path.join(path, "Level_0", "Cell_H")) as f:
        cellh = f.read()
    fods = re.findall(r"FabOnDisk: (\S+) (\d+)", cellh)
    for fname, off in fods:
        with open(os.path.join(path, "Level_0", fname), "rb") as f:
            f.seek(int(off))
            hdr = f.readline().decode()
            m = re.search(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \(\d+,\d+,\d+\)\) (\d+)", hdr)
            blo = [int(m.group(i)) for i in (1, 2, 3)]
            bhi = [int(m.group(i)) for i in (4, 5, 6)]
            nc = int(m.group(7))
            bn = [bhi[d] - blo[d] + 1 for d in range(3)]
            data = np.frombuffer(f.read(8 * nc * bn[0] * bn[1] * bn[2]), dtype="<f8").reshape(nc, bn[2], bn[1], bn[0])
            out[:, blo[2] - lo[2]:bhi[2] - lo[2] + 1, blo[1] - lo[1]:bhi[1] - lo[1] + 1, blo[0] - lo[0]:bhi[0] - lo[0] + 1] = data
    return out[:ncomp_keep], time


def run_reference(ncell, box, nsteps, threads=8):
    tmp = tempfile.mkdtemp(prefix="qkgold_")
    try:
        with open(os.path.join(tmp, "in"), "w") as f:
            f.write(BASE_INPUT)
            f.write(f"amr.n_cell = {ncell} {ncell} {ncell}\namr.max_grid_size = {box}\namr.blocking_factor = {box}\n")
            f.write(f"max_timesteps = {nsteps}\nplotfile_interval = {nsteps}\ncheckpoint_interval = -1\n")
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        log = subprocess.run([EXE, "in"], cwd=tmp, env=env, capture_output=True, text=True).stdout
        dts = [float(x) for x in re.findall(r"ADVANCE with time = \S+ dt = (\S+)", log)]
        retries = len(re.findall(r"Re-trying hydro advance", log))
        state, time = read_plotfile(os.path.join(tmp, f"plt{nsteps:05d}"))
        return state, time, dts, retries
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


CASES = [  # (name, ncell, box, nsteps)
    ("sedov16_b16_s5", 16, 16, 5),
    ("sedov32_b16_s10", 32, 16, 10),
    ("sedov32_b32_s30", 32, 32, 30),
]

if __name__ == "__main__":
    if not os.path.exists(EXE):
        sys.exit(f"{EXE} missing: make -C oracle/ref_build sedov")
    for name, ncell, box, nsteps in CASES:
        state, time, dts, retries = run_reference(ncell, box, nsteps)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), state=state, time=time, dts_printed=np.array(dts), ncell=ncell, box=box,
                            nsteps=nsteps, retries=retries)
        print(name, state.shape, "t =", repr(time), "retries", retries, "sum(E) =", repr(state[4].sum()))
