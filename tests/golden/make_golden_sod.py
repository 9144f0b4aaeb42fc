#!/usr/bin/env python3
"""Generate tests/golden/sod*.npz by running the UNMODIFIED reference's Sod shock tube executable
(oracle/_ref/test_hydro_shocktube = src/problems/HydroShocktube/test_hydro_shocktube.cpp built for AMREX_SPACEDIM=1 by
oracle/ref_build/Makefile) on a uniform level (amr.max_level = 0) and reading its own plotfiles.  config C1 of BASELINE.json:
PPM + HLLC, reconstruct_eint = true (the HydroSystem_Traits default), gamma = 1.4, cfl 0.6, Dirichlet walls, t_end = 0.4.

Run in the build container only:   python tests/golden/make_golden_sod.py
The fixtures also carry the exact solution the reference's own pass criterion reads (extern/ppm1d/output: x, rho, P, v).
"""
import glob
import os
import re
import shutil
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(ROOT, "oracle", "_ref", "test_hydro_shocktube")
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

INPUT = """
geometry.prob_lo     =  0.0  0.0  0.0
geometry.prob_hi     =  5.0  1.0  1.0
geometry.is_periodic =  0    1    1
amr.v = 1
amr.max_level = 0
amr.blocking_factor = 16
do_reflux = 0
do_subcycle = 0
checkpoint_interval = -1
cfl = 0.6
hydro.reconstruction_order = 3
"""


def read_plotfile_1d(path, ncomp_keep=6):
    with open(os.path.join(path, "Header")) as f:
        lines = f.read().split("\n")
    ncomp = int(lines[1])
    time = float(lines[2 + ncomp + 1])
    dom = re.findall(r"\(\((-?\d+)\) \((-?\d+)\) \((-?\d+)\)\)", lines[2 + ncomp + 1 + 5])
    lo, hi = int(dom[0][0]), int(dom[0][1])
    out = np.zeros((ncomp, hi - lo + 1))
    with open(os.path.join(path, "Level_0", "Cell_H")) as f:
        cellh = f.read()
    for fname, off in re.findall(r"FabOnDisk: (\S+) (\d+)", cellh):
        with open(os.path.join(path, "Level_0", fname), "rb") as f:
            f.seek(int(off))
            hdr = f.readline().decode()
            m = re.search(r"\(\((-?\d+)\) \((-?\d+)\) \((-?\d+)\)\) (\d+)", hdr)
            blo, bhi, nc = int(m.group(1)), int(m.group(2)), int(m.group(4))
            data = np.frombuffer(f.read(8 * nc * (bhi - blo + 1)), dtype="<f8").reshape(nc, bhi - blo + 1)
            out[:, blo - lo:bhi - lo + 1] = data
    return out[:ncomp_keep], time


def run_reference(ncell, box, plot_every):
    tmp = tempfile.mkdtemp(prefix="qksod_")
    try:
        run = os.path.join(tmp, "run")
        os.makedirs(run)
        os.makedirs(os.path.join(tmp, "extern", "ppm1d"))
        shutil.copy(os.path.join(REF, "extern", "ppm1d", "output"), os.path.join(tmp, "extern", "ppm1d", "output"))  # read by the test at exit
        with open(os.path.join(run, "in"), "w") as f:
            f.write(INPUT + f"amr.n_cell = {ncell} 16 16\namr.max_grid_size = {box}\nplotfile_interval = {plot_every}\n")
        log = subprocess.run([EXE, "in"], cwd=run, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="4")).stdout
        m = re.search(r"[Rr]elative rms L1 error norm = (\S+)", log)
        err = float(m.group(1)) if m else float("nan")
        dts = [float(x) for x in re.findall(r"ADVANCE with time = \S+ dt = (\S+)", log)]
        retry_steps = [m.start() for m in re.finditer(r"Re-trying hydro advance", log)]
        # number of dt-halving retries before each plotfile: count retry messages preceding the plotfile's step banner
        plots = {}
        for p in sorted(glob.glob(os.path.join(run, "plt*"))):
            step = int(os.path.basename(p)[3:])
            plots[step] = read_plotfile_1d(p)
        nretries = len(retry_steps)
        return plots, err, dts, nretries, log
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    exact = np.loadtxt(os.path.join(REF, "extern", "ppm1d", "output"), skiprows=2)  # i, x, rho, P, v  (test_hydro_shocktube.cpp:183-214)
    ex = dict(exact_x=exact[:, 1], exact_rho=exact[:, 2], exact_P=exact[:, 3], exact_v=exact[:, 4])
    plots, err, dts, nretries, log = run_reference(256, 128, 40)
    last = max(plots)
    s40, t40 = plots[40]
    # retries that happened within the first 40 coarse steps
    pos40 = [m.start() for m in re.finditer(r"Coarse STEP 41 ", log)]
    r40 = len(re.findall(r"Re-trying hydro advance", log[:pos40[0]])) if pos40 else nretries
    np.savez_compressed(os.path.join(OUT, "sod256_s40.npz"), state=s40, time=t40, ncell=256, box=128, nsteps=40, retries=r40)
    sl, tl = plots[last]
    np.savez_compressed(os.path.join(OUT, "sod256_full.npz"), state=sl, time=tl, ncell=256, box=128, nsteps=last, ref_l1_error=err, retries=nretries, **ex)
    print("sod256:", last, "steps, reference L1 error", err, "retries", r40, nretries)
    plots, err, dts, nretries, log = run_reference(1024, 128, 100000)
    last = max(plots)
    sl, tl = plots[last]
    np.savez_compressed(os.path.join(OUT, "sod1024_full.npz"), state=sl, time=tl, ncell=1024, box=128, nsteps=last, ref_l1_error=err, retries=nretries, **ex)
    print("sod1024:", last, "steps, reference L1 error", err, "retries", nretries)


if __name__ == "__main__":
    main()
