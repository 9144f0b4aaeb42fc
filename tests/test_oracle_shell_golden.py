"""Config C4 (RadhydroShell: hydro PLM + radiation subcycles + matter-radiation source terms) end to end: the oracle's level
driver -- radhydro time step, hydro advance with retries, subcycleRadiationAtLevel = transport stage 1, source terms, transport
stage 2, source terms, ten substeps per hydro step -- against state dumps of the reference's own problem file
(tests/golden/shell*.npz, shell_hashes.json, made by tests/golden/make_golden_shell.py).  Bar: bit-exact on all 10 components
after every coarse step, identical time, substep count and Newton-Raphson statistics."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200.capi import QK_RAD_SOURCE_NCOUNTERS, make_level_desc, qk_array4
from quokka_b200.problems import ShellProblem

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shell_energy_source(prob):
    """radEnergySource per box from the oracle's restatement of SetRadEnergySource (host FABs, valid cells)"""
    o = ol.oracle()
    fabs = []
    dx = (C.c_double * 3)(*prob.dx)
    lo = (C.c_double * 3)(0.0, 0.0, 0.0)
    hi = (C.c_double * 3)(prob.prob_hi, prob.prob_hi, prob.prob_hi)
    for bx in prob.boxes:
        f = ol.HostFab(bx, 1)
        o.orc_shell_rad_energy_source(C.byref(f.desc()), C.byref(bx), dx, lo, hi)
        fabs.append(f)
    return fabs


def gather(o, L, prob):
    out = np.zeros((prob.ncomp,) + tuple(reversed(prob.ncell)))
    ng = prob.nghost
    for b, bx in enumerate(prob.boxes):
        d = o.orc_level_state(L, 0, b)
        nz, ny, nx = bx.grown(ng).shape()
        buf = np.ctypeslib.as_array(C.cast(d.p, C.POINTER(C.c_double)), shape=(prob.ncomp, nz, ny, nx))
        out[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1] = buf[:, ng:nz - ng, ng:ny - ng, ng:nx - ng]
    return out


def run_oracle_shell(prob, nsteps):
    """yields (state, time, nsub, counters of the last substep pair) after every coarse step"""
    hp, rp, sp = prob.params(), prob.rad_params(), prob.rad_source_params()
    desc, keep = make_level_desc(prob.domain, prob.periodic, prob.dx, prob.nghost, prob.ncomp, prob.boxes, [0] * len(prob.boxes), 0, prob.bc_lo,
                                 prob.bc_hi)
    o = ol.oracle()
    L = o.orc_level_create(C.byref(desc))
    try:
        for b, bx in enumerate(prob.boxes):
            d = o.orc_level_state(L, 0, b)
            n = prob.ncomp * bx.grown(prob.nghost).ncells()
            np.ctypeslib.as_array(C.cast(d.p, C.POINTER(C.c_double)), shape=(n,))[:] = prob.initial_state(bx).ravel()
        src = shell_energy_source(prob)
        esrc = (qk_array4 * len(src))(*[f.desc() for f in src])
        t = 0.0
        for _ in range(nsteps):
            dt = o.orc_compute_timestep_radhydro(L, C.byref(hp), C.byref(rp), prob.max_substeps, prob.cfl, t, prob.stop_time)
            assert o.orc_step_with_retries(L, C.byref(hp), dt, prob.cfl) == 0
            cnt = (C.c_int64 * QK_RAD_SOURCE_NCOUNTERS)()
            with np.errstate(all="ignore"):
                nsub = o.orc_rad_subcycle_level(L, C.byref(hp), C.byref(rp), C.byref(sp), esrc, dt, prob.rad_cfl, cnt)
            t += dt
            yield gather(o, L, prob), t, dt, nsub, list(cnt)
    finally:
        o.orc_level_destroy(L)


def test_oracle_matches_reference_shell_run():
    g = np.load(os.path.join(GOLD, "shell16_b8_s3.npz"))
    ref = g["states"]
    prob = ShellProblem(int(g["ncell"]), int(g["box"]), initial=ref[0])
    ncell = int(g["ncell"]) ** 3
    for n, (state, t, dt, nsub, cnt) in enumerate(run_oracle_shell(prob, ref.shape[0] - 1)):
        assert dt == float(g["dts_printed"][n]) or abs(dt - float(g["dts_printed"][n])) <= 1e-10 * dt  # the log prints 11 digits
        assert t == float(g["times"][n + 1])
        assert nsub == int(g["nsub"][n])
        assert np.array_equal(state, ref[n + 1]), f"step {n + 1}: max rel diff {np.abs(state / ref[n + 1] - 1).max()}"
        assert cnt[4] == 0 and cnt[6] == 0  # the reference aborts on either (QuokkaSimulation.hpp:1682-1690)
        assert cnt[0] >= 2 * nsub * ncell


def test_oracle_matches_reference_shell_digest():
    h = json.load(open(os.path.join(GOLD, "shell_hashes.json")))["shell32_b16_s2"]
    init = np.load(os.path.join(GOLD, "shell32_b16_initial.npz"))["state"]
    assert hashlib.sha256(np.ascontiguousarray(init).tobytes()).hexdigest() == h["sha256_initial"]
    prob = ShellProblem(h["ncell"], h["box"], initial=init)
    last = None
    for last in run_oracle_shell(prob, h["nsteps"]):
        pass
    state, t, dt, nsub, cnt = last
    assert repr(t) == h["time"] and nsub == h["nsub"][-1]
    assert hashlib.sha256(np.ascontiguousarray(state).tobytes()).hexdigest() == h["sha256_final"]


def test_shell_run_does_not_depend_on_the_box_decomposition():
    """size-independent property: the same 16^3 problem in one 16^3 box, eight 8^3 boxes, or 16 x 8 x 8 slabs gives the same bits
    (ghost fill, per-box transport with in-place stage 2, cell-local source terms)"""
    g = np.load(os.path.join(GOLD, "shell16_b8_s3.npz"))
    ref = g["states"]
    for box in (16, (16, 8, 8)):
        prob = ShellProblem(16, box, initial=ref[0])
        for n, (state, t, dt, nsub, cnt) in enumerate(run_oracle_shell(prob, 2)):
            assert np.array_equal(state, ref[n + 1]), (box, n)
