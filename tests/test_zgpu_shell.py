"""GPU parity of the radiation subcycle and of config C4 (RadhydroShell) end to end, through the C ABI.

* qk_rad_subcycle (transport stage 1, source terms, transport stage 2, source terms, nsub substeps) against the oracle's
  orc_rad_subcycle_level on seeded states;
* the C++ driver with radiation enabled (qk_sim_enable_radiation: hydro PLM advance + subcycleRadiationAtLevel, radhydro time step)
  against the state dumps of the REFERENCE's own RadhydroShell problem file (tests/golden/shell16_b8_s3.npz): dt, time and
  substep count exact, all 10 components within the tolerance the source terms carry.

Tolerance: the hydro and transport kernels are bit-exact; the source terms differ from the CPU reference in the rounding of
T^4 (tests/test_rad_source_host.py), so the bar is L-infinity per component <= 1e-10 of the component's maximum."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from quokka_b200 import capi
from quokka_b200.capi import QK_RAD_SOURCE_NCOUNTERS, check, make_level_desc, qk_array4, qk_box
from quokka_b200.problems import ShellProblem, chop_domain
from test_oracle_shell_golden import GOLD, run_oracle_shell, shell_energy_source
from test_rad_source_host import trait_set

pytestmark = pytest.mark.gpu
TOL = 1e-10


def linf(a, b):
    return [float(np.abs(a[c] - b[c]).max() / max(np.abs(b[c]).max(), 1e-300)) for c in range(a.shape[0])]


@pytest.mark.parametrize("name,dt_hydro,dxv", [("shell", 1.2e10, 3.0e18), ("beta0", 2.0e-3, 0.02)])  # nsub = 3 and 2
def test_rad_subcycle_vs_oracle(name, dt_hydro, dxv):
    from quokka_b200.device import DevMultiFab

    hp, rp, sp, gen = trait_set(name)
    rp.reconstruction_order = 2
    ncell, ng, nc = (16, 8, 8), 4, 10
    boxes = chop_domain(list(ncell), 8)
    domain = qk_box.make((0, 0, 0), tuple(n - 1 for n in ncell))
    dx = [dxv] * 3
    bc = [capi.QK_BC_INT_DIR] * (3 * nc)
    desc, keep = make_level_desc(domain, (1, 1, 1), dx, ng, nc, boxes, [0] * len(boxes), 0, bc, bc)
    states = [ol.random_radhydro_cons(b.grown(ng), hp, rp, sp, seed=9 + n, T0=gen["T0"], rho0=gen["rho0"], vmax=gen["vmax"], spread=0.3, fmax=0.5)
              for n, b in enumerate(boxes)]
    # oracle
    o = ol.oracle()
    L = o.orc_level_create(C.byref(desc))
    for which in (0, 1):
        for b, st in enumerate(states):
            d = o.orc_level_state(L, which, b)
            np.ctypeslib.as_array(C.cast(d.p, C.POINTER(C.c_double)), shape=st.shape)[...] = st
    co = (C.c_int64 * QK_RAD_SOURCE_NCOUNTERS)()
    with np.errstate(all="ignore"):
        nsub_o = o.orc_rad_subcycle_level(L, C.byref(hp), C.byref(rp), C.byref(sp), None, dt_hydro, 0.3, co)
    want = []
    for b, st in enumerate(states):
        d = o.orc_level_state(L, 0, b)
        want.append(np.ctypeslib.as_array(C.cast(d.p, C.POINTER(C.c_double)), shape=st.shape)[:, ng:-ng, ng:-ng, ng:-ng].copy())
    o.orc_level_destroy(L)
    assert nsub_o >= 2
    # GPU
    lib = capi.load()
    lev = C.c_void_p()
    check(lib.qk_level_create(C.byref(desc), C.byref(lev)))
    Uold = DevMultiFab(boxes, nc, ngrow=ng, host=states)
    Unew = DevMultiFab(boxes, nc, ngrow=ng, host=states)
    Utmp = DevMultiFab(boxes, nc, ngrow=ng, fill=0.0)
    cg = (C.c_int64 * QK_RAD_SOURCE_NCOUNTERS)()
    nsub = C.c_int()
    check(lib.qk_rad_subcycle(lev, C.byref(hp), C.byref(rp), C.byref(sp), Uold.descs, Unew.descs, Utmp.descs, None, dt_hydro, 0.3, cg, C.byref(nsub), None))
    got = [a[:, ng:-ng, ng:-ng, ng:-ng] for a in Unew.numpy()]
    lib.qk_level_destroy(lev)
    assert nsub.value == nsub_o
    g, w = np.concatenate([x.reshape(nc, -1) for x in got], axis=1), np.concatenate([x.reshape(nc, -1) for x in want], axis=1)
    assert np.isfinite(g).all()
    err = linf(g, w)
    assert max(err) <= TOL, err
    assert np.array_equal(g[0], w[0])  # density is not touched
    if co[4] == 0 and co[6] == 0:
        assert cg[0] == co[0] or abs(cg[0] - co[0]) <= max(2, co[0] // 1000)


def test_config_c4_shell_against_the_reference_dumps():
    from quokka_b200.device import DevMultiFab
    from quokka_b200.simulation import HydroSimulation

    g = np.load(os.path.join(GOLD, "shell16_b8_s3.npz"))
    ref = g["states"]
    prob = ShellProblem(int(g["ncell"]), int(g["box"]), initial=ref[0])
    sim = HydroSimulation(prob)
    src = shell_energy_source(prob)  # host FABs from the oracle's restatement of SetRadEnergySource: an INPUT of the run
    esrc = DevMultiFab(prob.boxes, 1, ngrow=0, host=[f.a for f in src])
    sim.enableRadiation(prob.rad_params(), prob.rad_source_params(), esrc, rad_cfl=prob.rad_cfl, max_substeps=prob.max_substeps)
    sim.setInitialConditions()
    worst = 0.0
    for n in range(ref.shape[0] - 1):
        dt = sim.computeTimestep()
        assert abs(dt - float(g["dts_printed"][n])) <= 1e-10 * dt  # the reference's log prints 11 digits
        assert sim.advanceSingleTimestepAtLevel(dt) == 0
        assert sim.radiationSubsteps == int(g["nsub"][n])
        assert sim.time == float(g["times"][n + 1])
        got = sim.gather_global()
        assert np.isfinite(got).all()
        err = linf(got, ref[n + 1])
        worst = max(worst, max(err))
        assert max(err) <= TOL, (n, err)
    sim.close()
    print(f"C4 shell 16^3, 3 coarse steps (30 radiation substeps): worst L-inf / max = {worst:.3e}")
