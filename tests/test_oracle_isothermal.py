"""The isothermal EOS (EOS_Traits::gamma = 1, HydroSystem::is_eos_isothermal(), src/hydro/hydro_system.hpp:133) in the oracle's level driver.
The operators' isothermal branches are pinned bit-exactly to the reference's templates by tests/test_oracle_vs_ref.py (harness problem 3);
here the whole time loop (compute dt, RK2 stages, FOFC, retries) runs on an isothermal problem and must keep what the scheme conserves."""
import numpy as np

from quokka_b200.problems import IsothermalWaveProblem
from test_oracle_golden import run_oracle_problem


def test_isothermal_run_conserves_mass_and_momentum():
    prob = IsothermalWaveProblem(16, 8)
    init = np.zeros((prob.ncomp, 16, 16, 16))
    ng = prob.nghost
    for bx in prob.boxes:
        init[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1] = prob.initial_state(bx)[:, ng:-ng, ng:-ng, ng:-ng]
    state, t, retries = run_oracle_problem(prob, 12)
    assert t > 0 and np.isfinite(state).all() and (state[0] > 0).all()
    assert np.abs(state[0] - init[0]).max() > 1e-2  # the gas moved
    for c in range(4):  # periodic box, conservative update: totals change by rounding only
        scale = np.abs(init[c]).sum()
        assert abs(state[c].sum() - init[c].sum()) < 1e-12 * scale, c
    # the time step follows c_s = cs_isothermal: dt <= cfl dx / (cs + |v|) (hydro_system.hpp:242-246, simulation.hpp:703-720)
    vmax = (np.sqrt(init[1] ** 2 + init[2] ** 2 + init[3] ** 2) / init[0]).max()
    assert t < 12 * prob.cfl * prob.dx[0] / prob.cs_isothermal and t > 1.0 * prob.cfl * prob.dx[0] / (prob.cs_isothermal + 3 * vmax)
