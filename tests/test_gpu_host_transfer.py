"""Host <-> device transfer of the valid cells only (qk_sim_set_state_valid / qk_sim_get_state_valid, the e2e leg of bench.py): a round trip
returns the same bits, and a step driven from valid-cell host buffers equals a step on the resident state bit for bit (the ghost cells are
filled by the step itself, src/simulation.hpp:1704-1785)."""
import numpy as np
import pytest

from quokka_b200.problems import SedovProblem
from quokka_b200.simulation import HydroSimulation

pytestmark = pytest.mark.gpu


def test_valid_cell_round_trip_and_step():
    prob = SedovProblem(32, 16)
    a = HydroSimulation(prob)
    a.setInitialConditions()
    b = HydroSimulation(prob)
    b.setInitialConditions()
    ng = prob.nghost
    full = a.download()
    valid = a.download_valid()
    for f, v in zip(full, valid):
        assert np.array_equal(f[:, ng:-ng, ng:-ng, ng:-ng], v)
    b.download_valid()  # the host now owns b's state
    for _ in range(3):
        # a: resident state; b: host-resident state, valid cells up and down every step (ghost cells of the device copy are stale on purpose)
        b.upload_valid()
        ra = a.advanceSingleTimestepAtLevel(a.computeTimestep())
        rb = b.advanceSingleTimestepAtLevel(b.computeTimestep())
        assert ra == 0 and rb == 0
        vb = b.download_valid()
        for gid, (va, vv) in enumerate(zip(a.state_valid().values(), vb)):
            assert np.array_equal(va, vv), f"box {gid}"
    assert a.time == b.time
    assert b.valid_bytes() < b.h2d_bytes()
    a.close()
    b.close()
