import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quokka_b200 import capi
from quokka_b200.problems import SedovProblem
from quokka_b200.simulation import HydroSimulation
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = capi.load()
sim = HydroSimulation(SedovProblem(n, 128 if n >= 128 else n))
sim.setInitialConditions()
sim.evolve(3)
l0 = lib.qk_launch_count()
nd, el, ms = sim.evolve(steps)
print(f"Sedov {n}^3: {nd} steps, wall {el*1e3:.1f} ms, device {ms:.1f} ms, {n**3*nd/el/1e6:.1f} Mupdates/s, launches/step {(lib.qk_launch_count()-l0)/nd:.0f}, scratch {lib.qk_level_scratch_bytes(lib.qk_sim_level(sim.handle))/2**30:.2f} GiB")
