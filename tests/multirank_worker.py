"""torchrun worker: Sedov N^3 in B^3 boxes over WORLD_SIZE ranks (one GPU each, NCCL ghost exchange through
libquokka_b200's communicator), K steps; every rank bit-compares its boxes with the C oracle's single-process run.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multirank_worker.py 64 32 6"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    n, b, steps = (int(x) for x in sys.argv[1:4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from quokka_b200.problems import SedovProblem
    from quokka_b200.simulation import Communicator, HydroSimulation
    from test_oracle_golden import run_oracle_sedov

    def bcast(x):
        o = [x]
        dist.broadcast_object_list(o, src=0)
        return o[0]

    comm = Communicator(rank, world, bcast)
    prob = SedovProblem(n, b)
    sim = HydroSimulation(prob, nranks=world, rank=rank, comm=comm)
    sim.setInitialConditions()
    nd, _, _ = sim.evolve(steps)
    assert nd == steps
    mine = sim.state_valid()
    t = sim.time
    ref, t_ref, _ = run_oracle_sedov(n, b, steps)
    ok = (t == t_ref)
    for gid, a in mine.items():
        bx = prob.boxes[gid]
        want = ref[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1]
        ok = ok and np.array_equal(a, want)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sim.close()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK_OK" if int(flag.item()) == 1 else "MULTIRANK_MISMATCH", world, len(mine), t)
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main_shell():
    """config C4 over WORLD_SIZE ranks: RadhydroShell 16^3 in 8^3 boxes from the reference's step-0 dump, K coarse steps (hydro + ten
    radiation substeps with source terms), every rank compares its boxes with the reference's own dumps (bar 1e-10 of max|comp|)
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multirank_worker.py shell 2"""
    steps = int(sys.argv[2])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from quokka_b200.device import DevMultiFab
    from quokka_b200.problems import ShellProblem
    from quokka_b200.simulation import Communicator, HydroSimulation
    from test_oracle_shell_golden import GOLD, shell_energy_source

    def bcast(x):
        o = [x]
        dist.broadcast_object_list(o, src=0)
        return o[0]

    comm = Communicator(rank, world, bcast)
    g = np.load(os.path.join(GOLD, "shell16_b8_s3.npz"))
    ref = g["states"]
    prob = ShellProblem(int(g["ncell"]), int(g["box"]), initial=ref[0])
    sim = HydroSimulation(prob, nranks=world, rank=rank, comm=comm)
    src = shell_energy_source(prob)
    esrc = DevMultiFab(sim.local_boxes, 1, ngrow=0, host=[src[i].a for i in sim.local_ids])
    sim.enableRadiation(prob.rad_params(), prob.rad_source_params(), esrc, rad_cfl=prob.rad_cfl, max_substeps=prob.max_substeps)
    sim.setInitialConditions()
    ok = True
    for n in range(steps):
        dt = sim.computeTimestep()
        ok = ok and sim.advanceSingleTimestepAtLevel(dt) == 0 and sim.radiationSubsteps == int(g["nsub"][n])
    ok = ok and sim.time == float(g["times"][steps])
    want_all = ref[steps]
    scale = np.abs(want_all).reshape(10, -1).max(axis=1)
    worst = 0.0
    for gid, a in sim.state_valid().items():
        bx = prob.boxes[gid]
        want = want_all[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1]
        worst = max(worst, float((np.abs(a - want).reshape(10, -1).max(axis=1) / scale).max()))
    ok = ok and worst <= 1e-10
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sim.close()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK_OK" if int(flag.item()) == 1 else "MULTIRANK_MISMATCH", world, worst)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main_shell() if sys.argv[1] == "shell" else main()
