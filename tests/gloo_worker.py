"""torchrun worker (CPU, gloo): the N-rank ghost exchange of the level's copy-tag plan -- local copies, then per-peer
packed messages sent with torch.distributed, in the layout csrc/qk_level.cu's pack/unpack kernels use -- and the two
scalar reductions of a time step (max signal speed -> dt, ncells_bad sum), checked against the analytic field.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gloo_worker.py 32 16 1"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    n, b, periodic = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    from host_plan_lib import HostLevel
    from quokka_b200.problems import SedovProblem, distribute

    prob = SedovProblem(n, b)
    prob.periodic = (periodic,) * 3
    owner = distribute(prob.boxes, world)
    L = HostLevel(prob, owner, rank)
    L.fill_local()
    reqs, recv = [], {}
    for peer in L.peers():
        recv[peer] = torch.empty(L.recv_size(peer), dtype=torch.float64)
        reqs.append(dist.irecv(recv[peer], src=peer))
    sends = []
    for peer in L.peers():
        sends.append(torch.from_numpy(L.pack(peer)))
        reqs.append(dist.isend(sends[-1], dst=peer))
    for r in reqs:
        r.wait()
    for peer, buf in recv.items():
        L.unpack(peer, buf.numpy())
    L.check_ghosts()
    # reductions: ParallelDescriptor::ReduceRealMax / ReduceLongSum (AMReX_ParallelDescriptor.cpp:1091,1659)
    vmax = torch.tensor([float(max(L.local_ids))])
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    assert vmax.item() == len(prob.boxes) - 1
    nbad = torch.tensor([len(L.local_ids)], dtype=torch.int64)
    dist.all_reduce(nbad, op=dist.ReduceOp.SUM)
    assert nbad.item() == len(prob.boxes)
    L.close()
    dist.barrier()
    if rank == 0:
        print("GLOO_EXCHANGE_OK", world, len(prob.boxes))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
