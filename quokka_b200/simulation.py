"""Host-side mirror of the reference's simulation surface for ONE uniform level, bound to the C++ driver
inside libquokka_b200.so (csrc/qk_sim.cu).  Method names follow AMRSimulation<problem_t> /
QuokkaSimulation<problem_t> (src/simulation.hpp:141-408, src/QuokkaSimulation.hpp:64-282):
setInitialConditions, computeTimestep, advanceSingleTimestepAtLevel, evolve.  No numerics here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, make_level_desc, qk_array4


class Communicator:
    """NCCL communicator of the library (csrc/qk_comm.cpp).  `bcast_bytes(buf: bytearray|None) -> bytes` is the
    host application's bootstrap (torch.distributed.broadcast_object_list in bench.py; MPI_Bcast in Quokka)."""

    def __init__(self, rank: int, nranks: int, bcast_bytes):
        self.lib = capi.load()
        ident = None
        if rank == 0:
            buf = (C.c_char * 128)()
            check(self.lib.qk_comm_unique_id(C.cast(buf, C.c_void_p)), "qk_comm_unique_id")
            ident = bytes(buf.raw)
        ident = bcast_bytes(ident)
        self._id = (C.c_char * 128).from_buffer_copy(ident)
        self.handle = C.c_void_p()
        check(self.lib.qk_comm_create(C.cast(self._id, C.c_void_p), rank, nranks, C.byref(self.handle)), "qk_comm_create")
        self.rank, self.nranks = rank, nranks

    def close(self):
        if self.handle:
            self.lib.qk_comm_destroy(self.handle)
            self.handle = C.c_void_p()


class HydroSimulation:
    """QuokkaSimulation<problem_t> restricted to one uniform level, hydro only."""

    def __init__(self, problem, nranks: int = 1, rank: int = 0, comm: Communicator | None = None, owner=None, params=None):
        from .problems import distribute

        self.lib = capi.load()
        self.problem = problem
        self.rank, self.nranks = rank, nranks
        self.owner = list(owner) if owner is not None else distribute(problem.boxes, nranks)
        self.prm = params if params is not None else problem.params()
        self.desc, self._keep = make_level_desc(problem.domain, problem.periodic, problem.dx, problem.nghost, problem.ncomp, problem.boxes,
                                                self.owner, rank, problem.bc_lo, problem.bc_hi)
        self.handle = C.c_void_p()
        self.comm = comm
        check(self.lib.qk_sim_create(C.byref(self.desc), C.byref(self.prm), problem.cfl, comm.handle if comm else None, C.byref(self.handle)),
              "qk_sim_create")
        self.local_ids = [i for i, o in enumerate(self.owner) if o == rank]
        self.local_boxes = [problem.boxes[i] for i in self.local_ids]
        self._pinned = None

    # -- state transfer (whole FABs incl. ghost cells, AMReX layout) ------------------------------------------
    def _host_buffers(self):
        """pinned host staging, one (ncomp, nz, ny, nx) array per local box"""
        if self._pinned is None:
            import torch

            self._pinned = []
            for b, bx in enumerate(self.local_boxes):
                nz, ny, nx = bx.grown(self.problem.nghost).shape()
                t = torch.empty((self.problem.ncomp, nz, ny, nx), dtype=torch.float64).pin_memory()
                self._pinned.append(t)
        return self._pinned

    def setInitialConditions(self):
        """AMRSimulation::setInitialConditions (simulation.hpp:638) -> problem.initial_state on each local box."""
        bufs = self._host_buffers()
        for b, bx in enumerate(self.local_boxes):
            bufs[b].numpy()[...] = self.problem.initial_state(bx)
        self.upload()
        self.lib.qk_sim_reset_clock(self.handle, 0.0, 1.0e100)

    def upload(self):
        for b, t in enumerate(self._host_buffers()):
            check(self.lib.qk_sim_set_state(self.handle, b, t.data_ptr()), "qk_sim_set_state")

    def download(self):
        for b, t in enumerate(self._host_buffers()):
            check(self.lib.qk_sim_get_state(self.handle, b, t.data_ptr()), "qk_sim_get_state")
        check(self.lib.qk_sim_sync(self.handle), "qk_sim_sync")
        return [t.numpy() for t in self._pinned]

    def h2d_bytes(self):
        return sum(int(t.numel()) * 8 for t in self._host_buffers())

    # -- state transfer, valid cells only (what a host-resident caller owns; the step fills the ghost cells) --
    def _host_valid_buffers(self):
        if getattr(self, "_pinned_valid", None) is None:
            import torch

            self._pinned_valid = []
            for bx in self.local_boxes:
                nz, ny, nx = bx.shape()
                self._pinned_valid.append(torch.empty((self.problem.ncomp, nz, ny, nx), dtype=torch.float64).pin_memory())
        return self._pinned_valid

    def upload_valid(self):
        for b, t in enumerate(self._host_valid_buffers()):
            check(self.lib.qk_sim_set_state_valid(self.handle, b, t.data_ptr()), "qk_sim_set_state_valid")

    def download_valid(self):
        for b, t in enumerate(self._host_valid_buffers()):
            check(self.lib.qk_sim_get_state_valid(self.handle, b, t.data_ptr()), "qk_sim_get_state_valid")
        check(self.lib.qk_sim_sync(self.handle), "qk_sim_sync")
        return [t.numpy() for t in self._pinned_valid]

    def valid_bytes(self):
        return sum(int(t.numel()) * 8 for t in self._host_valid_buffers())

    def state_valid(self):
        """state_new_cc_ on the valid cells of the local boxes -> dict box_id -> (ncomp, nz, ny, nx)"""
        ng = self.problem.nghost
        out = {}
        for gid, a in zip(self.local_ids, self.download()):
            out[gid] = a[:, ng:a.shape[1] - ng, ng:a.shape[2] - ng, ng:a.shape[3] - ng].copy()
        return out

    def gather_global(self):
        """(ncomp, NZ, NY, NX) of the whole domain -- single-rank runs only"""
        assert self.nranks == 1
        p = self.problem
        out = np.zeros((p.ncomp,) + tuple(reversed(p.ncell)))
        for gid, a in self.state_valid().items():
            bx = p.boxes[gid]
            out[:, bx.lo[2]:bx.hi[2] + 1, bx.lo[1]:bx.hi[1] + 1, bx.lo[0]:bx.hi[0] + 1] = a
        return out

    def state_desc(self, which, b):
        d = qk_array4()
        check(self.lib.qk_sim_state_desc(self.handle, which, b, C.byref(d)), "qk_sim_state_desc")
        return d

    # -- time stepping ---------------------------------------------------------------------------------------
    @property
    def time(self):
        return self.lib.qk_sim_time(self.handle)

    @property
    def cellUpdates(self):
        return self.lib.qk_sim_cell_updates(self.handle)

    @property
    def retries(self):
        return self.lib.qk_sim_retries(self.handle)

    def enableRadiation(self, rad_params, source_params=None, rad_energy_source=None, rad_cfl=0.3, max_substeps=10):
        """Physics_Traits::is_radiation_enabled: subcycleRadiationAtLevel after every hydro advance (QuokkaSimulation.hpp:690-694).
        rad_energy_source: a DevMultiFab (one component, no ghost cells) over the local boxes, kept alive here."""
        self._rad = (rad_params, source_params, rad_energy_source)
        check(self.lib.qk_sim_enable_radiation(self.handle, C.byref(rad_params), C.byref(source_params) if source_params is not None else None,
                                               rad_energy_source.descs if rad_energy_source is not None else None, rad_cfl, max_substeps),
              "qk_sim_enable_radiation")

    @property
    def radiationSubsteps(self):
        return self.lib.qk_sim_last_rad_substeps(self.handle)

    def computeTimestep(self, stop_time=None):
        dt = C.c_double()
        check(self.lib.qk_sim_compute_timestep(self.handle, self.problem.stop_time if stop_time is None else stop_time, C.byref(dt)),
              "qk_sim_compute_timestep")
        return dt.value

    def advanceSingleTimestepAtLevel(self, dt):
        r = C.c_int()
        check(self.lib.qk_sim_step(self.handle, dt, C.byref(r)), "qk_sim_step")
        return r.value

    def evolve(self, max_steps, stop_time=None):
        """returns (steps_done, elapsed_s, device_ms)"""
        n, el, ms = C.c_int(), C.c_double(), C.c_double()
        check(self.lib.qk_sim_evolve(self.handle, max_steps, self.problem.stop_time if stop_time is None else stop_time, C.byref(n), C.byref(el),
                                     C.byref(ms)), "qk_sim_evolve")
        return n.value, el.value, ms.value

    def sync(self):
        check(self.lib.qk_sim_sync(self.handle), "qk_sim_sync")

    def close(self):
        if self.handle:
            self.lib.qk_sim_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
