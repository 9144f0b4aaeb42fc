"""Device-side FABs for the Python harness: torch is used ONLY as the allocator / copy engine of
float64 (int32) CUDA buffers; every numerical operation goes through libquokka_b200.so.

A DevFab is one AMReX FAB (x fastest, component-major, extern/amrex/Src/Base/AMReX_Array4.H:59-68)
on the GPU plus its qk_array4 descriptor.  DevMultiFab is a list of them over a list of boxes,
with the ctypes descriptor arrays the C ABI takes.
"""
from __future__ import annotations

import numpy as np

from .capi import qk_array4, qk_box, qk_iarray4


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("quokka_b200 needs a CUDA device (there is no CPU fallback)")
    return torch


class DevFab:
    def __init__(self, box: qk_box, ncomp: int, dtype="f64", fill=0.0, host: np.ndarray | None = None):
        torch = _torch()
        self.box = box
        self.ncomp = ncomp
        nz, ny, nx = box.shape()
        tdt = torch.float64 if dtype == "f64" else torch.int32
        if host is not None:
            assert host.shape == (ncomp, nz, ny, nx), (host.shape, (ncomp, nz, ny, nx))
            self.t = torch.from_numpy(np.ascontiguousarray(host)).to("cuda")
        else:
            self.t = torch.full((ncomp, nz, ny, nx), fill, dtype=tdt, device="cuda")
        self.is_int = dtype != "f64"

    def desc(self):
        d = qk_iarray4() if self.is_int else qk_array4()
        nz, ny, nx = self.box.shape()
        d.p = self.t.data_ptr()
        d.jstride = nx
        d.kstride = nx * ny
        d.nstride = nx * ny * nz
        d.begin[:] = list(self.box.lo)
        d.end[:] = [self.box.hi[i] + 1 for i in range(3)]
        d.ncomp = self.ncomp
        return d

    def numpy(self) -> np.ndarray:
        return self.t.cpu().numpy()

    def view(self, box: qk_box) -> np.ndarray:
        """host copy (ncomp, nz, ny, nx) of the sub-box"""
        o = [box.lo[d] - self.box.lo[d] for d in range(3)]
        s = [box.hi[d] - box.lo[d] + 1 for d in range(3)]
        return self.t[:, o[2]:o[2] + s[2], o[1]:o[1] + s[1], o[0]:o[0] + s[0]].cpu().numpy()


class DevMultiFab:
    """A MultiFab's local part: one DevFab per box, grown by `ngrow`, optionally nodal in `face_dir`."""

    def __init__(self, boxes, ncomp, ngrow=0, face_dir=None, dtype="f64", fill=0.0, host=None):
        self.valid = list(boxes)
        self.fabs = []
        for i, b in enumerate(self.valid):
            g = b.grown(ngrow, face_dir)
            self.fabs.append(DevFab(g, ncomp, dtype=dtype, fill=fill, host=None if host is None else host[i]))
        self.ncomp = ncomp
        self.is_int = dtype != "f64"
        self._refresh()

    def _refresh(self):
        n = len(self.fabs)
        T = qk_iarray4 if self.is_int else qk_array4
        self.descs = (T * n)(*[f.desc() for f in self.fabs])
        self.boxes_c = (qk_box * n)(*self.valid)

    def __len__(self):
        return len(self.fabs)

    def numpy(self):
        return [f.numpy() for f in self.fabs]
