"""numpy emulation of the ghost exchange driven by the library's copy-tag plan (host logic only, no GPU):
pack / unpack exactly as csrc/qk_level.cu's k_tags does (message of peer p = for each tag in plan order: ncomp x ncells
doubles at tag.offset * ncomp), so that CPU tests can run the N-rank exchange over gloo."""
import ctypes as C

import numpy as np

from quokka_b200 import capi


def analytic(i, j, k, n, ncell):
    """unique value per (wrapped) cell and component"""
    return (i % ncell[0]) + 1.0e3 * (j % ncell[1]) + 1.0e6 * (k % ncell[2]) + 1.0e9 * (n + 1)


class HostLevel:
    def __init__(self, problem, owner, rank):
        self.lib = capi.load()
        self.p = problem
        self.rank = rank
        self.owner = list(owner)
        self.desc, self._keep = capi.make_level_desc(problem.domain, problem.periodic, problem.dx, problem.nghost, problem.ncomp, problem.boxes, owner, rank,
                                                      problem.bc_lo, problem.bc_hi)
        self.h = C.c_void_p()
        capi.check(self.lib.qk_level_create(C.byref(self.desc), C.byref(self.h)), "qk_level_create")
        self.local_ids = [i for i, o in enumerate(owner) if o == rank]
        ids = (C.c_int32 * max(1, len(self.local_ids)))()
        assert self.lib.qk_level_nlocal(self.h) == len(self.local_ids)
        self.lib.qk_level_local_ids(self.h, ids)
        assert list(ids)[:len(self.local_ids)] == self.local_ids
        self.fabs = {}
        ng = problem.nghost
        for gid in self.local_ids:
            g = problem.boxes[gid].grown(ng)
            nz, ny, nx = g.shape()
            a = np.full((problem.ncomp, nz, ny, nx), np.nan)
            kk, jj, ii = np.meshgrid(np.arange(g.lo[2], g.hi[2] + 1), np.arange(g.lo[1], g.hi[1] + 1), np.arange(g.lo[0], g.hi[0] + 1), indexing="ij")
            for n in range(problem.ncomp):
                a[n, ng:-ng, ng:-ng, ng:-ng] = analytic(ii, jj, kk, n, problem.ncell)[ng:-ng, ng:-ng, ng:-ng]
            self.fabs[gid] = a

    def tags(self, which):
        fn = self.lib.qk_level_remote_tags if which == "remote" else self.lib.qk_level_local_tags
        n = fn(self.h, None, 0)
        arr = (capi.qk_copy_tag * max(1, n))()
        assert fn(self.h, arr, n) == n
        return [arr[i] for i in range(n)]

    def _slice(self, gid, lo, hi):
        g = self.p.boxes[gid].grown(self.p.nghost)
        return (slice(None), slice(lo[2] - g.lo[2], hi[2] - g.lo[2] + 1), slice(lo[1] - g.lo[1], hi[1] - g.lo[1] + 1), slice(lo[0] - g.lo[0], hi[0] - g.lo[0] + 1))

    def read_src(self, t):
        return self.fabs[t.src_box][self._slice(t.src_box, t.src_region.lo, t.src_region.hi)]

    def write_dst(self, t, vals):
        lo = [t.src_region.lo[d] + t.shift[d] for d in range(3)]
        hi = [t.src_region.hi[d] + t.shift[d] for d in range(3)]
        self.fabs[t.dst_box][self._slice(t.dst_box, lo, hi)] = vals

    def fill_local(self):
        for t in self.tags("local"):
            assert t.src_rank == self.rank and t.dst_rank == self.rank
            self.write_dst(t, self.read_src(t))

    def pack(self, peer):
        """message to `peer`: tags in plan order, ncomp x ncells each, at offset*ncomp"""
        mine = [t for t in self.tags("remote") if t.src_rank == self.rank and t.dst_rank == peer]
        total = sum(t.ncells for t in mine)
        buf = np.zeros(total * self.p.ncomp)
        for t in mine:
            v = self.read_src(t)
            assert v[0].size == t.ncells
            buf[t.offset * self.p.ncomp:(t.offset + t.ncells) * self.p.ncomp] = v.reshape(-1)
        return buf

    def recv_size(self, peer):
        return sum(t.ncells for t in self.tags("remote") if t.dst_rank == self.rank and t.src_rank == peer) * self.p.ncomp

    def unpack(self, peer, buf):
        for t in self.tags("remote"):
            if t.dst_rank == self.rank and t.src_rank == peer:
                n = [t.src_region.hi[d] - t.src_region.lo[d] + 1 for d in range(3)]
                v = buf[t.offset * self.p.ncomp:(t.offset + t.ncells) * self.p.ncomp].reshape(self.p.ncomp, n[2], n[1], n[0])
                self.write_dst(t, v)

    def peers(self):
        s = set()
        for t in self.tags("remote"):
            s.add(t.dst_rank if t.src_rank == self.rank else t.src_rank)
        return sorted(s)

    def check_ghosts(self):
        """every ghost cell that (after periodic wrap) lies inside the domain holds the analytic value; the others
        (physical-boundary ghosts) are untouched"""
        p = self.p
        ng = p.nghost
        for gid in self.local_ids:
            g = p.boxes[gid].grown(ng)
            kk, jj, ii = np.meshgrid(np.arange(g.lo[2], g.hi[2] + 1), np.arange(g.lo[1], g.hi[1] + 1), np.arange(g.lo[0], g.hi[0] + 1), indexing="ij")
            inside = np.ones(ii.shape, bool)
            for d, x in enumerate((ii, jj, kk)):
                if not p.periodic[d]:
                    inside &= (x >= 0) & (x < p.ncell[d])
            a = self.fabs[gid]
            for n in range(p.ncomp):
                want = analytic(ii, jj, kk, n, p.ncell)
                assert np.array_equal(a[n][inside], want[inside]), (gid, n)
                assert np.isnan(a[n][~inside]).all(), (gid, n)

    def close(self):
        self.lib.qk_level_destroy(self.h)
