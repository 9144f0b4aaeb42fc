"""Host logic of the multi-rank path, on CPU: the copy-tag plan (FabArray::FillBoundary semantics,
extern/amrex/Src/Base/AMReX_FabArrayBase.cpp FB::define_fb: every ghost cell that lies in another box's valid region,
periodic images included, corners included), its per-peer message layout, and the box -> rank map.  The N-rank
exchange itself runs over gloo in tests/test_gloo_exchange.py."""
import numpy as np
import pytest

from host_plan_lib import HostLevel
from quokka_b200.problems import SedovProblem, chop_domain, distribute


class Prob(SedovProblem):
    def __init__(self, ncell, box, periodic):
        super().__init__(ncell, box)
        self.periodic = periodic


@pytest.mark.parametrize("periodic", [(0, 0, 0), (1, 1, 1), (1, 0, 1)])
@pytest.mark.parametrize("ncell,box,nranks", [((32, 32, 32), 16, 1), ((32, 32, 32), 16, 2), ((24, 16, 8), 8, 3), ((16, 16, 16), 16, 1), ((32, 16, 16), 16, 2)])
def test_plan_fills_every_interior_ghost(periodic, ncell, box, nranks):
    p = Prob(ncell, box, periodic)
    owner = distribute(p.boxes, nranks)
    levels = [HostLevel(p, owner, r) for r in range(nranks)]
    for L in levels:
        L.fill_local()
    # the "network": sender packs, receiver unpacks; sizes must agree without any negotiation
    for src in levels:
        for peer in src.peers():
            msg = src.pack(peer)
            assert msg.size == levels[peer].recv_size(src.rank)
            levels[peer].unpack(src.rank, msg)
    for L in levels:
        L.check_ghosts()
        L.close()


def test_single_periodic_box_copies_from_itself():
    p = Prob((16, 16, 16), 16, (1, 1, 1))
    L = HostLevel(p, [0], 0)
    tags = L.tags("local")
    assert len(tags) == 26 and not L.tags("remote")
    L.fill_local()
    L.check_ghosts()
    L.close()


def test_chop_domain_matches_amrex_maxsize_order():
    # BoxArray(domain).maxSize(128) on 256^3: 8 boxes, x fastest (AMReX_BoxList.cpp maxSize / BoxArray ordering)
    b = chop_domain(256, 128)
    assert len(b) == 8
    assert [tuple(x.lo) for x in b][:3] == [(0, 0, 0), (128, 0, 0), (0, 128, 0)]
    assert all(x.ncells() == 128 ** 3 for x in b)
    assert sum(x.ncells() for x in chop_domain((24, 16, 8), 8)) == 24 * 16 * 8


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_distribute_is_balanced_and_compact(nranks):
    boxes = chop_domain(512, 128)
    owner = distribute(boxes, nranks)
    counts = np.bincount(owner, minlength=nranks)
    assert (counts == 64 // nranks).all()
    if nranks == 8:  # SFC: each rank owns one 2x2x2 block of boxes
        for r in range(8):
            mine = [boxes[i] for i, o in enumerate(owner) if o == r]
            for d in range(3):
                assert max(b.hi[d] for b in mine) - min(b.lo[d] for b in mine) + 1 == 256


@pytest.mark.parametrize("ncell,box", [((32, 1, 1), 16), ((16, 2, 3), 8), ((16, 16, 1), 8)])
def test_periodic_direction_thinner_than_the_ghost_width(ncell, box):
    # one-cell-thick quasi-1-D / 2-D domains: amrex::Periodicity::shiftIntVect(nghost) enumerates ceil(nghost / length) images on
    # either side, so all four ghost layers of a 1-cell periodic direction are copies of that one cell (qk_plan_tags)
    p = Prob(ncell, box, (1, 1, 1))
    L = HostLevel(p, [0] * len(p.boxes), 0)
    L.fill_local()
    L.check_ghosts()
    L.close()
