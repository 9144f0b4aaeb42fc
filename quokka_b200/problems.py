"""Problem set-ups of the reference's hot-path configs, as host-side data (no numerics of the path).

Sedov blast: src/problems/HydroBlast3D/test_hydro3d_blast.cpp (octant symmetry, reflecting walls,
gamma = 1.4, reconstruct_eint = false, cfl 0.3) with tests/blast_unigrid_*.in geometry.
"""
from __future__ import annotations

import numpy as np

from .capi import (QK_BC_REFLECT_EVEN, QK_BC_REFLECT_ODD, hydro_params, qk_box)


def chop_domain(ncell, max_grid_size):
    """amrex::BoxArray(domain).maxSize(max_grid_size): boxes in x-fastest order
    (extern/amrex/Src/Base/AMReX_BoxArray.cpp maxSize -> BoxList::maxSize)."""
    n = [int(c) for c in (ncell if hasattr(ncell, "__len__") else (ncell,) * 3)]
    m = [int(c) for c in (max_grid_size if hasattr(max_grid_size, "__len__") else (max_grid_size,) * 3)]
    cuts = []
    for d in range(3):
        nb = -(-n[d] // m[d])
        # AMReX chops into nb nearly-equal chunks; for the power-of-two configs they are equal
        base, rem = divmod(n[d], nb)
        edges = [0]
        for i in range(nb):
            edges.append(edges[-1] + base + (1 if i < rem else 0))
        cuts.append(edges)
    boxes = []
    for kz in range(len(cuts[2]) - 1):
        for jy in range(len(cuts[1]) - 1):
            for ix in range(len(cuts[0]) - 1):
                boxes.append(qk_box.make((cuts[0][ix], cuts[1][jy], cuts[2][kz]), (cuts[0][ix + 1] - 1, cuts[1][jy + 1] - 1, cuts[2][kz + 1] - 1)))
    return boxes


def distribute(boxes, nranks):
    """Box -> rank map.  The reference uses AMReX's SFC strategy (AMReX_DistributionMapping.cpp:42); for the
    uniform power-of-two configs (equal-weight boxes in Morton order) SFC assigns contiguous Z-order
    chunks, which is what this reproduces."""
    def morton(b):
        x, y, z = (b.lo[0] // max(1, b.hi[0] - b.lo[0] + 1), b.lo[1] // max(1, b.hi[1] - b.lo[1] + 1), b.lo[2] // max(1, b.hi[2] - b.lo[2] + 1))
        code = 0
        for bit in range(10):
            code |= ((x >> bit) & 1) << (3 * bit) | ((y >> bit) & 1) << (3 * bit + 1) | ((z >> bit) & 1) << (3 * bit + 2)
        return code
    order = sorted(range(len(boxes)), key=lambda i: morton(boxes[i]))
    owner = [0] * len(boxes)
    per = len(boxes) / nranks
    for pos, i in enumerate(order):
        owner[i] = min(nranks - 1, int(pos / per))
    return owner


class SedovProblem:
    """HydroBlast3D (test_hydro3d_blast.cpp:22-116,222-262)."""

    gamma = 1.4
    cfl = 0.3
    stop_time = 1.0
    ncomp = 6
    nghost = 4
    E_blast = 0.851072 / 8.0  # octant (test_hydro3d_blast.cpp:56-60)
    rho0 = 1.0

    def __init__(self, ncell, max_grid_size, prob_hi=1.2):
        self.ncell = [int(c) for c in (ncell if hasattr(ncell, "__len__") else (ncell,) * 3)]
        self.domain = qk_box.make((0, 0, 0), tuple(c - 1 for c in self.ncell))
        hi = prob_hi if hasattr(prob_hi, "__len__") else (prob_hi,) * 3
        self.dx = [(float(hi[d]) - 0.0) / self.ncell[d] for d in range(3)]
        self.boxes = chop_domain(self.ncell, max_grid_size)
        self.periodic = (0, 0, 0)
        # reflect_odd on the normal momentum, reflect_even otherwise (test_hydro3d_blast.cpp:224-254)
        lo = []
        for n in range(self.ncomp):
            for d in range(3):
                lo.append(QK_BC_REFLECT_ODD if n == 1 + d else QK_BC_REFLECT_EVEN)
        self.bc_lo = lo
        self.bc_hi = list(lo)

    def params(self, **kw):
        return hydro_params(gamma=self.gamma, reconstruct_eint=0, **kw)

    def initial_state(self, box: qk_box, ng=None) -> np.ndarray:
        """(ncomp, nz, ny, nx) on the grown box; ghost cells zero (filled later by the BC fill)."""
        ng = self.nghost if ng is None else ng
        g = box.grown(ng)
        nz, ny, nx = g.shape()
        a = np.zeros((self.ncomp, nz, ny, nx))
        cell_vol = self.dx[0] * self.dx[1] * self.dx[2]
        v = a[:, ng:nz - ng, ng:ny - ng, ng:nx - ng]
        v[0] = self.rho0
        v[4] = 1.0e-10 * (self.E_blast / cell_vol)
        if box.lo[0] <= 0 <= box.hi[0] and box.lo[1] <= 0 <= box.hi[1] and box.lo[2] <= 0 <= box.hi[2]:
            v[4, 0 - box.lo[2], 0 - box.lo[1], 0 - box.lo[0]] = self.E_blast / cell_vol
        return a


class IsothermalWaveProblem:
    """An isothermal-EOS problem (EOS_Traits::gamma = 1 with cs_isothermal, src/hydro/EOS.hpp:32-37; the EOS of the reference's
    BinaryOrbitCIC / StarCluster / RadForce problems): colliding density enhancements in a periodic box, strong enough to shock.
    Pressure is rho cs^2, there are no energy fluxes (hydro_system.hpp:1083-1087); the energy components only see the dual-energy sync."""

    gamma = 1.0
    cs_isothermal = 1.3
    cfl = 0.3
    stop_time = 1.0
    ncomp = 6
    nghost = 4

    def __init__(self, ncell, max_grid_size):
        self.ncell = [int(c) for c in (ncell if hasattr(ncell, "__len__") else (ncell,) * 3)]
        self.domain = qk_box.make((0, 0, 0), tuple(c - 1 for c in self.ncell))
        self.dx = [1.0 / c for c in self.ncell]
        self.boxes = chop_domain(self.ncell, max_grid_size)
        self.periodic = (1, 1, 1)
        self.bc_lo = [0] * (3 * self.ncomp)  # QK_BC_INT_DIR: interior / periodic
        self.bc_hi = list(self.bc_lo)

    def params(self, **kw):
        return hydro_params(gamma=self.gamma, reconstruct_eint=0, cs_isothermal=self.cs_isothermal, **kw)

    def initial_state(self, box: qk_box, ng=None) -> np.ndarray:
        ng = self.nghost if ng is None else ng
        g = box.grown(ng)
        nz, ny, nx = g.shape()
        a = np.zeros((self.ncomp, nz, ny, nx))
        z, y, x = np.meshgrid(*[(np.arange(box.lo[d], box.hi[d] + 1) + 0.5) * self.dx[d] for d in (2, 1, 0)], indexing="ij")
        rho = 1.0 + 0.8 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 0.5 * np.exp(-((x - 0.5) ** 2 + (y - 0.4) ** 2 + (z - 0.6) ** 2) / 0.01)
        vx = 2.0 * np.sin(2 * np.pi * (y + z))
        vy = -1.5 * np.cos(2 * np.pi * x)
        vz = 1.0 * np.sin(4 * np.pi * x) * np.sin(2 * np.pi * y)
        v = a[:, ng:nz - ng, ng:ny - ng, ng:nx - ng]
        v[0] = rho
        v[1], v[2], v[3] = rho * vx, rho * vy, rho * vz
        eint = rho * self.cs_isothermal ** 2  # any positive value: the isothermal EOS never reads it
        v[4] = eint + 0.5 * rho * (vx * vx + vy * vy + vz * vz)
        v[5] = eint
        return a


class SodProblem:
    """HydroShocktube (src/problems/HydroShocktube/test_hydro_shocktube.cpp:28-91, tests/shocktube.in), config C1 of
    BASELINE.json, on a uniform level.  The reference builds it with AMREX_SPACEDIM = 1; here the tube is 4 cells thick and
    periodic in y and z (all transverse differences are exactly zero, so the bits are the 1-D build's).  The Dirichlet walls
    (setCustomBoundaryConditions :93-141) hold the initial left/right states; until a wave reaches a wall (t = 0.4 ends before)
    first-order extrapolation of the undisturbed boundary cell gives the same ghost values."""

    gamma = 1.4
    cfl = 0.6
    stop_time = 0.4
    ncomp = 6
    nghost = 4
    rho_L, P_L, rho_R, P_R = 10.0, 100.0, 1.0, 1.0

    def __init__(self, ncell_x, max_grid_size=128, thickness=4):
        from .capi import QK_BC_FOEXTRAP, QK_BC_INT_DIR

        self.ncell = [int(ncell_x), thickness, thickness]
        self.domain = qk_box.make((0, 0, 0), tuple(c - 1 for c in self.ncell))
        self.dx = [5.0 / self.ncell[0], 1.0 / thickness, 1.0 / thickness]
        self.boxes = chop_domain(self.ncell, (max_grid_size, thickness, thickness))
        self.periodic = (0, 1, 1)
        lo = []
        for n in range(self.ncomp):
            lo += [QK_BC_FOEXTRAP, QK_BC_INT_DIR, QK_BC_INT_DIR]
        self.bc_lo = lo
        self.bc_hi = list(lo)

    def params(self, **kw):
        return hydro_params(gamma=self.gamma, reconstruct_eint=1, **kw)

    def initial_state(self, box: qk_box, ng=None) -> np.ndarray:
        ng = self.nghost if ng is None else ng
        g = box.grown(ng)
        nz, ny, nx = g.shape()
        a = np.zeros((self.ncomp, nz, ny, nx))
        x = 0.0 + (np.arange(box.lo[0], box.hi[0] + 1) + 0.5) * self.dx[0]
        left = x < 2.0
        rho = np.where(left, self.rho_L, self.rho_R)
        P = np.where(left, self.P_L, self.P_R)
        v = a[:, ng:nz - ng, ng:ny - ng, ng:nx - ng] if ng else a
        v[0] = rho
        v[4] = P / (self.gamma - 1.0) + 0.5 * rho * (0.0 * 0.0)
        v[5] = P / (self.gamma - 1.0)
        return a


class ShellProblem:
    """RadhydroShell (src/problems/RadhydroShell/test_radhydro_shell.cpp:34-93,127-135,396-431, tests/radhydro_shell*.in), config C4
    of BASELINE.json on a uniform periodic level: gamma = 5/3, mu = 2.2 m_u, reduced speed of light 860 * 2e5 cm/s, constant opacity
    20 cm^2/g, beta_order 1, PLM (minmod) hydro + PLM (MC) radiation, RK2, cfl 0.3 (hydro and radiation), density floor 1e-8 rho_0,
    at most 10 radiation substeps per hydro step.  The initial condition (an interpolation table and libm pow calls) is taken
    from the reference's own step-0 plotfile (tests/golden/shell*.npz), not re-derived."""

    gamma = 5.0 / 3.0
    cfl = 0.3
    rad_cfl = 0.3  # radiationCflNumber_ (QuokkaSimulation.hpp:125)
    max_substeps = 10  # maxSubsteps_ (:126)
    ncomp = 10
    nghost = 4
    a_rad = 7.5646e-15
    c_light = 2.99792458e10
    c_hat = 860.0 * 2.0e5
    kappa0 = 20.0
    prob_hi = 3.086e19
    r_0 = 5.0 * 3.086e18
    rho_0 = ((1 - 0.5) * (1.0e6 * 2.0e33)) / ((4.0 / 3.0) * np.pi * r_0 * r_0 * r_0)
    stop_time = 0.125 * (r_0 / 2.0e5)

    def __init__(self, ncell, max_grid_size, initial=None):
        from .capi import QK_BC_INT_DIR

        self.ncell = [int(ncell)] * 3 if np.isscalar(ncell) else [int(c) for c in ncell]
        self.domain = qk_box.make((0, 0, 0), tuple(c - 1 for c in self.ncell))
        self.dx = [self.prob_hi / c for c in self.ncell]
        self.boxes = chop_domain(self.ncell, max_grid_size)
        self.periodic = (1, 1, 1)
        self.bc_lo = [QK_BC_INT_DIR] * (3 * self.ncomp)
        self.bc_hi = list(self.bc_lo)
        self.initial = initial  # (10, nz, ny, nx) valid cells of the whole domain

    def params(self, **kw):
        from .capi import M_U, K_B

        return hydro_params(gamma=self.gamma, reconstruct_eint=0, recon_order=2, density_floor=1.0e-8 * self.rho_0, mean_molecular_weight=2.2 * M_U,
                            boltzmann_constant=K_B, **kw)

    def rad_params(self):
        from .capi import rad_params

        return rad_params(c_light=self.c_light, c_hat=self.c_hat, Erad_floor=0.0, ngroups=1, nstart=6, recon_order=2, integrator_order=2)

    def rad_source_params(self):
        from .capi import rad_source_params

        return rad_source_params(radiation_constant=self.a_rad, kappa_P=self.kappa0, beta_order=1)

    def initial_state(self, box: qk_box, ng=None) -> np.ndarray:
        ng = self.nghost if ng is None else ng
        g = box.grown(ng)
        nz, ny, nx = g.shape()
        a = np.zeros((self.ncomp, nz, ny, nx))
        a[:, ng:nz - ng, ng:ny - ng, ng:nx - ng] = self.initial[:, box.lo[2]:box.hi[2] + 1, box.lo[1]:box.hi[1] + 1, box.lo[0]:box.hi[0] + 1]
        return a
