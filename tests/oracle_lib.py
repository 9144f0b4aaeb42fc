"""Test-side loader of the CPU oracle (oracle/liboracle.so) and, where it was built, of the reference
itself (oracle/_ref/libquokka_ref.so).  TEST INFRASTRUCTURE: never imported by the product package.

Host arrays are numpy float64 of shape (ncomp, nz, ny, nx), C-contiguous == AMReX FAB order
(x fastest, component-major; extern/amrex/Src/Base/AMReX_Array4.H:59-68).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from quokka_b200.capi import (qk_array4, qk_box, qk_hydro_params, qk_iarray4, qk_level_desc, qk_rad_params, qk_rad_source_params)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libquokka_ref.so")

_A4P = C.POINTER(qk_array4)
_IA4P = C.POINTER(qk_iarray4)
_BXP = C.POINTER(qk_box)
_PRM = C.POINTER(qk_hydro_params)
_D3 = C.POINTER(C.c_double)
_RPRM = C.POINTER(qk_rad_params)
_RSPRM = C.POINTER(qk_rad_source_params)
_I64P = C.POINTER(C.c_int64)


def build_oracle() -> str:
    src = os.path.join(ROOT, "oracle", "quokka_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_oracle = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is not None:
        return _oracle
    lib = C.CDLL(build_oracle())
    sig = {
        "orc_conserved_to_primitive": (None, [_PRM, _A4P, _A4P, _BXP]),
        "orc_flattening_coefficients": (None, [_PRM, C.c_int, _A4P, _A4P, _BXP]),
        "orc_reconstruct_states": (None, [C.c_int, C.c_int, C.c_int, _A4P, _A4P, _A4P, _BXP, C.c_int]),
        "orc_flatten_shocks": (None, [C.c_int, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _BXP, C.c_int]),
        "orc_compute_fluxes": (None, [_PRM, C.c_int, C.c_int, _A4P, _A4P, _A4P, _A4P, _A4P, _BXP]),
        "orc_saxpy": (None, [_A4P, C.c_double, _A4P, _BXP, C.c_int]),
        "orc_rhs_from_fluxes": (None, [_A4P, _A4P, _A4P, _A4P, _D3, _BXP, C.c_int]),
        "orc_add_internal_energy_pdv": (None, [_PRM, _A4P, _A4P, _D3, _A4P, _A4P, _A4P, _IA4P, _BXP]),
        "orc_predict_step": (C.c_int64, [_PRM, _A4P, _A4P, _A4P, C.c_double, C.c_int, _IA4P, _BXP]),
        "orc_enforce_limits": (None, [_PRM, _A4P, _BXP]),
        "orc_sync_dual_energy": (C.c_int64, [_PRM, _A4P, _BXP]),
        "orc_replace_fluxes": (None, [C.c_int, _A4P, _A4P, _IA4P, _BXP, C.c_int]),
        "orc_max_signal_speed": (C.c_double, [_PRM, C.c_int, _A4P, _BXP]),
        "orc_eos_pressure": (C.c_double, [_PRM, C.c_double, C.c_double]),
        "orc_eos_sound_speed": (C.c_double, [_PRM, C.c_double, C.c_double]),
        "orc_eos_eint_from_pres": (C.c_double, [_PRM, C.c_double, C.c_double]),
        "orc_eos_tgas_from_eint": (C.c_double, [_PRM, C.c_double, C.c_double]),
        "orc_eos_eint_from_tgas": (C.c_double, [_PRM, C.c_double, C.c_double]),
        "orc_level_create": (C.c_void_p, [C.POINTER(qk_level_desc)]),
        "orc_level_destroy": (None, [C.c_void_p]),
        "orc_level_nboxes": (C.c_int, [C.c_void_p]),
        "orc_level_state": (qk_array4, [C.c_void_p, C.c_int, C.c_int]),
        "orc_level_box": (qk_box, [C.c_void_p, C.c_int]),
        "orc_fill_boundary": (None, [C.c_void_p, _A4P, C.c_int, C.c_int]),
        "orc_advance_hydro_level": (C.c_int, [C.c_void_p, _PRM, C.c_double, C.c_double, _I64P, _I64P]),
        "orc_compute_timestep": (C.c_double, [C.c_void_p, _PRM, C.c_double, C.c_double, C.c_double]),
        "orc_step_with_retries": (C.c_int, [C.c_void_p, _PRM, C.c_double, C.c_double]),
        "orc_rad_conserved_to_primitive": (None, [_RPRM, _A4P, _A4P, _BXP]),
        "orc_rad_compute_fluxes": (None, [_RPRM, C.c_int, _A4P, _A4P, _A4P, _A4P, _A4P, _BXP]),
        "orc_rad_predict_step": (None, [_RPRM, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_double, _D3, _BXP]),
        "orc_rad_add_fluxes_rk2": (None, [_RPRM, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_double, _D3, _BXP]),
        "orc_rad_advance_level": (None, [C.c_void_p, _RPRM, C.c_double]),
        "orc_rad_add_source_terms": (None, [_PRM, _RPRM, _RSPRM, _A4P, _A4P, _BXP, C.c_double, C.c_int, _I64P]),
        "orc_rad_num_substeps": (C.c_int, [C.c_void_p, _RPRM, C.c_double, C.c_double]),
        "orc_rad_subcycle_level": (C.c_int, [C.c_void_p, _PRM, _RPRM, _RSPRM, _A4P, C.c_double, C.c_double, _I64P]),
        "orc_compute_timestep_radhydro": (C.c_double, [C.c_void_p, _PRM, _RPRM, C.c_int, C.c_double, C.c_double, C.c_double]),
        "orc_shell_rad_energy_source": (None, [_A4P, _BXP, _D3, _D3, _D3]),
        "orc_interp_cons_lin_minmax": (None, [_A4P, C.c_int, _A4P, C.c_int, C.c_int, _BXP, _BXP, _BXP, C.POINTER(C.c_int), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32)]),
        "orc_pre_interp_state": (None, [_A4P, _BXP]),
        "orc_post_interp_state": (None, [_A4P, _BXP]),
        "orc_average_down": (None, [_A4P, C.c_int, _A4P, C.c_int, C.c_int, _BXP, C.POINTER(C.c_int)]),
        "orc_level_swap": (None, [C.c_void_p]),
        "orc_time_interp": (C.c_int, [_A4P, C.c_int, _A4P, _A4P, C.c_int, C.c_int, _BXP, C.c_double, C.c_double, C.c_double]),
        "orc_tag_pressure_gradient": (None, [_PRM, _A4P, C.c_char_p, _BXP, C.c_double, C.c_double]),
        "orc_tag_gradient_x": (None, [_A4P, C.c_int, C.c_char_p, _BXP, C.c_double, C.c_double, C.c_double]),
        "orc_fixup_state": (None, [_PRM, _A4P, _BXP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _oracle = lib
    return lib


_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    """The reference's own templates behind extern "C" (oracle/ref_build/ref_harness.cpp)."""
    global _ref
    if _ref is not None:
        return _ref
    lib = C.CDLL(REF_SO)
    lib.ref_nvar.argtypes = [C.c_int]
    lib.ref_cons_to_prim.argtypes = [C.c_int, _BXP, _A4P, _A4P, C.c_int]
    lib.ref_flattening_coefficients.argtypes = [C.c_int, C.c_int, _BXP, _A4P, _A4P, C.c_int, C.c_int]
    lib.ref_reconstruct.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, C.c_int, C.c_int, C.c_int]
    lib.ref_flatten_shocks.argtypes = [C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_int, C.c_int, C.c_int]
    lib.ref_compute_fluxes.argtypes = [C.c_int, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_int, C.c_double]
    lib.ref_update_op.argtypes = [C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _IA4P, _D3, C.c_double, C.c_double, C.c_double, _D3]
    lib.ref_rad_params.argtypes = [C.c_int, _RPRM]
    lib.ref_rad_cons_to_prim.argtypes = [C.c_int, _BXP, _A4P, _A4P, C.c_int]
    lib.ref_rad_compute_fluxes.argtypes = [C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_int]
    if hasattr(lib, "ref_rad_compute_fluxes_wsc"):
        lib.ref_rad_compute_fluxes_wsc.argtypes = [C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_int, C.POINTER(C.c_double)]
    lib.ref_rad_update.argtypes = [C.c_int, C.c_int, _BXP] + [_A4P] * 9 + [C.c_double, _D3]
    lib.ref_interp_cons_lin_minmax.argtypes = [_A4P, _A4P, C.c_int, _BXP, _BXP, _BXP, C.POINTER(C.c_int), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.ref_pre_post_interp_state.argtypes = [C.c_int, _BXP, _A4P]
    lib.ref_average_down.argtypes = [_A4P, _A4P, C.c_int, _BXP, C.POINTER(C.c_int)]
    if hasattr(lib, "ref_time_interp"):
        lib.ref_time_interp.argtypes = [_BXP, _A4P, _A4P, _A4P, C.c_int, C.c_double, C.c_double, C.c_double]
    lib.ref_rad_source_params.argtypes = [C.c_int, _PRM, _RPRM, _RSPRM]
    lib.ref_rad_add_source_terms.argtypes = [C.c_int, _BXP, _A4P, _A4P, C.c_double, C.c_int, _I64P]
    _ref = lib
    return lib


# ---- numpy <-> qk_array4 ---------------------------------------------------------------------------
class HostFab:
    """A numpy-backed FAB over an index box (inclusive lo/hi)."""

    def __init__(self, box: qk_box, ncomp: int, dtype=np.float64, fill=0.0):
        self.box = box
        self.ncomp = ncomp
        nz, ny, nx = box.shape()
        self.a = np.full((ncomp, nz, ny, nx), fill, dtype=dtype)

    def desc(self):
        d = qk_array4() if self.a.dtype == np.float64 else qk_iarray4()
        nz, ny, nx = self.box.shape()
        d.p = self.a.ctypes.data
        d.jstride = nx
        d.kstride = nx * ny
        d.nstride = nx * ny * nz
        d.begin[:] = list(self.box.lo)
        d.end[:] = [self.box.hi[i] + 1 for i in range(3)]
        d.ncomp = self.ncomp
        return d

    def view(self, box: qk_box):
        """numpy view (ncomp, nz, ny, nx) of the sub-box."""
        o = [box.lo[d] - self.box.lo[d] for d in range(3)]
        s = [box.hi[d] - box.lo[d] + 1 for d in range(3)]
        return self.a[:, o[2]:o[2] + s[2], o[1]:o[1] + s[1], o[0]:o[0] + s[0]]


def face_box(valid: qk_box, d: int, ng: int = 0) -> qk_box:
    """amrex::convert(box, e_d) grown by ng (cell-index convention: hi[d]+1 is the last face)."""
    lo = [valid.lo[i] - ng for i in range(3)]
    hi = [valid.hi[i] + ng for i in range(3)]
    hi[d] += 1
    return qk_box.make(lo, hi)


def grow(valid: qk_box, ng: int) -> qk_box:
    return valid.grown(ng)


def random_cons(box: qk_box, nscalars=0, seed=12345, kind="smooth"):
    """Seeded conserved states (rho, m, E, Eaux [, scalars]) on `box`: SURVEY.md 8(d) microbenchmark
    inputs (rho in U(0.1,10), v in U(-1,1)^3, P in U(0.1,10)); kind='shocked' adds jumps."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = box.shape()
    shp = (nz, ny, nx)
    rho = rng.uniform(0.1, 10.0, shp)
    v = rng.uniform(-1.0, 1.0, (3,) + shp)
    P = rng.uniform(0.1, 10.0, shp)
    if kind == "smooth":
        # low-pass so that PPM takes the smooth branches too
        for arr in (rho, P, v[0], v[1], v[2]):
            for ax in range(3):
                arr[...] = (np.roll(arr, 1, ax) + 2 * arr + np.roll(arr, -1, ax)) / 4
    elif kind == "shocked":
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        m = (x + y + z) % 7 < 3
        P[m] *= 50.0
        rho[m] *= 4.0
    return rho, v, P, rng


def cons_from_prim(rho, v, P, gamma, rng, nscalars=0, eaux_jitter=True):
    ke = 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    eint = P / (gamma - 1.0)
    comps = [rho, rho * v[0], rho * v[1], rho * v[2], eint + ke, eint * (1 + (1e-3 * rng.standard_normal(rho.shape) if eaux_jitter else 0))]
    for _ in range(nscalars):
        comps.append(rho * rng.uniform(0.0, 1.0, rho.shape))
    return np.stack(comps)


def random_rad_cons(box: qk_box, prm, seed=777, kind="smooth", ncomp=None):
    """Seeded radiation states in a full state FAB (hydro components = 1): E_r log-uniform over 4 decades, reduced flux
    |f| in [0, 0.95] with random directions; kind='beam' adds cells with |f| -> 1 and inadmissible cells (E_r <= 0,
    |f| > 1) so that the first-order fallback, the wave-speed floor and amendRadState are exercised."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = box.shape()
    shp = (nz, ny, nx)
    nc = ncomp if ncomp is not None else prm.nstart + 4 * prm.ngroups
    a = np.ones((nc,) + shp)
    for g in range(prm.ngroups):
        E = 10.0 ** rng.uniform(-2.0, 2.0, shp)
        if kind == "smooth":
            for ax in range(3):
                E = (np.roll(E, 1, ax) + 2 * E + np.roll(E, -1, ax)) / 4
        fmag = rng.uniform(0.0, 0.95, shp)
        mu = rng.uniform(-1.0, 1.0, shp)
        phi = rng.uniform(0.0, 2 * np.pi, shp)
        n = np.stack([np.sqrt(1 - mu ** 2) * np.cos(phi), np.sqrt(1 - mu ** 2) * np.sin(phi), mu])
        if kind == "beam":
            z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
            m = (x + 2 * y + 3 * z) % 11 == 0
            fmag[m] = 1.0 - 1e-12
            fmag[(x + y + z) % 13 == 0] = 0.0
            fmag[(x * y + z) % 17 == 0] = 1.3
            E[(x + y * z) % 19 == 0] *= -1.0
            n[:, (x + 3 * y + z) % 7 == 0] = np.array([1.0, 0.0, 0.0])[:, None]
        s = prm.nstart + 4 * g
        a[s] = E
        for m_ in range(3):
            a[s + 1 + m_] = fmag * n[m_] * prm.c_light * np.abs(E)
    return a


def random_radhydro_cons(box: qk_box, hp, rp, sp, seed=4242, T0=1.0, rho0=1.0, vmax=1.0, fmax=0.9, spread=1.0):
    """Seeded coupled gas + radiation states (one photon group) for the matter-radiation source terms: rho log-uniform over
    2*spread decades around rho0, gas and radiation temperatures independently log-uniform over 2*spread decades around T0
    (so cells heat and cool), |v| <= vmax, reduced flux |f| in [0, fmax] with random directions.  Eint from the gamma law
    (c_v = k_B / (mu (gamma - 1))); for gamma == 1 the energies are arbitrary positive numbers."""
    rng = np.random.default_rng(seed)
    shp = box.shape()
    rho = rho0 * 10.0 ** rng.uniform(-spread, spread, shp)
    Tg = T0 * 10.0 ** rng.uniform(-spread, spread, shp)
    Tr = T0 * 10.0 ** rng.uniform(-spread, spread, shp)
    v = rng.uniform(-vmax, vmax, (3,) + shp)
    if hp.gamma != 1.0:
        eint = rho * hp.boltzmann_constant * Tg / (hp.mean_molecular_weight * (hp.gamma - 1.0))
    else:
        eint = rho * Tg
    ke = 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    E = sp.radiation_constant * Tr ** 4
    fmag = rng.uniform(0.0, fmax, shp)
    mu = rng.uniform(-1.0, 1.0, shp)
    phi = rng.uniform(0.0, 2 * np.pi, shp)
    n = np.stack([np.sqrt(1 - mu ** 2) * np.cos(phi), np.sqrt(1 - mu ** 2) * np.sin(phi), mu])
    a = np.zeros((rp.nstart + 4,) + shp)
    a[0] = rho
    for m in range(3):
        a[1 + m] = rho * v[m]
        a[rp.nstart + 1 + m] = fmag * n[m] * rp.c_light * E
    a[4] = eint + ke
    a[5] = eint
    a[rp.nstart] = E
    return a
