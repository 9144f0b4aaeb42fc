"""ctypes image of include/quokka_b200.h and loader of libquokka_b200.so (the product library).

The library is the drop-in boundary: plain pointers and sizes, no torch types.  This module only
mirrors the structs and declares argtypes; it contains no numerics and NO fallback: if the CUDA
library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

QK_MAX_SCALARS = 8
QK_OK = 0
QK_ERR_NO_DEVICE = -1
QK_ERR_BAD_ARG, QK_ERR_UNSUPPORTED, QK_ERR_NOMEM = -2, -3, -4
QK_X1, QK_X2, QK_X3 = 0, 1, 2
QK_HLLC, QK_LLF = 0, 1
QK_MINMOD, QK_MC = 0, 1
QK_BC_INT_DIR, QK_BC_REFLECT_ODD, QK_BC_REFLECT_EVEN, QK_BC_FOEXTRAP, QK_BC_EXT_DIR = 0, -1, 1, 2, 3
QK_ARITH_EXACT, QK_ARITH_FAST = 0, 1


class qk_array4(C.Structure):
    """== amrex::Array4<double> (AMReX_Array4.H:59-68)."""

    _fields_ = [
        ("p", C.c_void_p),
        ("jstride", C.c_int64),
        ("kstride", C.c_int64),
        ("nstride", C.c_int64),
        ("begin", C.c_int32 * 3),
        ("end", C.c_int32 * 3),
        ("ncomp", C.c_int32),
    ]


class qk_iarray4(C.Structure):
    _fields_ = qk_array4._fields_


class qk_box(C.Structure):
    _fields_ = [("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3)]

    @classmethod
    def make(cls, lo, hi):
        b = cls()
        b.lo[:] = [int(x) for x in lo]
        b.hi[:] = [int(x) for x in hi]
        return b

    def grown(self, n, dir_hi=None):
        lo = [self.lo[d] - n for d in range(3)]
        hi = [self.hi[d] + n for d in range(3)]
        if dir_hi is not None:
            hi[dir_hi] += 1
        return qk_box.make(lo, hi)

    def shape(self):
        return tuple(self.hi[d] - self.lo[d] + 1 for d in (2, 1, 0))  # (nz, ny, nx)

    def ncells(self):
        s = self.shape()
        return s[0] * s[1] * s[2]


class qk_hydro_params(C.Structure):
    _fields_ = [
        ("gamma", C.c_double),
        ("mean_molecular_weight", C.c_double),
        ("boltzmann_constant", C.c_double),
        ("small_temp", C.c_double),
        ("small_dens", C.c_double),
        ("density_floor", C.c_double),
        ("temp_floor", C.c_double),
        ("K_visc", C.c_double),
        ("small_x", C.c_double),
        ("reconstruct_eint", C.c_int32),
        ("nscalars", C.c_int32),
        ("nmscalars", C.c_int32),
        ("reconstruction_order", C.c_int32),
        ("use_dual_energy", C.c_int32),
        ("integrator_order", C.c_int32),
        ("abort_on_fofc_failure", C.c_int32),
        ("arith", C.c_int32),
        ("cs_isothermal", C.c_double),
    ]


class qk_rad_params(C.Structure):
    """run-time image of RadSystem_Traits<problem_t> (src/radiation/radiation_system.hpp:73-82)"""

    _fields_ = [
        ("c_light", C.c_double),
        ("c_hat", C.c_double),
        ("Erad_floor", C.c_double),
        ("ngroups", C.c_int32),
        ("nstart", C.c_int32),
        ("reconstruction_order", C.c_int32),
        ("integrator_order", C.c_int32),
        ("arith", C.c_int32),
        ("use_wavespeed_correction", C.c_int32),
        ("kappa_F", C.c_double),
        ("cell_dx", C.c_double * 3),
    ]


def rad_params(c_light=1.0, c_hat=1.0, Erad_floor=0.0, ngroups=1, nstart=6, recon_order=3, integrator_order=2, arith=QK_ARITH_EXACT,
               wavespeed_correction=0, kappa_F=0.0, cell_dx=(1.0, 1.0, 1.0)) -> qk_rad_params:
    p = qk_rad_params()
    p.use_wavespeed_correction, p.kappa_F = int(wavespeed_correction), kappa_F
    p.cell_dx[:] = list(cell_dx)
    p.c_light, p.c_hat, p.Erad_floor = c_light, c_hat, Erad_floor
    p.ngroups, p.nstart, p.reconstruction_order, p.integrator_order = ngroups, nstart, recon_order, integrator_order
    p.arith = arith
    return p


class qk_rad_source_params(C.Structure):
    """run-time image of what RadSystem::AddSourceTermsSingleGroup takes from RadSystem_Traits<problem_t> and the problem's
    opacity specialisations (src/radiation/source_terms_single_group.hpp:9-565, radiation_system.hpp:73-82,1141-1153)"""

    _fields_ = [
        ("radiation_constant", C.c_double),
        ("kappa_P", C.c_double),
        ("kappa_E", C.c_double),
        ("kappa_F", C.c_double),
        ("beta_order", C.c_int32),
        ("opacity_model", C.c_int32),
    ]


QK_OPACITY_CONSTANT = 0
QK_RAD_SOURCE_NCOUNTERS = 7


def rad_source_params(radiation_constant=7.5657e-15, kappa_P=1.0, kappa_E=None, kappa_F=None, beta_order=1) -> qk_rad_source_params:
    """kappa_E / kappa_F default to kappa_P as the reference's ComputeEnergyMeanOpacity / ComputeFluxMeanOpacity do (:1146-1154)."""
    p = qk_rad_source_params()
    p.radiation_constant = radiation_constant
    p.kappa_P = kappa_P
    p.kappa_E = kappa_P if kappa_E is None else kappa_E
    p.kappa_F = kappa_P if kappa_F is None else kappa_F
    p.beta_order, p.opacity_model = beta_order, QK_OPACITY_CONSTANT
    return p


K_B = 1.3806488e-16  # Microphysics constants/fundamental_constants.H:22
M_U = 1.6605390666e-24  # :55


def hydro_params(gamma=1.4, reconstruct_eint=0, nscalars=0, nmscalars=0, recon_order=3, arith=QK_ARITH_EXACT,
                 density_floor=0.0, temp_floor=0.0, mean_molecular_weight=M_U, boltzmann_constant=K_B,
                 cs_isothermal=float("nan")) -> qk_hydro_params:
    """Defaults = the reference's defaults (simulation.hpp:172-173, QuokkaSimulation.hpp:107-131,165-166)."""
    p = qk_hydro_params()
    p.gamma = gamma
    p.mean_molecular_weight = mean_molecular_weight
    p.boltzmann_constant = boltzmann_constant
    p.small_temp = 1e-10
    p.small_dens = 1e-100
    p.density_floor = density_floor
    p.temp_floor = temp_floor
    p.K_visc = 0.0
    p.small_x = 1e-30  # network_rp::small_x default (extern/Microphysics/networks/_parameters)
    p.reconstruct_eint = reconstruct_eint
    p.nscalars = nscalars
    p.nmscalars = nmscalars
    p.reconstruction_order = recon_order
    p.use_dual_energy = 1
    p.integrator_order = 2
    p.abort_on_fofc_failure = 1
    p.arith = arith
    p.cs_isothermal = cs_isothermal  # EOS_Traits default NAN (src/hydro/EOS.hpp:34); read only when gamma == 1
    return p


class qk_level_desc(C.Structure):
    _fields_ = [
        ("domain", qk_box),
        ("periodic", C.c_int32 * 3),
        ("dx", C.c_double * 3),
        ("nghost", C.c_int32),
        ("ncomp", C.c_int32),
        ("nboxes_global", C.c_int32),
        ("boxes_global", C.POINTER(qk_box)),
        ("owner", C.POINTER(C.c_int32)),
        ("my_rank", C.c_int32),
        ("bc_lo", C.POINTER(C.c_int32)),
        ("bc_hi", C.POINTER(C.c_int32)),
    ]


class qk_copy_tag(C.Structure):
    _fields_ = [
        ("src_box", C.c_int32),
        ("dst_box", C.c_int32),
        ("src_rank", C.c_int32),
        ("dst_rank", C.c_int32),
        ("src_region", qk_box),
        ("shift", C.c_int32 * 3),
        ("offset", C.c_int64),
        ("ncells", C.c_int64),
    ]


def make_level_desc(domain: qk_box, periodic, dx, nghost, ncomp, boxes, owner, my_rank, bc_lo, bc_hi):
    """Returns (desc, keepalive) -- keepalive holds the ctypes arrays the desc points into."""
    d = qk_level_desc()
    d.domain = domain
    d.periodic[:] = [int(x) for x in periodic]
    d.dx[:] = [float(x) for x in dx]
    d.nghost = nghost
    d.ncomp = ncomp
    nb = len(boxes)
    d.nboxes_global = nb
    barr = (qk_box * nb)(*boxes)
    oarr = (C.c_int32 * nb)(*[int(o) for o in owner])
    lo = (C.c_int32 * (3 * ncomp))(*[int(x) for x in bc_lo])
    hi = (C.c_int32 * (3 * ncomp))(*[int(x) for x in bc_hi])
    d.boxes_global = barr
    d.owner = oarr
    d.my_rank = my_rank
    d.bc_lo = lo
    d.bc_hi = hi
    return d, (barr, oarr, lo, hi)


class qk_carray4(C.Structure):
    """amrex::Array4<char> (TagBox)"""
    _fields_ = [("p", C.c_void_p), ("jstride", C.c_int64), ("kstride", C.c_int64), ("nstride", C.c_int64), ("begin", C.c_int32 * 3), ("end", C.c_int32 * 3),
                ("ncomp", C.c_int32)]


_A4P = C.POINTER(qk_array4)
_CA4P = C.POINTER(qk_carray4)
_IA4P = C.POINTER(qk_iarray4)
_BXP = C.POINTER(qk_box)
_PRM = C.POINTER(qk_hydro_params)
_D3 = C.POINTER(C.c_double)
_RPRM = C.POINTER(qk_rad_params)
_I64P = C.POINTER(C.c_int64)
_VP = C.c_void_p
_RSPRM = C.POINTER(qk_rad_source_params)

# name -> (restype, argtypes): EVERY symbol include/quokka_b200.h declares
SYMBOLS = {
    "qk_abi_version": (C.c_int, []),
    "qk_device_count": (C.c_int, []),
    "qk_error_string": (C.c_char_p, [C.c_int]),
    "qk_launch_count": (C.c_int64, []),
    "qk_prof_enable": (C.c_int, [C.c_int]),
    "qk_prof_report": (C.c_int, [C.c_char_p, C.c_int]),
    "qk_selftest_division": (C.c_int, [C.c_uint64, C.c_int, C.c_int64, _I64P, _I64P]),
    "qk_hydro_conserved_to_primitive": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _A4P, C.c_int, _VP]),
    "qk_hydro_flattening_coefficients": (C.c_int, [_PRM, C.c_int, C.c_int, _BXP, _A4P, _A4P, C.c_int, _VP]),
    "qk_reconstruct_states": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, C.c_int, C.c_int, _VP]),
    "qk_hydro_flatten_shocks": (C.c_int, [C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_int, C.c_int, _VP]),
    "qk_hydro_compute_fluxes": (C.c_int, [_PRM, C.c_int, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _VP]),
    "qk_hydro_flux_function": (C.c_int, [_PRM, C.c_int, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _VP]),
    "qk_saxpy": (C.c_int, [C.c_int, _BXP, _A4P, C.c_double, _A4P, C.c_int, _VP]),
    "qk_hydro_rhs_from_fluxes": (C.c_int, [C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _D3, C.c_int, _VP]),
    "qk_hydro_add_internal_energy_pdv": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _A4P, _D3, _A4P, _A4P, _A4P, _IA4P, _VP]),
    "qk_hydro_predict_step": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _A4P, _A4P, C.c_double, C.c_int, _IA4P, _I64P, _VP]),
    "qk_hydro_enforce_limits": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _VP]),
    "qk_hydro_sync_dual_energy": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _I64P, _VP]),
    "qk_hydro_replace_fluxes": (C.c_int, [C.c_int, C.c_int, _BXP, _A4P, _A4P, _IA4P, C.c_int, _VP]),
    "qk_hydro_max_signal_speed": (C.c_int, [_PRM, C.c_int, C.c_int, _BXP, _A4P, _D3, _VP]),
    "qk_rad_conserved_to_primitive": (C.c_int, [_RPRM, C.c_int, _BXP, _A4P, _A4P, C.c_int, _VP]),
    "qk_rad_compute_fluxes": (C.c_int, [_RPRM, C.c_int, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _VP]),
    "qk_rad_predict_step": (C.c_int, [_RPRM, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_double, _D3, _VP]),
    "qk_rad_add_fluxes_rk2": (C.c_int, [_RPRM, C.c_int, _BXP, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, _A4P, C.c_double, _D3, _VP]),
    "qk_rad_add_source_terms": (C.c_int, [_PRM, _RPRM, _RSPRM, C.c_int, C.c_int, _BXP, _A4P, _A4P, C.c_double, _I64P, _VP]),
    "qk_amr_interp_cons_lin_minmax": (C.c_int, [C.c_int, _A4P, C.c_int, _A4P, C.c_int, C.c_int, _BXP, _BXP, _BXP, C.POINTER(C.c_int), C.POINTER(C.c_int32),
                                                C.POINTER(C.c_int32), _VP]),
    "qk_amr_pre_interp_state": (C.c_int, [C.c_int, _BXP, _A4P, _VP]),
    "qk_amr_post_interp_state": (C.c_int, [C.c_int, _BXP, _A4P, _VP]),
    "qk_amr_average_down": (C.c_int, [C.c_int, _A4P, C.c_int, _A4P, C.c_int, C.c_int, _BXP, C.POINTER(C.c_int), _VP]),
    "qk_amr_time_interp": (C.c_int, [C.c_int, _A4P, C.c_int, _A4P, _A4P, C.c_int, C.c_int, _BXP, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int), _VP]),
    "qk_tag_pressure_gradient": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _CA4P, C.c_double, C.c_double, _I64P, _VP]),
    "qk_tag_gradient_x": (C.c_int, [C.c_int, _BXP, _A4P, C.c_int, _CA4P, C.c_double, C.c_double, C.c_double, _I64P, _VP]),
    "qk_hydro_fixup_state": (C.c_int, [_PRM, C.c_int, _BXP, _A4P, _VP]),
    "qk_rad_subcycle": (C.c_int, [_VP, _PRM, _RPRM, _RSPRM, _A4P, _A4P, _A4P, _A4P, C.c_double, C.c_double, _I64P, C.POINTER(C.c_int), _VP]),
    "qk_rad_advance_stage": (C.c_int, [_VP, _RPRM, C.c_int, _A4P, _A4P, _A4P, C.c_double, _VP]),
    "qk_level_create": (C.c_int, [C.POINTER(qk_level_desc), C.POINTER(_VP)]),
    "qk_level_destroy": (None, [_VP]),
    "qk_level_nlocal": (C.c_int, [_VP]),
    "qk_level_local_ids": (C.c_int, [_VP, C.POINTER(C.c_int32)]),
    "qk_level_remote_tags": (C.c_int, [_VP, C.POINTER(qk_copy_tag), C.c_int]),
    "qk_level_local_tags": (C.c_int, [_VP, C.POINTER(qk_copy_tag), C.c_int]),
    "qk_fill_boundary_local": (C.c_int, [_VP, _A4P, C.c_int, C.c_int, _VP]),
    "qk_pack_ghosts": (C.c_int, [_VP, C.c_int, _A4P, C.c_int, C.c_int, _VP, _I64P, _VP]),
    "qk_unpack_ghosts": (C.c_int, [_VP, C.c_int, _A4P, C.c_int, C.c_int, _VP, _VP]),
    "qk_fill_physical_bc": (C.c_int, [_VP, _A4P, C.c_int, C.c_int, _VP]),
    "qk_hydro_advance_stage": (C.c_int, [_VP, _PRM, C.c_int, _A4P, _A4P, _A4P, C.c_double, _I64P, _VP]),
    "qk_hydro_advance_stage_faithful": (C.c_int, [_VP, _PRM, C.c_int, _A4P, _A4P, _A4P, C.c_double, _I64P, _VP]),
    "qk_hydro_advance_stage_keep_fluxes": (C.c_int, [_VP, _PRM, C.c_int, _A4P, _A4P, _A4P, C.c_double, _I64P, _VP]),
    "qk_level_stage_fluxes": (C.c_int, [_VP, C.c_int, _A4P]),
    "qk_level_scratch_bytes": (C.c_int64, [_VP]),
    "qk_comm_unique_id": (C.c_int, [_VP]),
    "qk_comm_create": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(_VP)]),
    "qk_comm_destroy": (None, [_VP]),
    "qk_comm_rank": (C.c_int, [_VP]),
    "qk_comm_nranks": (C.c_int, [_VP]),
    "qk_level_set_comm": (C.c_int, [_VP, _VP]),
    "qk_fill_boundary": (C.c_int, [_VP, _A4P, C.c_int, C.c_int, _VP]),
    "qk_sim_create": (C.c_int, [C.POINTER(qk_level_desc), _PRM, C.c_double, _VP, C.POINTER(_VP)]),
    "qk_sim_destroy": (None, [_VP]),
    "qk_sim_nlocal": (C.c_int, [_VP]),
    "qk_sim_level": (_VP, [_VP]),
    "qk_sim_stream": (_VP, [_VP]),
    "qk_sim_time": (C.c_double, [_VP]),
    "qk_sim_cell_updates": (C.c_int64, [_VP]),
    "qk_sim_retries": (C.c_int64, [_VP]),
    "qk_sim_box_doubles": (C.c_int64, [_VP, C.c_int]),
    "qk_sim_state_desc": (C.c_int, [_VP, C.c_int, C.c_int, _A4P]),
    "qk_sim_set_state": (C.c_int, [_VP, C.c_int, _VP]),
    "qk_sim_box_valid_doubles": (C.c_int64, [_VP, C.c_int]),
    "qk_sim_set_state_valid": (C.c_int, [_VP, C.c_int, _VP]),
    "qk_sim_get_state_valid": (C.c_int, [_VP, C.c_int, _VP]),
    "qk_sim_get_state": (C.c_int, [_VP, C.c_int, _VP]),
    "qk_sim_sync": (C.c_int, [_VP]),
    "qk_sim_reset_clock": (None, [_VP, C.c_double, C.c_double]),
    "qk_sim_compute_timestep": (C.c_int, [_VP, C.c_double, _D3]),
    "qk_sim_step": (C.c_int, [_VP, C.c_double, C.POINTER(C.c_int)]),
    "qk_sim_enable_radiation": (C.c_int, [_VP, _RPRM, _RSPRM, _A4P, C.c_double, C.c_int]),
    "qk_sim_last_rad_substeps": (C.c_int, [_VP]),
    "qk_sim_rad_cell_updates": (C.c_int64, [_VP]),
    "qk_sim_evolve": (C.c_int, [_VP, C.c_int, C.c_double, C.POINTER(C.c_int), _D3, _D3]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libquokka_b200.so")
_lib = None


def load(path: str | None = None) -> C.CDLL:
    """Load libquokka_b200.so and bind every symbol of the header.  Raises if the library is missing:
    there is deliberately no CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  quokka_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.qk_abi_version() != 1:
        raise RuntimeError("libquokka_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().qk_error_string(code)
        raise RuntimeError(f"libquokka_b200 {what} failed: code {code} ({msg.decode() if msg else '?'})")
