/* quokka_b200.h -- C ABI of libquokka_b200.so: the B200-native (sm_100a) implementation of Quokka's
 * per-patch hydro update hot path (PPM -> shock flattening -> HLLC -> RK2 update with dual energy)
 * and, in the same style, the two-moment radiation transport sweep.
 *
 * Quokka has no FFI: its seam is the static-function surface of HydroSystem<problem_t> /
 * HyperbolicSystem<problem_t> / RadSystem<problem_t> as called from QuokkaSimulation<problem_t>
 * (reference: src/QuokkaSimulation.hpp:1096-1278 hydro, :1806-1848 radiation).  Every entry point
 * below names the reference operator (file:line) it replaces.  The compile-time traits of the
 * reference (EOS_Traits<>, HydroSystem_Traits<>, Physics_Traits<>) become the run-time
 * qk_hydro_params; amrex::Array4<double> becomes the layout-identical POD qk_array4, so a
 * maintainer can pass `reinterpret_cast<qk_array4 const*>(&mf.arrays()[i])`-style views without
 * copying.
 *
 * Conventions
 *  - all pointers inside qk_array4 are DEVICE pointers; descriptor arrays themselves are HOST arrays
 *  - a "MultiFab" is `nboxes` descriptors + `nboxes` valid boxes (cell-centred, inclusive bounds)
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), except where a
 *    scalar is returned to the host (qk_*_reduce_*, ncells_bad), which synchronise that stream
 *  - return value: 0 = ok, >0 = cudaError_t, <0 = QK_ERR_*; kernels never abort: bad cells are
 *    reported through redoFlag / ncells_bad exactly as the reference does
 *    (src/QuokkaSimulation.hpp:1146,1234)
 *  - there is NO CPU fallback: without a CUDA device every compute entry returns QK_ERR_NO_DEVICE
 */
#ifndef QUOKKA_B200_H_
#define QUOKKA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QK_ABI_VERSION 1
#define QK_MAX_SCALARS 8
#define QK_MAX_GROUPS 8

enum {
	QK_OK = 0,
	QK_ERR_NO_DEVICE = -1,
	QK_ERR_BAD_ARG = -2,
	QK_ERR_UNSUPPORTED = -3,
	QK_ERR_NOMEM = -4,
	/* the implicit matter-radiation solve did not converge in some cell (the reference aborts: "Newton-Raphson iteration for
	 * matter-radiation coupling failed to converge!", src/QuokkaSimulation.hpp:1700-1711) or the radiation subcycle needs more
	 * than maxSubsteps_ + 1 substeps (AMREX_ALWAYS_ASSERT, :1597) */
	QK_ERR_NOT_CONVERGED = -5
};

/* flux direction: FluxDir::{X1,X2,X3}, src/util/ArrayView_3d.hpp:18 */
enum { QK_X1 = 0, QK_X2 = 1, QK_X3 = 2 };
/* RiemannSolver::{HLLC,LLF}, src/hydro/hydro_system.hpp:43 (HLLD/MHD is out of scope) */
enum { QK_HLLC = 0, QK_LLF = 1 };
/* reconstruction order as QuokkaSimulation::reconstructionOrder_ (src/QuokkaSimulation.hpp:1498-1506):
 * 1 donor cell, 2 PLM, 3 PPM.  SlopeLimiter::{minmod,MC}, src/hyperbolic_system.hpp:38 */
enum { QK_MINMOD = 0, QK_MC = 1 };
/* amrex::BCType values used by the path (extern/amrex/Src/Base/AMReX_BC_TYPES.H) */
enum { QK_BC_INT_DIR = 0, QK_BC_REFLECT_ODD = -1, QK_BC_REFLECT_EVEN = 1, QK_BC_FOEXTRAP = 2, QK_BC_EXT_DIR = 3 };
/* arithmetic mode: EXACT reproduces the reference's IEEE operation order with FMA contraction off
 * (the reference builds with --fmad=false, CMakeLists.txt:31) and is bit-identical to it;
 * FAST allows contraction and algebraically-equivalent EOS shortcuts (drift is reported) */
enum { QK_ARITH_EXACT = 0, QK_ARITH_FAST = 1 };

/* == amrex::Array4<double> (extern/amrex/Src/Base/AMReX_Array4.H:59-68): Fortran order, x fastest,
 * component-major: p[(i-begin[0]) + (j-begin[1])*jstride + (k-begin[2])*kstride + n*nstride];
 * `end` is one past the last index, as in AMReX. */
typedef struct qk_array4 {
	double *p;
	int64_t jstride, kstride, nstride;
	int32_t begin[3];
	int32_t end[3];
	int32_t ncomp;
} qk_array4;

/* == amrex::Array4<int> (redoFlag, src/hyperbolic_system.hpp:34) */
typedef struct qk_iarray4 {
	int32_t *p;
	int64_t jstride, kstride, nstride;
	int32_t begin[3];
	int32_t end[3];
	int32_t ncomp;
} qk_iarray4;

/* == amrex::Box (cell-centred), inclusive bounds */
typedef struct qk_box {
	int32_t lo[3];
	int32_t hi[3];
} qk_box;

/* run-time image of EOS_Traits<problem_t> (src/hydro/EOS.hpp:32-37), HydroSystem_Traits<problem_t>
 * (src/hydro/hydro_system.hpp:38-41), Physics_Traits<problem_t> and the hydro.* / top-level
 * parameters the path reads (src/QuokkaSimulation.hpp:107-131, src/simulation.hpp:172-173) */
typedef struct qk_hydro_params {
	double gamma;		      /* EOS_Traits::gamma -> eos_rp::eos_gamma (QuokkaSimulation.hpp:160) */
	double mean_molecular_weight; /* EOS_Traits::mean_molecular_weight [g] */
	double boltzmann_constant;    /* EOS_Traits::boltzmann_constant */
	double small_temp;	      /* eos_init small_temp = 1e-10 (QuokkaSimulation.hpp:165) */
	double small_dens;	      /* eos_init small_dens = 1e-100 (QuokkaSimulation.hpp:166) */
	double density_floor;	      /* densityFloor_ (simulation.hpp:172) */
	double temp_floor;	      /* tempFloor_ (simulation.hpp:173) */
	double K_visc;		      /* artificialViscosityK_ (QuokkaSimulation.hpp:120); != 0 takes the one-kernel-per-operator path */
	double small_x;		      /* network_rp::small_x, mass-scalar floor (hydro_system.hpp:730) */
	int32_t reconstruct_eint;     /* HydroSystem_Traits::reconstruct_eint */
	int32_t nscalars;	      /* Physics_Traits::numPassiveScalars (includes mass scalars) */
	int32_t nmscalars;	      /* Physics_Traits::numMassScalars */
	int32_t reconstruction_order; /* reconstructionOrder_: 1|2|3 */
	int32_t use_dual_energy;      /* useDualEnergy_ */
	int32_t integrator_order;     /* integratorOrder_: 1|2 */
	int32_t abort_on_fofc_failure; /* abortOnFofcFailure_ */
	int32_t arith;		      /* QK_ARITH_* */
	double cs_isothermal;	      /* EOS_Traits::cs_isothermal (src/hydro/EOS.hpp:34), read only when gamma == 1: the isothermal EOS
				       * (HydroSystem::is_eos_isothermal(), hydro_system.hpp:133) runs on the one-kernel-per-operator path */
} qk_hydro_params;

/* ---- library / device --------------------------------------------------------------------- */
int qk_abi_version(void);
/* number of CUDA devices visible (0 => every compute entry returns QK_ERR_NO_DEVICE) */
int qk_device_count(void);
const char *qk_error_string(int code);
/* total kernels launched by this library since load (for bench.py "gpu_launches") */
int64_t qk_launch_count(void);
/* optional device timing per kernel class (CUDA events on the launching stream; used by bench.py for the roofline
 * line).  qk_prof_report writes "name launches total_ms" lines and returns the bytes needed. */
int qk_prof_enable(int on);
int qk_prof_report(char *buf, int buflen);
/* device self-test of the shared-reciprocal FP64 division used by the fused sweeps (csrc/qk_div.cuh): compares
 * npairs quotients bit for bit with the compiler's IEEE division; mode 0 random bit patterns, 1 moderate exponents,
 * 2 with zeros / subnormals / infinities / NaNs mixed in.  *bad_rcp counts refined reciprocals != 1.0/b. */
int qk_selftest_division(uint64_t seed, int mode, int64_t npairs, int64_t *bad_div, int64_t *bad_rcp);

/* ---- per-operator entry points (one per reference operator; parity harness + drop-in) ------ */

/* HydroSystem::ConservedToPrimitive(cons_mf, primVar_mf, nghost)  src/hydro/hydro_system.hpp:138-196 */
int qk_hydro_conserved_to_primitive(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim,
				    int nghost, void *stream);

/* HydroSystem::ComputeFlatteningCoefficients<DIR>(primVar_mf, x1Chi_mf, nghost)  hydro_system.hpp:531-626 */
int qk_hydro_flattening_coefficients(const qk_hydro_params *prm, int dir, int nboxes, const qk_box *valid, const qk_array4 *prim,
				     const qk_array4 *chi, int nghost, void *stream);

/* HyperbolicSystem::ReconstructStates{Constant,PLM<limiter>,PPM}<DIR>(q, left, right, nghost, nvars)
 * src/hyperbolic_system.hpp:129-181,183-247,295-433.  left/right are nodal in `dir`. */
int qk_reconstruct_states(int order, int limiter, int dir, int nboxes, const qk_box *valid, const qk_array4 *q, const qk_array4 *left,
			  const qk_array4 *right, int nghost, int nvars, void *stream);

/* HydroSystem::FlattenShocks<DIR>(q, chi1, chi2, chi3, left, right, nghost, nvars)  hydro_system.hpp:628-694 */
int qk_hydro_flatten_shocks(int dir, int nboxes, const qk_box *valid, const qk_array4 *q, const qk_array4 *chi1, const qk_array4 *chi2,
			    const qk_array4 *chi3, const qk_array4 *left, const qk_array4 *right, int nghost, int nvars, void *stream);

/* HydroSystem::ComputeFluxes<RIEMANN,DIR>(flux, faceVel, left, right, primVar, K_visc)  hydro_system.hpp:852-1112
 * (+ Riemann::HLLC src/hydro/HLLC.hpp:21-153, Riemann::LLF src/hydro/LLF.hpp:15-43).
 * Computes on the faces of `valid` (nodal in dir: hi[dir]+1 included). */
int qk_hydro_compute_fluxes(const qk_hydro_params *prm, int solver, int dir, int nboxes, const qk_box *valid, const qk_array4 *flux,
			    const qk_array4 *facevel, const qk_array4 *left, const qk_array4 *right, const qk_array4 *prim, void *stream);

/* hydroFluxFunction<DIR> fused (Reconstruct + FlattenShocks + ComputeFluxes<HLLC>)  QuokkaSimulation.hpp:1492-1517,
 * and hydroFOFluxFunction<DIR> (donor cell + LLF) :1559-1568 when fo != 0.
 * chi1..3 may be NULL when fo != 0.  No left/right states are materialised. */
int qk_hydro_flux_function(const qk_hydro_params *prm, int fo, int dir, int nboxes, const qk_box *valid, const qk_array4 *prim,
			   const qk_array4 *chi1, const qk_array4 *chi2, const qk_array4 *chi3, const qk_array4 *flux, const qk_array4 *facevel,
			   void *stream);

/* MultiFab::Saxpy(dst, a, src, 0, 0, ncomp, 0) on the valid faces/cells  QuokkaSimulation.hpp:1105-1108 */
int qk_saxpy(int nboxes, const qk_box *region, const qk_array4 *dst, double a, const qk_array4 *src, int ncomp, void *stream);

/* HydroSystem::ComputeRhsFromFluxes(rhs, fluxArray, dx, nvars)  hydro_system.hpp:448-473 */
int qk_hydro_rhs_from_fluxes(int nboxes, const qk_box *valid, const qk_array4 *rhs, const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz,
			     const double dx[3], int nvars, void *stream);

/* HydroSystem::AddInternalEnergyPdV(rhs, consVar, dx, faceVelArray, redoFlag)  hydro_system.hpp:775-814 */
int qk_hydro_add_internal_energy_pdv(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *rhs, const qk_array4 *cons,
				     const double dx[3], const qk_array4 *vx, const qk_array4 *vy, const qk_array4 *vz, const qk_iarray4 *redo,
				     void *stream);

/* HydroSystem::PredictStep(consVarOld, consVarNew, rhs, dt, nvars, redoFlag)  hydro_system.hpp:475-497;
 * *ncells_bad = redoFlag.sum(0) (QuokkaSimulation.hpp:1146) if ncells_bad != NULL (synchronises) */
int qk_hydro_predict_step(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons_old, const qk_array4 *cons_new,
			  const qk_array4 *rhs, double dt, int nvars, const qk_iarray4 *redo, int64_t *ncells_bad, void *stream);

/* HydroSystem::EnforceLimits(densityFloor, tempFloor, state)  hydro_system.hpp:698-773 */
int qk_hydro_enforce_limits(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, void *stream);

/* HydroSystem::SyncDualEnergy(consVar)  hydro_system.hpp:816-850.  Cells with rho<=0 (where the
 * reference calls amrex::Abort) are counted into *nabort instead (may be NULL). */
int qk_hydro_sync_dual_energy(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, int64_t *nabort, void *stream);

/* QuokkaSimulation::replaceFluxes(fluxes, FOfluxes, redoFlag) for one direction  QuokkaSimulation.hpp:1324-1368 */
int qk_hydro_replace_fluxes(int dir, int nboxes, const qk_box *valid, const qk_array4 *flux, const qk_array4 *fo_flux, const qk_iarray4 *redo,
			    int ncomp, void *stream);

/* HydroSystem::ComputeMaxSignalSpeed + MultiFab::norminf (hydro_system.hpp:223-252, simulation.hpp:709-710)
 * == HydroSystem::maxSignalSpeedLocal up to the |v| formula: which=0 uses the ComputeMaxSignalSpeed
 * expression, which=1 the maxSignalSpeedLocal one (hydro_system.hpp:198-221).  Synchronises. */
int qk_hydro_max_signal_speed(const qk_hydro_params *prm, int which, int nboxes, const qk_box *valid, const qk_array4 *cons, double *max_out,
			      void *stream);

/* ---- two-moment (M1) radiation transport sweep: RadSystem<problem_t>, src/radiation/radiation_system.hpp --------------
 * run-time image of RadSystem_Traits<problem_t> (:73-82) + Physics_Indices<problem_t>::radFirstIndex (src/physics_info.hpp:40)
 * + the radiation knobs of QuokkaSimulation (radiationReconstructionOrder_, src/QuokkaSimulation.hpp:125).  The state
 * MultiFab holds, per photon group g, (E_r, F_x, F_y, F_z) at components nstart + 4 g .. nstart + 4 g + 3 (:181). */
typedef struct qk_rad_params {
	double c_light;		      /* RadSystem_Traits::c_light */
	double c_hat;		      /* RadSystem_Traits::c_hat (reduced speed of light) */
	double Erad_floor;	      /* RadSystem_Traits::Erad_floor (total; divided by ngroups inside, :211) */
	int32_t ngroups;	      /* Physics_Traits::nGroups */
	int32_t nstart;		      /* nstartHyperbolic_ = radFirstIndex = 6 + numPassiveScalars */
	int32_t reconstruction_order; /* radiationReconstructionOrder_: 1 donor cell | 2 PLM(MC) | 3 PPM (QuokkaSimulation.hpp:1942-1957) */
	int32_t integrator_order;     /* 1 forward Euler | 2 RK2 (IMEX PD-ARS transport part, IMEX_a32 = 0.5, :52) */
	int32_t arith;		      /* QK_ARITH_EXACT (0, the default of a zero-initialised struct): bit-identical to the oracle; QK_ARITH_FAST: relaxed
				       * transport sweeps.  qk_rad_subcycle overrides it with hydro->arith. */
	int32_t use_wavespeed_correction; /* radiation.use_wavespeed_correction (src/QuokkaSimulation.hpp:133, radiation_system.hpp:1018-1022,1100-1109): on faces
				       * with even i+j+k the diffusive term of the ENERGY flux is scaled by min(1, 1/tau_cell), tau_cell = the harmonic mean of
				       * dl rho kappa_F of the two cells (ComputeCellOpticalDepth :803-871).  Constant flux-mean opacity only. */
	double kappa_F;		      /* ComputeFluxMeanOpacity (constant); read only with use_wavespeed_correction */
	double cell_dx[3];	      /* cell sizes for qk_rad_compute_fluxes with use_wavespeed_correction (the stage-level entries take the level's) */
} qk_rad_params;

/* RadSystem::ConservedToPrimitive(cons, primVar, ghostRange)  radiation_system.hpp:589-614.  prim has 4*ngroups components
 * (E_r, f_x, f_y, f_z per group, primVarIndex :183-188); computed on `valid` grown by nghost. */
int qk_rad_conserved_to_primitive(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim, int nghost,
				  void *stream);
/* RadSystem::ComputeFluxes<DIR>(x1Flux, x1FluxDiffusive, left, right, x1FluxRange, consVar, dx, use_wavespeed_correction = false)
 * :985-1139 (+ ComputeRadPressure :918-983, ComputeEddingtonTensor :873-916, ComputeEddingtonFactor :773-790).
 * flux / flux_diffusive: 4*ngroups components, nodal in dir; flux_diffusive may be NULL (the reference computes it but no
 * consumer reads it, :669,715-716). */
int qk_rad_compute_fluxes(const qk_rad_params *prm, int dir, int nboxes, const qk_box *valid, const qk_array4 *flux, const qk_array4 *flux_diffusive,
			  const qk_array4 *left, const qk_array4 *right, const qk_array4 *cons, void *stream);
/* RadSystem::PredictStep(consVarOld, consVarNew, fluxArray, ., dt, dx, indexRange, .)  :667-710 (isStateValid / amendRadState :624-665) */
int qk_rad_predict_step(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons_old, const qk_array4 *cons_new,
			const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz, double dt, const double dx[3], void *stream);
/* RadSystem::AddFluxesRK2(U_new, U0, U1, fluxArrayOld, fluxArray, ., ., dt, dx, indexRange, .)  :712-771 */
int qk_rad_add_fluxes_rk2(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *u_new, const qk_array4 *u0, const qk_array4 *u1,
			  const qk_array4 *fx_old, const qk_array4 *fy_old, const qk_array4 *fz_old, const qk_array4 *fx, const qk_array4 *fy,
			  const qk_array4 *fz, double dt, const double dx[3], void *stream);

/* ---- matter-radiation coupling source terms (single photon group): RadSystem<problem_t>::AddSourceTermsSingleGroup,
 * src/radiation/source_terms_single_group.hpp:9-565, called twice per radiation substep through operatorSplitSourceTerms
 * (src/QuokkaSimulation.hpp:1638,1656,1860-1885).  Run-time image of what the reference takes from RadSystem_Traits<problem_t>
 * (radiation_constant, beta_order; src/radiation/radiation_system.hpp:73-82) and from the problem's opacity specialisations
 * ComputePlanckOpacity / ComputeEnergyMeanOpacity / ComputeFluxMeanOpacity (:1141-1153; constants in
 * src/problems/RadhydroShell/test_radhydro_shell.cpp:127-135).  The solver hyper-parameters are the reference's compile-time
 * values (:34-44: include_work_term_in_source = true, enable_dE_constrain = true, force_rad_floor_in_iteration = false,
 * add_line_cooling_to_radiation_in_jac = false; no dust model, zero net cooling and cosmic-ray heating, IMEX_a32 = 0.5). */
enum { QK_OPACITY_CONSTANT = 0 };
typedef struct qk_rad_source_params {
	double radiation_constant; /* RadSystem_Traits::radiation_constant (a_rad) */
	double kappa_P;		   /* ComputePlanckOpacity(rho, T)      [cm^2/g], constant */
	double kappa_E;		   /* ComputeEnergyMeanOpacity(rho, T) */
	double kappa_F;		   /* ComputeFluxMeanOpacity(rho, T) */
	int32_t beta_order;	   /* RadSystem_Traits::beta_order: 0|1|2|3 */
	int32_t opacity_model;	   /* QK_OPACITY_CONSTANT */
} qk_rad_source_params;

/* counters[0..3] = iteration_counter {cells solved, sum of Newton-Raphson iterations, max Newton-Raphson iterations, 0},
 * counters[4..6] = iteration_failure_counter {Newton-Raphson not converged, dust temperature (always 0), outer work-term
 * iteration not converged}   (src/QuokkaSimulation.hpp:1620-1625). */
#define QK_RAD_SOURCE_NCOUNTERS 7
/* RadSystem::AddSourceTermsSingleGroup(consVar, radEnergySource, indexRange, dt_radiation, stage, ., p_iteration_counter,
 * p_iteration_failure_counter): updates gas momentum, gas energy, gas internal energy, E_r and F_r of `cons` in place on the
 * valid boxes.  hydro supplies the EOS (gamma, mean molecular weight, k_B, small_temp/small_dens); gamma == 1 takes the
 * reference's isothermal branch (flux update only).  rad_energy_source: one component per box (SetRadEnergySource,
 * QuokkaSimulation.hpp:1866-1872) or NULL for zero.  counters: HOST array of QK_RAD_SOURCE_NCOUNTERS, accumulated into
 * (the call then synchronises the stream), or NULL (fully asynchronous).  Arithmetic: the reference's operation order with
 * contraction off; T^3, T^4 and lorentz^3, which the reference takes from std::pow, are formed in double-double and rounded
 * once, so results agree with the CPU reference to the last bit except where libm's pow is itself not
 * correctly rounded; after the Newton-Raphson solve (residual tolerance 1e-11 E_tot) the stated parity bar is 1e-10 of the cell's
 * energy / momentum scale (measured <= 1e-14).  hydro->arith == QK_ARITH_FAST selects the relaxed form (closed-form EOS, reciprocal
 * products; <= 3e-12). */
int qk_rad_add_source_terms(const qk_hydro_params *hydro, const qk_rad_params *prm, const qk_rad_source_params *src, int stage, int nboxes,
			    const qk_box *valid, const qk_array4 *cons, const qk_array4 *rad_energy_source, double dt_radiation, int64_t *counters,
			    void *stream);

/* ---- coarse <-> fine transfer operators of the AMR ghost fill -------------------------------------------------
 * amrex::mf_linear_slope_minmax_interp, the cell-centred interpolater Quokka selects with amr_interpolation_method = 1
 * (getAmrInterpolaterCellCentered, src/simulation.hpp:1389-1407; MFCellConsLinMinmaxLimitInterp::interp,
 * extern/amrex/Src/AmrCore/AMReX_MFInterpolater.cpp:332-418 with AMReX_MFInterp_3D_C.H:7-109,246-262 and AMReX_MFInterp_C.H:10-90):
 * conservative linear interpolation whose slopes are limited so that no component gets a new extremum, with ONE limiter per
 * direction for all ncomp components (linear combinations of the components are preserved).  Per box pair p: crse[p] must
 * cover CoarseBox(fine_region[p]) = coarsen(fine_region[p]) grown by one cell in every refined direction (ghost cells of the
 * coarse data already filled); the cells of fine_region[p] that lie inside dest_domain are written, components
 * fcomp .. fcomp + ncomp - 1 from ccomp .. ccomp + ncomp - 1 (ncomp <= 16).  cdomain: the coarse level's domain; bc_lo / bc_hi:
 * amrex::BCType per [3 * comp + dim] (one-sided slopes at ext_dir / hoextrap walls).  One launch for all pairs (16 per launch);
 * no slope MultiFab.  Bit-identical to AMReX. */
int qk_amr_interp_cons_lin_minmax(int npatch, const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *fine_region,
				  const qk_box *dest_domain, const qk_box *cdomain, const int ratio[3], const int32_t *bc_lo, const int32_t *bc_hi,
				  void *stream);
/* amrex::average_down(S_fine, S_crse, scomp, ncomp, ratio) (extern/amrex/Src/Base/AMReX_MultiFabUtil_3D_C.H:345-375; AverageDownTo,
 * src/simulation.hpp:1309-1343): crse(i,j,k) = mean of its ratio^3 fine cells on the coarse boxes cbx[p]. */
int qk_amr_average_down(int npatch, const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *cbx, const int ratio[3],
			void *stream);

/* QuokkaSimulation::PreInterpState / PostInterpState (src/QuokkaSimulation.hpp:804-841), the hooks FillPatcher runs on the coarse data
 * before and on the fine data after the interpolation: gas total energy (component 4) -> specific internal energy (E - KE) / rho,
 * and back E = rho e + KE, on the cells of bx[b] of state[b].  One launch for all boxes. */
int qk_amr_pre_interp_state(int nboxes, const qk_box *bx, const qk_array4 *state, void *stream);
int qk_amr_post_interp_state(int nboxes, const qk_box *bx, const qk_array4 *state, void *stream);

/* Time interpolation of the coarse data of the fine-level ghost fill: amrex::FillPatcher::fill (extern/amrex/Src/AmrCore/AMReX_FillPatcher.H:
 * 340-387; the same expression in FillPatchSingleLevel, AMReX_FillPatchUtil_I.H:140-175), which AMRSimulation::fillBoundaryConditions
 * reaches through FillPatchWithData (src/simulation.hpp:1789-1858) with the old and new coarse states that GetData returns (:1860-1905).
 * With teps = |t1 - t0| * 1e-3:  time within teps of t0 -> copy of src0 (returns 0); within teps of t1 -> copy of src1 (returns 1); else
 *     dst(i,j,k,dcomp+n) = alpha * src0(i,j,k,scomp+n) + beta * src1(i,j,k,scomp+n),  alpha = (t1-time)/(t1-t0), beta = (time-t0)/(t1-t0)
 * (two roundings of the products and one of the sum, no contraction; returns 2) on region[p] of each patch.  src1 may be NULL when only one
 * coarse time level exists (copy of src0).  *which (may be NULL) receives the branch taken.  One launch for all patches. */
int qk_amr_time_interp(int npatch, const qk_array4 *dst, int dcomp, const qk_array4 *src0, const qk_array4 *src1, int scomp, int ncomp,
		       const qk_box *region, double t0, double t1, double time, int *which, void *stream);

/* ---- regrid support ------------------------------------------------------------------------------------------------------
 * amrex::TagBox is BaseFab<char>: the tag arrays have the layout of qk_array4 with 1-byte elements; TagBox::CLEAR = 0, TagBox::BUF = 1, TagBox::SET = 2
 * (extern/amrex/Src/AmrCore/AMReX_TagBox.H:33). */
#define QK_TAG_SET 2
typedef struct qk_carray4 {
	char *p;
	int64_t jstride, kstride, nstride;
	int32_t begin[3], end[3];
	int32_t ncomp;
} qk_carray4;
/* QuokkaSimulation<SedovProblem>::ErrorEst (src/problems/HydroBlast3D/test_hydro3d_blast.cpp:118-151): with P = HydroSystem::ComputePressure
 * of the conserved state (needs one filled ghost cell), tag(i,j,k) = SET where
 *     max over x,y,z of max(|P(+1) - P|, |P - P(-1)|) / P > eta_threshold   and   P > P_min.
 * Cells that do not satisfy the criterion are left untouched (the reference never clears).  *ntagged (may be NULL: asynchronous) receives
 * the number of cells this call set. */
int qk_tag_pressure_gradient(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, const qk_carray4 *tags,
			     double eta_threshold, double P_min, int64_t *ntagged, void *stream);
/* QuokkaSimulation<ShocktubeProblem>::ErrorEst (src/problems/HydroShocktube/test_hydro_shocktube.cpp:146-171): component `comp`,
 *     sqrt(del^2) / q > eta_threshold and q >= q_min,   del = (q(i+1,j,k) - q(i-1,j,k)) / (2.0 * dx)        (x direction only) */
int qk_tag_gradient_x(int nboxes, const qk_box *valid, const qk_array4 *state, int comp, const qk_carray4 *tags, double dx, double eta_threshold,
		      double q_min, int64_t *ntagged, void *stream);
/* QuokkaSimulation::FixupState (src/QuokkaSimulation.hpp:761-770), called after reflux / average-down (src/simulation.hpp:1311): EnforceLimits
 * followed by SyncDualEnergy on the valid cells, one launch each for all boxes. */
int qk_hydro_fixup_state(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, void *stream);

/* ---- level object: fused path + ghost fill -------------------------------------------------- */

/* Description of the boxes of ONE AMR level owned by this rank (a MultiFab's local part) and of
 * the whole level (for box<->box ghost copies): the run-time image of BoxArray +
 * DistributionMapping + Geometry + BCRec as used by fillBoundaryConditions (simulation.hpp:1704-1785). */
typedef struct qk_level_desc {
	qk_box domain;	     /* geom.Domain() */
	int32_t periodic[3]; /* geom.isPeriodic(d) */
	double dx[3];	     /* geom.CellSizeArray() */
	int32_t nghost;	     /* nghost_cc_ = 4 */
	int32_t ncomp;	     /* components of the state MultiFab */
	int32_t nboxes_global;
	const qk_box *boxes_global; /* BoxArray */
	const int32_t *owner;	    /* DistributionMapping: rank of each global box */
	int32_t my_rank;
	/* BCRec per component: bc_lo[n*3+d], bc_hi[n*3+d] = QK_BC_* */
	const int32_t *bc_lo;
	const int32_t *bc_hi;
} qk_level_desc;

typedef struct qk_level qk_level; /* opaque */

/* one remote copy the caller must transport (pack on src rank -> send -> unpack on dst rank) */
typedef struct qk_copy_tag {
	int32_t src_box, dst_box; /* global box ids */
	int32_t src_rank, dst_rank;
	qk_box src_region; /* cells to read in the source box's index space */
	int32_t shift[3];  /* dst index = src index + shift (periodic wrap) */
	int64_t offset;	   /* offset in doubles of this tag inside the (src_rank -> dst_rank) message, per component */
	int64_t ncells;
} qk_copy_tag;

int qk_level_create(const qk_level_desc *desc, qk_level **out);
void qk_level_destroy(qk_level *lev);
int qk_level_nlocal(const qk_level *lev);
/* global ids of the local boxes, in local order (== MFIter order) */
int qk_level_local_ids(const qk_level *lev, int32_t *ids);
/* copy tags whose src or dst is this rank and src_rank != dst_rank; returns count (tags may be NULL) */
int qk_level_remote_tags(const qk_level *lev, qk_copy_tag *tags, int max_tags);
/* copy tags whose src and dst boxes both live on this rank (offset is unused); returns count */
int qk_level_local_tags(const qk_level *lev, qk_copy_tag *tags, int max_tags);

/* FabArray::FillBoundary, same-rank part (FB_local_copy_gpu, extern/amrex/Src/Base/AMReX_FBI.H:272) */
int qk_fill_boundary_local(qk_level *lev, const qk_array4 *state, int scomp, int ncomp, void *stream);
/* pack_send_buffer_gpu / unpack_recv_buffer_gpu (AMReX_FBI.H:730,790) for the message to/from `peer`:
 * buffer layout = for each tag (in qk_level_remote_tags order restricted to that peer): ncomp x ncells doubles */
int qk_pack_ghosts(qk_level *lev, int peer, const qk_array4 *state, int scomp, int ncomp, double *buf, int64_t *ndoubles, void *stream);
int qk_unpack_ghosts(qk_level *lev, int peer, const qk_array4 *state, int scomp, int ncomp, const double *buf, void *stream);
/* PhysBCFunct<GpuBndryFuncFab<..>> with amrex::FilccCell (simulation.hpp:1760-1765;
 * extern/amrex/Src/Base/AMReX_FilCC_3D_C.H): reflect_even/odd, foextrap; ext_dir cells are left for the caller */
int qk_fill_physical_bc(qk_level *lev, const qk_array4 *state, int scomp, int ncomp, void *stream);

/* One RK stage of QuokkaSimulation::advanceHydroAtLevel (src/QuokkaSimulation.hpp:1099-1198 stage 1,
 * :1202-1285 stage 2) for all local boxes: computeHydroFluxes (K1-K5) -> RK2 flux average ->
 * ComputeRhsFromFluxes -> AddInternalEnergyPdV -> PredictStep -> [FOFC: computeFOHydroFluxes,
 * replaceFluxes, redo] -> EnforceLimits -> SyncDualEnergy.
 *   stage 1: Uout = U0 + dt L(U0)                       (Ustage == U0, ghost-filled)
 *   stage 2: Uout = U0 + dt L(0.5 F(U0) + 0.5 F(Ustage)) (Ustage = stage-1 result, ghost-filled)
 * The level keeps 0.5*F(U0) and 0.5*faceVel(U0) between the two calls.  *ncells_bad receives the
 * redoFlag.sum() that survived FOFC (0 => success).  Synchronises `stream` once (for ncells_bad). */
int qk_hydro_advance_stage(qk_level *lev, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
			   const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream);

/* One stage of the radiation transport substep for all local boxes, fused (no left/right/flux arrays exist):
 *   stage 1 = advanceRadiationForwardEuler's transport part (src/QuokkaSimulation.hpp:1791-1822):
 *             Uout = PredictStep(U0, F(U0))                          (Ustage == U0, ghost-filled)
 *   stage 2 = advanceRadiationMidpointRK2's (:1824-1862):
 *             Uout = AddFluxesRK2(U0, Ustage, F(U0), F(Ustage))       (Ustage = stage-1 result, ghost-filled)
 * Only components nstart .. nstart + 4*ngroups - 1 are read and written.  The level keeps the stage-1 flux divergence
 * between the two calls (the reference re-evaluates F(U0); its coefficient 0.5 - IMEX_a32 is exactly 0, the product is still
 * formed so that non-finite values and signed zeros propagate identically). */
int qk_rad_advance_stage(qk_level *lev, const qk_rad_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout,
			 double dt, void *stream);

/* QuokkaSimulation::subcycleRadiationAtLevel (src/QuokkaSimulation.hpp:1577-1700) for a uniform level, hydro enabled, no flux
 * registers: nsub = computeNumberOfRadiationSubsteps (:397-406) = ceil(dt_hydro / (rad_cfl dx_min / c_hat)) IMEX PD-ARS substeps
 * of dt_hydro / nsub, each
 *     [i > 0: swapRadiationState, radiation components new -> old]                                  (:1570-1574,1604-1610)
 *     ghost fill of U_old; transport stage 1 U_old -> U_new (advanceRadiationForwardEuler)          (:1791-1822)
 *     source terms of stage 1 on U_new                                                              (:1628-1641)
 *     ghost fill of U_new; transport stage 2 (advanceRadiationMidpointRK2)                          (:1824-1862)
 *     source terms of stage 2 on U_new                                                              (:1649-1658)
 * U_old = state_old_cc_ (pre-step state; its radiation components are overwritten from the second substep on, as in the
 * reference), U_new = state_new_cc_ (hydro components already advanced; receives the result), U_tmp = a third MultiFab of the
 * same shape owned by the caller: the fused transport stage may not write the array it reads its neighbours from, so stage 2
 * goes U_new -> U_tmp and the four radiation components are copied back.  Only the radiation components' ghost cells are
 * filled (the transport reads nothing else outside the valid boxes).  src may be NULL: transport only.
 * rad_energy_source: one component per box or NULL.  counters as qk_rad_add_source_terms (accumulated over all substeps; the
 * call synchronises when non-NULL).  nsub_out (may be NULL) receives nsub. */
int qk_rad_subcycle(qk_level *lev, const qk_hydro_params *hydro, const qk_rad_params *prm, const qk_rad_source_params *src, const qk_array4 *U_old,
		    const qk_array4 *U_new, const qk_array4 *U_tmp, const qk_array4 *rad_energy_source, double dt_hydro, double rad_cfl,
		    int64_t *counters, int *nsub_out, void *stream);

/* The same stage through the FAITHFUL path only: one kernel per reference operator, fluxes materialised as
 * MultiFabs exactly as QuokkaSimulation.hpp:1403-1490 does.  qk_hydro_advance_stage runs the fused sweep kernels
 * and falls back to this path when a cell is flagged (FOFC); both produce identical bits. */
int qk_hydro_advance_stage_faithful(qk_level *lev, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
				    const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream);

/* The stage for a level with FLUX REGISTERS (AMR with do_reflux; src/QuokkaSimulation.hpp:1195-1198,1280-1283): the fused sweeps of
 * qk_hydro_advance_stage, which additionally store the stage's own face fluxes -- F(U0) in stage 1, F(U1) in stage 2 -- for
 * incrementFluxRegisters (src/simulation.hpp:1345-1387).  Same state bits as qk_hydro_advance_stage.  A stage the fused kernels cannot
 * take (rows that cannot be bulk-copied, an uninstantiated trait set) or that flags a cell (FOFC) runs on the faithful path; either way
 * qk_level_stage_fluxes returns the stage's flux arrays afterwards. */
int qk_hydro_advance_stage_keep_fluxes(qk_level *lev, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage,
				       const qk_array4 *Uout, double dt, int64_t *ncells_bad, void *stream);

/* Face fluxes of the LAST stage that kept them (qk_hydro_advance_stage_keep_fluxes or qk_hydro_advance_stage_faithful) in direction dir (0|1|2),
 * one descriptor per local box (qk_level_nlocal, MFIter order):
 * nodal in dir, 6 + nscalars components, no padding.  These are the `fluxArrays` QuokkaSimulation::advanceHydroAtLevel hands to
 * incrementFluxRegisters (src/QuokkaSimulation.hpp:1195-1198,1280-1283 -> src/simulation.hpp:1345-1387, YAFluxRegister::CrseAdd /
 * FineAdd): after stage 1 the stage's fluxes with the first-order replacements of FOFC applied, after stage 2 F(U1) as evaluated
 * (the reference passes the uncorrected stage-2 array there; the corrected average lives in flux_rk2).  The pointers stay valid
 * until the level is destroyed; the contents until the next stage.  QK_ERR_BAD_ARG before the first flux-keeping stage. */
int qk_level_stage_fluxes(const qk_level *lev, int dir, qk_array4 *out);

/* bytes of device scratch currently held by the level */
int64_t qk_level_scratch_bytes(const qk_level *lev);

/* ---- rank <-> rank transport: NCCL over NVLink/NVSwitch in place of AMReX's MPI calls --------------------
 * (FabArray::FillBoundary MPI_Isend/Irecv: extern/amrex/Src/Base/AMReX_FabArrayCommI.H:7-165;
 *  ParallelDescriptor::ReduceRealMax / ReduceLongSum: AMReX_ParallelDescriptor.cpp:1091,1659,1746).
 * One process per GPU.  The 128-byte id is created on rank 0 (qk_comm_unique_id) and handed to the other ranks by
 * whatever bootstrap the host application has (MPI_Bcast in Quokka; torch.distributed in bench.py). */
typedef struct qk_comm qk_comm; /* opaque */
int qk_comm_unique_id(void *id128);
int qk_comm_create(const void *id128, int rank, int nranks, qk_comm **out);
void qk_comm_destroy(qk_comm *comm);
int qk_comm_rank(const qk_comm *comm);
int qk_comm_nranks(const qk_comm *comm);
/* attach (or detach with NULL) the communicator used by qk_fill_boundary and the stage's ncells_bad sum */
int qk_level_set_comm(qk_level *lev, qk_comm *comm);

/* AMRSimulation::fillBoundaryConditions on level 0 (src/simulation.hpp:1752-1765): FillBoundary(periodicity)
 * = same-rank copies + packed neighbour messages over the communicator, then the physical BC fill. */
int qk_fill_boundary(qk_level *lev, const qk_array4 *state, int scomp, int ncomp, void *stream);

/* ---- single-level time-step driver (C++ host code inside the library) ---------------------------------------
 * The uniform-grid image of AMRSimulation::evolve/computeTimestep (src/simulation.hpp:722-977) and
 * QuokkaSimulation::advanceSingleTimestepAtLevel/advanceHydroAtLevelWithRetries/advanceHydroAtLevel
 * (src/QuokkaSimulation.hpp:653-707,885-990,1032-1322).  It owns state_new/state_old/state_inter on the device
 * in AMReX FAB layout (ghost cells included) and is what bench.py and the whole-run parity tests drive. */
typedef struct qk_sim qk_sim; /* opaque */
int qk_sim_create(const qk_level_desc *desc, const qk_hydro_params *prm, double cfl, qk_comm *comm, qk_sim **out);
void qk_sim_destroy(qk_sim *sim);
int qk_sim_nlocal(const qk_sim *sim);
qk_level *qk_sim_level(qk_sim *sim);
void *qk_sim_stream(qk_sim *sim);
double qk_sim_time(const qk_sim *sim);
int64_t qk_sim_cell_updates(const qk_sim *sim); /* cellUpdates_, simulation.hpp:1285 */
int64_t qk_sim_retries(const qk_sim *sim);
int64_t qk_sim_box_doubles(const qk_sim *sim, int local_box);
/* descriptor of state_new (which=0), state_old (1), state_inter (2) of a local box (device pointer) */
int qk_sim_state_desc(qk_sim *sim, int which, int local_box, qk_array4 *out);
/* host <-> device copy of state_new of one local box (whole FAB incl. ghosts); stream-ordered, see qk_sim_sync */
int qk_sim_set_state(qk_sim *sim, int local_box, const double *host);
int qk_sim_get_state(qk_sim *sim, int local_box, double *host);
/* the same for the VALID cells only: host = contiguous (ncomp, nz, ny, nx) array of the box's valid cells (what a host-resident caller
 * owns; the step fills the ghost cells itself, src/simulation.hpp:1704-1785).  Copies run on their own stream and overlap the (un)packing
 * kernel of the previous box; qk_sim_sync waits for both.  qk_sim_box_valid_doubles = ncomp * number of valid cells. */
int64_t qk_sim_box_valid_doubles(const qk_sim *sim, int local_box);
int qk_sim_set_state_valid(qk_sim *sim, int local_box, const double *host_valid);
int qk_sim_get_state_valid(qk_sim *sim, int local_box, double *host_valid);
int qk_sim_sync(qk_sim *sim);
void qk_sim_reset_clock(qk_sim *sim, double t, double dt_prev);
int qk_sim_compute_timestep(qk_sim *sim, double stop_time, double *dt_out);
/* one coarse step; *retries = number of dt-halving retries used, -1 if all 6 failed (the reference aborts there) */
int qk_sim_step(qk_sim *sim, double dt, int *retries);
/* Physics_Traits::is_radiation_enabled for this simulation: every qk_sim_step then runs subcycleRadiationAtLevel (qk_rad_subcycle)
 * after the hydro advance (src/QuokkaSimulation.hpp:690-694) and qk_sim_compute_timestep takes std::max(c_hat / max_substeps,
 * hydro signal speed) (computeMaxSignalLocal :408-441).  src == NULL: transport only.  rad_energy_source: one-component device
 * FABs per local box (caller-owned, must outlive the simulation) or NULL.  The level must carry the radiation components. */
int qk_sim_enable_radiation(qk_sim *sim, const qk_rad_params *rad, const qk_rad_source_params *src, const qk_array4 *rad_energy_source, double rad_cfl,
			    int max_substeps);
int qk_sim_last_rad_substeps(const qk_sim *sim);
int64_t qk_sim_rad_cell_updates(const qk_sim *sim); /* radiationCellUpdates_, src/QuokkaSimulation.hpp:1697 */
int qk_sim_evolve(qk_sim *sim, int max_steps, double stop_time, int *steps_done, double *elapsed_s, double *device_ms);

#ifdef __cplusplus
}
#endif
#endif /* QUOKKA_B200_H_ */
