// quokka_b200_amrex.hpp -- header-only C++17 shim between Quokka's operator surface and the C ABI of
// libquokka_b200.so (include/quokka_b200.h).
//
// `quokka::b200::HydroSystemB200<problem_t>` has the same static-function surface as the reference's
// `HydroSystem<problem_t>` / `HyperbolicSystem<problem_t>` (src/hydro/hydro_system.hpp:66-135,
// src/hyperbolic_system.hpp:85-125) for every operator `QuokkaSimulation<problem_t>::advanceHydroAtLevel` calls
// (src/QuokkaSimulation.hpp:1096-1278, 1403-1568): same names, same argument order and meaning, MultiFabs in and out.
// A maintainer switches a call site by replacing `HydroSystem<problem_t>::` with `HydroSystemB200<problem_t>::`
// (INTEGRATION.md shows the patch).  The compile-time traits of the problem become the run-time qk_hydro_params.
//
// Nothing here computes: every function builds Array4 views of the local FABs and forwards to the library.  AMReX must
// be built with GPU support for the pointers to be device pointers; kernels are enqueued on amrex::Gpu::gpuStream().
// This file is written for this repository; it includes the reference's headers only for the trait/enum declarations.
#pragma once

#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "AMReX_Array4.H"
#include "AMReX_BCRec.H"
#include "AMReX_Geometry.H"
#include "AMReX_GpuDevice.H"
#include "AMReX_MFInterpolater.H"
#include "AMReX_MultiFab.H"
#include "AMReX_MultiFabUtil.H"
#include "AMReX_iMultiFab.H"

#include "radiation/radiation_system.hpp" // RadSystem_Traits, RadSystem<problem_t> constants
#include "hydro/hydro_system.hpp" // RiemannSolver, FluxDir, SlopeLimiter, EOS_Traits, HydroSystem_Traits, Physics_Traits
#include "quokka_b200.h"

namespace quokka::b200
{

inline void check(int rc, const char *what)
{
	if (rc != QK_OK) {
		// the reference reports operator failures with amrex::Abort
		amrex::Abort(std::string("libquokka_b200: ") + what + " failed: " + qk_error_string(rc));
	}
}

// amrex::Array4<T> and qk_array4 have the same members in the same order; copy field by field so that no layout
// assumption beyond the names is needed
template <typename T> inline auto view(amrex::Array4<T> const &a) -> qk_array4
{
	qk_array4 v;
	v.p = const_cast<double *>(reinterpret_cast<double const *>(a.p));
	v.jstride = a.jstride;
	v.kstride = a.kstride;
	v.nstride = a.nstride;
	v.begin[0] = a.begin.x;
	v.begin[1] = a.begin.y;
	v.begin[2] = a.begin.z;
	v.end[0] = a.end.x;
	v.end[1] = a.end.y;
	v.end[2] = a.end.z;
	v.ncomp = a.ncomp;
	return v;
}
template <typename T> inline auto iview(amrex::Array4<T> const &a) -> qk_iarray4
{
	qk_iarray4 v;
	v.p = const_cast<int32_t *>(reinterpret_cast<int32_t const *>(a.p));
	v.jstride = a.jstride;
	v.kstride = a.kstride;
	v.nstride = a.nstride;
	v.begin[0] = a.begin.x;
	v.begin[1] = a.begin.y;
	v.begin[2] = a.begin.z;
	v.end[0] = a.end.x;
	v.end[1] = a.end.y;
	v.end[2] = a.end.z;
	v.ncomp = a.ncomp;
	return v;
}

inline auto to_box(amrex::Box const &b) -> qk_box
{
	const amrex::Box c = amrex::enclosedCells(b);
	qk_box q;
	for (int d = 0; d < 3; ++d) {
		q.lo[d] = (d < AMREX_SPACEDIM) ? c.smallEnd(d) : 0;
		q.hi[d] = (d < AMREX_SPACEDIM) ? c.bigEnd(d) : 0;
	}
	return q;
}

// the local part of a MultiFab as the (descriptors, valid boxes) pair the ABI takes, in MFIter order
struct MFView {
	std::vector<qk_array4> arr;
	std::vector<qk_box> valid;
	explicit MFView(amrex::MultiFab const &mf)
	{
		for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi) {
			arr.push_back(view(mf.const_array(mfi)));
			valid.push_back(to_box(mfi.validbox()));
		}
	}
	[[nodiscard]] auto n() const -> int { return static_cast<int>(arr.size()); }
};
struct iMFView {
	std::vector<qk_iarray4> arr;
	explicit iMFView(amrex::iMultiFab const &mf)
	{
		for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi) {
			arr.push_back(iview(mf.const_array(mfi)));
		}
	}
};

inline auto stream() -> void *
{
#ifdef AMREX_USE_GPU
	return static_cast<void *>(amrex::Gpu::gpuStream());
#else
	return nullptr; // host-only AMReX build: the library refuses with QK_ERR_NO_DEVICE unless the FABs live in device memory
#endif
}

// run-time image of the problem's traits (include/quokka_b200.h qk_hydro_params); the simulation-level knobs
// (floors, reconstruction order, ...) are filled by the caller from QuokkaSimulation's members
template <typename problem_t> inline auto make_params() -> qk_hydro_params
{
	qk_hydro_params p{};
	p.gamma = quokka::EOS_Traits<problem_t>::gamma;
	p.mean_molecular_weight = quokka::EOS_Traits<problem_t>::mean_molecular_weight;
	p.boltzmann_constant = quokka::EOS_Traits<problem_t>::boltzmann_constant;
	p.small_temp = 1e-10;  // eos_init arguments, src/QuokkaSimulation.hpp:165-166
	p.small_dens = 1e-100;
	p.density_floor = 0.0;
	p.temp_floor = 0.0;
	p.K_visc = 0.0;
	p.small_x = network_rp::small_x;
	p.reconstruct_eint = HydroSystem_Traits<problem_t>::reconstruct_eint ? 1 : 0;
	p.nscalars = Physics_Traits<problem_t>::numPassiveScalars;
	p.nmscalars = Physics_Traits<problem_t>::numMassScalars;
	p.reconstruction_order = 3;
	p.use_dual_energy = 1;
	p.integrator_order = 2;
	p.abort_on_fofc_failure = 1;
	p.arith = QK_ARITH_EXACT;
	// src/hydro/EOS.hpp:34; problem traits that are not isothermal need not define it (the reference reads it only inside
	// `if constexpr (is_eos_isothermal())`, hydro_system.hpp:132-133)
	if constexpr (quokka::EOS_Traits<problem_t>::gamma == 1.0) {
		p.cs_isothermal = quokka::EOS_Traits<problem_t>::cs_isothermal;
	} else {
		p.cs_isothermal = NAN;
	}
	return p;
}

template <typename problem_t> class HydroSystemB200
{
      public:
	static constexpr int nvar_ = HydroSystem<problem_t>::nvar_;

	// HydroSystem::ConservedToPrimitive  src/hydro/hydro_system.hpp:138-196
	static void ConservedToPrimitive(amrex::MultiFab const &cons_mf, amrex::MultiFab &primVar_mf, int nghost)
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView c(cons_mf);
		MFView q(primVar_mf);
		check(qk_hydro_conserved_to_primitive(&prm, c.n(), c.valid.data(), c.arr.data(), q.arr.data(), nghost, stream()), "ConservedToPrimitive");
	}

	// HydroSystem::ComputeFlatteningCoefficients<DIR>  :531-626
	template <FluxDir DIR> static void ComputeFlatteningCoefficients(amrex::MultiFab const &primVar_mf, amrex::MultiFab &x1Chi_mf, int nghost)
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView q(primVar_mf);
		MFView chi(x1Chi_mf);
		check(qk_hydro_flattening_coefficients(&prm, static_cast<int>(DIR), q.n(), q.valid.data(), q.arr.data(), chi.arr.data(), nghost, stream()),
		      "ComputeFlatteningCoefficients");
	}

	// HyperbolicSystem::ReconstructStatesConstant / PLM<limiter> / PPM  src/hyperbolic_system.hpp:129-181,183-247,295-433
	template <FluxDir DIR>
	static void ReconstructStatesConstant(amrex::MultiFab const &q_mf, amrex::MultiFab &leftState_mf, amrex::MultiFab &rightState_mf, int nghost, int nvars)
	{
		reconstruct<DIR>(1, QK_MINMOD, q_mf, leftState_mf, rightState_mf, nghost, nvars);
	}
	template <FluxDir DIR, SlopeLimiter limiter>
	static void ReconstructStatesPLM(amrex::MultiFab const &q_mf, amrex::MultiFab &leftState_mf, amrex::MultiFab &rightState_mf, int nghost, int nvars)
	{
		reconstruct<DIR>(2, limiter == SlopeLimiter::MC ? QK_MC : QK_MINMOD, q_mf, leftState_mf, rightState_mf, nghost, nvars);
	}
	template <FluxDir DIR>
	static void ReconstructStatesPPM(amrex::MultiFab const &q_mf, amrex::MultiFab &leftState_mf, amrex::MultiFab &rightState_mf, int nghost, int nvars)
	{
		reconstruct<DIR>(3, QK_MINMOD, q_mf, leftState_mf, rightState_mf, nghost, nvars);
	}

	// HydroSystem::FlattenShocks<DIR>  hydro_system.hpp:628-694
	template <FluxDir DIR>
	static void FlattenShocks(amrex::MultiFab const &q_mf, amrex::MultiFab const &x1Chi_mf, amrex::MultiFab const &x2Chi_mf, amrex::MultiFab const &x3Chi_mf,
				  amrex::MultiFab &x1LeftState_mf, amrex::MultiFab &x1RightState_mf, int nghost, int nvars)
	{
		MFView q(q_mf);
		MFView c1(x1Chi_mf);
		MFView c2(x2Chi_mf);
		MFView c3(x3Chi_mf);
		MFView l(x1LeftState_mf);
		MFView r(x1RightState_mf);
		check(qk_hydro_flatten_shocks(static_cast<int>(DIR), q.n(), q.valid.data(), q.arr.data(), c1.arr.data(), c2.arr.data(), c3.arr.data(),
					      l.arr.data(), r.arr.data(), nghost, nvars, stream()),
		      "FlattenShocks");
	}

	// HydroSystem::ComputeFluxes<RIEMANN, DIR>  :852-1112 (HLLC, LLF; HLLD/MHD is out of scope)
	template <RiemannSolver RIEMANN, FluxDir DIR>
	static void ComputeFluxes(amrex::MultiFab &x1Flux_mf, amrex::MultiFab &x1FaceVel_mf, amrex::MultiFab const &x1LeftState_mf,
				  amrex::MultiFab const &x1RightState_mf, amrex::MultiFab const &primVar_mf, amrex::Real K_visc)
	{
		static_assert(RIEMANN == RiemannSolver::HLLC || RIEMANN == RiemannSolver::LLF, "HLLD is not provided by libquokka_b200");
		qk_hydro_params prm = make_params<problem_t>();
		prm.K_visc = K_visc;
		MFView q(primVar_mf);
		MFView f(x1Flux_mf);
		MFView v(x1FaceVel_mf);
		MFView l(x1LeftState_mf);
		MFView r(x1RightState_mf);
		check(qk_hydro_compute_fluxes(&prm, RIEMANN == RiemannSolver::HLLC ? QK_HLLC : QK_LLF, static_cast<int>(DIR), q.n(), q.valid.data(),
					      f.arr.data(), v.arr.data(), l.arr.data(), r.arr.data(), q.arr.data(), stream()),
		      "ComputeFluxes");
	}

	// HydroSystem::ComputeRhsFromFluxes  :448-473
	static void ComputeRhsFromFluxes(amrex::MultiFab &rhs_mf, std::array<amrex::MultiFab, AMREX_SPACEDIM> const &fluxArray,
					 amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx, int nvars)
	{
		static_assert(AMREX_SPACEDIM == 3, "libquokka_b200 operates on 3-D FABs (1-D/2-D problems use one-cell-thick boxes)");
		MFView rhs(rhs_mf);
		MFView fx(fluxArray[0]);
		MFView fy(fluxArray[1]);
		MFView fz(fluxArray[2]);
		const double d[3] = {dx[0], dx[1], dx[2]};
		check(qk_hydro_rhs_from_fluxes(rhs.n(), rhs.valid.data(), rhs.arr.data(), fx.arr.data(), fy.arr.data(), fz.arr.data(), d, nvars, stream()),
		      "ComputeRhsFromFluxes");
	}

	// HydroSystem::AddInternalEnergyPdV  :775-814
	static void AddInternalEnergyPdV(amrex::MultiFab &rhs_mf, amrex::MultiFab const &consVar_mf, amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx,
					 std::array<amrex::MultiFab, AMREX_SPACEDIM> const &faceVelArray, amrex::iMultiFab const &redoFlag_mf)
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView rhs(rhs_mf);
		MFView cons(consVar_mf);
		MFView vx(faceVelArray[0]);
		MFView vy(faceVelArray[1]);
		MFView vz(faceVelArray[2]);
		iMFView redo(redoFlag_mf);
		const double d[3] = {dx[0], dx[1], dx[2]};
		check(qk_hydro_add_internal_energy_pdv(&prm, rhs.n(), rhs.valid.data(), rhs.arr.data(), cons.arr.data(), d, vx.arr.data(), vy.arr.data(),
						       vz.arr.data(), redo.arr.data(), stream()),
		      "AddInternalEnergyPdV");
	}

	// HydroSystem::PredictStep  :475-497 (fills redoFlag; the caller sums it as the reference does, QuokkaSimulation.hpp:1146)
	static void PredictStep(amrex::MultiFab const &consVarOld, amrex::MultiFab &consVarNew, amrex::MultiFab const &rhs, double dt, int nvars,
				amrex::iMultiFab &redoFlag_mf)
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView u0(consVarOld);
		MFView u1(consVarNew);
		MFView r(rhs);
		iMFView redo(redoFlag_mf);
		check(qk_hydro_predict_step(&prm, u0.n(), u0.valid.data(), u0.arr.data(), u1.arr.data(), r.arr.data(), dt, nvars, redo.arr.data(), nullptr,
					    stream()),
		      "PredictStep");
	}

	// HydroSystem::EnforceLimits  :698-773
	static void EnforceLimits(amrex::Real densityFloor, amrex::Real tempFloor, amrex::MultiFab &state_mf)
	{
		qk_hydro_params prm = make_params<problem_t>();
		prm.density_floor = densityFloor;
		prm.temp_floor = tempFloor;
		MFView s(state_mf);
		check(qk_hydro_enforce_limits(&prm, s.n(), s.valid.data(), s.arr.data(), stream()), "EnforceLimits");
	}

	// HydroSystem::SyncDualEnergy  :816-850
	static void SyncDualEnergy(amrex::MultiFab &consVar_mf)
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView s(consVar_mf);
		int64_t nabort = 0;
		check(qk_hydro_sync_dual_energy(&prm, s.n(), s.valid.data(), s.arr.data(), &nabort, stream()), "SyncDualEnergy");
		if (nabort > 0) {
			amrex::Abort("density is negative in SyncDualEnergy! abort!!"); // hydro_system.hpp:832
		}
	}

	// HydroSystem::maxSignalSpeedLocal  :198-221 (local max; the caller reduces over ranks as the reference does)
	static auto maxSignalSpeedLocal(amrex::MultiFab const &cons) -> amrex::Real
	{
		const qk_hydro_params prm = make_params<problem_t>();
		MFView c(cons);
		double m = 0.0;
		check(qk_hydro_max_signal_speed(&prm, 1, c.n(), c.valid.data(), c.arr.data(), &m, stream()), "maxSignalSpeedLocal");
		return m;
	}

	// QuokkaSimulation::replaceFluxes for one direction  src/QuokkaSimulation.hpp:1324-1368
	template <FluxDir DIR> static void ReplaceFluxes(amrex::MultiFab &flux, amrex::MultiFab const &FOflux, amrex::iMultiFab const &redoFlag_mf, int ncomp)
	{
		MFView f(flux);
		MFView fo(FOflux);
		iMFView redo(redoFlag_mf);
		check(qk_hydro_replace_fluxes(static_cast<int>(DIR), f.n(), f.valid.data(), f.arr.data(), fo.arr.data(), redo.arr.data(), ncomp, stream()),
		      "replaceFluxes");
	}

      private:
	template <FluxDir DIR>
	static void reconstruct(int order, int limiter, amrex::MultiFab const &q_mf, amrex::MultiFab &leftState_mf, amrex::MultiFab &rightState_mf, int nghost,
				int nvars)
	{
		MFView q(q_mf);
		MFView l(leftState_mf);
		MFView r(rightState_mf);
		check(qk_reconstruct_states(order, limiter, static_cast<int>(DIR), q.n(), q.valid.data(), q.arr.data(), l.arr.data(), r.arr.data(), nghost, nvars,
					    stream()),
		      "ReconstructStates");
	}
};

// ---- radiation: RadSystem<problem_t>'s static functions take Array4s and a Box (one FAB at a time, as
// QuokkaSimulation::fluxFunction<DIR> / advanceRadiation* call them, src/QuokkaSimulation.hpp:1791-1862, 1942-1986) ----------
template <typename problem_t> inline auto make_rad_params() -> qk_rad_params
{
	qk_rad_params p{};
	p.c_light = RadSystem<problem_t>::c_light_;
	p.c_hat = RadSystem<problem_t>::c_hat_;
	p.Erad_floor = RadSystem_Traits<problem_t>::Erad_floor;
	p.ngroups = RadSystem<problem_t>::nGroups_;
	p.nstart = RadSystem<problem_t>::nstartHyperbolic_;
	p.reconstruction_order = 3; // radiationReconstructionOrder_, filled by the caller
	p.integrator_order = 2;
	return p;
}

// constant-opacity problems only (the reference's default specialisations and RadhydroShell's,
// src/radiation/radiation_system.hpp:1141-1153, src/problems/RadhydroShell/test_radhydro_shell.cpp:127-135): the opacity
// functions are sampled once; a problem whose opacities depend on rho or T keeps the reference's kernel
template <typename problem_t> inline auto make_rad_source_params() -> qk_rad_source_params
{
	qk_rad_source_params p{};
	p.radiation_constant = RadSystem<problem_t>::radiation_constant_;
	p.kappa_P = RadSystem<problem_t>::ComputePlanckOpacity(1.0, 1.0);
	p.kappa_E = RadSystem<problem_t>::ComputeEnergyMeanOpacity(1.0, 1.0);
	p.kappa_F = RadSystem<problem_t>::ComputeFluxMeanOpacity(1.0, 1.0);
	p.beta_order = RadSystem<problem_t>::beta_order_;
	p.opacity_model = QK_OPACITY_CONSTANT;
	return p;
}

template <typename problem_t> class RadSystemB200
{
      public:
	using arrayconst_t = amrex::Array4<const amrex::Real> const;
	using array_t = amrex::Array4<amrex::Real> const;

	// RadSystem::ConservedToPrimitive(cons, primVar, indexRange)  src/radiation/radiation_system.hpp:589-614
	static void ConservedToPrimitive(arrayconst_t &cons, array_t &primVar, amrex::Box const &indexRange)
	{
		const qk_rad_params prm = make_rad_params<problem_t>();
		const qk_array4 c = view(cons);
		const qk_array4 q = view(primVar);
		const qk_box bx = to_box(indexRange);
		check(qk_rad_conserved_to_primitive(&prm, 1, &bx, &c, &q, 0, stream()), "RadSystem::ConservedToPrimitive");
	}

	// RadSystem::ComputeFluxes<DIR>(x1Flux, x1FluxDiffusive, left, right, x1FluxRange, consVar, dx, use_wavespeed_correction)  :985-1139
	template <FluxDir DIR>
	static void ComputeFluxes(array_t &x1Flux_in, array_t &x1FluxDiffusive_in, arrayconst_t &x1LeftState_in, arrayconst_t &x1RightState_in,
				  amrex::Box const &indexRange, arrayconst_t &consVar_in, amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx,
				  bool const use_wavespeed_correction)
	{
		qk_rad_params prm = make_rad_params<problem_t>();
		if (use_wavespeed_correction) {
			// ComputeCellOpticalDepth (:803-871) with a CONSTANT flux-mean opacity, one photon group (what the library provides): the
			// problem's ComputeFluxMeanOpacity is sampled once; a density- or temperature-dependent opacity keeps the stock kernel
			prm.use_wavespeed_correction = 1;
			prm.kappa_F = RadSystem<problem_t>::ComputeFluxMeanOpacity(1.0, 1.0);
			for (int d = 0; d < AMREX_SPACEDIM; ++d) {
				prm.cell_dx[d] = dx[d];
			}
		}
		const qk_array4 f = view(x1Flux_in);
		const qk_array4 fd = view(x1FluxDiffusive_in);
		const qk_array4 l = view(x1LeftState_in);
		const qk_array4 r = view(x1RightState_in);
		const qk_array4 c = view(consVar_in);
		const qk_box bx = to_box(indexRange); // nodal in DIR -> enclosed cells; the library iterates over their faces
		check(qk_rad_compute_fluxes(&prm, static_cast<int>(DIR), 1, &bx, &f, &fd, &l, &r, &c, stream()), "RadSystem::ComputeFluxes");
	}

	// RadSystem::PredictStep  :667-710
	static void PredictStep(arrayconst_t &consVarOld, array_t &consVarNew, amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> fluxArray,
				amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> /*fluxDiffusiveArray*/, double dt_in,
				amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx_in, amrex::Box const &indexRange, int /*nvars*/)
	{
		static_assert(AMREX_SPACEDIM == 3, "libquokka_b200 operates on 3-D FABs");
		const qk_rad_params prm = make_rad_params<problem_t>();
		const qk_array4 u0 = view(consVarOld);
		const qk_array4 un = view(consVarNew);
		const qk_array4 fx = view(fluxArray[0]);
		const qk_array4 fy = view(fluxArray[1]);
		const qk_array4 fz = view(fluxArray[2]);
		const qk_box bx = to_box(indexRange);
		const double dx[3] = {dx_in[0], dx_in[1], dx_in[2]};
		check(qk_rad_predict_step(&prm, 1, &bx, &u0, &un, &fx, &fy, &fz, dt_in, dx, stream()), "RadSystem::PredictStep");
	}

	// RadSystem::AddFluxesRK2  :712-771
	static void AddFluxesRK2(array_t &U_new, arrayconst_t &U0, arrayconst_t &U1, amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> fluxArrayOld,
				 amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> fluxArray, amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> /*fluxDiffusiveArrayOld*/,
				 amrex::GpuArray<arrayconst_t, AMREX_SPACEDIM> /*fluxDiffusiveArray*/, double dt_in,
				 amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx_in, amrex::Box const &indexRange, int /*nvars*/)
	{
		static_assert(AMREX_SPACEDIM == 3, "libquokka_b200 operates on 3-D FABs");
		const qk_rad_params prm = make_rad_params<problem_t>();
		const qk_array4 un = view(U_new);
		const qk_array4 u0 = view(U0);
		const qk_array4 u1 = view(U1);
		const qk_array4 fo[3] = {view(fluxArrayOld[0]), view(fluxArrayOld[1]), view(fluxArrayOld[2])};
		const qk_array4 fn[3] = {view(fluxArray[0]), view(fluxArray[1]), view(fluxArray[2])};
		const qk_box bx = to_box(indexRange);
		const double dx[3] = {dx_in[0], dx_in[1], dx_in[2]};
		check(qk_rad_add_fluxes_rk2(&prm, 1, &bx, &un, &u0, &u1, &fo[0], &fo[1], &fo[2], &fn[0], &fn[1], &fn[2], dt_in, dx, stream()),
		      "RadSystem::AddFluxesRK2");
	}

	// RadSystem::AddSourceTermsSingleGroup(consVar, radEnergySource, indexRange, dt, stage, dustGasCoeff, p_iteration_counter,
	// p_iteration_failure_counter)  src/radiation/source_terms_single_group.hpp:9-565, one FAB as operatorSplitSourceTerms calls it
	// (src/QuokkaSimulation.hpp:1876).  The counters are HOST ints here (the reference passes device pointers it copies back,
	// :1620-1625,1661); nullptr keeps the call asynchronous.
	static void AddSourceTermsSingleGroup(array_t &consVar, arrayconst_t &radEnergySource, amrex::Box const &indexRange, amrex::Real dt, int stage,
					      double /*dustGasCoeff*/, int *h_iteration_counter, int *h_iteration_failure_counter)
	{
		static_assert(!RadSystem<problem_t>::enable_dust_gas_thermal_coupling_model_, "libquokka_b200: the dust-gas coupling model is not provided");
		const qk_hydro_params hp = make_params<problem_t>();
		const qk_rad_params prm = make_rad_params<problem_t>();
		const qk_rad_source_params sp = make_rad_source_params<problem_t>();
		const qk_array4 c = view(consVar);
		const qk_array4 e = view(radEnergySource);
		const qk_box bx = to_box(indexRange);
		int64_t cnt[QK_RAD_SOURCE_NCOUNTERS] = {0, 0, 0, 0, 0, 0, 0};
		const bool want = (h_iteration_counter != nullptr) || (h_iteration_failure_counter != nullptr);
		check(qk_rad_add_source_terms(&hp, &prm, &sp, stage, 1, &bx, &c, &e, dt, want ? cnt : nullptr, stream()), "RadSystem::AddSourceTermsSingleGroup");
		if (h_iteration_counter != nullptr) {
			h_iteration_counter[0] += static_cast<int>(cnt[0]);
			h_iteration_counter[1] += static_cast<int>(cnt[1]);
			h_iteration_counter[2] = std::max(h_iteration_counter[2], static_cast<int>(cnt[2]));
		}
		if (h_iteration_failure_counter != nullptr) {
			for (int n = 0; n < 3; ++n) {
				h_iteration_failure_counter[n] += static_cast<int>(cnt[4 + n]);
			}
		}
	}

	// the same for the whole level in ONE launch: replaces both MFIter loops around operatorSplitSourceTerms
	// (src/QuokkaSimulation.hpp:1631-1658); radEnergySource may be nullptr (SetRadEnergySource's default is zero, :582-587)
	static void AddSourceTermsSingleGroup(amrex::MultiFab &state, amrex::MultiFab const *radEnergySource, amrex::Real dt, int stage, int64_t *counters)
	{
		const qk_hydro_params hp = make_params<problem_t>();
		const qk_rad_params prm = make_rad_params<problem_t>();
		const qk_rad_source_params sp = make_rad_source_params<problem_t>();
		MFView s(state);
		if (radEnergySource != nullptr) {
			MFView e(*radEnergySource);
			check(qk_rad_add_source_terms(&hp, &prm, &sp, stage, s.n(), s.valid.data(), s.arr.data(), e.arr.data(), dt, counters, stream()),
			      "RadSystem::AddSourceTermsSingleGroup");
		} else {
			check(qk_rad_add_source_terms(&hp, &prm, &sp, stage, s.n(), s.valid.data(), s.arr.data(), nullptr, dt, counters, stream()),
			      "RadSystem::AddSourceTermsSingleGroup");
		}
	}
};

// ---- stage-level drop-in -----------------------------------------------------------------------------------------
// One qk_level per AMR level mirrors BoxArray + DistributionMapping + Geometry + BCRec; it is rebuilt when the grids
// change (after regrid).  advanceStage replaces the body of one RK stage of advanceHydroAtLevel
// (src/QuokkaSimulation.hpp:1099-1198 / :1202-1285): *all* of K1-K10 and the FOFC logic run inside the library.
class LevelB200
{
      public:
	LevelB200(amrex::BoxArray const &ba, amrex::DistributionMapping const &dm, amrex::Geometry const &geom, amrex::Vector<amrex::BCRec> const &bcs, int nghost,
		  int ncomp)
	{
		const int nb = static_cast<int>(ba.size());
		boxes_.resize(nb);
		owner_.resize(nb);
		for (int i = 0; i < nb; ++i) {
			boxes_[i] = to_box(ba[i]);
			owner_[i] = dm[i];
		}
		bc_lo_.resize(3 * ncomp);
		bc_hi_.resize(3 * ncomp);
		for (int n = 0; n < ncomp; ++n) {
			for (int d = 0; d < 3; ++d) {
				bc_lo_[3 * n + d] = (d < AMREX_SPACEDIM) ? bcs[n].lo(d) : QK_BC_INT_DIR;
				bc_hi_[3 * n + d] = (d < AMREX_SPACEDIM) ? bcs[n].hi(d) : QK_BC_INT_DIR;
			}
		}
		qk_level_desc d{};
		d.domain = to_box(geom.Domain());
		for (int k = 0; k < 3; ++k) {
			d.periodic[k] = (k < AMREX_SPACEDIM) ? static_cast<int>(geom.isPeriodic(k)) : 1;
			d.dx[k] = (k < AMREX_SPACEDIM) ? geom.CellSize(k) : 1.0;
		}
		d.nghost = nghost;
		d.ncomp = ncomp;
		d.nboxes_global = nb;
		d.boxes_global = boxes_.data();
		d.owner = owner_.data();
		d.my_rank = amrex::ParallelDescriptor::MyProc();
		d.bc_lo = bc_lo_.data();
		d.bc_hi = bc_hi_.data();
		check(qk_level_create(&d, &lev_), "qk_level_create");
	}
	~LevelB200() { qk_level_destroy(lev_); }
	LevelB200(LevelB200 const &) = delete;
	auto operator=(LevelB200 const &) -> LevelB200 & = delete;

	// ghost fill of level 0 / a uniform level: FillBoundary + physical BCs (simulation.hpp:1752-1765)
	void fillBoundary(amrex::MultiFab &state, int scomp, int ncomp)
	{
		MFView s(state);
		check(qk_fill_boundary(lev_, s.arr.data(), scomp, ncomp, stream()), "fillBoundary");
	}

	// returns ncells_bad after the first-order flux correction (0 = success), as redoFlag.sum() does in the reference
	auto advanceStage(qk_hydro_params const &prm, int stage, amrex::MultiFab const &U0, amrex::MultiFab const &Ustage, amrex::MultiFab &Uout, double dt)
	    -> int64_t
	{
		MFView u0(U0);
		MFView us(Ustage);
		MFView uo(Uout);
		int64_t bad = 0;
		check(qk_hydro_advance_stage(lev_, &prm, stage, u0.arr.data(), us.arr.data(), uo.arr.data(), dt, &bad, stream()), "advanceStage");
		return bad;
	}
	// the same stage for levels with flux registers (fused sweeps that also store the stage's face fluxes; the one-kernel-per-operator
	// path only when the fused kernels cannot take the stage or a cell is flagged): afterwards stageFluxes(d) are the
	// face fluxes the reference hands to incrementFluxRegisters (src/QuokkaSimulation.hpp:1195-1198: stage 1 after the FOFC
	// replacement; :1280-1283: stage 2's own F(U1))
	auto advanceStageWithFluxes(qk_hydro_params const &prm, int stage, amrex::MultiFab const &U0, amrex::MultiFab const &Ustage, amrex::MultiFab &Uout,
				    double dt) -> int64_t
	{
		MFView u0(U0);
		MFView us(Ustage);
		MFView uo(Uout);
		int64_t bad = 0;
		check(qk_hydro_advance_stage_keep_fluxes(lev_, &prm, stage, u0.arr.data(), us.arr.data(), uo.arr.data(), dt, &bad, stream()), "advanceStageWithFluxes");
		return bad;
	}
	// descriptors of the last stage's face-flux arrays in direction d, one per local box in MFIter order (nodal in d, 6 + nscalars
	// components, contiguous Fortran order: an alias amrex::FArrayBox(surroundingNodes(validbox, d), ncomp, p) views them)
	[[nodiscard]] auto stageFluxes(int d) const -> std::vector<qk_array4>
	{
		std::vector<qk_array4> out(static_cast<std::size_t>(qk_level_nlocal(lev_)));
		check(qk_level_stage_fluxes(lev_, d, out.data()), "stageFluxes");
		return out;
	}
	// one stage of the radiation transport substep (advanceRadiationForwardEuler / MidpointRK2 transport part), fused
	void advanceRadiationStage(qk_rad_params const &prm, int stage, amrex::MultiFab const &U0, amrex::MultiFab const &Ustage, amrex::MultiFab &Uout, double dt)
	{
		MFView u0(U0);
		MFView us(Ustage);
		MFView uo(Uout);
		check(qk_rad_advance_stage(lev_, &prm, stage, u0.arr.data(), us.arr.data(), uo.arr.data(), dt, stream()), "advanceRadiationStage");
	}
	// QuokkaSimulation::subcycleRadiationAtLevel (src/QuokkaSimulation.hpp:1577-1700) without flux registers: all substeps, transport
	// stages, ghost fills and (single-group, constant-opacity) source terms inside the library.  U_tmp: a MultiFab shaped like the
	// state (state_inter_cc_[lev] is free at this point).  Returns the number of substeps.
	template <typename problem_t>
	auto subcycleRadiation(qk_rad_params const &prm, amrex::MultiFab &state_old, amrex::MultiFab &state_new, amrex::MultiFab &U_tmp,
			       amrex::MultiFab const *radEnergySource, double dt_lev_hydro, double radiationCflNumber, int64_t *counters) -> int
	{
		const qk_hydro_params hp = make_params<problem_t>();
		const qk_rad_source_params sp = make_rad_source_params<problem_t>();
		MFView uo(state_old);
		MFView un(state_new);
		MFView ut(U_tmp);
		int nsub = 0;
		if (radEnergySource != nullptr) {
			MFView e(*radEnergySource);
			check(qk_rad_subcycle(lev_, &hp, &prm, &sp, uo.arr.data(), un.arr.data(), ut.arr.data(), e.arr.data(), dt_lev_hydro, radiationCflNumber,
					      counters, &nsub, stream()),
			      "subcycleRadiation");
		} else {
			check(qk_rad_subcycle(lev_, &hp, &prm, &sp, uo.arr.data(), un.arr.data(), ut.arr.data(), nullptr, dt_lev_hydro, radiationCflNumber, counters,
					      &nsub, stream()),
			      "subcycleRadiation");
		}
		return nsub;
	}
	// the same with the parameter blocks filled by the caller (arithmetic mode, floors, reconstruction order)
	auto subcycleRadiation(qk_hydro_params const &hp, qk_rad_params const &prm, qk_rad_source_params const &sp, amrex::MultiFab &state_old,
			       amrex::MultiFab &state_new, amrex::MultiFab &U_tmp, amrex::MultiFab const *radEnergySource, double dt_lev_hydro,
			       double radiationCflNumber, int64_t *counters) -> int
	{
		MFView uo(state_old);
		MFView un(state_new);
		MFView ut(U_tmp);
		std::vector<qk_array4> src;
		if (radEnergySource != nullptr) {
			src = MFView(*radEnergySource).arr;
		}
		int nsub = 0;
		check(qk_rad_subcycle(lev_, &hp, &prm, &sp, uo.arr.data(), un.arr.data(), ut.arr.data(), src.empty() ? nullptr : src.data(), dt_lev_hydro,
				      radiationCflNumber, counters, &nsub, stream()),
		      "subcycleRadiation");
		return nsub;
	}
	[[nodiscard]] auto handle() const -> qk_level * { return lev_; }

      private:
	qk_level *lev_ = nullptr;
	std::vector<qk_box> boxes_;
	std::vector<int32_t> owner_, bc_lo_, bc_hi_;
};

// ---- AMR transfer operators --------------------------------------------------------------------------------
// AMReX's own plugin interface for coarse->fine interpolation is the virtual class amrex::MFInterpolater; Quokka hands an
// instance to FillPatcher / FillPatchTwoLevels / InterpFromCoarseLevel through getAmrInterpolaterCellCentered()
// (src/simulation.hpp:1389-1407).  MFInterpB200 is a drop-in for amrex::mf_linear_slope_minmax_interp (amr_interpolation_method
// = 1): return &quokka::b200::mf_interp_b200 there.  One launch for all local FABs; no slope MultiFab.
class MFInterpB200 final : public amrex::MFInterpolater
{
      public:
	auto CoarseBox(amrex::Box const &fine, int ratio) -> amrex::Box override { return amrex::mf_linear_slope_minmax_interp.CoarseBox(fine, ratio); }
	auto CoarseBox(amrex::Box const &fine, amrex::IntVect const &ratio) -> amrex::Box override
	{
		return amrex::mf_linear_slope_minmax_interp.CoarseBox(fine, ratio);
	}
	void interp(amrex::MultiFab const &crsemf, int ccomp, amrex::MultiFab &finemf, int fcomp, int nc, amrex::IntVect const &ng, amrex::Geometry const &cgeom,
		    amrex::Geometry const & /*fgeom*/, amrex::Box const &dest_domain, amrex::IntVect const &ratio, amrex::Vector<amrex::BCRec> const &bcs,
		    int bcomp) override
	{
		static_assert(AMREX_SPACEDIM == 3, "libquokka_b200 operates on 3-D FABs");
		std::vector<qk_array4> c;
		std::vector<qk_array4> f;
		std::vector<qk_box> region;
		for (amrex::MFIter mfi(finemf); mfi.isValid(); ++mfi) {
			c.push_back(view(crsemf.const_array(mfi)));
			f.push_back(view(finemf.const_array(mfi)));
			region.push_back(to_box(amrex::grow(mfi.validbox(), ng)));
		}
		std::vector<int32_t> lo(3 * static_cast<std::size_t>(nc));
		std::vector<int32_t> hi(3 * static_cast<std::size_t>(nc));
		for (int n = 0; n < nc; ++n) {
			for (int d = 0; d < 3; ++d) {
				lo[3 * n + d] = bcs[bcomp + n].lo(d);
				hi[3 * n + d] = bcs[bcomp + n].hi(d);
			}
		}
		const qk_box dest = to_box(dest_domain);
		const qk_box cdom = to_box(cgeom.Domain());
		const int rr[3] = {ratio[0], ratio[1], ratio[2]};
		check(qk_amr_interp_cons_lin_minmax(static_cast<int>(c.size()), c.data(), ccomp, f.data(), fcomp, nc, region.data(), &dest, &cdom, rr, lo.data(),
						    hi.data(), stream()),
		      "MFInterpB200::interp");
	}
};
inline MFInterpB200 mf_interp_b200; // NOLINT(cppcoreguidelines-avoid-non-const-global-variables): mirrors amrex::mf_linear_slope_minmax_interp

// QuokkaSimulation::PreInterpState / PostInterpState (src/QuokkaSimulation.hpp:804-841) with the hooks' own signature, so that they can
// be passed to fillBoundaryConditions / FillPatchWithData in their place (:798,1076): one launch for all FABs of mf
inline void PreInterpStateB200(amrex::MultiFab &mf, int /*scomp*/, int /*ncomp*/)
{
	MFView s(mf);
	check(qk_amr_pre_interp_state(s.n(), s.valid.data(), s.arr.data(), stream()), "PreInterpState");
}
inline void PostInterpStateB200(amrex::MultiFab &mf, int /*scomp*/, int /*ncomp*/)
{
	MFView s(mf);
	check(qk_amr_post_interp_state(s.n(), s.valid.data(), s.arr.data(), stream()), "PostInterpState");
}

// amrex::average_down(S_fine, S_crse, scomp, ncomp, ratio) for fine and coarse MultiFabs with matching (coarsened) BoxArrays, as
// AverageDownTo builds them (src/simulation.hpp:1309-1343 averages into a coarsened copy and ParallelCopies it).
inline void average_down_b200(amrex::MultiFab const &S_fine, amrex::MultiFab &S_crse_on_fine_layout, int scomp, int ncomp, amrex::IntVect const &ratio)
{
	MFView f(S_fine);
	MFView c(S_crse_on_fine_layout);
	const int rr[3] = {ratio[0], ratio[1], ratio[2]};
	check(qk_amr_average_down(c.n(), c.arr.data(), scomp, f.arr.data(), scomp, ncomp, c.valid.data(), rr, stream()), "average_down_b200");
}

} // namespace quokka::b200
