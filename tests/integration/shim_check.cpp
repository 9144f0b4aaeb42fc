// Compile-only check (g++ -fsyntax-only against the reference's own headers under /root/reference): every static function of
// HydroSystemB200<problem_t> instantiates with the argument lists QuokkaSimulation<problem_t> uses for HydroSystem<problem_t>.
#include "quokka_b200_amrex.hpp"

struct SedovLike {
};
template <> struct quokka::EOS_Traits<SedovLike> {
	static constexpr double gamma = 1.4;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct HydroSystem_Traits<SedovLike> {
	static constexpr bool reconstruct_eint = false;
};
template <> struct Physics_Traits<SedovLike> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

// the call sequence of QuokkaSimulation::hydroFluxFunction / advanceHydroAtLevel (src/QuokkaSimulation.hpp:1403-1517, 1096-1198)
template <typename Hydro> void stage(amrex::MultiFab &state_old, amrex::MultiFab &state_new, amrex::iMultiFab &redoFlag, amrex::Real dt,
				     amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx)
{
	const int ng = 4;
	const int nvars = Hydro::nvar_;
	auto const &ba = state_old.boxArray();
	auto const &dm = state_old.DistributionMap();
	amrex::MultiFab prim(ba, dm, nvars, ng);
	std::array<amrex::MultiFab, 3> chi, flux, fvel;
	amrex::MultiFab left, right, rhs(ba, dm, nvars, 0);
	Hydro::ConservedToPrimitive(state_old, prim, ng);
	Hydro::template ComputeFlatteningCoefficients<FluxDir::X1>(prim, chi[0], 2);
	Hydro::template ComputeFlatteningCoefficients<FluxDir::X2>(prim, chi[1], 2);
	Hydro::template ComputeFlatteningCoefficients<FluxDir::X3>(prim, chi[2], 2);
	Hydro::template ReconstructStatesPPM<FluxDir::X1>(prim, left, right, 1, nvars);
	Hydro::template ReconstructStatesPLM<FluxDir::X2, SlopeLimiter::minmod>(prim, left, right, 1, nvars);
	Hydro::template ReconstructStatesConstant<FluxDir::X3>(prim, left, right, 1, nvars);
	Hydro::template FlattenShocks<FluxDir::X1>(prim, chi[0], chi[1], chi[2], left, right, 1, nvars);
	Hydro::template ComputeFluxes<RiemannSolver::HLLC, FluxDir::X1>(flux[0], fvel[0], left, right, prim, 0.0);
	Hydro::template ComputeFluxes<RiemannSolver::LLF, FluxDir::X2>(flux[1], fvel[1], left, right, prim, 0.0);
	Hydro::ComputeRhsFromFluxes(rhs, flux, dx, nvars);
	Hydro::AddInternalEnergyPdV(rhs, state_old, dx, fvel, redoFlag);
	Hydro::PredictStep(state_old, state_new, rhs, dt, nvars, redoFlag);
	Hydro::EnforceLimits(0.0, 0.0, state_new);
	Hydro::SyncDualEnergy(state_new);
}

template void stage<HydroSystem<SedovLike>>(amrex::MultiFab &, amrex::MultiFab &, amrex::iMultiFab &, amrex::Real, amrex::GpuArray<amrex::Real, AMREX_SPACEDIM>);
template void stage<quokka::b200::HydroSystemB200<SedovLike>>(amrex::MultiFab &, amrex::MultiFab &, amrex::iMultiFab &, amrex::Real,
							      amrex::GpuArray<amrex::Real, AMREX_SPACEDIM>);

auto stage_level(quokka::b200::LevelB200 &lev, amrex::MultiFab &U0, amrex::MultiFab &U1, amrex::MultiFab &Unew, double dt) -> int64_t
{
	qk_hydro_params prm = quokka::b200::make_params<SedovLike>();
	lev.fillBoundary(U0, 0, U0.nComp());
	int64_t bad = lev.advanceStage(prm, 1, U0, U0, U1, dt);
	lev.fillBoundary(U1, 0, U1.nComp());
	bad += lev.advanceStage(prm, 2, U0, U1, Unew, dt);
	return bad;
}

// ---- radiation: the call sequence of QuokkaSimulation::fluxFunction<DIR> + advanceRadiation* on one FAB ----
struct RadLike {
};
template <> struct quokka::EOS_Traits<RadLike> {
	static constexpr double gamma = 5. / 3.;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<RadLike> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = true;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};
template <> struct RadSystem_Traits<RadLike> {
	static constexpr double c_light = 1.0;
	static constexpr double c_hat = 1.0;
	static constexpr double radiation_constant = 1.0;
	static constexpr double Erad_floor = 0.;
	static constexpr int beta_order = 0;
};

template <typename Rad>
void rad_box(amrex::FArrayBox &cons, amrex::FArrayBox &prim, amrex::FArrayBox &l, amrex::FArrayBox &r, std::array<amrex::FArrayBox, 3> &f,
	     std::array<amrex::FArrayBox, 3> &fd, amrex::FArrayBox &unew, amrex::FArrayBox &u1, amrex::Box const &bx, double dt,
	     amrex::GpuArray<amrex::Real, AMREX_SPACEDIM> dx)
{
	Rad::ConservedToPrimitive(cons.const_array(), prim.array(), amrex::grow(bx, 4));
	Rad::template ComputeFluxes<FluxDir::X1>(f[0].array(), fd[0].array(), l.const_array(), r.const_array(), amrex::surroundingNodes(bx, 0), cons.const_array(), dx,
						 false);
	Rad::template ComputeFluxes<FluxDir::X3>(f[2].array(), fd[2].array(), l.const_array(), r.const_array(), amrex::surroundingNodes(bx, 2), cons.const_array(), dx,
						 false);
	Rad::PredictStep(cons.const_array(), unew.array(), {f[0].const_array(), f[1].const_array(), f[2].const_array()},
			 {fd[0].const_array(), fd[1].const_array(), fd[2].const_array()}, dt, dx, bx, 4);
	Rad::AddFluxesRK2(unew.array(), cons.const_array(), u1.const_array(), {f[0].const_array(), f[1].const_array(), f[2].const_array()},
			  {f[0].const_array(), f[1].const_array(), f[2].const_array()}, {fd[0].const_array(), fd[1].const_array(), fd[2].const_array()},
			  {fd[0].const_array(), fd[1].const_array(), fd[2].const_array()}, dt, dx, bx, 4);
}
template void rad_box<RadSystem<RadLike>>(amrex::FArrayBox &, amrex::FArrayBox &, amrex::FArrayBox &, amrex::FArrayBox &, std::array<amrex::FArrayBox, 3> &,
					  std::array<amrex::FArrayBox, 3> &, amrex::FArrayBox &, amrex::FArrayBox &, amrex::Box const &, double,
					  amrex::GpuArray<amrex::Real, AMREX_SPACEDIM>);
template void rad_box<quokka::b200::RadSystemB200<RadLike>>(amrex::FArrayBox &, amrex::FArrayBox &, amrex::FArrayBox &, amrex::FArrayBox &,
							    std::array<amrex::FArrayBox, 3> &, std::array<amrex::FArrayBox, 3> &, amrex::FArrayBox &,
							    amrex::FArrayBox &, amrex::Box const &, double, amrex::GpuArray<amrex::Real, AMREX_SPACEDIM>);

// matter-radiation source terms: the reference's Array4 call (QuokkaSimulation.hpp:1876) and the level-wide form
template <typename Rad> void rad_source_box(amrex::FArrayBox &cons, amrex::FArrayBox &esrc, amrex::Box const &bx, double dt, int *it, int *fail)
{
	Rad::AddSourceTermsSingleGroup(cons.array(), esrc.const_array(), bx, dt, 1, 0.0, it, fail);
	Rad::AddSourceTermsSingleGroup(cons.array(), esrc.const_array(), bx, dt, 2, 0.0, it, fail);
}
template void rad_source_box<RadSystem<RadLike>>(amrex::FArrayBox &, amrex::FArrayBox &, amrex::Box const &, double, int *, int *);
template void rad_source_box<quokka::b200::RadSystemB200<RadLike>>(amrex::FArrayBox &, amrex::FArrayBox &, amrex::Box const &, double, int *, int *);
void rad_source_level(amrex::MultiFab &state, double dt, int64_t *counters)
{
	quokka::b200::RadSystemB200<RadLike>::AddSourceTermsSingleGroup(state, nullptr, dt, 1, counters);
	quokka::b200::RadSystemB200<RadLike>::AddSourceTermsSingleGroup(state, nullptr, dt, 2, counters);
}

// AMR: the interpolater is an amrex::MFInterpolater, i.e. it can be handed to AMReX wherever mf_linear_slope_minmax_interp is
amrex::MFInterpolater *amr_mapper(bool ours) { return ours ? static_cast<amrex::MFInterpolater *>(&quokka::b200::mf_interp_b200) : &amrex::mf_linear_slope_minmax_interp; }
void amr_average_down(amrex::MultiFab const &fine, amrex::MultiFab &crse, amrex::IntVect const &ratio) { quokka::b200::average_down_b200(fine, crse, 0, 6, ratio); }

void amr_hooks(amrex::MultiFab &mf)
{
	void (*pre)(amrex::MultiFab &, int, int) = &quokka::b200::PreInterpStateB200; // the type of QuokkaSimulation<P>::PreInterpState
	void (*post)(amrex::MultiFab &, int, int) = &quokka::b200::PostInterpStateB200;
	pre(mf, 0, 6);
	post(mf, 0, 6);
}

int rad_subcycle(quokka::b200::LevelB200 &lev, amrex::MultiFab &Uold, amrex::MultiFab &Unew, amrex::MultiFab &Utmp, double dt_hydro, int64_t *counters)
{
	qk_rad_params prm = quokka::b200::make_rad_params<RadLike>();
	return lev.subcycleRadiation<RadLike>(prm, Uold, Unew, Utmp, nullptr, dt_hydro, 0.3, counters);
}

void rad_level(quokka::b200::LevelB200 &lev, amrex::MultiFab &U0, amrex::MultiFab &U1, amrex::MultiFab &Unew, double dt)
{
	qk_rad_params prm = quokka::b200::make_rad_params<RadLike>();
	lev.fillBoundary(U0, 0, U0.nComp());
	lev.advanceRadiationStage(prm, 1, U0, U0, U1, dt);
	lev.fillBoundary(U1, 0, U1.nComp());
	lev.advanceRadiationStage(prm, 2, U0, U1, Unew, dt);
}
