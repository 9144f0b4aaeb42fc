// quokka_b200_driver.hpp -- the stage-level drop-in EXECUTED inside the reference's own time-step loop.
//
// `oracle/ref_build/apply_b200_patch.py` writes a patched copy of src/QuokkaSimulation.hpp whose only changes are
//   (1) `#include "quokka_b200_driver.hpp"` after the reference's own includes                  [part 1: DriverState]
//   (2) three member declarations + one `quokka::b200::DriverState b200_` member in class QuokkaSimulation
//   (3) a two-line hook at the top of advanceHydroAtLevel (src/QuokkaSimulation.hpp:1032) and of
//       subcycleRadiationAtLevel (:1577) that forwards to the member functions defined HERE
//   (4) `#define QUOKKA_B200_DRIVER_IMPL` + a second include at the end of the header            [part 2: definitions]
// Everything else of the reference -- AMRSimulation::evolve, computeTimestep, timeStepWithSubcycling, FillPatcher /
// YAFluxRegister / regrid / plotfiles, the problem file -- runs unmodified around it.  Run-time switches (ParmParse):
//   b200.enabled = 1|0      (0: the stock path, same executable)
//   b200.arith   = exact|relaxed
//   b200.fill    = 1|0      (1: level-0 ghost fill by qk_fill_boundary when no ext_dir BC is present; 0: always AMReX's)
//   b200.radiation = 1|0 (0: radiation subcycle stays the reference's; needed when SetRadEnergySource depends on time)
//   b200.fused_amr = 1|0    (1: levels with flux registers use the fused sweeps + captured face fluxes when available)
//
// Written for this repository; it names the reference's members because that is the boundary.
#ifndef QUOKKA_B200_DRIVER_STATE_HPP_
#define QUOKKA_B200_DRIVER_STATE_HPP_

#include <memory>
#include <string>
#include <vector>

#include "AMReX_ParmParse.H"
#include "AMReX_YAFluxRegister.H"
#include "quokka_b200_amrex.hpp"

namespace quokka::b200
{
struct DriverState {
	struct Slot {
		amrex::BoxArray ba;
		amrex::DistributionMapping dm;
		std::unique_ptr<LevelB200> lev;
	};
	std::vector<Slot> slots;
	int parsed = 0;
	int enabled = 1;
	int arith = QK_ARITH_EXACT;
	int lib_fill = 1;
	int fused_amr = 1;
	int radiation = 1;
	int64_t stages_fused = 0, stages_faithful = 0, rad_subcycles = 0;

	void parse()
	{
		if (parsed != 0) {
			return;
		}
		parsed = 1;
		amrex::ParmParse pp("b200");
		pp.query("enabled", enabled);
		std::string a = "exact";
		pp.query("arith", a);
		arith = (a == "relaxed" || a == "fast") ? QK_ARITH_FAST : QK_ARITH_EXACT;
		pp.query("fill", lib_fill);
		pp.query("fused_amr", fused_amr);
		pp.query("radiation", radiation);
		amrex::Print() << "[b200] libquokka_b200 driver: enabled = " << enabled << ", arith = " << (arith == QK_ARITH_FAST ? "relaxed" : "exact")
			       << ", fill = " << lib_fill << "\n";
	}
	[[nodiscard]] auto on() -> bool
	{
		parse();
		return enabled != 0;
	}
	// one qk_level per AMR level, rebuilt when the grids change (regrid): BoxArray / DistributionMapping equality is a cheap
	// reference comparison in AMReX when nothing changed
	auto level(int lev, amrex::BoxArray const &ba, amrex::DistributionMapping const &dm, amrex::Geometry const &geom, amrex::Vector<amrex::BCRec> const &bcs,
		   int nghost, int ncomp) -> LevelB200 &
	{
		if (static_cast<int>(slots.size()) <= lev) {
			slots.resize(lev + 1);
		}
		Slot &s = slots[lev];
		if (!s.lev || !(s.ba == ba) || !(s.dm == dm)) {
			s.lev.reset(); // free the old level's scratch before the new one allocates
			s.lev = std::make_unique<LevelB200>(ba, dm, geom, bcs, nghost, ncomp);
			s.ba = ba;
			s.dm = dm;
		}
		return *s.lev;
	}
};
} // namespace quokka::b200
#endif // QUOKKA_B200_DRIVER_STATE_HPP_

#if defined(QUOKKA_B200_DRIVER_IMPL) && !defined(QUOKKA_B200_DRIVER_IMPL_DONE)
#define QUOKKA_B200_DRIVER_IMPL_DONE

// can the library fill this level's ghost cells itself?  level 0 only (finer levels need FillPatcher), and no user-defined
// (ext_dir) boundary: those cells are written by the problem's setCustomBoundaryConditions functor
template <typename problem_t> auto QuokkaSimulation<problem_t>::b200CanFill(int lev) -> bool
{
	if (lev != 0 || b200_.lib_fill == 0) {
		return false;
	}
	for (auto const &bc : BCs_cc_) {
		for (int d = 0; d < AMREX_SPACEDIM; ++d) {
			if (bc.lo(d) == amrex::BCType::ext_dir || bc.hi(d) == amrex::BCType::ext_dir) {
				return false;
			}
		}
	}
	return true;
}

// QuokkaSimulation::advanceHydroAtLevel (src/QuokkaSimulation.hpp:1032-1322) with both RK stages inside libquokka_b200.
// Kept from the reference, in its order: Strang-split sources (:1047,1318), ghost fill of the old and of the intermediate
// state (:1076,1204; AMReX's FillPatcher on refined levels), FOFC failure -> `return false` (:1171-1184,1258-1270), flux
// register increments with the STAGE's fluxes scaled by fluxScaleFactor * dt (:1195-1198,1280-1283; stage 1 after FOFC
// replacement, stage 2 the raw F(U1) exactly as the reference passes `fluxArrays`), forward-Euler copy (:1285), isCflViolated.
template <typename problem_t>
auto QuokkaSimulation<problem_t>::advanceHydroAtLevelB200(amrex::MultiFab &state_old_cc_tmp, amrex::YAFluxRegister *fr_as_crse, amrex::YAFluxRegister *fr_as_fine,
							  int lev, amrex::Real time, amrex::Real dt_lev) -> bool
{
	BL_PROFILE("QuokkaSimulation::advanceHydroAtLevelB200()");
	namespace b2 = quokka::b200;

	const amrex::Real fluxScaleFactor = (integratorOrder_ == 2) ? 0.5 : 1.0;

	if (!addStrangSplitSourcesWithBuiltin(state_old_cc_tmp, lev, time, 0.5 * dt_lev)) {
		return false;
	}

	amrex::MultiFab state_inter_cc_(grids[lev], dmap[lev], Physics_Indices<problem_t>::nvarTotal_cc, nghost_cc_);
	state_inter_cc_.setVal(0);

	b2::LevelB200 &L = b200_.level(lev, grids[lev], dmap[lev], geom[lev], BCs_cc_, nghost_cc_, Physics_Indices<problem_t>::nvarTotal_cc);
	qk_hydro_params prm = b2::make_params<problem_t>();
	prm.density_floor = densityFloor_;
	prm.temp_floor = tempFloor_;
	prm.K_visc = artificialViscosityK_;
	prm.reconstruction_order = reconstructionOrder_;
	prm.use_dual_energy = useDualEnergy_;
	prm.integrator_order = integratorOrder_;
	prm.abort_on_fofc_failure = abortOnFofcFailure_;
	prm.arith = b200_.arith;

	const bool need_fluxes = (do_reflux == 1) && (fr_as_crse != nullptr || fr_as_fine != nullptr);
	const bool lib_fill = b200CanFill(lev);
	auto fill = [&](amrex::MultiFab &mf, amrex::Real t) {
		if (lib_fill) {
			L.fillBoundary(mf, 0, mf.nComp());
		} else {
			fillBoundaryConditions(mf, mf, lev, t, quokka::centering::cc, quokka::direction::na, PreInterpState, PostInterpState);
		}
	};
	// YAFluxRegister::CrseAdd / FineAdd read the face fluxes of one box through FArrayBox pointers (simulation.hpp:1345-1365):
	// alias FABs over the library's flux scratch, no copy
	auto reflux = [&](amrex::Real dt_flux) {
		std::array<std::vector<qk_array4>, 3> fl;
		for (int d = 0; d < 3; ++d) {
			fl[d] = L.stageFluxes(d);
		}
		int li = 0;
		for (amrex::MFIter mfi(state_new_cc_[lev]); mfi.isValid(); ++mfi, ++li) {
			std::array<amrex::FArrayBox, AMREX_SPACEDIM> fabs;
			for (int d = 0; d < AMREX_SPACEDIM; ++d) {
				fabs[d] = amrex::FArrayBox(amrex::surroundingNodes(mfi.validbox(), d), ncompHydro_, fl[d][li].p);
			}
			incrementFluxRegisters(mfi, fr_as_crse, fr_as_fine, fabs, lev, dt_flux);
		}
	};
	auto stage = [&](int s, amrex::MultiFab const &Ustage, amrex::MultiFab &Uout) -> int64_t {
		if (need_fluxes) {
			++b200_.stages_faithful;
			return L.advanceStageWithFluxes(prm, s, state_old_cc_tmp, Ustage, Uout, dt_lev);
		}
		++b200_.stages_fused;
		return L.advanceStage(prm, s, state_old_cc_tmp, Ustage, Uout, dt_lev);
	};

	// Stage 1 of RK2-SSP
	fill(state_old_cc_tmp, time);
	{
		const int64_t ncells_bad = stage(1, state_old_cc_tmp, state_inter_cc_);
		if (ncells_bad > 0) {
			if (Verbose()) {
				amrex::Print() << "[FOFC-1] failed for " << ncells_bad << " cells on level " << lev << "\n";
			}
			if (abortOnFofcFailure_ != 0) {
				return false;
			}
		}
		if (need_fluxes) {
			reflux(fluxScaleFactor * dt_lev);
		}
	}
	amrex::Gpu::streamSynchronizeAll();

	// Stage 2 of RK2-SSP
	if (integratorOrder_ == 2) {
		fill(state_inter_cc_, time + dt_lev);
		const int64_t ncells_bad = stage(2, state_inter_cc_, state_new_cc_[lev]);
		if (ncells_bad > 0) {
			if (Verbose()) {
				amrex::Print() << "[FOFC-2] failed for " << ncells_bad << " cells on level " << lev << "\n";
			}
			if (abortOnFofcFailure_ != 0) {
				return false;
			}
		}
		if (need_fluxes) {
			reflux(fluxScaleFactor * dt_lev);
		}
	} else {
		amrex::Copy(state_new_cc_[lev], state_inter_cc_, 0, 0, ncompHydro_, 0);
	}
	amrex::Gpu::streamSynchronizeAll();

	auto burn_success_second = addStrangSplitSourcesWithBuiltin(state_new_cc_[lev], lev, time + dt_lev, 0.5 * dt_lev);
	return (!isCflViolated(lev, time, dt_lev) && burn_success_second);
}

// QuokkaSimulation::subcycleRadiationAtLevel (src/QuokkaSimulation.hpp:1577-1720) for a level without flux registers: substep
// count, swapRadiationState, both transport stages, both source-term solves and the ghost fills inside qk_rad_subcycle; the
// reference's assertions and its three convergence aborts are kept.
template <typename problem_t> void QuokkaSimulation<problem_t>::subcycleRadiationAtLevelB200(int lev, amrex::Real time, amrex::Real dt_lev_hydro)
{
	BL_PROFILE("QuokkaSimulation::subcycleRadiationAtLevelB200()");
	namespace b2 = quokka::b200;
	if constexpr (Physics_Traits<problem_t>::is_radiation_enabled) {
		b2::LevelB200 &L = b200_.level(lev, grids[lev], dmap[lev], geom[lev], BCs_cc_, nghost_cc_, Physics_Indices<problem_t>::nvarTotal_cc);
		qk_hydro_params hp = b2::make_params<problem_t>();
		hp.density_floor = densityFloor_;
		hp.temp_floor = tempFloor_;
		hp.arith = b200_.arith;
		qk_rad_params rp = b2::make_rad_params<problem_t>();
		rp.reconstruction_order = radiationReconstructionOrder_;
		rp.integrator_order = 2;
		if (use_wavespeed_correction_) { // radiation.use_wavespeed_correction (src/QuokkaSimulation.hpp:133,1960): constant flux-mean opacity
			rp.use_wavespeed_correction = 1;
			rp.kappa_F = RadSystem<problem_t>::ComputeFluxMeanOpacity(1.0, 1.0);
		}
		const qk_rad_source_params sp = b2::make_rad_source_params<problem_t>();
		amrex::MultiFab U_tmp(grids[lev], dmap[lev], Physics_Indices<problem_t>::nvarTotal_cc, nghost_cc_);
		// operatorSplitSourceTerms (:1860-1885) evaluates RadSystem::SetRadEnergySource box by box before every solve, at time_subcycle +
		// dt_radiation.  Here it is evaluated ONCE per coarse step (at the first substep's time) into a one-component MultiFab that all
		// substeps read: exact for time-independent sources (RadhydroShell's Gaussian star, the reference's default of zero); a problem
		// whose source depends on time keeps the stock path (b200.radiation = 0).
		amrex::MultiFab radEnergySource(grids[lev], dmap[lev], 1, 0);
		radEnergySource.setVal(0.);
		{
			const int nsub_expected = computeNumberOfRadiationSubsteps(lev, dt_lev_hydro);
			const amrex::Real dt_radiation = dt_lev_hydro / static_cast<double>(nsub_expected);
			auto const &dx = geom[lev].CellSizeArray();
			auto const &prob_lo = geom[lev].ProbLoArray();
			auto const &prob_hi = geom[lev].ProbHiArray();
			for (amrex::MFIter iter(radEnergySource); iter.isValid(); ++iter) {
				RadSystem<problem_t>::SetRadEnergySource(radEnergySource.array(iter), iter.validbox(), dx, prob_lo, prob_hi, time + dt_radiation);
			}
		}
		int64_t counters[QK_RAD_SOURCE_NCOUNTERS] = {};
		const int nsubSteps =
		    L.subcycleRadiation(hp, rp, sp, state_old_cc_[lev], state_new_cc_[lev], U_tmp, &radEnergySource, dt_lev_hydro, radiationCflNumber_, counters);
		if (Verbose() != 0) {
			amrex::Print() << "\tRadiation substeps: " << nsubSteps << "\tdt: " << dt_lev_hydro / nsubSteps << "\n";
		}
		AMREX_ALWAYS_ASSERT(nsubSteps >= 1);
		AMREX_ALWAYS_ASSERT(nsubSteps <= (maxSubsteps_ + 1));
		long nf_coupling = counters[4];
		long nf_outer = counters[6];
		amrex::ParallelDescriptor::ReduceLongSum(nf_coupling);
		amrex::ParallelDescriptor::ReduceLongSum(nf_outer);
		if (nf_coupling > 0) {
			amrex::Abort("Newton-Raphson iteration for matter-radiation coupling failed to converge!");
		}
		if (nf_outer > 0) {
			amrex::Abort("Outer iteration for matter-radiation coupling failed to converge!");
		}
		radiationCellUpdates_ += static_cast<amrex::Long>(nsubSteps) * CountCells(lev);
		++b200_.rad_subcycles;
	}
}

#endif // QUOKKA_B200_DRIVER_IMPL
