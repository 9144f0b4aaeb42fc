// oracle/ref_build/errorest_harness.cpp -- TEST INFRASTRUCTURE (written for this repo, not reference source).
//
// Pins the oracle's regrid-tagging restatement to the reference itself: this translation unit compiles the reference's own
// Sedov problem file where it lies (traits, initial condition, ErrorEst specialisation, all untouched; its problem_main is renamed away),
// builds a QuokkaSimulation<SedovProblem> on one box, overwrites the level-0 state with caller-provided data and calls the reference's
// QuokkaSimulation<SedovProblem>::ErrorEst (src/problems/HydroBlast3D/test_hydro3d_blast.cpp:118-151) and FixupState
// (src/QuokkaSimulation.hpp:761-770) on it.  Built into oracle/_ref/libquokka_ref_sedov.so; used by tests/test_oracle_regrid_vs_ref.py.
#include <cstdint>
#include <cstring>
#include <memory>

#define problem_main reference_problem_main_unused
#include "problems/HydroBlast3D/test_hydro3d_blast.cpp"
#undef problem_main

#include "AMReX_TagBox.H"

#include "../../include/quokka_b200.h"

namespace
{
struct Probe : public QuokkaSimulation<SedovProblem> {
	using QuokkaSimulation<SedovProblem>::QuokkaSimulation;
	auto state(int lev) -> amrex::MultiFab & { return state_new_cc_[lev]; }
	void fixup(int lev) { FixupState(lev); }
};
std::unique_ptr<Probe> g_sim;
int g_n = 0;

void ensure(int n)
{
	if (g_sim && g_n == n) {
		return;
	}
	if (!g_sim) {
		int argc = 1;
		static char arg0[] = "errorest_harness";
		static char *argv_s[] = {arg0, nullptr};
		char **argv = argv_s;
		amrex::Initialize(argc, argv, false, MPI_COMM_WORLD, [n]() {
			amrex::ParmParse pa("amrex");
			pa.add("verbose", 0);
			pa.add("signal_handling", 0);
			amrex::ParmParse pg("geometry");
			pg.addarr("prob_lo", std::vector<amrex::Real>{0., 0., 0.});
			pg.addarr("prob_hi", std::vector<amrex::Real>{1.2, 1.2, 1.2});
			pg.addarr("is_periodic", std::vector<int>{0, 0, 0});
			amrex::ParmParse pr("amr");
			pr.add("v", 0);
			pr.addarr("n_cell", std::vector<int>{n, n, n});
			pr.add("max_level", 0);
			pr.add("max_grid_size", n);
			pr.add("blocking_factor", n);
			amrex::ParmParse pp;
			pp.add("do_reflux", 0);
			pp.add("do_subcycle", 0);
			pp.add("plotfile_interval", -1);
			pp.add("checkpoint_interval", -1);
		});
	} else {
		return; // one grid size per process: the AmrCore geometry is fixed by the first call
	}
	const int ncomp_cc = Physics_Indices<SedovProblem>::nvarTotal_cc;
	amrex::Vector<amrex::BCRec> BCs_cc(ncomp_cc);
	for (int nn = 0; nn < ncomp_cc; ++nn) {
		for (int i = 0; i < AMREX_SPACEDIM; ++i) {
			BCs_cc[nn].setLo(i, amrex::BCType::reflect_even);
			BCs_cc[nn].setHi(i, amrex::BCType::reflect_even);
		}
	}
	g_sim = std::make_unique<Probe>(BCs_cc);
	g_sim->setInitialConditions();
	g_n = n;
}

// state: 6 components on the n^3 box grown by `ng` >= 1 cells (caller's ghost values are used as they are)
void copy_in(amrex::MultiFab &mf, const qk_array4 *src, bool ghosts)
{
	for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi) {
		auto const &a = mf.array(mfi);
		const amrex::Box bx = ghosts ? amrex::grow(mfi.validbox(), 1) : mfi.validbox();
		for (int n = 0; n < 6; ++n) {
			for (int k = bx.smallEnd(2); k <= bx.bigEnd(2); ++k) {
				for (int j = bx.smallEnd(1); j <= bx.bigEnd(1); ++j) {
					for (int i = bx.smallEnd(0); i <= bx.bigEnd(0); ++i) {
						a(i, j, k, n) = src->p[(i - src->begin[0]) + (j - src->begin[1]) * src->jstride + (k - src->begin[2]) * src->kstride + n * src->nstride];
					}
				}
			}
		}
	}
}
} // namespace

extern "C" {
// tags_out: n^3 chars (x fastest), TagBox::CLEAR / SET as the reference leaves them
int ref_sedov_error_est(int n, const qk_array4 *cons, char *tags_out)
{
	ensure(n);
	if (g_n != n) {
		return -1;
	}
	amrex::MultiFab &S = g_sim->state(0);
	copy_in(S, cons, true);
	amrex::TagBoxArray tags(S.boxArray(), S.DistributionMap(), 0);
	tags.setVal(amrex::TagBox::CLEAR);
	g_sim->ErrorEst(0, tags, 0.0, 0);
	for (amrex::MFIter mfi(tags); mfi.isValid(); ++mfi) {
		auto const &t = tags.const_array(mfi);
		const amrex::Box bx = mfi.validbox();
		for (int k = 0; k < n; ++k) {
			for (int j = 0; j < n; ++j) {
				for (int i = 0; i < n; ++i) {
					tags_out[i + n * (j + static_cast<int64_t>(n) * k)] = t(bx.smallEnd(0) + i, bx.smallEnd(1) + j, bx.smallEnd(2) + k);
				}
			}
		}
	}
	return 0;
}
// QuokkaSimulation::FixupState on the valid cells, in place
int ref_sedov_fixup_state(int n, const qk_array4 *cons)
{
	ensure(n);
	if (g_n != n) {
		return -1;
	}
	amrex::MultiFab &S = g_sim->state(0);
	copy_in(S, cons, false);
	g_sim->fixup(0);
	for (amrex::MFIter mfi(S); mfi.isValid(); ++mfi) {
		auto const &a = S.const_array(mfi);
		const amrex::Box bx = mfi.validbox();
		for (int nn = 0; nn < 6; ++nn) {
			for (int k = bx.smallEnd(2); k <= bx.bigEnd(2); ++k) {
				for (int j = bx.smallEnd(1); j <= bx.bigEnd(1); ++j) {
					for (int i = bx.smallEnd(0); i <= bx.bigEnd(0); ++i) {
						cons->p[(i - cons->begin[0]) + (j - cons->begin[1]) * cons->jstride + (k - cons->begin[2]) * cons->kstride + nn * cons->nstride] =
						    a(i, j, k, nn);
					}
				}
			}
		}
	}
	return 0;
}
}
