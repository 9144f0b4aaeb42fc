// oracle/ref_build/shell_golden.cpp -- TEST INFRASTRUCTURE (written for this repo, not reference source).
//
// The reference's RadhydroShell problem (config C4) hard-codes `plotfileInterval_ = -1` and `maxTimesteps_ = 50` AFTER its
// simulation object has read the inputs file (src/problems/RadhydroShell/test_radhydro_shell.cpp:426-431), so the stock
// executable cannot write the state it computes.  This translation unit compiles the reference's own problem file where it
// lies -- traits, opacities, energy source, initial conditions, all untouched -- and supplies a problem_main() that sets the
// same run-time parameters (:396-424: cfl 0.3, density floor 1e-8 rho_0, PLM for hydro and radiation, RK2, periodic box) but
// leaves the step count and the plotfile interval to the inputs file.  Used only by tests/golden/make_golden_shell.py.
#define problem_main reference_problem_main_unused
#include "problems/RadhydroShell/test_radhydro_shell.cpp"
#undef problem_main

auto problem_main() -> int
{
	const int ncomp_cc = Physics_Indices<ShellProblem>::nvarTotal_cc;
	amrex::Vector<amrex::BCRec> BCs_cc(ncomp_cc);
	for (int n = 0; n < ncomp_cc; ++n) {
		for (int i = 0; i < AMREX_SPACEDIM; ++i) {
			BCs_cc[n].setLo(i, amrex::BCType::int_dir); // simulate_full_box: periodic
			BCs_cc[n].setHi(i, amrex::BCType::int_dir);
		}
	}
	QuokkaSimulation<ShellProblem> sim(BCs_cc); // reads max_timesteps, plotfile_interval, ... from the inputs file
	sim.cflNumber_ = 0.3;
	sim.densityFloor_ = 1.0e-8 * rho_0;
	sim.pressureFloor_ = 1.0e-8 * P_0;
	sim.reconstructionOrder_ = 2;
	sim.radiationReconstructionOrder_ = 2;
	sim.integratorOrder_ = 2;
	sim.stopTime_ = 0.125 * (r_0 / a0);
	sim.checkpointInterval_ = -1;
	sim.setInitialConditions();
	sim.computeAfterTimestep();
	sim.evolve();
	return 0;
}
