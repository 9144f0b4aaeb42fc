// qk_sweep_keepf.cu -- the TMA-staged fused stage kernels (exact arithmetic, --fmad=false) instantiated with KEEPF = true: the same sweeps,
// which additionally store the stage's own face fluxes F (SweepBox::fo, tight nodal arrays).  Levels with flux registers (AMR with
// do_reflux) take this path: incrementFluxRegisters (src/simulation.hpp:1345-1387) reads stage 1's F(U0) and stage 2's F(U1)
// (src/QuokkaSimulation.hpp:1195-1198,1280-1283) straight from those arrays, so refined levels no longer need the one-kernel-per-operator path.
#include "qk_sweep_kernels.cuh"

int qk_sweep_stage_keepf(int ns, bool reint, int order, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5],
			 int stage, bool dual, cudaStream_t s)
{
	if (order == 2)
		return sweep_stage_dispatch_plm<0, true>(ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, s);
	return sweep_stage_dispatch<0, true>(ns, reint, ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, true, s);
}
