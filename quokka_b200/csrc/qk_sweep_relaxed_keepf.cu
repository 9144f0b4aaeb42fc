// qk_sweep_relaxed_keepf.cu -- as qk_sweep_keepf.cu for the RELAXED arithmetic mode (compiled with FMA contraction, csrc/Makefile).  The relaxed
// stage keeps R(U0) per cell instead of the half-flux arrays, so the kept fluxes F(U0) / F(U1) are the only face arrays it writes.
#include "qk_sweep_kernels.cuh"

int qk_sweep_stage_relaxed_keepf(int ns, bool reint, int order, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb,
				 const int maxn[5], int stage, bool dual, cudaStream_t s)
{
	if (order == 2)
		return sweep_stage_dispatch_plm<1, true>(ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, s);
	return sweep_stage_dispatch<1, true>(ns, reint, ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, true, s);
}
