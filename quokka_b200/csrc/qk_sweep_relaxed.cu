// qk_sweep_relaxed.cu -- the fused stage kernels instantiated with RELAXED arithmetic (QK_ARITH_FAST): this translation
// unit is compiled with FMA contraction ON (csrc/Makefile) and uses the closed-form gamma-law EOS and unchecked ~1-ulp
// reciprocals of qk_relaxed.cuh.  Results are not the reference's bits; the drift is bounded by tests/test_gpu_relaxed.py.
#include "qk_sweep_kernels.cuh"

int qk_sweep_stage_relaxed_plm(int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5], int stage, bool dual,
			       cudaStream_t s)
{
	return sweep_stage_dispatch_plm<1>(ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, s);
}

int qk_sweep_stage_relaxed(int ns, bool reint, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5], int stage,
			   bool dual, bool tma, cudaStream_t s)
{
	return sweep_stage_dispatch<1>(ns, reint, ng, d_counters, c, static_cast<const SweepBox *>(boxes), static_cast<const unsigned char *>(tmaps), nb, maxn, stage, dual, tma, s);
}
