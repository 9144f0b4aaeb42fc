// qk_rad_relaxed.cu -- the TMA-staged radiation transport sweeps instantiated with RELAXED arithmetic (qk_rad_params::arith == QK_ARITH_FAST):
// this translation unit is compiled with FMA contraction ON (csrc/Makefile) and uses ~1-ulp reciprocals, Newton square roots without a slow
// path and the f^2 form of the closure (qk_rad_kernels.cuh).  Results are not the oracle's bits; tests/test_gpu_radiation.py bounds the drift.
#include "qk_rad_kernels.cuh"

int qk_rad_stage_relaxed(int order, const void *rad_const, const void *boxes, const void *maps, int nb, const int maxn[3], int stage, bool keep, bool fix, int g,
			 double dtdx, double dtdy, double dtdz, cudaStream_t s)
{
	return dispatch_rad_tma<1>(order, *static_cast<const RadConst *>(rad_const), static_cast<const RadBox2 *>(boxes), static_cast<const RadMaps *>(maps), nb, maxn,
				   stage, keep, fix, g, dtdx, dtdy, dtdz, s);
}
