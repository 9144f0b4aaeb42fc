// qk_common.cuh -- shared device/host helpers for libquokka_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quokka_b200.h"

#define QK_MAXV (6 + QK_MAX_SCALARS)

// ---- Array4 views (extern/amrex/Src/Base/AMReX_Array4.H:135-144) ---------------------------------
struct A4 {
	double *__restrict__ p;
	int64_t js, ks, ns;
	int bx, by, bz; // begin
	__host__ __device__ A4() {}
	__host__ __device__ explicit A4(const qk_array4 &a) : p(a.p), js(a.jstride), ks(a.kstride), ns(a.nstride), bx(a.begin[0]), by(a.begin[1]), bz(a.begin[2]) {}
	__device__ __forceinline__ int64_t off(int i, int j, int k) const { return (int64_t)(i - bx) + (int64_t)(j - by) * js + (int64_t)(k - bz) * ks; }
	__device__ __forceinline__ double &operator()(int i, int j, int k, int n) const { return p[off(i, j, k) + n * ns]; }
	__device__ __forceinline__ double ld(int i, int j, int k, int n) const { return __ldg(p + off(i, j, k) + n * ns); }
};
struct IA4 {
	int32_t *__restrict__ p;
	int64_t js, ks;
	int bx, by, bz;
	int ex, ey, ez;
	__host__ __device__ IA4() {}
	__host__ __device__ explicit IA4(const qk_iarray4 &a)
	    : p(a.p), js(a.jstride), ks(a.kstride), bx(a.begin[0]), by(a.begin[1]), bz(a.begin[2]), ex(a.end[0]), ey(a.end[1]), ez(a.end[2])
	{
	}
	__device__ __forceinline__ int32_t &operator()(int i, int j, int k) const { return p[(int64_t)(i - bx) + (int64_t)(j - by) * js + (int64_t)(k - bz) * ks]; }
};

struct Box3 {
	int lo[3], hi[3];
	__host__ __device__ Box3() {}
	__host__ __device__ explicit Box3(const qk_box &b)
	{
		for (int d = 0; d < 3; ++d) {
			lo[d] = b.lo[d];
			hi[d] = b.hi[d];
		}
	}
	__host__ __device__ int len(int d) const { return hi[d] - lo[d] + 1; }
	__host__ __device__ int64_t ncells() const { return (int64_t)len(0) * len(1) * len(2); }
	__host__ __device__ Box3 grown(int n) const
	{
		Box3 b = *this;
		for (int d = 0; d < 3; ++d) {
			b.lo[d] -= n;
			b.hi[d] += n;
		}
		return b;
	}
	__host__ __device__ Box3 face(int dir) const
	{
		Box3 b = *this;
		b.hi[dir] += 1;
		return b;
	}
};

// std::min / std::max semantics (first argument on ties / unordered), as the reference's host+device code
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
__device__ __forceinline__ int sgnd(double v) { return (0.0 < v) - (v < 0.0); }

// ---- host-side plumbing ---------------------------------------------------------------------------
extern int64_t g_qk_launches;
#define QK_LAUNCHED() (++g_qk_launches)
#define QK_CUDA(x)                                                                                                                                   \
	do {                                                                                                                                         \
		cudaError_t e_ = (x);                                                                                                                \
		if (e_ != cudaSuccess)                                                                                                               \
			return (int)e_;                                                                                                              \
	} while (0)
#define QK_KERNEL_CHECK()                                                                                                                            \
	do {                                                                                                                                         \
		QK_LAUNCHED();                                                                                                                       \
		cudaError_t e_ = cudaGetLastError();                                                                                                 \
		if (e_ != cudaSuccess)                                                                                                               \
			return (int)e_;                                                                                                              \
	} while (0)

int qk_require_device();

// ---- optional per-kernel-class device timing (bench.py roofline): CUDA events on the launching stream ----
extern bool g_qk_prof_on;
void qk_prof_mark(const char *name, int launches, cudaStream_t s, bool begin);
struct ProfScope {
	const char *name;
	cudaStream_t s;
	int64_t l0;
	ProfScope(const char *n, cudaStream_t st) : name(n), s(st), l0(g_qk_launches)
	{
		if (g_qk_prof_on)
			qk_prof_mark(name, 0, s, true);
	}
	~ProfScope()
	{
		if (g_qk_prof_on)
			qk_prof_mark(name, (int)(g_qk_launches - l0), s, false);
	}
};
