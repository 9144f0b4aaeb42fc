// qk_rad_source.cu -- matter-radiation coupling source terms, single photon group: qk_rad_add_source_terms =
// RadSystem<problem_t>::AddSourceTermsSingleGroup (src/radiation/source_terms_single_group.hpp:9-565) as called by
// QuokkaSimulation::operatorSplitSourceTerms (src/QuokkaSimulation.hpp:1860-1885) twice per radiation substep (:1638,1656).
//
//   k_rad_source   one thread per cell, ALL boxes of the level in one launch (the reference launches box by box under MFIter,
//                  :1631,1650): lane <-> x, so the ten component loads and nine component stores of a warp are contiguous
//                  256-byte rows; the implicit solve (Newton-Raphson on (E_gas, R) inside the lagged work-term iteration) runs
//                  entirely in registers (qk_rad_source.cuh).  Iteration counters are reduced per warp, then per CTA in shared
//                  memory, and reach global memory as one atomic per CTA and counter.
//
// Cell-local and compute-bound (thirteen IEEE divisions per Newton-Raphson iteration in the exact arithmetic mode, one in the
// relaxed mode; 5-20 iterations per cell): the roof is the FP64 pipe, not HBM; algorithmic traffic is 80 B read + 72 B written
// per cell (+8 with an energy source array).  The kernel is instantiated per register cap: measured on B200, 64 registers
// (8 CTAs of 128 threads per SM) is the optimum -- the solve is one dependent chain per thread and needs warps to hide it.
// qk_rad_subcycle (below) is subcycleRadiationAtLevel: transport stages, source terms, ghost fills and component copies.
// DESIGN.md section 3.
#include "qk_level.h"
#include "qk_div.cuh"
#include "qk_rad_source.cuh"

#include <algorithm>
#include <stdlib.h>

namespace
{
constexpr int SRC_TPB = 128;
constexpr int SRC_MAXBOX = 24; // boxes per launch (kernel-parameter table, 24 * 128 B)

// shared-reciprocal IEEE division (qk_div.cuh): the reciprocal of a denominator is refined once, every quotient over it is the
// compiler's own three closing instructions; operands outside the compiler's fast-path domain take its `/`.  Bit-identical to
// DivPlain (tests/test_zgpu_rad_source.py::test_shared_reciprocal_division_is_bit_identical, tests/test_gpu_division.py).
struct DivShared {
	typedef QkRcp R;
	__device__ __forceinline__ static R rcp(double b) { return qk_rcp(b); }
	__device__ __forceinline__ static double div(double a, const R &r) { return qk_div(a, r); }
	// call-constant denominators: (b, refined 1/b) pairs in shared memory, filled once per CTA -- no registers are held for
	// them across the Newton-Raphson loop; the host has checked that every b is finite and normal (else DivPlain runs)
	struct CTab {
		const double2 *sm;
	};
	__device__ __forceinline__ static double divc(double a, const CTab &t, int idx)
	{
		const double2 by = t.sm[idx];
		QkRcp r;
		r.b = by.x;
		r.y = by.y;
		r.ok = true;
		return qk_div(a, r);
	}
};
__device__ __forceinline__ void make_ctab(const qk_rsrc::Const &k, qk_rsrc::DivPlain::CTab &ct, double2 *) { qk_rsrc::const_denoms(k, ct.b); }
__device__ __forceinline__ void make_ctab(const qk_rsrc::Const &k, DivShared::CTab &ct, double2 *sm)
{
	if (threadIdx.x < qk_rsrc::C_N) {
		double b[qk_rsrc::C_N];
		qk_rsrc::const_denoms(k, b);
		double mine = b[0];
#pragma unroll
		for (int n = 1; n < qk_rsrc::C_N; ++n)
			mine = (threadIdx.x == n) ? b[n] : mine;
		const QkRcp r = qk_rcp(mine);
		sm[threadIdx.x] = make_double2(r.b, r.y);
	}
	__syncthreads();
	ct.sm = sm;
}

struct SrcBox {
	A4 cons, esrc; // esrc.p == nullptr: zero source
	int lo[3], n[3];
	int64_t total;
};
struct SrcTable {
	SrcBox b[SRC_MAXBOX];
};

template <class D, int MINB, bool RELAXED> __global__ void __launch_bounds__(SRC_TPB, MINB) k_rad_source(const qk_rsrc::Const k, const SrcTable tab, const int nstart, int *__restrict__ counters)
{
	const SrcBox &B = tab.b[blockIdx.y];
	const unsigned t = blockIdx.x * SRC_TPB + threadIdx.x; // the host refuses boxes of 2^31 cells or more
	const bool active = (t < (unsigned)B.total);
	qk_rsrc::CellOut out;
	out.solves = out.nr_iters = out.nr_max = out.fail_nr = out.fail_outer = 0;
	__shared__ double2 s_ct[qk_rsrc::C_N];
	typename D::CTab ct;
	make_ctab(k, ct, s_ct);
	if (active) {
		const unsigned jk = t / (unsigned)B.n[0];
		const int i = B.lo[0] + (int)(t - jk * (unsigned)B.n[0]);
		const int kk = (int)(jk / (unsigned)B.n[1]);
		const int j = B.lo[1] + (int)(jk - (unsigned)kk * (unsigned)B.n[1]);
		const int kz = B.lo[2] + kk;
		double *__restrict__ p = B.cons.p + B.cons.off(i, j, kz);
		const int64_t ns = B.cons.ns;
		double *__restrict__ pr = p + (int64_t)nstart * ns;
		qk_rsrc::CellIn in;
		in.rho = p[0];
		in.mom[0] = p[ns];
		in.mom[1] = p[2 * ns];
		in.mom[2] = p[3 * ns];
		in.Egastot = p[4 * ns];
		in.Erad = pr[0];
		in.F[0] = pr[ns];
		in.F[1] = pr[2 * ns];
		in.F[2] = pr[3 * ns];
		in.src = (B.esrc.p != nullptr) ? B.esrc.p[B.esrc.off(i, j, kz)] : 0.0;
		qk_rsrc::source_cell<D, RELAXED>(k, ct, in, out);
		p[ns] = out.mom[0];
		p[2 * ns] = out.mom[1];
		p[3 * ns] = out.mom[2];
		if (k.gamma != 1.0) { // :558-561
			p[4 * ns] = out.Egastot;
			p[5 * ns] = out.Eint;
			pr[0] = out.Erad;
		}
		pr[ns] = out.F[0];
		pr[2 * ns] = out.F[1];
		pr[3 * ns] = out.F[2];
	}
	if (counters != nullptr) { // uniform across the grid
		__shared__ int sh[5];
		if (threadIdx.x < 5)
			sh[threadIdx.x] = 0;
		__syncthreads();
		const unsigned full = 0xffffffffu;
		const int s0 = __reduce_add_sync(full, out.solves);
		const int s1 = __reduce_add_sync(full, out.nr_iters);
		const int s2 = __reduce_max_sync(full, out.nr_max);
		const int s3 = __reduce_add_sync(full, out.fail_nr);
		const int s4 = __reduce_add_sync(full, out.fail_outer);
		if ((threadIdx.x & 31) == 0) {
			atomicAdd(&sh[0], s0);
			atomicAdd(&sh[1], s1);
			atomicMax(&sh[2], s2);
			atomicAdd(&sh[3], s3);
			atomicAdd(&sh[4], s4);
		}
		__syncthreads();
		// counters[0..3] = iteration_counter, [4..6] = iteration_failure_counter (src/QuokkaSimulation.hpp:1620-1625)
		if (threadIdx.x == 0 && sh[0] != 0) {
			atomicAdd(&counters[0], sh[0]);
			atomicAdd(&counters[1], sh[1]);
			atomicMax(&counters[2], sh[2]);
			if (sh[3])
				atomicAdd(&counters[4], sh[3]);
			if (sh[4])
				atomicAdd(&counters[6], sh[4]);
		}
	}
}

// radiation components of src -> dst on the valid boxes, all boxes in one launch (swapRadiationState; stage-2 result back)
struct CopyBox {
	A4 dst, src;
	int lo[3], n[3];
	unsigned total;
};
struct CopyTable {
	CopyBox b[SRC_MAXBOX];
};
__global__ void __launch_bounds__(256) k_copy_comps(const CopyTable tab, const int c0, const int nc)
{
	const CopyBox &B = tab.b[blockIdx.y];
	const unsigned t = blockIdx.x * 256u + threadIdx.x;
	if (t >= B.total)
		return;
	const unsigned jk = t / (unsigned)B.n[0];
	const int i = B.lo[0] + (int)(t - jk * (unsigned)B.n[0]);
	const int kk = (int)(jk / (unsigned)B.n[1]);
	const int j = B.lo[1] + (int)(jk - (unsigned)kk * (unsigned)B.n[1]);
	const int k = B.lo[2] + kk;
	const double *__restrict__ ps = B.src.p + B.src.off(i, j, k);
	double *__restrict__ pd = B.dst.p + B.dst.off(i, j, k);
	for (int n = c0; n < c0 + nc; ++n)
		pd[n * B.dst.ns] = ps[n * B.src.ns];
}
int copy_comps(const qk_level *L, const qk_array4 *dst, const qk_array4 *src, int c0, int nc, cudaStream_t s)
{
	const int nboxes = (int)L->valid.size();
	for (int b0 = 0; b0 < nboxes; b0 += SRC_MAXBOX) {
		const int nb = (nboxes - b0 < SRC_MAXBOX) ? (nboxes - b0) : SRC_MAXBOX;
		CopyTable tab;
		unsigned most = 0;
		for (int b = 0; b < nb; ++b) {
			CopyBox &B = tab.b[b];
			B.dst = A4(dst[b0 + b]);
			B.src = A4(src[b0 + b]);
			int64_t tot = 1;
			for (int d = 0; d < 3; ++d) {
				B.lo[d] = L->valid[b0 + b].lo[d];
				B.n[d] = L->valid[b0 + b].hi[d] - L->valid[b0 + b].lo[d] + 1;
				tot *= B.n[d];
			}
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			B.total = (unsigned)tot;
			most = (B.total > most) ? B.total : most;
		}
		if (most == 0)
			continue;
		k_copy_comps<<<dim3((most + 255u) / 256u, (unsigned)nb), 256, 0, s>>>(tab, c0, nc);
		QK_KERNEL_CHECK();
	}
	return 0;
}

int *g_dcount = nullptr; // 8 ints on the device
int *g_hcount = nullptr; // pinned mirror
} // namespace

extern "C" int qk_rad_add_source_terms(const qk_hydro_params *hydro, const qk_rad_params *prm, const qk_rad_source_params *src, int stage, int nboxes,
				       const qk_box *valid, const qk_array4 *cons, const qk_array4 *rad_energy_source, double dt_radiation,
				       int64_t *counters, void *stream)
{
	if (!hydro || !prm || !src || (nboxes > 0 && (!valid || !cons)) || nboxes < 0 || (stage != 1 && stage != 2))
		return QK_ERR_BAD_ARG;
	// single group with user opacities = OpacityModel::single_group (radiation_system.hpp:63-64,216-222); multi-group
	// (AddSourceTermsMultiGroup), the dust model and beta_order outside 0..3 (static_assert :113) are not built
	if (prm->ngroups != 1 || src->opacity_model != QK_OPACITY_CONSTANT || src->beta_order < 0 || src->beta_order > 3 || prm->nstart < 6)
		return QK_ERR_UNSUPPORTED;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	cudaStream_t s = (cudaStream_t)stream;
	const qk_rsrc::Const k = qk_rsrc::make_const(hydro, prm, src, dt_radiation, stage);
	ProfScope prof_("rad_source_terms", s);
	// QK_RADSRC_PLAIN_DIV=0: the shared-reciprocal instantiation (same bits; measured slower here)
	const char *pd = getenv("QK_RADSRC_PLAIN_DIV");
	bool plain_div = !(pd != nullptr && pd[0] == '0');
	// QK_RADSRC_MINB: resident CTAs per SM the kernel is compiled for (register cap 168 / 128 / 96 / 80); tuning knob
	const char *mb = getenv("QK_RADSRC_MINB");
	const int minb = (mb != nullptr) ? atoi(mb) : 8;
	{
		double b[qk_rsrc::C_N];
		qk_rsrc::const_denoms(k, b);
		for (int n = 0; n < qk_rsrc::C_N; ++n)
			if (!(fabs(b[n]) >= 2.3e-308 && fabs(b[n]) < 1.0e306)) // outside qk_rcp's fast-path domain
				plain_div = true;
	}
	int *dcount = nullptr;
	if (counters) {
		if (!g_dcount) {
			QK_CUDA(cudaMalloc(&g_dcount, 8 * sizeof(int)));
			QK_CUDA(cudaMallocHost(&g_hcount, 8 * sizeof(int)));
		}
		dcount = g_dcount;
		QK_CUDA(cudaMemsetAsync(dcount, 0, 8 * sizeof(int), s));
	}
	for (int b0 = 0; b0 < nboxes; b0 += SRC_MAXBOX) {
		const int nb = (nboxes - b0 < SRC_MAXBOX) ? (nboxes - b0) : SRC_MAXBOX;
		SrcTable tab;
		int64_t most = 0;
		for (int b = 0; b < nb; ++b) {
			SrcBox &B = tab.b[b];
			B.cons = A4(cons[b0 + b]);
			if (cons[b0 + b].ncomp < prm->nstart + 4)
				return QK_ERR_BAD_ARG;
			if (rad_energy_source)
				B.esrc = A4(rad_energy_source[b0 + b]);
			else {
				B.esrc = B.cons;
				B.esrc.p = nullptr;
			}
			B.total = 1;
			for (int d = 0; d < 3; ++d) {
				B.lo[d] = valid[b0 + b].lo[d];
				B.n[d] = valid[b0 + b].hi[d] - valid[b0 + b].lo[d] + 1;
				if (B.n[d] < 0)
					B.n[d] = 0;
				B.total *= B.n[d];
			}
			if (B.total > most)
				most = B.total;
		}
		if (most == 0)
			continue;
		const dim3 grid((unsigned)((most + SRC_TPB - 1) / SRC_TPB), (unsigned)nb);
		if (most >= (int64_t(1) << 31))
			return QK_ERR_UNSUPPORTED;
#define QK_SRC_LAUNCH(DIV, MINB) k_rad_source<DIV, MINB, false><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount)
		if (hydro->arith == QK_ARITH_FAST) { // relaxed arithmetic (qk_rad_source.cuh): closed-form EOS, reciprocal products
			if (minb == 6)
				k_rad_source<qk_rsrc::DivPlain, 6, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
			else if (minb == 5)
				k_rad_source<qk_rsrc::DivPlain, 5, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
			else if (minb == 4)
				k_rad_source<qk_rsrc::DivPlain, 4, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
			else if (minb == 10)
				k_rad_source<qk_rsrc::DivPlain, 10, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
			else if (minb == 12)
				k_rad_source<qk_rsrc::DivPlain, 12, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
			else
				k_rad_source<qk_rsrc::DivPlain, 8, true><<<grid, SRC_TPB, 0, s>>>(k, tab, prm->nstart, dcount);
		} else if (plain_div) {
			switch (minb) {
			case 3: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 3); break;
			case 5: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 5); break;
			case 4: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 4); break;
			case 6: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 6); break;
			case 7: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 7); break;
			case 10: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 10); break;
			default: QK_SRC_LAUNCH(qk_rsrc::DivPlain, 8); break;
			}
		} else {
			switch (minb) {
			case 3: QK_SRC_LAUNCH(DivShared, 3); break;
			case 5: QK_SRC_LAUNCH(DivShared, 5); break;
			default: QK_SRC_LAUNCH(DivShared, 4); break;
			}
		}
#undef QK_SRC_LAUNCH
		QK_KERNEL_CHECK();
	}
	if (counters) {
		QK_CUDA(cudaMemcpyAsync(g_hcount, dcount, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
		QK_CUDA(cudaStreamSynchronize(s));
		for (int n = 0; n < QK_RAD_SOURCE_NCOUNTERS; ++n) {
			if (n == 2)
				counters[2] = (counters[2] < g_hcount[2]) ? g_hcount[2] : counters[2];
			else
				counters[n] += g_hcount[n];
		}
	}
	return 0;
}

#define QK_TRY_(x)                                                                                                                                   \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

extern "C" int qk_rad_subcycle(qk_level *L, const qk_hydro_params *hydro, const qk_rad_params *prm, const qk_rad_source_params *src, const qk_array4 *U_old,
			       const qk_array4 *U_new, const qk_array4 *U_tmp, const qk_array4 *rad_energy_source, double dt_hydro, double rad_cfl,
			       int64_t *counters, int *nsub_out, void *stream)
{
	if (!L || !prm || !U_old || !U_new || !U_tmp || (src && !hydro) || !(dt_hydro > 0.0) || !(rad_cfl > 0.0))
		return QK_ERR_BAD_ARG;
	if (!L->has_device)
		return QK_ERR_NO_DEVICE;
	cudaStream_t s = (cudaStream_t)stream;
	// computeNumberOfRadiationSubsteps  src/QuokkaSimulation.hpp:397-406
	const double dx_min = std::min({L->dx[0], L->dx[1], L->dx[2]});
	const double dtrad_tmp = rad_cfl * (dx_min / prm->c_hat);
	const int nsub = (int)ceil(dt_hydro / dtrad_tmp);
	if (nsub < 1)
		return QK_ERR_BAD_ARG;
	const double dt_radiation = dt_hydro / static_cast<double>(nsub);
	if (nsub_out)
		*nsub_out = nsub;
	const int ns = prm->nstart, nh = 4 * prm->ngroups, nb = (int)L->valid.size();
	// one arithmetic mode for the whole subcycle: the hydro block's (relaxed transport sweeps with relaxed source terms)
	qk_rad_params rp = *prm;
	if (hydro)
		rp.arith = hydro->arith;
	prm = &rp;
	for (int i = 0; i < nsub; ++i) {
		if (i > 0)
			QK_TRY_(copy_comps(L, U_old, U_new, ns, nh, s)); // swapRadiationState
		QK_TRY_(qk_fill_boundary(L, U_old, ns, nh, stream));
		QK_TRY_(qk_rad_advance_stage(L, prm, 1, U_old, U_old, U_new, dt_radiation, stream));
		if (src)
			QK_TRY_(qk_rad_add_source_terms(hydro, prm, src, 1, nb, L->valid.data(), U_new, rad_energy_source, dt_radiation, counters, stream));
		QK_TRY_(qk_fill_boundary(L, U_new, ns, nh, stream));
		QK_TRY_(qk_rad_advance_stage(L, prm, 2, U_old, U_new, U_tmp, dt_radiation, stream));
		QK_TRY_(copy_comps(L, U_new, U_tmp, ns, nh, s));
		if (src)
			QK_TRY_(qk_rad_add_source_terms(hydro, prm, src, 2, nb, L->valid.data(), U_new, rad_energy_source, dt_radiation, counters, stream));
	}
	return 0;
}
