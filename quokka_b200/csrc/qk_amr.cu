// qk_amr.cu -- coarse <-> fine transfer operators of the AMR ghost fill on the device:
//   qk_amr_interp_cons_lin_minmax = amrex::mf_linear_slope_minmax_interp (the cell-centred interpolater Quokka selects,
//                                   src/simulation.hpp:1389-1407; AMReX_MFInterpolater.cpp:332-418) on a list of box pairs
//   qk_amr_average_down           = amrex::average_down (AMReX_MultiFabUtil_3D_C.H:345-375)
// One launch covers every box pair of the call (patch table as a kernel parameter, 16 pairs per launch).  k_amr_interp: one
// thread per COARSE cell, lane <-> x: the 27-point coarse stencil is read through L1/L2 (neighbouring lanes share 2/3 of it),
// the limiter of all components is reduced in registers, and the cell's ratio^3 children are written directly -- AMReX's
// intermediate slope MultiFab (3 * ncomp components written and read back) does not exist.  HBM-bound streaming: per coarse
// cell and component 8 B read + ratio^3 * 8 B written.  Arithmetic in qk_amr.cuh (bit-identical to AMReX, no libm).
#include "qk_common.cuh"
#include "qk_amr.cuh"

namespace
{
constexpr int AMR_TPB = 128;
constexpr int AMR_MAXPATCH = 16;

struct InterpPatch {
	qk_amr::V4 crse, fine;
	qk_amr::Box region; // fine cells to fill
	int clo[3], cn[3];  // coarsen(region): first coarse cell and extent
	unsigned total;
};
struct InterpTable {
	InterpPatch p[AMR_MAXPATCH];
};
__global__ void __launch_bounds__(AMR_TPB) k_amr_interp(const InterpTable tab, const qk_amr::InterpParams P)
{
	const InterpPatch &T = tab.p[blockIdx.y];
	const unsigned t = blockIdx.x * AMR_TPB + threadIdx.x;
	if (t >= T.total)
		return;
	const unsigned jk = t / (unsigned)T.cn[0];
	const int ic = T.clo[0] + (int)(t - jk * (unsigned)T.cn[0]);
	const int kk = (int)(jk / (unsigned)T.cn[1]);
	const int jc = T.clo[1] + (int)(jk - (unsigned)kk * (unsigned)T.cn[1]);
	const int kc = T.clo[2] + kk;
	qk_amr::interp_coarse_cell(T.crse, T.fine, ic, jc, kc, T.region, P);
}

struct AvgPatch {
	qk_amr::V4 crse, fine;
	int lo[3], n[3];
	unsigned total;
};
struct AvgTable {
	AvgPatch p[AMR_MAXPATCH];
};
struct AvgParams {
	int ratio[3];
	int ccomp, fcomp, ncomp;
};
__global__ void __launch_bounds__(AMR_TPB) k_amr_avgdown(const AvgTable tab, const AvgParams P)
{
	const AvgPatch &T = tab.p[blockIdx.y];
	const unsigned t = blockIdx.x * AMR_TPB + threadIdx.x;
	if (t >= T.total)
		return;
	const unsigned jk = t / (unsigned)T.n[0];
	const int i = T.lo[0] + (int)(t - jk * (unsigned)T.n[0]);
	const int kk = (int)(jk / (unsigned)T.n[1]);
	const int j = T.lo[1] + (int)(jk - (unsigned)kk * (unsigned)T.n[1]);
	const int k = T.lo[2] + kk;
	for (int n = 0; n < P.ncomp; ++n)
		qk_amr::at(T.crse, i, j, k, n + P.ccomp) = qk_amr::avgdown_cell(T.fine, i, j, k, n + P.fcomp, P.ratio);
}

template <bool POST> __global__ void __launch_bounds__(256) k_amr_prepost(const AvgTable tab)
{
	const AvgPatch &T = tab.p[blockIdx.y]; // only .crse, .lo, .n, .total are used
	const unsigned t = blockIdx.x * 256u + threadIdx.x;
	if (t >= T.total)
		return;
	const unsigned jk = t / (unsigned)T.n[0];
	const int i = T.lo[0] + (int)(t - jk * (unsigned)T.n[0]);
	const int kk = (int)(jk / (unsigned)T.n[1]);
	const int j = T.lo[1] + (int)(jk - (unsigned)kk * (unsigned)T.n[1]);
	qk_amr::prepost_cell<POST>(T.crse, i, j, T.lo[2] + kk);
}

bool contains(const qk_array4 &a, const int lo[3], const int hi[3])
{
	for (int d = 0; d < 3; ++d)
		if (lo[d] < a.begin[d] || hi[d] >= a.end[d])
			return false;
	return true;
}
} // namespace

extern "C" int qk_amr_interp_cons_lin_minmax(int npatch, const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp,
					     const qk_box *fine_region, const qk_box *dest_domain, const qk_box *cdomain, const int ratio[3],
					     const int32_t *bc_lo, const int32_t *bc_hi, void *stream)
{
	if (npatch < 0 || (npatch > 0 && (!crse || !fine || !fine_region)) || !dest_domain || !cdomain || !ratio || !bc_lo || !bc_hi || ncomp < 1 ||
	    ccomp < 0 || fcomp < 0)
		return QK_ERR_BAD_ARG;
	if (ncomp > QK_AMR_MAXCOMP || ratio[0] < 1 || ratio[1] < 1 || ratio[2] < 1)
		return QK_ERR_UNSUPPORTED;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	cudaStream_t s = (cudaStream_t)stream;
	qk_amr::InterpParams P;
	for (int d = 0; d < 3; ++d) {
		P.cdomain.lo[d] = cdomain->lo[d];
		P.cdomain.hi[d] = cdomain->hi[d];
		P.dest.lo[d] = dest_domain->lo[d];
		P.dest.hi[d] = dest_domain->hi[d];
		P.ratio[d] = ratio[d];
	}
	P.ccomp = ccomp;
	P.fcomp = fcomp;
	P.ncomp = ncomp;
	for (int n = 0; n < 3 * QK_AMR_MAXCOMP; ++n) {
		P.bc_lo[n] = (n < 3 * ncomp) ? bc_lo[n] : 0;
		P.bc_hi[n] = (n < 3 * ncomp) ? bc_hi[n] : 0;
	}
	ProfScope prof_("amr_interp", s);
	for (int p0 = 0; p0 < npatch; p0 += AMR_MAXPATCH) {
		const int np = (npatch - p0 < AMR_MAXPATCH) ? (npatch - p0) : AMR_MAXPATCH;
		InterpTable tab;
		unsigned most = 0;
		for (int p = 0; p < np; ++p) {
			InterpPatch &T = tab.p[p];
			const qk_array4 &c = crse[p0 + p], &f = fine[p0 + p];
			const qk_box &r = fine_region[p0 + p];
			if (c.ncomp < ccomp + ncomp || f.ncomp < fcomp + ncomp)
				return QK_ERR_BAD_ARG;
			T.crse = qk_amr::view(c);
			T.fine = qk_amr::view(f);
			int64_t tot = 1;
			int need_lo[3], need_hi[3];
			for (int d = 0; d < 3; ++d) {
				T.region.lo[d] = r.lo[d];
				T.region.hi[d] = r.hi[d];
				T.clo[d] = qk_amr::coarsen(r.lo[d], ratio[d]);
				T.cn[d] = qk_amr::coarsen(r.hi[d], ratio[d]) - T.clo[d] + 1;
				if (T.cn[d] < 0)
					T.cn[d] = 0;
				tot *= T.cn[d];
				const int g = (ratio[d] > 1) ? 1 : 0; // CoarseBox: coarsen(fine) grown by 1 where refined
				need_lo[d] = T.clo[d] - g;
				need_hi[d] = T.clo[d] + T.cn[d] - 1 + g;
			}
			if (tot > 0 && (!contains(c, need_lo, need_hi) || !contains(f, r.lo, r.hi)))
				return QK_ERR_BAD_ARG; // the coarse FAB must cover CoarseBox(fine_region), the fine FAB the region
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			T.total = (unsigned)tot;
			most = (T.total > most) ? T.total : most;
		}
		if (most == 0)
			continue;
		k_amr_interp<<<dim3((most + AMR_TPB - 1) / AMR_TPB, (unsigned)np), AMR_TPB, 0, s>>>(tab, P);
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_amr_average_down(int npatch, const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *cbx,
				   const int ratio[3], void *stream)
{
	if (npatch < 0 || (npatch > 0 && (!crse || !fine || !cbx)) || !ratio || ncomp < 1 || ccomp < 0 || fcomp < 0)
		return QK_ERR_BAD_ARG;
	if (ratio[0] < 1 || ratio[1] < 1 || ratio[2] < 1)
		return QK_ERR_UNSUPPORTED;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	cudaStream_t s = (cudaStream_t)stream;
	AvgParams P;
	for (int d = 0; d < 3; ++d)
		P.ratio[d] = ratio[d];
	P.ccomp = ccomp;
	P.fcomp = fcomp;
	P.ncomp = ncomp;
	ProfScope prof_("amr_average_down", s);
	for (int p0 = 0; p0 < npatch; p0 += AMR_MAXPATCH) {
		const int np = (npatch - p0 < AMR_MAXPATCH) ? (npatch - p0) : AMR_MAXPATCH;
		AvgTable tab;
		unsigned most = 0;
		for (int p = 0; p < np; ++p) {
			AvgPatch &T = tab.p[p];
			const qk_array4 &c = crse[p0 + p], &f = fine[p0 + p];
			const qk_box &b = cbx[p0 + p];
			if (c.ncomp < ccomp + ncomp || f.ncomp < fcomp + ncomp)
				return QK_ERR_BAD_ARG;
			T.crse = qk_amr::view(c);
			T.fine = qk_amr::view(f);
			int64_t tot = 1;
			int flo[3], fhi[3];
			for (int d = 0; d < 3; ++d) {
				T.lo[d] = b.lo[d];
				T.n[d] = b.hi[d] - b.lo[d] + 1;
				if (T.n[d] < 0)
					T.n[d] = 0;
				tot *= T.n[d];
				flo[d] = b.lo[d] * ratio[d];
				fhi[d] = (b.hi[d] + 1) * ratio[d] - 1;
			}
			if (tot > 0 && (!contains(c, b.lo, b.hi) || !contains(f, flo, fhi)))
				return QK_ERR_BAD_ARG;
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			T.total = (unsigned)tot;
			most = (T.total > most) ? T.total : most;
		}
		if (most == 0)
			continue;
		k_amr_avgdown<<<dim3((most + AMR_TPB - 1) / AMR_TPB, (unsigned)np), AMR_TPB, 0, s>>>(tab, P);
		QK_KERNEL_CHECK();
	}
	return 0;
}

static int prepost(bool post, int nboxes, const qk_box *bx, const qk_array4 *state, void *stream)
{
	if (nboxes < 0 || (nboxes > 0 && (!bx || !state)))
		return QK_ERR_BAD_ARG;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	cudaStream_t s = (cudaStream_t)stream;
	ProfScope prof_(post ? "amr_post_interp" : "amr_pre_interp", s);
	for (int p0 = 0; p0 < nboxes; p0 += AMR_MAXPATCH) {
		const int np = (nboxes - p0 < AMR_MAXPATCH) ? (nboxes - p0) : AMR_MAXPATCH;
		AvgTable tab;
		unsigned most = 0;
		for (int p = 0; p < np; ++p) {
			AvgPatch &T = tab.p[p];
			const qk_array4 &c = state[p0 + p];
			const qk_box &b = bx[p0 + p];
			if (c.ncomp < 5)
				return QK_ERR_BAD_ARG;
			T.crse = qk_amr::view(c);
			T.fine = T.crse;
			int64_t tot = 1;
			for (int d = 0; d < 3; ++d) {
				T.lo[d] = b.lo[d];
				T.n[d] = b.hi[d] - b.lo[d] + 1;
				if (T.n[d] < 0)
					T.n[d] = 0;
				tot *= T.n[d];
			}
			if (tot > 0 && !contains(c, b.lo, b.hi))
				return QK_ERR_BAD_ARG;
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			T.total = (unsigned)tot;
			most = (T.total > most) ? T.total : most;
		}
		if (most == 0)
			continue;
		const dim3 grid((most + 255u) / 256u, (unsigned)np);
		if (post)
			k_amr_prepost<true><<<grid, 256, 0, s>>>(tab);
		else
			k_amr_prepost<false><<<grid, 256, 0, s>>>(tab);
		QK_KERNEL_CHECK();
	}
	return 0;
}
extern "C" int qk_amr_pre_interp_state(int nboxes, const qk_box *bx, const qk_array4 *state, void *stream) { return prepost(false, nboxes, bx, state, stream); }
extern "C" int qk_amr_post_interp_state(int nboxes, const qk_box *bx, const qk_array4 *state, void *stream) { return prepost(true, nboxes, bx, state, stream); }
