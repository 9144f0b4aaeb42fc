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
#include "qk_physics.cuh"

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

// ---------------------------------------------------------------------------------------------------------------
// time interpolation of the coarse data (FillPatcher::fill), regrid tagging (ErrorEst), FixupState
// ---------------------------------------------------------------------------------------------------------------
namespace
{
struct TiPatch {
	qk_amr::V4 dst, s0, s1;
	int lo[3], n[3];
	unsigned total;
};
struct TiTable {
	TiPatch p[AMR_MAXPATCH];
};
__global__ void __launch_bounds__(256) k_amr_time_interp(const TiTable tab, int which, double alpha, double beta, int dcomp, int scomp, int ncomp)
{
	const TiPatch &T = tab.p[blockIdx.y];
	const unsigned t = blockIdx.x * 256u + threadIdx.x;
	if (t >= T.total)
		return;
	const unsigned jk = t / (unsigned)T.n[0];
	const int i = T.lo[0] + (int)(t - jk * (unsigned)T.n[0]);
	const int kk = (int)(jk / (unsigned)T.n[1]);
	const int j = T.lo[1] + (int)(jk - (unsigned)kk * (unsigned)T.n[1]);
	const int k = T.lo[2] + kk;
	for (int n = 0; n < ncomp; ++n) {
		const double a0 = (which != 1) ? qk_amr::at(T.s0, i, j, k, scomp + n) : 0.0;
		const double a1 = (which != 0) ? qk_amr::at(T.s1, i, j, k, scomp + n) : 0.0;
		qk_amr::at(T.dst, i, j, k, dcomp + n) = qk_amr::time_interp_value(which, alpha, beta, a0, a1);
	}
}

struct C4 { // amrex::Array4<char>
	char *p;
	int64_t js, ks;
	int b[3];
};
struct TagPatch {
	qk_amr::V4 u;
	C4 tag;
	int lo[3], n[3];
	unsigned total;
};
struct TagTable {
	TagPatch p[AMR_MAXPATCH];
};
// MODE 0: pressure-gradient criterion (Sedov); MODE 1: x-gradient of one component (Sod)
template <int MODE>
__global__ void __launch_bounds__(256) k_tag(const TagTable tab, HydroConst c, int comp, double dx, double eta, double qmin, unsigned long long *count)
{
	const TagPatch &T = tab.p[blockIdx.y];
	const unsigned t = blockIdx.x * 256u + threadIdx.x;
	bool set = false;
	if (t < T.total) {
		const unsigned jk = t / (unsigned)T.n[0];
		const int i = T.lo[0] + (int)(t - jk * (unsigned)T.n[0]);
		const int kk = (int)(jk / (unsigned)T.n[1]);
		const int j = T.lo[1] + (int)(jk - (unsigned)kk * (unsigned)T.n[1]);
		const int k = T.lo[2] + kk;
		if (MODE == 0) {
			const int di[7] = {0, 1, -1, 0, 0, 0, 0}, dj[7] = {0, 0, 0, 1, -1, 0, 0}, dk[7] = {0, 0, 0, 0, 0, 1, -1};
			double P[7];
#pragma unroll
			for (int m = 0; m < 7; ++m) {
				const int ii = i + di[m], jj = j + dj[m], kq = k + dk[m];
				P[m] = cons_pressure(c, qk_amr::at(T.u, ii, jj, kq, 0), qk_amr::at(T.u, ii, jj, kq, 1), qk_amr::at(T.u, ii, jj, kq, 2),
						     qk_amr::at(T.u, ii, jj, kq, 3), qk_amr::at(T.u, ii, jj, kq, 4));
			}
			set = qk_amr::tag_pressure_gradient(P, eta, qmin);
		} else {
			set = qk_amr::tag_gradient_x(qk_amr::at(T.u, i - 1, j, k, comp), qk_amr::at(T.u, i, j, k, comp), qk_amr::at(T.u, i + 1, j, k, comp), dx, eta,
						     qmin);
		}
		if (set)
			T.tag.p[(int64_t)(i - T.tag.b[0]) + (int64_t)(j - T.tag.b[1]) * T.tag.js + (int64_t)(k - T.tag.b[2]) * T.tag.ks] = QK_TAG_SET; // TagBox::SET
	}
	if (count) {
		const int n = __syncthreads_count(set ? 1 : 0);
		if (threadIdx.x == 0 && n)
			atomicAdd(count, (unsigned long long)n);
	}
}
unsigned long long *g_tag_dev = nullptr, *g_tag_host = nullptr;
} // namespace

extern "C" int qk_amr_time_interp(int npatch, const qk_array4 *dst, int dcomp, const qk_array4 *src0, const qk_array4 *src1, int scomp, int ncomp,
				  const qk_box *region, double t0, double t1, double time, int *which_out, void *stream)
{
	if (npatch < 0 || (npatch > 0 && (!dst || !src0 || !region)) || ncomp < 1 || dcomp < 0 || scomp < 0)
		return QK_ERR_BAD_ARG;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	const int which = qk_amr::time_interp_branch(t0, t1, time, src1 != nullptr);
	if (which_out)
		*which_out = which;
	const double alpha = (which == 2) ? (t1 - time) / (t1 - t0) : 0.0, beta = (which == 2) ? (time - t0) / (t1 - t0) : 0.0;
	cudaStream_t s = (cudaStream_t)stream;
	ProfScope prof_("amr_time_interp", s);
	for (int p0 = 0; p0 < npatch; p0 += AMR_MAXPATCH) {
		const int np = (npatch - p0 < AMR_MAXPATCH) ? (npatch - p0) : AMR_MAXPATCH;
		TiTable tab;
		unsigned most = 0;
		for (int p = 0; p < np; ++p) {
			TiPatch &T = tab.p[p];
			const qk_box &b = region[p0 + p];
			T.dst = qk_amr::view(dst[p0 + p]);
			T.s0 = qk_amr::view(src0[p0 + p]);
			T.s1 = qk_amr::view(src1 ? src1[p0 + p] : src0[p0 + p]);
			int64_t tot = 1;
			for (int d = 0; d < 3; ++d) {
				T.lo[d] = b.lo[d];
				T.n[d] = b.hi[d] - b.lo[d] + 1;
				if (T.n[d] < 0)
					T.n[d] = 0;
				tot *= T.n[d];
			}
			if (tot > 0 && (!contains(dst[p0 + p], b.lo, b.hi) || !contains(src0[p0 + p], b.lo, b.hi) || (src1 && !contains(src1[p0 + p], b.lo, b.hi)) ||
					dcomp + ncomp > dst[p0 + p].ncomp || scomp + ncomp > src0[p0 + p].ncomp))
				return QK_ERR_BAD_ARG;
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			T.total = (unsigned)tot;
			most = (T.total > most) ? T.total : most;
		}
		if (most == 0)
			continue;
		k_amr_time_interp<<<dim3((most + 255u) / 256u, (unsigned)np), 256, 0, s>>>(tab, which, alpha, beta, dcomp, scomp, ncomp);
		QK_KERNEL_CHECK();
	}
	return 0;
}

static int tag_common(int mode, const HydroConst &c, int nboxes, const qk_box *valid, const qk_array4 *state, int comp, const qk_carray4 *tags, double dx,
		      double eta, double qmin, int64_t *ntagged, void *stream)
{
	if (nboxes < 0 || (nboxes > 0 && (!valid || !state || !tags)))
		return QK_ERR_BAD_ARG;
	{
		const int r = qk_require_device();
		if (r != 0)
			return r;
	}
	cudaStream_t s = (cudaStream_t)stream;
	if (ntagged) {
		if (!g_tag_dev) {
			QK_CUDA(cudaMalloc(&g_tag_dev, 8));
			QK_CUDA(cudaMallocHost(&g_tag_host, 8));
		}
		QK_CUDA(cudaMemsetAsync(g_tag_dev, 0, 8, s));
	}
	ProfScope prof_("error_est", s);
	for (int p0 = 0; p0 < nboxes; p0 += AMR_MAXPATCH) {
		const int np = (nboxes - p0 < AMR_MAXPATCH) ? (nboxes - p0) : AMR_MAXPATCH;
		TagTable tab;
		unsigned most = 0;
		for (int p = 0; p < np; ++p) {
			TagPatch &T = tab.p[p];
			const qk_box &b = valid[p0 + p];
			const qk_array4 &u = state[p0 + p];
			const qk_carray4 &tg = tags[p0 + p];
			T.u = qk_amr::view(u);
			T.tag.p = tg.p;
			T.tag.js = tg.jstride;
			T.tag.ks = tg.kstride;
			int64_t tot = 1;
			int glo[3], ghi[3];
			for (int d = 0; d < 3; ++d) {
				T.tag.b[d] = tg.begin[d];
				T.lo[d] = b.lo[d];
				T.n[d] = b.hi[d] - b.lo[d] + 1;
				if (T.n[d] < 0)
					T.n[d] = 0;
				tot *= T.n[d];
				const int g = (mode == 0 || d == 0) ? 1 : 0; // one filled ghost cell where the stencil reaches
				glo[d] = b.lo[d] - g;
				ghi[d] = b.hi[d] + g;
				if (tot > 0 && (b.lo[d] < tg.begin[d] || b.hi[d] >= tg.end[d]))
					return QK_ERR_BAD_ARG;
			}
			if (tot > 0 && (!contains(u, glo, ghi) || u.ncomp < ((mode == 0) ? 5 : comp + 1)))
				return QK_ERR_BAD_ARG;
			if (tot >= (int64_t(1) << 31))
				return QK_ERR_UNSUPPORTED;
			T.total = (unsigned)tot;
			most = (T.total > most) ? T.total : most;
		}
		if (most == 0)
			continue;
		const dim3 grid((most + 255u) / 256u, (unsigned)np);
		if (mode == 0)
			k_tag<0><<<grid, 256, 0, s>>>(tab, c, comp, dx, eta, qmin, ntagged ? g_tag_dev : nullptr);
		else
			k_tag<1><<<grid, 256, 0, s>>>(tab, c, comp, dx, eta, qmin, ntagged ? g_tag_dev : nullptr);
		QK_KERNEL_CHECK();
	}
	if (ntagged) {
		QK_CUDA(cudaMemcpyAsync(g_tag_host, g_tag_dev, 8, cudaMemcpyDeviceToHost, s));
		QK_CUDA(cudaStreamSynchronize(s));
		*ntagged = (int64_t)g_tag_host[0];
	}
	return 0;
}

extern "C" int qk_tag_pressure_gradient(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, const qk_carray4 *tags,
					double eta_threshold, double P_min, int64_t *ntagged, void *stream)
{
	if (!prm)
		return QK_ERR_BAD_ARG;
	return tag_common(0, make_hydro_const(prm), nboxes, valid, cons, 0, tags, 0.0, eta_threshold, P_min, ntagged, stream);
}

extern "C" int qk_tag_gradient_x(int nboxes, const qk_box *valid, const qk_array4 *state, int comp, const qk_carray4 *tags, double dx, double eta_threshold,
				 double q_min, int64_t *ntagged, void *stream)
{
	if (comp < 0 || !(dx > 0.0))
		return QK_ERR_BAD_ARG;
	HydroConst c{};
	return tag_common(1, c, nboxes, valid, state, comp, tags, dx, eta_threshold, q_min, ntagged, stream);
}

extern "C" int qk_hydro_fixup_state(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, void *stream)
{
	const int rc = qk_hydro_enforce_limits(prm, nboxes, valid, state, stream);
	if (rc != 0)
		return rc;
	return qk_hydro_sync_dual_energy(prm, nboxes, valid, state, nullptr, stream);
}
