// qk_amr.cuh -- coarse <-> fine transfer operators of the AMR ghost fill, per coarse cell:
//   * the cell-centred interpolater Quokka selects with amr_interpolation_method = 1 (getAmrInterpolaterCellCentered,
//     src/simulation.hpp:1389-1407) = amrex::mf_linear_slope_minmax_interp: MFCellConsLinMinmaxLimitInterp::interp
//     (extern/amrex/Src/AmrCore/AMReX_MFInterpolater.cpp:332-418), mf_cell_cons_lin_interp_limit_minmax_llslope and
//     mf_cell_cons_lin_interp (AMReX_MFInterp_3D_C.H:7-109,246-262), mf_compute_slopes_{x,y,z} (AMReX_MFInterp_C.H:10-90);
//   * amrex::average_down = amrex_avgdown (extern/amrex/Src/Base/AMReX_MultiFabUtil_3D_C.H:345-375).
//
// AMReX runs the interpolation as two passes with a slope MultiFab (3 * ncomp components) between them.  Here ONE thread owns
// a coarse cell: it finds the per-direction limiter over all components (pass 1, nothing kept but three numbers), then forms
// each component's limited slopes again and writes the cell's ratio^3 fine children directly -- no slope array exists.
// Everything is QK_AHD (host + device) so that tests/host_src/amr_host.cpp runs the same arithmetic on the CPU against the
// oracle; there is no libm call and no contraction (--fmad=false), so CPU and GPU results are bit-identical to AMReX's.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/quokka_b200.h"

#ifdef __CUDACC__
#define QK_AHD __host__ __device__ __forceinline__
#else
#define QK_AHD static inline
#endif

#define QK_AMR_MAXCOMP 16 // components per call (6 + QK_MAX_SCALARS gas components, or the radiation block)

namespace qk_amr
{
struct V4 { // amrex::Array4<double>
	double *p;
	int64_t js, ks, ns;
	int b[3], e[3]; // begin, end (exclusive)
};
QK_AHD double &at(const V4 &a, int i, int j, int k, int n) { return a.p[(int64_t)(i - a.b[0]) + (int64_t)(j - a.b[1]) * a.js + (int64_t)(k - a.b[2]) * a.ks + n * a.ns]; }
static inline V4 view(const qk_array4 &q)
{
	V4 v;
	v.p = q.p;
	v.js = q.jstride;
	v.ks = q.kstride;
	v.ns = q.nstride;
	for (int d = 0; d < 3; ++d) {
		v.b[d] = q.begin[d];
		v.e[d] = q.end[d];
	}
	return v;
}
struct Box {
	int lo[3], hi[3];
};
QK_AHD double mn(double a, double b) { return (b < a) ? b : a; } // amrex::min = std::min
QK_AHD double mx(double a, double b) { return (a < b) ? b : a; }
QK_AHD int coarsen(int i, int r) { return (i < 0) ? -((-i + r - 1) / r) : i / r; } // amrex::coarsen: floor division

struct InterpParams {
	Box cdomain, dest; // coarse domain; fine cells outside `dest` are not written
	int ratio[3];
	int ccomp, fcomp, ncomp;
	int32_t bc_lo[3 * QK_AMR_MAXCOMP], bc_hi[3 * QK_AMR_MAXCOMP]; // [3 * comp + dim], amrex::BCType values
};

// mf_compute_slopes_<dir>
QK_AHD double slope_dir(const V4 &u, int i, int j, int k, int nu, int dir, const Box &dom, int bclo, int bchi)
{
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	const int idx = (dir == 0) ? i : (dir == 1) ? j : k;
#define QK_U(o) at(u, i + (o)*e0, j + (o)*e1, k + (o)*e2, nu)
	double dc = 0.5 * (QK_U(1) - QK_U(-1));
	if (idx == dom.lo[dir] && (bclo == QK_BC_EXT_DIR || bclo == 4 /* BCType::hoextrap */)) {
		if (idx + 2 < u.e[dir])
			dc = -(16. / 15.) * QK_U(-1) + 0.5 * QK_U(0) + (2. / 3.) * QK_U(1) - 0.1 * QK_U(2);
		else
			dc = 0.25 * (QK_U(1) + 5. * QK_U(0) - 6. * QK_U(-1));
	}
	if (idx == dom.hi[dir] && (bchi == QK_BC_EXT_DIR || bchi == 4)) {
		if (idx - 2 >= u.b[dir])
			dc = (16. / 15.) * QK_U(1) - 0.5 * QK_U(0) - (2. / 3.) * QK_U(-1) + 0.1 * QK_U(-2);
		else
			dc = -0.25 * (QK_U(-1) + 5. * QK_U(0) - 6. * QK_U(1));
	}
#undef QK_U
	return dc;
}

// unlimited (dc) and min-max-limited (sl) slopes of component nu in coarse cell (i,j,k)  AMReX_MFInterp_3D_C.H:23-87
QK_AHD void cell_slopes(const V4 &u, int i, int j, int k, int ns, const InterpParams &P, double dc[3], double sl[3])
{
	const int nu = ns + P.ccomp;
	const double uc = at(u, i, j, k, nu);
	for (int d = 0; d < 3; ++d) {
		dc[d] = 0.;
		sl[d] = 0.;
		if (P.ratio[d] > 1) {
			const int e0 = (d == 0), e1 = (d == 1), e2 = (d == 2);
			dc[d] = slope_dir(u, i, j, k, nu, d, P.cdomain, P.bc_lo[3 * ns + d], P.bc_hi[3 * ns + d]);
			const double df = 2.0 * (at(u, i + e0, j + e1, k + e2, nu) - uc);
			const double db = 2.0 * (uc - at(u, i - e0, j - e1, k - e2, nu));
			double sd = (df * db >= 0.0) ? mn(fabs(df), fabs(db)) : 0.;
			sd = copysign(1., dc[d]) * mn(sd, fabs(dc[d]));
			sl[d] = sd;
		}
	}
	double alpha = 1.0;
	if (sl[0] != 0.0 || sl[1] != 0.0 || sl[2] != 0.0) {
		const double dumax = fabs(sl[0]) * (double)(P.ratio[0] - 1) / (double)(2 * P.ratio[0]) + fabs(sl[1]) * (double)(P.ratio[1] - 1) / (double)(2 * P.ratio[1]) +
				     fabs(sl[2]) * (double)(P.ratio[2] - 1) / (double)(2 * P.ratio[2]);
		double umax = uc, umin = uc;
		const int il = P.ratio[0] > 1, jl = P.ratio[1] > 1, kl = P.ratio[2] > 1;
		for (int ko = -kl; ko <= kl; ++ko)
			for (int jo = -jl; jo <= jl; ++jo)
				for (int io = -il; io <= il; ++io) {
					const double v = at(u, i + io, j + jo, k + ko, nu);
					umin = mn(umin, v);
					umax = mx(umax, v);
				}
		if (dumax * alpha > (umax - uc))
			alpha = (umax - uc) / dumax;
		if (dumax * alpha > (uc - umin))
			alpha = (uc - umin) / dumax;
	}
	for (int d = 0; d < 3; ++d)
		sl[d] *= alpha;
}

// all fine children of coarse cell (ic,jc,kc) that lie in `region` (and in P.dest), all components
QK_AHD void interp_coarse_cell(const V4 &crse, const V4 &fine, int ic, int jc, int kc, const Box &region, const InterpParams &P)
{
	double sf[3] = {1.0, 1.0, 1.0};
	for (int ns = 0; ns < P.ncomp; ++ns) { // :90-98
		double dc[3], sl[3];
		cell_slopes(crse, ic, jc, kc, ns, P, dc, sl);
		for (int d = 0; d < 3; ++d)
			if (dc[d] != 0.0)
				sf[d] = mn(sf[d], sl[d] / dc[d]);
	}
	int lo[3], hi[3];
	const int c[3] = {ic, jc, kc};
	for (int d = 0; d < 3; ++d) {
		lo[d] = c[d] * P.ratio[d];
		hi[d] = lo[d] + P.ratio[d] - 1;
		const int rlo = (region.lo[d] > P.dest.lo[d]) ? region.lo[d] : P.dest.lo[d];
		const int rhi = (region.hi[d] < P.dest.hi[d]) ? region.hi[d] : P.dest.hi[d];
		lo[d] = (lo[d] < rlo) ? rlo : lo[d];
		hi[d] = (hi[d] > rhi) ? rhi : hi[d];
	}
	if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2])
		return;
	for (int ns = 0; ns < P.ncomp; ++ns) {
		double dc[3], sl[3];
		cell_slopes(crse, ic, jc, kc, ns, P, dc, sl);
		// slope(i,j,k,ns + d ncomp) = dc[d] (unlimited), then *= sf[d]  :33,102-106; zero where ratio == 1
		const double sx = (P.ratio[0] > 1) ? dc[0] * sf[0] : 0.0 * sf[0];
		const double sy = (P.ratio[1] > 1) ? dc[1] * sf[1] : 0.0 * sf[1];
		const double sz = (P.ratio[2] > 1) ? dc[2] * sf[2] : 0.0 * sf[2];
		const double uc = at(crse, ic, jc, kc, P.ccomp + ns);
		for (int k = lo[2]; k <= hi[2]; ++k) {
			const double zoff = ((double)(k - kc * P.ratio[2]) + 0.5) / (double)P.ratio[2] - 0.5;
			for (int j = lo[1]; j <= hi[1]; ++j) {
				const double yoff = ((double)(j - jc * P.ratio[1]) + 0.5) / (double)P.ratio[1] - 0.5;
				for (int i = lo[0]; i <= hi[0]; ++i) {
					const double xoff = ((double)(i - ic * P.ratio[0]) + 0.5) / (double)P.ratio[0] - 0.5;
					at(fine, i, j, k, P.fcomp + ns) = uc + xoff * sx + yoff * sy + zoff * sz; // :258-261
				}
			}
		}
	}
}

// QuokkaSimulation::PreInterpState / PostInterpState  src/QuokkaSimulation.hpp:804-841 for one cell: gas total energy (component 4)
// <-> specific internal energy, the variable the reference interpolates instead of E
template <bool POST> QK_AHD void prepost_cell(const V4 &c, int i, int j, int k)
{
	const double rho = at(c, i, j, k, 0), px = at(c, i, j, k, 1), py = at(c, i, j, k, 2), pz = at(c, i, j, k, 3);
	const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
	double &E = at(c, i, j, k, 4);
	if (POST) {
		const double Eint = rho * E;
		E = Eint + kinetic_energy;
	} else {
		E = (E - kinetic_energy) / rho;
	}
}

// amrex_avgdown for one coarse cell and component
QK_AHD double avgdown_cell(const V4 &fine, int i, int j, int k, int n, const int ratio[3])
{
	const double volfrac = 1.0 / (double)(ratio[0] * ratio[1] * ratio[2]);
	double c = 0;
	for (int kr = 0; kr < ratio[2]; ++kr)
		for (int jr = 0; jr < ratio[1]; ++jr)
			for (int ir = 0; ir < ratio[0]; ++ir)
				c += at(fine, i * ratio[0] + ir, j * ratio[1] + jr, k * ratio[2] + kr, n);
	return volfrac * c;
}

// ---- time interpolation between two coarse states: amrex::FillPatcher::fill, AMReX_FillPatcher.H:340-387 ------------------------------
// which: 0 copy of src0, 1 copy of src1, 2 alpha * src0 + beta * src1 (products rounded separately, then the sum: no contraction)
QK_AHD int time_interp_branch(double t0, double t1, double time, bool have1)
{
	if (!have1)
		return 0;
	const double teps = fabs(t1 - t0) * 1.e-3;
	if (time > t0 - teps && time < t0 + teps)
		return 0;
	if (time > t1 - teps && time < t1 + teps)
		return 1;
	return 2;
}
QK_AHD double time_interp_value(int which, double alpha, double beta, double a0, double a1)
{
	if (which == 0)
		return a0;
	if (which == 1)
		return a1;
#ifdef __CUDA_ARCH__
	return __dadd_rn(__dmul_rn(alpha, a0), __dmul_rn(beta, a1));
#else
	return alpha * a0 + beta * a1; // host builds use -ffp-contract=off
#endif
}

// ---- regrid tagging ------------------------------------------------------------------------------------------------------------------
// QuokkaSimulation<SedovProblem>::ErrorEst, src/problems/HydroBlast3D/test_hydro3d_blast.cpp:118-151; P[0] = centre, then x+, x-, y+, y-, z+, z-
QK_AHD bool tag_pressure_gradient(const double P[7], double eta_threshold, double P_min)
{
	const double del_x = mx(fabs(P[1] - P[0]), fabs(P[0] - P[2]));
	const double del_y = mx(fabs(P[3] - P[0]), fabs(P[0] - P[4]));
	const double del_z = mx(fabs(P[5] - P[0]), fabs(P[0] - P[6]));
	// std::max({a, b, c}): the first of the largest
	double m = del_x;
	if (m < del_y)
		m = del_y;
	if (m < del_z)
		m = del_z;
	const double gradient_indicator = m / P[0];
	return (gradient_indicator > eta_threshold) && (P[0] > P_min);
}
// QuokkaSimulation<ShocktubeProblem>::ErrorEst, src/problems/HydroShocktube/test_hydro_shocktube.cpp:146-171
QK_AHD bool tag_gradient_x(double qm, double q0, double qp, double dx, double eta_threshold, double q_min)
{
	const double del_x = (qp - qm) / (2.0 * dx);
	const double gradient_indicator = sqrt(del_x * del_x) / q0;
	return (gradient_indicator > eta_threshold) && (q0 >= q_min);
}
} // namespace qk_amr
