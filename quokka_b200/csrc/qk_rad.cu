// qk_rad.cu -- two-moment (M1) radiation transport sweep: RadSystem<problem_t> of src/radiation/radiation_system.hpp
// as driven by QuokkaSimulation::advanceRadiationForwardEuler / advanceRadiationMidpointRK2 and fluxFunction<DIR>
// (src/QuokkaSimulation.hpp:1791-1862, 1903-1986).
//
//   per-operator kernels (parity harness, operator-level drop-in): k_rad_prim_op, k_rad_flux_op<DIR>, k_rad_predict_op,
//                                                                  k_rad_rk2_op
//   fused stage (qk_rad_advance_stage):  k_rad_prim   (E_r, F) -> (E_r, f = F / (c E_r)) on the valid box grown by 3, all boxes
//                                        k_rad_stage  one CTA per 32 x 4 x 4 tile of cells: every thread reconstructs and
//                                                     solves the HLL problem of the three LOW faces of its cell (PPM /
//                                                     PLM-MC / donor cell on the reduced-flux primitives, Levermore closure,
//                                                     frozen Eddington tensor), the tile's high faces are done by a second
//                                                     small pass, fluxes meet in shared memory, and the conservative
//                                                     update (PredictStep or AddFluxesRK2 + isStateValid / amendRadState)
//                                                     is written straight to the new state: no left/right/flux arrays.
//
// Arithmetic: the reference's IEEE operation order, --fmad=false.  Parity with the oracle is bit for bit
// (tests/test_gpu_radiation.py).
#include "qk_level.h"
#include "qk_kernels.cuh"
#include "qk_fast.cuh"
#include "qk_tma.cuh"

#include <algorithm>
#include <stdlib.h>

#include "qk_rad_kernels.cuh"

namespace
{

// ---------------------------------------------------------------------------------------------------------------
// per-operator kernels
// ---------------------------------------------------------------------------------------------------------------
// RadSystem::ConservedToPrimitive  :589-614
__global__ void __launch_bounds__(TPB) k_rad_prim_op(RadConst c, Iter it, A4 cons, A4 prim)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t o = cons.off(i, j, k), op = prim.off(i, j, k);
	for (int g = 0; g < c.ng; ++g) {
		const double *u = cons.p + o + (c.nstart + 4 * g) * cons.ns;
		const double E_r = u[0], Fx = u[cons.ns], Fy = u[2 * cons.ns], Fz = u[3 * cons.ns];
		double *q = prim.p + op + 4 * g * prim.ns;
		q[0] = E_r;
		q[prim.ns] = Fx / (c.c * E_r);
		q[2 * prim.ns] = Fy / (c.c * E_r);
		q[3 * prim.ns] = Fz / (c.c * E_r);
	}
}

// RadSystem::ComputeFluxes<DIR>  :985-1139 from materialised left/right states
template <int DIR> __global__ void __launch_bounds__(TPB) k_rad_flux_op(RadConst c, Iter it, A4 flux, A4 fdiff, bool have_fdiff, A4 left, A4 right, A4 cons)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t sc = (DIR == 0) ? 1 : (DIR == 1) ? cons.js : cons.ks;
	for (int g = 0; g < c.ng; ++g) {
		double L[4], R[4], F[4];
#pragma unroll
		for (int n = 0; n < 4; ++n) {
			L[n] = left(i, j, k, 4 * g + n);
			R[n] = right(i, j, k, 4 * g + n);
		}
		const double *cR = cons.p + cons.off(i, j, k) + (c.nstart + 4 * g) * cons.ns;
		const int64_t roff = (int64_t)(c.nstart + 4 * g) * cons.ns;
		rad_face_flux<DIR>(c, L, R, cR - sc, cR, cons.ns, F, i + j + k, roff);
		double Fd0 = F[0];
		if (have_fdiff && c.wsc && ((i + j + k) % 2) == 0) { // the "diffusive" flux is the expression with epsilon = 1 (:1130-1131)
			double Fd[4];
			rad_face_flux<DIR>(c, L, R, cR - sc, cR, cons.ns, Fd);
			Fd0 = Fd[0];
		}
#pragma unroll
		for (int n = 0; n < 4; ++n) {
			flux(i, j, k, 4 * g + n) = F[n];
			if (have_fdiff)
				fdiff(i, j, k, 4 * g + n) = (n == 0) ? Fd0 : F[n];
		}
	}
}

// RadSystem::PredictStep :667-710 (RK2 = false) / AddFluxesRK2 :712-771 (RK2 = true)
template <bool RK2>
__global__ void __launch_bounds__(TPB) k_rad_update_op(RadConst c, Iter it, A4 unew, A4 u0, A4 u1, A4 fxo, A4 fyo, A4 fzo, A4 fx, A4 fy, A4 fz, double dtdx,
						       double dtdy, double dtdz)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	double cons[4 * QK_MAX_GROUPS];
	const int nh = 4 * c.ng;
	for (int n = 0; n < nh; ++n) {
		const double FxU_1 = dtdx * (fx(i, j, k, n) - fx(i + 1, j, k, n));
		const double FyU_1 = dtdy * (fy(i, j, k, n) - fy(i, j + 1, k, n));
		const double FzU_1 = dtdz * (fz(i, j, k, n) - fz(i, j, k + 1, n));
		const double U_0 = u0(i, j, k, c.nstart + n);
		if (!RK2) {
			cons[n] = U_0 + (FxU_1 + FyU_1 + FzU_1);
		} else {
			const double IMEX_a32 = 0.5;
			const double U_1 = u1(i, j, k, c.nstart + n);
			const double FxU_0 = dtdx * (fxo(i, j, k, n) - fxo(i + 1, j, k, n));
			const double FyU_0 = dtdy * (fyo(i, j, k, n) - fyo(i, j + 1, k, n));
			const double FzU_0 = dtdz * (fzo(i, j, k, n) - fzo(i, j, k + 1, n));
			cons[n] = (1.0 - IMEX_a32) * U_0 + IMEX_a32 * U_1 + ((0.5 - IMEX_a32) * (FxU_0 + FyU_0 + FzU_0)) + (0.5 * (FxU_1 + FyU_1 + FzU_1));
		}
	}
	rad_validate<QK_MAX_GROUPS>(c, c.ng, cons);
	for (int n = 0; n < nh; ++n)
		unew(i, j, k, c.nstart + n) = cons[n];
}

// ---------------------------------------------------------------------------------------------------------------
// fused stage
// ---------------------------------------------------------------------------------------------------------------
constexpr int RTX = 32, RTY = 4, RTZ = 4;
constexpr int RSX = RTX + 1, RSY = RTY + 1, RSZ = RTZ + 1;
constexpr int RAD_SMEM_DOUBLES = 3 * 4 * RSX * RSY * RSZ;

struct RadBox {
	A4 U0, Us, Uo; // state_old, stage input (ghost-filled), stage output
	A4 prim;       // 4*NG reduced-flux primitives of Us (valid box grown by 3)
	A4 S0;	       // stage-1 flux divergence (FxU_0 + FyU_0 + FzU_0), 4*NG components on the valid cells
	int lo[3], hi[3];
};

__global__ void __launch_bounds__(256) k_rad_prim(RadConst c, const RadBox *__restrict__ boxes, int halo)
{
	const RadBox &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 1 + 2 * halo, ny = B.hi[1] - B.lo[1] + 1 + 2 * halo, nz = B.hi[2] - B.lo[2] + 1 + 2 * halo;
	const int64_t total = (int64_t)nx * ny * nz;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] - halo + (int)(t - jk * nx);
		const int k = B.lo[2] - halo + (int)(jk / ny);
		const int j = B.lo[1] - halo + (int)(jk - (jk / ny) * ny);
		const int64_t o = B.Us.off(i, j, k), op = B.prim.off(i, j, k);
		for (int g = 0; g < c.ng; ++g) {
			const double *u = B.Us.p + o + (c.nstart + 4 * g) * B.Us.ns;
			const double E_r = u[0], Fx = u[B.Us.ns], Fy = u[2 * B.Us.ns], Fz = u[3 * B.Us.ns];
			double *q = B.prim.p + op + 4 * g * B.prim.ns;
			q[0] = E_r;
			q[B.prim.ns] = Fx / (c.c * E_r);
			q[2 * B.prim.ns] = Fy / (c.c * E_r);
			q[3 * B.prim.ns] = Fz / (c.c * E_r);
		}
	}
}

// left state of face (cell-1 | cell) = a_plus of cell-1, right state = a_minus of cell; qp points at the cell on the high side
template <int ORDER> __device__ __forceinline__ void rad_recon_face(const double *qp, int64_t s, double &L, double &R)
{
	if (ORDER == 1) {
		L = qp[-s];
		R = qp[0];
	} else if (ORDER == 2) { // PLM, MC limiter, interface-centred (src/hyperbolic_system.hpp:243-246)
		const double qm2 = qp[-2 * s], qm1 = qp[-s], q0 = qp[0], qp1 = qp[s];
		L = qm1 + 0.25 * lim_MC(q0 - qm1, qm1 - qm2);
		R = q0 - 0.25 * lim_MC(qp1 - q0, q0 - qm1);
	} else {
		const double qm3 = qp[-3 * s], qm2 = qp[-2 * s], qm1 = qp[-s], q0 = qp[0], qp1 = qp[s], qp2 = qp[2 * s];
		const double if_m = ppm_iface(qm3, qm2, qm1, q0), if_0 = ppm_iface(qm2, qm1, q0, qp1), if_p = ppm_iface(qm1, q0, qp1, qp2);
		double am, ap;
		ppm_limit(qm2, qm1, q0, if_m, if_0, am, ap);
		L = ap;
		ppm_limit(qm1, q0, qp1, if_0, if_p, am, ap);
		R = am;
	}
}

template <int ORDER, int DIR> __device__ __forceinline__ void rad_face_of_cell(const RadConst &c, const RadBox &B, int g, int i, int j, int k, double *F)
{
	const A4 &q = B.prim;
	const int64_t s = (DIR == 0) ? 1 : (DIR == 1) ? q.js : q.ks;
	const double *qp = q.p + q.off(i, j, k) + 4 * g * q.ns;
	double L[4], R[4];
#pragma unroll
	for (int n = 0; n < 4; ++n)
		rad_recon_face<ORDER>(qp + n * q.ns, s, L[n], R[n]);
	const A4 &u = B.Us;
	const int64_t su = (DIR == 0) ? 1 : (DIR == 1) ? u.js : u.ks;
	const double *cR = u.p + u.off(i, j, k) + (c.nstart + 4 * g) * u.ns;
	rad_face_flux<DIR>(c, L, R, cR - su, cR, u.ns, F, i + j + k, (int64_t)(c.nstart + 4 * g) * u.ns);
}

__device__ __forceinline__ int rs_idx(int d, int n, int lz, int ly, int lx) { return (((d * 4 + n) * RSZ + lz) * RSY + ly) * RSX + lx; }

template <int ORDER, int STAGE, bool KEEP_S0, int NG>
__global__ void __launch_bounds__(RTX *RTY *RTZ) k_rad_stage(RadConst c, const RadBox *__restrict__ boxes, int tiles_y, int tiles_z, double dtdx, double dtdy,
							      double dtdz)
{
	extern __shared__ double sF[];
	const int box = blockIdx.z / tiles_z, tz = blockIdx.z - box * tiles_z;
	const RadBox &B = boxes[box];
	const int lx = threadIdx.x & 31, ly = (threadIdx.x >> 5) & 3, lz = threadIdx.x >> 7;
	const int i0 = B.lo[0] + blockIdx.x * RTX, j0 = B.lo[1] + blockIdx.y * RTY, k0 = B.lo[2] + tz * RTZ;
	if (i0 > B.hi[0] || j0 > B.hi[1] || k0 > B.hi[2])
		return; // whole CTA
	const int i = i0 + lx, j = j0 + ly, k = k0 + lz;
	const bool inx = (i <= B.hi[0]), iny = (j <= B.hi[1]), inz = (k <= B.hi[2]);
	const bool cell = inx && iny && inz;
	double cons[4 * NG];
#pragma unroll
	for (int g = 0; g < NG; ++g) {
		if (g > 0)
			__syncthreads(); // shared fluxes of the previous group have been consumed
		// pass 1: the three low faces of this thread's cell (a cell one past the box edge along d still owns the box's last d-face)
		double F[4];
		if ((i <= B.hi[0] + 1) && iny && inz) {
			rad_face_of_cell<ORDER, 0>(c, B, g, i, j, k, F);
#pragma unroll
			for (int n = 0; n < 4; ++n)
				sF[rs_idx(0, n, lz, ly, lx)] = F[n];
		}
		if (inx && (j <= B.hi[1] + 1) && inz) {
			rad_face_of_cell<ORDER, 1>(c, B, g, i, j, k, F);
#pragma unroll
			for (int n = 0; n < 4; ++n)
				sF[rs_idx(1, n, lz, ly, lx)] = F[n];
		}
		if (inx && iny && (k <= B.hi[2] + 1)) {
			rad_face_of_cell<ORDER, 2>(c, B, g, i, j, k, F);
#pragma unroll
			for (int n = 0; n < 4; ++n)
				sF[rs_idx(2, n, lz, ly, lx)] = F[n];
		}
		// pass 2: the faces on the high side of the tile
		const int t = threadIdx.x;
		if (t < RTY * RTZ) { // x faces at i0 + RTX
			const int py = t & 3, pz = t >> 2;
			const int fi = i0 + RTX, fj = j0 + py, fk = k0 + pz;
			if (fi <= B.hi[0] + 1 && fj <= B.hi[1] && fk <= B.hi[2]) {
				rad_face_of_cell<ORDER, 0>(c, B, g, fi, fj, fk, F);
#pragma unroll
				for (int n = 0; n < 4; ++n)
					sF[rs_idx(0, n, pz, py, RTX)] = F[n];
			}
		} else if (t >= 32 && t < 32 + RTX * RTZ) { // y faces at j0 + RTY
			const int u = t - 32, px = u & 31, pz = u >> 5;
			const int fi = i0 + px, fj = j0 + RTY, fk = k0 + pz;
			if (fi <= B.hi[0] && fj <= B.hi[1] + 1 && fk <= B.hi[2]) {
				rad_face_of_cell<ORDER, 1>(c, B, g, fi, fj, fk, F);
#pragma unroll
				for (int n = 0; n < 4; ++n)
					sF[rs_idx(1, n, pz, RTY, px)] = F[n];
			}
		} else if (t >= 32 + RTX * RTZ && t < 32 + RTX * RTZ + RTX * RTY) { // z faces at k0 + RTZ
			const int u = t - 32 - RTX * RTZ, px = u & 31, py = u >> 5;
			const int fi = i0 + px, fj = j0 + py, fk = k0 + RTZ;
			if (fi <= B.hi[0] && fj <= B.hi[1] && fk <= B.hi[2] + 1) {
				rad_face_of_cell<ORDER, 2>(c, B, g, fi, fj, fk, F);
#pragma unroll
				for (int n = 0; n < 4; ++n)
					sF[rs_idx(2, n, RTZ, py, px)] = F[n];
			}
		}
		__syncthreads();
		if (cell) {
			const int64_t o0 = B.U0.off(i, j, k) + (c.nstart + 4 * g) * B.U0.ns;
			const int64_t os = B.S0.off(i, j, k) + 4 * g * B.S0.ns;
#pragma unroll
			for (int n = 0; n < 4; ++n) {
				const double FxU = dtdx * (sF[rs_idx(0, n, lz, ly, lx)] - sF[rs_idx(0, n, lz, ly, lx + 1)]);
				const double FyU = dtdy * (sF[rs_idx(1, n, lz, ly, lx)] - sF[rs_idx(1, n, lz, ly + 1, lx)]);
				const double FzU = dtdz * (sF[rs_idx(2, n, lz, ly, lx)] - sF[rs_idx(2, n, lz + 1, ly, lx)]);
				const double div = FxU + FyU + FzU;
				const double U_0 = B.U0.p[o0 + n * B.U0.ns];
				if (STAGE == 1) {
					cons[4 * g + n] = U_0 + div; // PredictStep :694-698
					if (KEEP_S0)
						B.S0.p[os + n * B.S0.ns] = div;
				} else { // AddFluxesRK2 :757-759
					const double IMEX_a32 = 0.5;
					const double U_1 = B.Us.p[B.Us.off(i, j, k) + (c.nstart + 4 * g + n) * B.Us.ns];
					const double div0 = B.S0.p[os + n * B.S0.ns];
					cons[4 * g + n] = (1.0 - IMEX_a32) * U_0 + IMEX_a32 * U_1 + ((0.5 - IMEX_a32) * div0) + (0.5 * div);
				}
			}
		}
	}
	if (cell) {
		rad_validate<NG>(c, NG, cons);
		const int64_t oo = B.Uo.off(i, j, k) + c.nstart * B.Uo.ns;
#pragma unroll
		for (int n = 0; n < 4 * NG; ++n)
			B.Uo.p[oo + n * B.Uo.ns] = cons[n];
	}
}

// ---------------------------------------------------------------------------------------------------------------
// fused stage, direction-split form (the default): one parabola per cell per direction, no shared memory, no block barriers
//   k_rad_x   lane <-> cell x0-1+lane of a row tile of 30 cells: parabola of the own cell, left state and the next face's
//             flux by warp shuffle; writes acc = FxU = (dt/dx)(F_i - F_{i+1})
//   k_rad_m   y / z by MARCHING: lane <-> x (coalesced rows), each thread walks a 32-cell segment keeping a 5-row window of the
//             primitives, the previous cell's right state and the previous face's flux in registers; y adds FyU to acc, the z
//             instance forms (FxU + FyU) + FzU (the reference's association, :694-698 / :748-759) and carries the update
// One photon group per launch (groups are independent in the transport step); with several groups the admissibility fix-up, which
// looks at all groups of a cell (isStateValid :624-643), runs as k_rad_fix afterwards.
// ---------------------------------------------------------------------------------------------------------------
template <int ORDER> __global__ void __launch_bounds__(128) k_rad_x(RadConst c, const RadBox2 *__restrict__ boxes, int g, double dtdx)
{
	const RadBox2 &B = boxes[blockIdx.z];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int ny = B.hi[1] - B.lo[1] + 1, nz = B.hi[2] - B.lo[2] + 1;
	const int row = blockIdx.x * 4 + warp; // rows in grid.x: a 512 x 512 cross-section has more row groups than grid.y allows
	if (row >= ny * nz)
		return; // whole warp
	const int j = B.lo[1] + row % ny, k = B.lo[2] + row / ny;
	const int x0 = B.lo[0] + blockIdx.y * 30;
	if (x0 > B.hi[0])
		return;
	const int i = x0 - 1 + lane;
	const int ic = (i <= B.hi[0] + 1) ? i : B.hi[0] + 1; // cells lo-1 .. hi+1 carry a parabola
	const A4 &q = B.prim;
	const double *qp = q.p + q.off(ic, j, k) + 4 * g * q.ns;
	double am[4], ap[4], Ls[4];
#pragma unroll
	for (int n = 0; n < 4; ++n) {
		const double *p = qp + n * q.ns;
		rad_cell_parabola<ORDER>(p[-2], p[-1], p[0], p[1], p[2], am[n], ap[n]);
		Ls[n] = rshfl_up1(ap[n]);
	}
	const bool face_ok = (lane >= 1) && (i >= B.lo[0]) && (i <= B.hi[0] + 1);
	double F[4] = {0., 0., 0., 0.};
	if (face_ok) {
		const A4 &u = B.Us;
		const double *cR = u.p + u.off(i, j, k) + (c.nstart + 4 * g) * u.ns;
		rad_face_flux<0>(c, Ls, am, cR - 1, cR, u.ns, F, i + j + k, (int64_t)(c.nstart + 4 * g) * u.ns);
	}
	const bool upd = (lane >= 1) && (lane <= 30) && (i <= B.hi[0]);
	const A4 &a = B.acc;
	const int64_t oa = upd ? a.off(i, j, k) + 4 * g * a.ns : 0;
#pragma unroll
	for (int n = 0; n < 4; ++n) {
		const double Fn = rshfl_dn1(F[n]);
		if (upd)
			a.p[oa + n * a.ns] = dtdx * (F[n] - Fn);
	}
}

// STAGE / KEEP_S0 / FIX only matter when LAST
template <int DIR, int ORDER, bool LAST, int STAGE, bool KEEP_S0, bool FIX>
__global__ void __launch_bounds__(128, 4) k_rad_m(RadConst c, const RadBox2 *__restrict__ boxes, int nseg, int g, double dtd)
{
	constexpr int TD = (DIR == 1) ? 2 : 1;
	const int box = blockIdx.z / nseg, seg = blockIdx.z - box * nseg;
	const RadBox2 &B = boxes[box];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int i = B.lo[0] + blockIdx.x * 32 + lane;
	const int t = B.lo[TD] + blockIdx.y * 4 + warp;
	const int s0 = B.lo[DIR] + seg * RSEG;
	if (i > B.hi[0] || t > B.hi[TD] || s0 > B.hi[DIR])
		return;
	const int s1 = min(s0 + RSEG, B.hi[DIR] + 1); // cells s0 .. s1-1 are updated, faces s0 .. s1 evaluated
	const A4 &q = B.prim;
	const int64_t sN = (DIR == 1) ? q.js : q.ks;
	int idx[3];
	idx[0] = i;
	idx[TD] = t;
	idx[DIR] = s0 - 1;
	const double *qp = q.p + q.off(idx[0], idx[1], idx[2]) + 4 * g * q.ns;
	const A4 &u = B.Us;
	const int64_t suN = (DIR == 1) ? u.js : u.ks;
	const double *cu = u.p + u.off(idx[0], idx[1], idx[2]) + (c.nstart + 4 * g) * u.ns;
	const A4 &a = B.acc;
	const int64_t saN = (DIR == 1) ? a.js : a.ks;
	int64_t oa = a.off(idx[0], idx[1], idx[2]) + 4 * g * a.ns;
	const int64_t s0N = (DIR == 1) ? B.S0.js : B.S0.ks, u0N = (DIR == 1) ? B.U0.js : B.U0.ks, uoN = (DIR == 1) ? B.Uo.js : B.Uo.ks;
	int64_t os = B.S0.off(idx[0], idx[1], idx[2]) + 4 * g * B.S0.ns;
	int64_t o0 = B.U0.off(idx[0], idx[1], idx[2]) + (c.nstart + 4 * g) * B.U0.ns;
	int64_t oo = B.Uo.off(idx[0], idx[1], idx[2]) + (c.nstart + 4 * g) * B.Uo.ns;
	// 5-row window of the four primitives around cell s0-1
	double w[4][5];
#pragma unroll
	for (int n = 0; n < 4; ++n)
#pragma unroll
		for (int m = 0; m < 5; ++m)
			w[n][m] = qp[n * q.ns + (m - 2) * sN];
	double apL[4] = {0., 0., 0., 0.}, Fp[4] = {0., 0., 0., 0.};
	for (int s = s0 - 1; s <= s1; ++s) {
		double am[4], ap[4];
#pragma unroll
		for (int n = 0; n < 4; ++n)
			rad_cell_parabola<ORDER>(w[n][0], w[n][1], w[n][2], w[n][3], w[n][4], am[n], ap[n]);
		if (s >= s0) {
			double F[4];
			rad_face_flux<DIR>(c, apL, am, cu - suN, cu, u.ns, F, i + t + s, (int64_t)(c.nstart + 4 * g) * u.ns);
			if (s > s0) { // cell s-1: both faces known
				const int64_t oac = oa - saN;
				double cons[4];
#pragma unroll
				for (int n = 0; n < 4; ++n) {
					const double d = dtd * (Fp[n] - F[n]);
					const double sum = a.p[oac + n * a.ns] + d;
					if (!LAST) {
						a.p[oac + n * a.ns] = sum;
					} else {
						const double U_0 = B.U0.p[(o0 - u0N) + n * B.U0.ns];
						if (STAGE == 1) {
							cons[n] = U_0 + sum; // PredictStep :694-698
							if (KEEP_S0)
								B.S0.p[(os - s0N) + n * B.S0.ns] = sum;
						} else { // AddFluxesRK2 :757-759
							const double IMEX_a32 = 0.5;
							const double U_1 = cu[-suN + n * u.ns];
							const double div0 = B.S0.p[(os - s0N) + n * B.S0.ns];
							cons[n] = (1.0 - IMEX_a32) * U_0 + IMEX_a32 * U_1 + ((0.5 - IMEX_a32) * div0) + (0.5 * sum);
						}
					}
				}
				if (LAST) {
					if (FIX)
						rad_validate<1>(c, 1, cons);
#pragma unroll
					for (int n = 0; n < 4; ++n)
						B.Uo.p[(oo - uoN) + n * B.Uo.ns] = cons[n];
				}
			}
#pragma unroll
			for (int n = 0; n < 4; ++n)
				Fp[n] = F[n];
		}
#pragma unroll
		for (int n = 0; n < 4; ++n) {
			apL[n] = ap[n];
#pragma unroll
			for (int m = 0; m < 4; ++m)
				w[n][m] = w[n][m + 1];
		}
		qp += sN;
		cu += suN;
		oa += saN;
		os += s0N;
		o0 += u0N;
		oo += uoN;
		if (s < s1) { // next cell's +2 row
#pragma unroll
			for (int n = 0; n < 4; ++n)
				w[n][4] = qp[n * q.ns + 2 * sN];
		}
	}
}

// isStateValid / amendRadState over all groups of a cell, in place (several photon groups only)
__global__ void __launch_bounds__(256) k_rad_fix(RadConst c, const RadBox2 *__restrict__ boxes)
{
	const RadBox2 &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 1, ny = B.hi[1] - B.lo[1] + 1, nz = B.hi[2] - B.lo[2] + 1;
	const int64_t total = (int64_t)nx * ny * nz;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] + (int)(t - jk * nx);
		const int k = B.lo[2] + (int)(jk / ny);
		const int j = B.lo[1] + (int)(jk - (jk / ny) * ny);
		const int64_t o = B.Uo.off(i, j, k) + c.nstart * B.Uo.ns;
		double cons[4 * QK_MAX_GROUPS];
		for (int n = 0; n < 4 * c.ng; ++n)
			cons[n] = B.Uo.p[o + n * B.Uo.ns];
		rad_validate<QK_MAX_GROUPS>(c, c.ng, cons);
		for (int n = 0; n < 4 * c.ng; ++n)
			B.Uo.p[o + n * B.Uo.ns] = cons[n];
	}
}
} // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
#define QK_TRY(x)                                                                                                                                    \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

extern "C" int qk_rad_conserved_to_primitive(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim,
					     int nghost, void *stream)
{
	QK_TRY(check_rad(prm));
	const RadConst c = make_rad_const(prm);
	ProfScope prof_("rad_cons_to_prim", S(stream));
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(nghost));
		k_rad_prim_op<<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(cons[b]), A4(prim[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_rad_compute_fluxes(const qk_rad_params *prm, int dir, int nboxes, const qk_box *valid, const qk_array4 *flux,
				     const qk_array4 *flux_diffusive, const qk_array4 *left, const qk_array4 *right, const qk_array4 *cons, void *stream)
{
	QK_TRY(check_rad(prm));
	if (dir < 0 || dir > 2)
		return QK_ERR_BAD_ARG;
	if (prm->use_wavespeed_correction && prm->ngroups != 1)
		return QK_ERR_UNSUPPORTED; // the multigroup optical depth is not built
	const RadConst c = make_rad_const(prm);
	ProfScope prof_("rad_compute_fluxes", S(stream));
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).face(dir));
		const bool hd = (flux_diffusive != nullptr);
		const A4 fd = hd ? A4(flux_diffusive[b]) : A4(flux[b]);
		if (dir == 0)
			k_rad_flux_op<0><<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(flux[b]), fd, hd, A4(left[b]), A4(right[b]), A4(cons[b]));
		else if (dir == 1)
			k_rad_flux_op<1><<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(flux[b]), fd, hd, A4(left[b]), A4(right[b]), A4(cons[b]));
		else
			k_rad_flux_op<2><<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(flux[b]), fd, hd, A4(left[b]), A4(right[b]), A4(cons[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_rad_predict_step(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons_old, const qk_array4 *cons_new,
				   const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz, double dt, const double dx[3], void *stream)
{
	QK_TRY(check_rad(prm));
	const RadConst c = make_rad_const(prm);
	ProfScope prof_("rad_predict_step", S(stream));
	for (int b = 0; b < nboxes; ++b) {
		Iter it{Box3(valid[b])};
		k_rad_update_op<false><<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(cons_new[b]), A4(cons_old[b]), A4(cons_old[b]), A4(fx[b]), A4(fy[b]),
									   A4(fz[b]), A4(fx[b]), A4(fy[b]), A4(fz[b]), dt / dx[0], dt / dx[1], dt / dx[2]);
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_rad_add_fluxes_rk2(const qk_rad_params *prm, int nboxes, const qk_box *valid, const qk_array4 *u_new, const qk_array4 *u0,
				     const qk_array4 *u1, const qk_array4 *fx_old, const qk_array4 *fy_old, const qk_array4 *fz_old, const qk_array4 *fx,
				     const qk_array4 *fy, const qk_array4 *fz, double dt, const double dx[3], void *stream)
{
	QK_TRY(check_rad(prm));
	const RadConst c = make_rad_const(prm);
	ProfScope prof_("rad_add_fluxes_rk2", S(stream));
	for (int b = 0; b < nboxes; ++b) {
		Iter it{Box3(valid[b])};
		k_rad_update_op<true><<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(u_new[b]), A4(u0[b]), A4(u1[b]), A4(fx_old[b]), A4(fy_old[b]), A4(fz_old[b]),
									  A4(fx[b]), A4(fy[b]), A4(fz[b]), dt / dx[0], dt / dx[1], dt / dx[2]);
		QK_KERNEL_CHECK();
	}
	return 0;
}

// qk_rad_relaxed.cu: the TMA-staged sweeps instantiated with relaxed arithmetic (compiled with FMA contraction)
int qk_rad_stage_relaxed(int order, const void *rad_const, const void *boxes, const void *maps, int nb, const int maxn[3], int stage, bool keep, bool fix, int g,
			 double dtdx, double dtdy, double dtdz, cudaStream_t s);

// ---- fused stage ------------------------------------------------------------------------------------------------
struct RadState {
	int nh = 0; // 4 * ngroups the scratch was built for
	std::vector<qk_array4> prim, S0, acc;
	RadBox *d_boxes = nullptr, *h_boxes = nullptr; // ring of 8 tables
	RadBox2 *d_boxes2 = nullptr, *h_boxes2 = nullptr;
	int ring = 0;
	cudaEvent_t ev[8];
	bool ev_used[8];
	bool s0_valid = false;
	std::vector<RadMaps> maps; // tensor-map descriptors of the stage being launched, one entry per chunk of TMAP_MAXB boxes
};

void qk_rad_free(qk_level *L)
{
	if (!L->rad)
		return;
	RadState *R = L->rad;
	if (R->d_boxes)
		cudaFree(R->d_boxes);
	if (R->d_boxes2)
		cudaFree(R->d_boxes2);
	if (R->h_boxes2)
		cudaFreeHost(R->h_boxes2);
	if (R->h_boxes) {
		cudaFreeHost(R->h_boxes);
		for (int i = 0; i < 8; ++i)
			cudaEventDestroy(R->ev[i]);
	}
	delete R;
	L->rad = nullptr;
}

static int rad_setup(qk_level *L, int nh)
{
	if (L->rad && L->rad->nh == nh)
		return 0;
	if (L->rad)
		return QK_ERR_UNSUPPORTED;
	RadState *R = new RadState();
	L->rad = R;
	const int nb = (int)L->valid.size();
	QK_TRY(L->alloc_fabs(R->prim, nh, 3, -1));
	QK_TRY(L->alloc_fabs(R->S0, nh, 0, -1));
	QK_TRY(L->alloc_fabs(R->acc, nh, 0, -1));
	QK_CUDA(cudaMalloc(&R->d_boxes, sizeof(RadBox) * nb * 8));
	QK_CUDA(cudaMallocHost(&R->h_boxes, sizeof(RadBox) * nb * 8));
	QK_CUDA(cudaMalloc(&R->d_boxes2, sizeof(RadBox2) * nb * 8));
	QK_CUDA(cudaMallocHost(&R->h_boxes2, sizeof(RadBox2) * nb * 8));
	for (int i = 0; i < 8; ++i) {
		QK_CUDA(cudaEventCreateWithFlags(&R->ev[i], cudaEventDisableTiming));
		R->ev_used[i] = false;
	}
	R->nh = nh;
	return 0;
}

template <int ORDER, int NG>
static int launch_rad_stage(const RadConst &c, const RadBox *tab, int nb, const int maxn[3], int stage, bool keep, double dtdx, double dtdy, double dtdz,
			    cudaStream_t s)
{
	const int tx = (maxn[0] + RTX - 1) / RTX, ty = (maxn[1] + RTY - 1) / RTY, tz = (maxn[2] + RTZ - 1) / RTZ;
	dim3 grid(tx, ty, tz * nb);
	const size_t smem = sizeof(double) * RAD_SMEM_DOUBLES;
#define QK_RAD_LAUNCH(ST, KP)                                                                                                                        \
	do {                                                                                                                                         \
		auto kern = k_rad_stage<ORDER, ST, KP, NG>;                                                                                          \
		static bool attr_set = false;                                                                                                        \
		if (!attr_set) {                                                                                                                     \
			QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                                 \
			attr_set = true;                                                                                                             \
		}                                                                                                                                    \
		kern<<<grid, RTX * RTY * RTZ, smem, s>>>(c, tab, ty, tz, dtdx, dtdy, dtdz);                                                          \
	} while (0)
	if (stage == 1 && keep)
		QK_RAD_LAUNCH(1, true);
	else if (stage == 1)
		QK_RAD_LAUNCH(1, false);
	else
		QK_RAD_LAUNCH(2, false);
#undef QK_RAD_LAUNCH
	QK_KERNEL_CHECK();
	return 0;
}

template <int NG>
static int dispatch_rad_order(int order, const RadConst &c, const RadBox *tab, int nb, const int maxn[3], int stage, bool keep, double dtdx, double dtdy,
			      double dtdz, cudaStream_t s)
{
	if (order == 3)
		return launch_rad_stage<3, NG>(c, tab, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
	if (order == 2)
		return launch_rad_stage<2, NG>(c, tab, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
	return launch_rad_stage<1, NG>(c, tab, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
}

// direction-split launches of one stage for photon group g
template <int ORDER>
static int launch_rad_split(const RadConst &c, const RadBox2 *tab, int nb, const int maxn[3], int stage, bool keep, bool fix, int g, double dtdx, double dtdy,
			    double dtdz, cudaStream_t s)
{
	{
		dim3 grid((maxn[1] * maxn[2] + 3) / 4, (maxn[0] + 29) / 30, nb);
		k_rad_x<ORDER><<<grid, 128, 0, s>>>(c, tab, g, dtdx);
		QK_KERNEL_CHECK();
	}
	{
		const int nseg = (maxn[1] + RSEG - 1) / RSEG;
		dim3 grid((maxn[0] + 31) / 32, (maxn[2] + 3) / 4, nb * nseg);
		k_rad_m<1, ORDER, false, 1, false, false><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdy);
		QK_KERNEL_CHECK();
	}
	{
		const int nseg = (maxn[2] + RSEG - 1) / RSEG;
		dim3 grid((maxn[0] + 31) / 32, (maxn[1] + 3) / 4, nb * nseg);
		if (stage == 1 && keep) {
			if (fix)
				k_rad_m<2, ORDER, true, 1, true, true><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
			else
				k_rad_m<2, ORDER, true, 1, true, false><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
		} else if (stage == 1) {
			if (fix)
				k_rad_m<2, ORDER, true, 1, false, true><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
			else
				k_rad_m<2, ORDER, true, 1, false, false><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
		} else {
			if (fix)
				k_rad_m<2, ORDER, true, 2, false, true><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
			else
				k_rad_m<2, ORDER, true, 2, false, false><<<grid, 128, 0, s>>>(c, tab, nseg, g, dtdz);
		}
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_rad_advance_stage(qk_level *L, const qk_rad_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout,
				    double dt, void *stream)
{
	if (!L || !U0 || !Ustage || !Uout || (stage != 1 && stage != 2))
		return QK_ERR_BAD_ARG;
	QK_TRY(check_rad(prm));
	if (!L->has_device)
		return QK_ERR_NO_DEVICE;
	const int ng = prm->ngroups;
	const bool tile_form = (getenv("QK_RAD_TILE") != nullptr); // the first-generation one-kernel form (kept for comparison)
	if (tile_form && ng != 1 && ng != 2 && ng != 4)
		return QK_ERR_UNSUPPORTED; // instantiated group counts of the tile kernel
	if (L->nghost < 3 || prm->nstart + 4 * ng > L->ncomp)
		return QK_ERR_BAD_ARG;
	cudaStream_t s = S(stream);
	QK_TRY(rad_setup(L, 4 * ng));
	RadState *R = L->rad;
	const int nb = (int)L->valid.size();
	if (stage == 2 && !R->s0_valid)
		return QK_ERR_BAD_ARG; // stage 2 needs the stage-1 flux divergence of the same step (the relaxed sweeps drop the term but keep the protocol)
	for (int b = 0; b < nb; ++b)
		if (Uout[b].p == Ustage[b].p || Uout[b].p == U0[b].p)
			return QK_ERR_BAD_ARG; // a tile reads its neighbours' cells of Ustage / U0 while other tiles write Uout
	const int slot = R->ring;
	R->ring = (R->ring + 1) % 8;
	if (R->ev_used[slot])
		QK_CUDA(cudaEventSynchronize(R->ev[slot]));
	RadBox *hb = R->h_boxes + (size_t)slot * nb;
	RadBox2 *hb2 = R->h_boxes2 + (size_t)slot * nb;
	int maxn[3] = {1, 1, 1};
	for (int b = 0; b < nb; ++b) {
		RadBox2 &B2 = hb2[b];
		B2.U0 = A4(U0[b]);
		B2.Us = A4(Ustage[b]);
		B2.Uo = A4(Uout[b]);
		B2.prim = A4(R->prim[b]);
		B2.S0 = A4(R->S0[b]);
		B2.acc = A4(R->acc[b]);
		B2.us_xend = Ustage[b].end[0];
		for (int d = 0; d < 3; ++d) {
			B2.lo[d] = L->valid[b].lo[d];
			B2.hi[d] = L->valid[b].hi[d];
		}
		RadBox &B = hb[b];
		B.U0 = A4(U0[b]);
		B.Us = A4(Ustage[b]);
		B.Uo = A4(Uout[b]);
		B.prim = A4(R->prim[b]);
		B.S0 = A4(R->S0[b]);
		for (int d = 0; d < 3; ++d) {
			B.lo[d] = L->valid[b].lo[d];
			B.hi[d] = L->valid[b].hi[d];
			maxn[d] = std::max(maxn[d], B.hi[d] - B.lo[d] + 1);
		}
	}
	RadBox *db = R->d_boxes + (size_t)slot * nb;
	RadBox2 *db2 = R->d_boxes2 + (size_t)slot * nb;
	QK_CUDA(cudaMemcpyAsync(db2, hb2, sizeof(RadBox2) * nb, cudaMemcpyHostToDevice, s));
	QK_CUDA(cudaMemcpyAsync(db, hb, sizeof(RadBox) * nb, cudaMemcpyHostToDevice, s));
	QK_CUDA(cudaEventRecord(R->ev[slot], s));
	R->ev_used[slot] = true;
	RadConst c = make_rad_const(prm);
	for (int d = 0; d < 3; ++d)
		c.dl[d] = L->dx[d]; // ComputeCellOpticalDepth of the stage uses the level's cell sizes
	if (tile_form && prm->use_wavespeed_correction)
		return QK_ERR_UNSUPPORTED;
	if (prm->use_wavespeed_correction && ng != 1)
		return QK_ERR_UNSUPPORTED; // the multigroup optical depth (DefineOpacityExponentsAndLowerValues) is not built
	// The TMA-staged sweeps (qk_rad_kernels.cuh) copy tiles of the caller's state and of the level's scratch through tensor maps: pitches must be
	// even (16-byte strides), the base 16-byte aligned, and the x sweep reads four ghost cells.  Anything else takes the first-generation
	// direction-split kernels below (global loads; exact arithmetic only).
	bool tma = !tile_form && (getenv("QK_RAD_V1") == nullptr) && (L->nghost >= 4);
	R->maps.resize((size_t)(nb + TMAP_MAXB - 1) / TMAP_MAXB);
	for (int b = 0; b < nb && tma; ++b) { // the tiles need the ghost cells to exist and every array to be describable (even pitches, aligned base)
		const int lo0 = L->valid[b].lo[0];
		TmapBytes *m = R->maps[b / TMAP_MAXB].m[b % TMAP_MAXB];
		tma = (lo0 - 4 >= Ustage[b].begin[0]) && (L->valid[b].lo[1] - 3 >= Ustage[b].begin[1]) && (L->valid[b].lo[2] - 3 >= Ustage[b].begin[2]) &&
		      qk_encode_tile(m[RM_US].b, Ustage[b], 32, 1, 1, 4) && qk_encode_tile(m[RM_USX].b, Ustage[b], 38, 1, 1, 4) &&
		      qk_encode_tile(m[RM_U0].b, U0[b], 32, 1, 1, 4) && qk_encode_tile(m[RM_ACC].b, R->acc[b], 32, 1, 1, 4) &&
		      qk_encode_tile(m[RM_S0].b, R->S0[b], 32, 1, 1, 4);
	}
	const bool relaxed = (prm->arith == QK_ARITH_FAST) && tma; // the relaxed arithmetic exists in the TMA-staged form only
	if (!tma) {
		ProfScope p("rad_prim", s);
		const int64_t cells = (int64_t)(maxn[0] + 6) * (maxn[1] + 6) * (maxn[2] + 6);
		dim3 grid((unsigned)std::min<int64_t>((cells + 255) / 256, 4096), nb);
		k_rad_prim<<<grid, 256, 0, s>>>(c, db, 3);
		QK_KERNEL_CHECK();
	}
	const bool keep = (stage == 1 && prm->integrator_order == 2);
	const double dtdx = dt / L->dx[0], dtdy = dt / L->dx[1], dtdz = dt / L->dx[2];
	int rc = 0;
	if (tma) {
		ProfScope p("rad_stage", s);
		const bool fix = (ng == 1);
		for (int g = 0; g < ng && rc == 0; ++g)
			rc = relaxed ? qk_rad_stage_relaxed(prm->reconstruction_order, &c, db2, R->maps.data(), nb, maxn, stage, keep, fix, g, dtdx, dtdy, dtdz, s)
				     : dispatch_rad_tma<0>(prm->reconstruction_order, c, db2, R->maps.data(), nb, maxn, stage, keep, fix, g, dtdx, dtdy, dtdz, s);
		if (rc == 0 && !fix) {
			const int64_t cells = (int64_t)maxn[0] * maxn[1] * maxn[2];
			dim3 grid((unsigned)std::min<int64_t>((cells + 255) / 256, 4096), nb);
			k_rad_fix<<<grid, 256, 0, s>>>(c, db2);
			QK_KERNEL_CHECK();
		}
	} else {
		ProfScope p("rad_stage", s);
		if (tile_form) {
			if (ng == 1)
				rc = dispatch_rad_order<1>(prm->reconstruction_order, c, db, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
			else if (ng == 2)
				rc = dispatch_rad_order<2>(prm->reconstruction_order, c, db, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
			else
				rc = dispatch_rad_order<4>(prm->reconstruction_order, c, db, nb, maxn, stage, keep, dtdx, dtdy, dtdz, s);
		} else {
			const bool fix = (ng == 1); // one group: the admissibility fix-up is cell-local to the z sweep's epilogue
			for (int g = 0; g < ng && rc == 0; ++g) {
				if (prm->reconstruction_order == 3)
					rc = launch_rad_split<3>(c, db2, nb, maxn, stage, keep, fix, g, dtdx, dtdy, dtdz, s);
				else if (prm->reconstruction_order == 2)
					rc = launch_rad_split<2>(c, db2, nb, maxn, stage, keep, fix, g, dtdx, dtdy, dtdz, s);
				else
					rc = launch_rad_split<1>(c, db2, nb, maxn, stage, keep, fix, g, dtdx, dtdy, dtdz, s);
			}
			if (rc == 0 && !fix) {
				const int64_t cells = (int64_t)maxn[0] * maxn[1] * maxn[2];
				dim3 grid((unsigned)std::min<int64_t>((cells + 255) / 256, 4096), nb);
				k_rad_fix<<<grid, 256, 0, s>>>(c, db2);
				QK_KERNEL_CHECK();
			}
		}
	}
	QK_TRY(rc);
	R->s0_valid = (stage == 1 && keep);
	return 0;
}
