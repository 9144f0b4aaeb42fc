// qk_sweep_kernels.cuh -- kernels and launchers of the fused stage, templated on the arithmetic mode ARITH (0 exact,
// 1 relaxed).  Included by exactly two translation units: qk_sweep.cu (ARITH = 0, compiled with --fmad=false) and
// qk_sweep_relaxed.cu (ARITH = 1, compiled with FMA contraction on).  See qk_sweep.cu for the description of the kernels.
#pragma once
#include "qk_relaxed.cuh"
#include "qk_level.h"

#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace
{
constexpr int SEG = 32; // cells per marching segment

struct SweepBox {
	A4 U0, Us, Uo; // state_old (ghost-filled), stage input (ghost-filled), stage output
	A4 prim;       // 6+NS primitives + chi_min as the last component; same index space as the state (ng ghosts)
	A4 chi3;       // chi_x, chi_y, chi_z (ng = 2)
	A4 rhs;        // 6+NS RHS components + div v as the last component (valid cells)
	A4 hF[3];      // per direction: 0.5*F(U0) (6+NS) + 0.5*faceVel(U0), nodal in that direction
	A4 fo[3];      // KEEPF instantiations only: the stage's own face fluxes F (6+NS), nodal in that direction, TIGHT rows (what
		       // incrementFluxRegisters reads through an alias FArrayBox, src/simulation.hpp:1345-1387)
	int lo[3], hi[3];
};

__device__ __forceinline__ bool nonfinite(double v) { return ((unsigned)__double2hiint(v) & 0x7ff00000u) == 0x7ff00000u; }

// ---- plain-`/` fallback for isolated quotients (taken when a fast quotient left its domain) ----
__device__ __noinline__ double slow_div(double a, double b) { return a / b; }

// ---------------------------------------------------------------------------------------------------------------
template <int ARITH, int NS, bool REINT> __global__ void __launch_bounds__(256) k_fprim(FastConst c, const SweepBox *__restrict__ boxes, int ng)
{
	const SweepBox &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 1 + 2 * ng, ny = B.hi[1] - B.lo[1] + 1 + 2 * ng, nz = B.hi[2] - B.lo[2] + 1 + 2 * ng;
	const int64_t total = (int64_t)nx * ny * nz;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] - ng + (int)(t - jk * nx);
		const int k = B.lo[2] - ng + (int)(jk / ny);
		const int j = B.lo[1] - ng + (int)(jk - (jk / ny) * ny);
		const int64_t o = B.Us.off(i, j, k), op = B.prim.off(i, j, k);
		double U[6 + NS], q[6 + NS];
#pragma unroll
		for (int n = 0; n < 6 + NS; ++n)
			U[n] = B.Us.p[o + n * B.Us.ns];
		if (ARITH == 0) {
			unsigned slow = 0;
			f_cons_to_prim<NS, REINT, true>(c, U, q, slow);
			if (slow)
				f_cons_to_prim<NS, REINT, false>(c, U, q, slow);
		} else {
			r_cons_to_prim<NS, REINT>(c, U, q);
		}
#pragma unroll
		for (int n = 0; n < 6 + NS; ++n)
			B.prim.p[op + n * B.prim.ns] = q[n];
	}
}

// chi along the three directions of one cell; rho c_s^2 is evaluated once (the reference recomputes it per direction)
template <int ARITH, int NS, bool REINT> __global__ void __launch_bounds__(256) k_fchi(FastConst c, const SweepBox *__restrict__ boxes)
{
	const SweepBox &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 5, ny = B.hi[1] - B.lo[1] + 5, nz = B.hi[2] - B.lo[2] + 5;
	const int64_t total = (int64_t)nx * ny * nz;
	const A4 &q = B.prim;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] - 2 + (int)(t - jk * nx);
		const int k = B.lo[2] - 2 + (int)(jk / ny);
		const int j = B.lo[1] - 2 + (int)(jk - (jk / ny) * ny);
		const int64_t o = q.off(i, j, k);
		const int64_t st[3] = {1, q.js, q.ks};
		double Pc[3][5];
		double vm1[3], vp1[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			vm1[d] = q.p[o - st[d] + (1 + d) * q.ns];
			vp1[d] = q.p[o + st[d] + (1 + d) * q.ns];
		}
		// chi is exactly 1 along every direction in which the flow does not converge (v(+1) >= v(-1), hydro_system.hpp:619-622): a cell that
		// converges along none of them needs neither its pressures nor its sound speed.  Same bits; gas at rest, expansion waves and the
		// interior of a blast take this exit.
		if (!((vp1[0] < vm1[0]) || (vp1[1] < vm1[1]) || (vp1[2] < vm1[2]))) {
			const int64_t oc1 = B.chi3.off(i, j, k);
#pragma unroll
			for (int d = 0; d < 3; ++d)
				B.chi3.p[oc1 + d * B.chi3.ns] = 1.0;
			continue;
		}
		const double rho = q.p[o];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
#pragma unroll
			for (int m = -2; m <= 2; ++m)
				Pc[d][m + 2] = q.p[o + m * st[d] + 4 * q.ns];
		}
		if (ARITH != 0) { // relaxed arithmetic (qk_relaxed.cuh): rho c_s^2 = gamma p in closed form
			if (REINT) {
#pragma unroll
				for (int d = 0; d < 3; ++d)
#pragma unroll
					for (int m = -2; m <= 2; ++m) {
						if (d > 0 && m == 0) {
							Pc[d][2] = Pc[0][2];
							continue;
						}
						const double r = q.p[o + m * st[d]];
						Pc[d][m + 2] = r_pressure_from_e(c, r, (r == 0.0) ? 0.0 : Pc[d][m + 2]);
					}
			}
			const double yKS = r_rcp(c.h.gamma * r_p_of_p(c, rho, Pc[0][2]));
			const int64_t ocr = B.chi3.off(i, j, k);
#pragma unroll
			for (int d = 0; d < 3; ++d)
				B.chi3.p[ocr + d * B.chi3.ns] = r_flatten_chi(c, Pc[d][0], Pc[d][1], Pc[d][3], Pc[d][4], yKS, vm1[d], vp1[d]);
			continue;
		}
		unsigned slow = 0;
		if (REINT) { // pressures from the specific internal energies (hydro_system.hpp:577-586)
#pragma unroll
			for (int d = 0; d < 3; ++d)
#pragma unroll
				for (int m = -2; m <= 2; ++m) {
					if (d > 0 && m == 0) {
						Pc[d][2] = Pc[0][2];
						continue;
					}
					const double r = q.p[o + m * st[d]];
					Pc[d][m + 2] = f_pressure_from_e<true>(c, r, (r == 0.0) ? 0.0 : div_d<true>(r * Pc[d][m + 2], r, slow), slow);
				}
		}
		const double cs = f_sound_speed<true>(c, rho, Pc[0][2], slow);
		const QkRcp RKS = rcp_f<true>((cs * cs) * rho, slow);
		double chi[3];
#pragma unroll
		for (int d = 0; d < 3; ++d)
			chi[d] = f_flatten_chi<true>(c, Pc[d][0], Pc[d][1], Pc[d][3], Pc[d][4], RKS, vm1[d], vp1[d], slow);
		if (slow) { // plain-division fallback (qk_physics.cuh)
			if (REINT) {
				for (int d = 0; d < 3; ++d)
					for (int m = -2; m <= 2; ++m) {
						const double r = q.p[o + m * st[d]];
						Pc[d][m + 2] = eos_pressure(c.h, r, r * q.p[o + m * st[d] + 4 * q.ns]);
					}
			}
			const double cs2 = eos_sound_speed(c.h, rho, Pc[0][2]);
			const double KS = (cs2 * cs2) * rho;
			for (int d = 0; d < 3; ++d)
				chi[d] = flatten_chi(Pc[d][0], Pc[d][1], Pc[d][3], Pc[d][4], KS, vm1[d], vp1[d]);
		}
		const int64_t oc = B.chi3.off(i, j, k);
#pragma unroll
		for (int d = 0; d < 3; ++d)
			B.chi3.p[oc + d * B.chi3.ns] = chi[d];
	}
}

// Relaxed arithmetic: the pre-minimised flattening coefficient of a cell in ONE pass -- the nine chi values (three cells along
// each axis) are evaluated from 7-point pressure and 5-point velocity lines held in registers, so the chi_x/chi_y/chi_z arrays
// are never written or re-read (k_fchi + k_fchimin: 0.48 ms per stage at 256^3).  With the closed-form EOS a chi evaluation is
// ~25 instructions; the exact mode, whose rho c_s^2 costs a sound-speed evaluation per cell, keeps the two-kernel form.
// (Round 2 tried a shared-memory tile form -- each directional chi evaluated once per cell, 3.8 evaluations per output instead of 9 -- and
// measured it SLOWER: 0.48 vs 0.42 ms per launch at 256^3; the index decode and the block barrier cost more than the saved evaluations.)
template <int NS, bool REINT> __global__ void __launch_bounds__(256) k_fchi9(FastConst c, const SweepBox *__restrict__ boxes)
{
	const SweepBox &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 3, ny = B.hi[1] - B.lo[1] + 3, nz = B.hi[2] - B.lo[2] + 3;
	const int64_t total = (int64_t)nx * ny * nz;
	const A4 &q = B.prim;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] - 1 + (int)(t - jk * nx);
		const int k = B.lo[2] - 1 + (int)(jk / ny);
		const int j = B.lo[1] - 1 + (int)(jk - (jk / ny) * ny);
		const int64_t o = q.off(i, j, k);
		const int64_t st[3] = {1, q.js, q.ks};
		double chi = 1.0;
		bool first = true;
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			double P[7], v[5], r3[3];
#pragma unroll
			for (int m = -2; m <= 2; ++m)
				v[m + 2] = q.p[o + m * st[d] + (1 + d) * q.ns];
			// chi is exactly 1 wherever the flow does not converge along d (v(+1) >= v(-1), hydro_system.hpp:619-622): none of the three
			// cells converging -> this direction contributes min(chi, 1) and its pressures are never read.  Gas at rest, expansion waves
			// and the interior of a blast take this exit; the value is the same bit for bit.
			if (!((v[2] < v[0]) || (v[3] < v[1]) || (v[4] < v[2]))) {
				chi = first ? 1.0 : dmin(chi, 1.0);
				first = false;
				continue;
			}
#pragma unroll
			for (int m = -3; m <= 3; ++m) {
				double pv = q.p[o + m * st[d] + 4 * q.ns];
				if (REINT) { // pressures from the specific internal energies (hydro_system.hpp:577-586)
					const double r = q.p[o + m * st[d]];
					pv = r_pressure_from_e(c, r, (r == 0.0) ? 0.0 : pv);
				}
				P[m + 3] = pv;
			}
#pragma unroll
			for (int m = -1; m <= 1; ++m)
				r3[m + 1] = q.p[o + m * st[d]];
#pragma unroll
			for (int m = -1; m <= 1; ++m) { // chi_d of cell + m along d, in the reference's order of the 9-point min (:655-669)
				const double yKS = r_rcp(c.h.gamma * r_p_of_p(c, r3[m + 1], P[m + 3]));
				const double x = r_flatten_chi(c, P[m + 1], P[m + 2], P[m + 4], P[m + 5], yKS, v[m + 1], v[m + 3]);
				chi = first ? x : dmin(chi, x);
				first = false;
			}
		}
		B.prim.p[o + (6 + NS) * B.prim.ns] = chi;
	}
}

// min over chi_x(i-1,i,i+1), chi_y(j-1,j,j+1), chi_z(k-1,k,k+1) in the reference's order (hydro_system.hpp:655-669)
template <int NS> __global__ void __launch_bounds__(256) k_fchimin(const SweepBox *__restrict__ boxes)
{
	const SweepBox &B = boxes[blockIdx.y];
	const int nx = B.hi[0] - B.lo[0] + 3, ny = B.hi[1] - B.lo[1] + 3, nz = B.hi[2] - B.lo[2] + 3;
	const int64_t total = (int64_t)nx * ny * nz;
	const A4 &x = B.chi3;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = B.lo[0] - 1 + (int)(t - jk * nx);
		const int k = B.lo[2] - 1 + (int)(jk / ny);
		const int j = B.lo[1] - 1 + (int)(jk - (jk / ny) * ny);
		const int64_t o = x.off(i, j, k);
		double chi = x.p[o - 1];
		chi = dmin(chi, x.p[o]);
		chi = dmin(chi, x.p[o + 1]);
		chi = dmin(chi, x.p[o - x.js + x.ns]);
		chi = dmin(chi, x.p[o + x.ns]);
		chi = dmin(chi, x.p[o + x.js + x.ns]);
		chi = dmin(chi, x.p[o - x.ks + 2 * x.ns]);
		chi = dmin(chi, x.p[o + 2 * x.ns]);
		chi = dmin(chi, x.p[o + x.ks + 2 * x.ns]);
		B.prim.p[B.prim.off(i, j, k) + (6 + NS) * B.prim.ns] = chi;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// per-cell pieces shared by the two sweep kernels
// ---------------------------------------------------------------------------------------------------------------
// transverse velocity-difference minima of one cell (hydro_system.hpp:1022-1033): mV = min(vV(+V)-vV, vV-vV(-V)), mW likewise
template <int DIR> __device__ __forceinline__ void cell_trans_min(const A4 &q, int64_t o, double &mV, double &mW)
{
	constexpr int aV = (DIR + 1) % 3, aW = (DIR + 2) % 3;
	const int64_t sV = (aV == 0) ? 1 : (aV == 1) ? q.js : q.ks;
	const int64_t sW = (aW == 0) ? 1 : (aW == 1) ? q.js : q.ks;
	const double *vV = q.p + o + (1 + aV) * q.ns;
	const double *vW = q.p + o + (1 + aW) * q.ns;
	const double v0 = vV[0], w0 = vW[0];
	mV = dmin(vV[sV] - v0, v0 - vV[-sV]);
	mW = dmin(vW[sW] - w0, w0 - vW[-sW]);
}

// stage epilogue of one cell: rhs (all three directions summed) -> new state (hydro_system.hpp:775-814, 475-497, 698-773, 816-850)
// BLEND (relaxed arithmetic, stage 2): r0 points at R(U0) of this cell (component stride 32) and the update uses 0.5 R(U0) + 0.5 R(U1)
template <int ARITH, int NS, int NMS, bool BLEND = false>
__device__ __forceinline__ void cell_epilogue(const FastConst &c, const double *U0, double *r, double divv, double *Un, int &bad, int &nonfin,
					      const double *r0 = nullptr)
{
	// AddInternalEnergyPdV: P from the OLD state; redoFlag is none on this path
	double P;
	if (ARITH == 0) {
		unsigned slow = 0;
		const QkRcp Rr = rcp_f<true>(U0[0], slow);
		const double vx = div_r<true>(U0[1], Rr, slow), vy = div_r<true>(U0[2], Rr, slow), vz = div_r<true>(U0[3], Rr, slow);
		const double ke = 0.5 * U0[0] * (vx * vx + vy * vy + vz * vz);
		const double Eint = U0[4] - ke;
		P = f_pressure_from_e<true>(c, U0[0], (U0[0] == 0.0) ? 0.0 : div_r<true>(Eint, Rr, slow), slow);
		if (slow)
			P = cons_pressure(c.h, U0[0], U0[1], U0[2], U0[3], U0[4]);
	} else {
		const double y = r_rcp(U0[0]);
		const double vx = U0[1] * y, vy = U0[2] * y, vz = U0[3] * y;
		const double Eint = U0[4] - 0.5 * U0[0] * (vx * vx + vy * vy + vz * vz);
		P = r_pressure_from_e(c, U0[0], (U0[0] == 0.0) ? 0.0 : Eint * y);
	}
	r[5] = r[5] + (-P * divv);
	if (BLEND) {
#pragma unroll
		for (int n = 0; n < 6 + NS; ++n)
			r[n] = 0.5 * r0[n * 32] + 0.5 * r[n];
	}
	// PredictStep
#pragma unroll
	for (int n = 0; n < 6 + NS; ++n)
		Un[n] = U0[n] + c.dt * r[n];
	bad = !(Un[0] > 0.);
#pragma unroll
	for (int n = 0; n < NMS; ++n)
		if (Un[6 + n] < 0.0)
			bad = 1;
	nonfin = 0;
#pragma unroll
	for (int n = 0; n < 6 + NS; ++n)
		nonfin |= nonfinite(Un[n]);
	// EnforceLimits
	const double rho = Un[0];
	double rho_new = rho;
	if (rho < c.h.dfloor) {
		rho_new = c.h.dfloor;
		Un[0] = rho_new;
#pragma unroll
		for (int n = 0; n < NS; ++n) {
			if (rho_new == 0.0)
				Un[6 + n] = 0.0;
			else
				Un[6 + n] *= rho / rho_new;
		}
	}
	if (NMS > 0) {
		double sp_sum = 0.0;
#pragma unroll
		for (int n = 0; n < NMS; ++n) {
			if (Un[6 + n] < 0.0)
				Un[6 + n] = c.h.small_x * rho_new;
			sp_sum += Un[6 + n];
		}
		if ((sp_sum > 2.2250738585072014e-308) && (rho_new > 2.2250738585072014e-308)) {
			sp_sum /= rho_new;
#pragma unroll
			for (int n = 0; n < NMS; ++n)
				Un[6 + n] /= sp_sum;
		}
	}
	// the temperature floors compare T >= small_temp > 0 (or NaN) with tempFloor: never taken unless tempFloor > 0
	if (c.h.tfloor > 0. && rho_new > 2.2250738585072014e-308) {
		const double v1 = Un[1] / rho_new, v2 = Un[2] / rho_new, v3 = Un[3] / rho_new;
		const double Ekin = 0.5 * rho_new * (v1 * v1 + v2 * v2 + v3 * v3);
		const double primTemp = eos_tgas_from_eint(c.h, rho_new, (Un[4] - Ekin));
		if (primTemp < c.h.tfloor)
			Un[4] = Ekin + eos_eint_from_tgas(c.h, rho_new, c.h.tfloor);
		const double auxTemp = eos_tgas_from_eint(c.h, rho_new, Un[5]);
		if (auxTemp < c.h.tfloor)
			Un[5] = eos_eint_from_tgas(c.h, rho_new, c.h.tfloor);
	}
	// SyncDualEnergy (cells with rho <= 0 are flagged above and the stage is redone by the faithful path)
	if (Un[0] > 0.) {
		const double num = Un[1] * Un[1] + Un[2] * Un[2] + Un[3] * Un[3], den = 2.0 * Un[0];
		double Ekin;
		if (ARITH == 0) {
			unsigned s2 = 0;
			Ekin = div_d<true>(num, den, s2);
			if (s2)
				Ekin = slow_div(num, den);
		} else {
			Ekin = num * r_rcp(den);
		}
		const double Eint_cons = Un[4] - Ekin;
		if (Eint_cons > 1.0e-3 * Un[4])
			Un[5] = Eint_cons;
		else
			Un[4] = Un[5] + Ekin;
	}
}

// order-preserving map of doubles to unsigned 64-bit for atomicMax (as qk_kernels.cuh)
__device__ __forceinline__ unsigned long long sig_key(double v)
{
	long long b = __double_as_longlong(v);
	return (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
// signal speeds of one updated cell: s0 = the ComputeMaxSignalSpeed expression (hydro_system.hpp:223-252, feeds computeTimestep),
// s1 = maxSignalSpeedLocal's (:198-221, feeds isCflViolated); the same per-cell values k_max_signal forms
template <int ARITH> __device__ __forceinline__ void cell_signal(const FastConst &c, const double *U, double &s0, double &s1)
{
	const double rho = U[0], px = U[1], py = U[2], pz = U[3], E = U[4];
	if (ARITH == 0) {
		// the quotients over rho share one refined reciprocal (qk_div.cuh); num/(2 rho) = 0.5 (num/rho) and 2 KE = num/rho are
		// exact binary scalings; anything outside the fast-path domain recomputes with the plain formulas of k_max_signal
		unsigned slow = 0;
		const QkRcp Rr = rcp_f<true>(rho, slow);
		const double vx = div_r<true>(px, Rr, slow), vy = div_r<true>(py, Rr, slow), vz = div_r<true>(pz, Rr, slow);
		const double Eint = E - 0.5 * rho * (vx * vx + vy * vy + vz * vz);
		const double P = f_pressure_from_e<true>(c, rho, (rho == 0.0) ? 0.0 : div_r<true>(Eint, Rr, slow), slow);
		const double cs = f_sound_speed<true>(c, rho, P, slow);
		s0 = fabs(cs + sqrt(vx * vx + vy * vy + vz * vz));
		const double twoKE = div_r<true>(px * px + py * py + pz * pz, Rr, slow);
		s1 = cs + sqrt(div_r<true>(twoKE, Rr, slow));
		const unsigned th = (unsigned)__double2hiint(twoKE) & 0x7fffffffu;
		if (slow || (th != 0u && th < 0x00300000u)) {
			const double P2 = cons_pressure(c.h, rho, px, py, pz, E);
			const double cs2 = eos_sound_speed(c.h, rho, P2);
			const double ux = px / rho, uy = py / rho, uz = pz / rho;
			s0 = fabs(cs2 + sqrt(ux * ux + uy * uy + uz * uz));
			const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
			s1 = cs2 + sqrt(2.0 * kinetic_energy / rho);
		}
	} else {
		const double y = r_rcp(rho);
		const double vsq = (px * px + py * py + pz * pz) * (y * y);
		const double P = r_pressure_from_e(c, rho, (rho == 0.0) ? 0.0 : (E - 0.5 * rho * vsq) * y);
		const double s = sqrt(c.h.gamma * r_p_of_p(c, rho, P) * y) + sqrt(vsq);
		s0 = fabs(s);
		s1 = s;
	}
}

// both signal-speed maxima of one box in one grid-stride pass (the dt / CFL reductions of a step): out[0] = ComputeMaxSignalSpeed
// + norminf, out[1] = maxSignalSpeedLocal.  Per-cell values are exactly k_max_signal's (shared-reciprocal quotients).
struct SigBoxes {
	static const int MAXB = 16;
	A4 u[MAXB];
	Box3 bx[MAXB];
};
template <int ARITH> __global__ void __launch_bounds__(256) k_signal(FastConst c, SigBoxes sb, unsigned long long *__restrict__ out)
{
	const Box3 &bx = sb.bx[blockIdx.y];
	const A4 &u = sb.u[blockIdx.y];
	const int nx = bx.len(0), ny = bx.len(1);
	const int64_t total = bx.ncells();
	double s0 = 0.0, s1 = -1.7976931348623157e308;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t jk = t / nx;
		const int i = bx.lo[0] + (int)(t - jk * nx);
		const int k = bx.lo[2] + (int)(jk / ny);
		const int j = bx.lo[1] + (int)(jk - (jk / ny) * ny);
		const int64_t o = u.off(i, j, k);
		const double U[5] = {u.p[o], u.p[o + u.ns], u.p[o + 2 * u.ns], u.p[o + 3 * u.ns], u.p[o + 4 * u.ns]};
		double a0, a1;
		cell_signal<ARITH>(c, U, a0, a1);
		s0 = dmax(s0, a0);
		s1 = dmax(s1, a1);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		s0 = dmax(s0, __shfl_xor_sync(0xffffffffu, s0, o));
		s1 = dmax(s1, __shfl_xor_sync(0xffffffffu, s1, o));
	}
	if ((threadIdx.x & 31) == 0) { // NaN never wins a '<' comparison, as in the reference's max reductions; look before the atomic
		const unsigned long long k0 = sig_key(s0), k1 = sig_key(s1);
		const volatile unsigned long long *cur = out;
		if (!(s0 != s0) && k0 > cur[0])
			atomicMax(out, k0);
		if (!(s1 != s1) && k1 > cur[1])
			atomicMax(out + 1, k1);
	}
}

// ---- arithmetic-mode dispatch: ARITH = 0 exact (shared-reciprocal fast path + plain-division twin), 1 relaxed ----
template <int ARITH, int DIR, int NS, int NMS, bool REINT>
__device__ __forceinline__ void hllc_face(const FastConst &c, const double *__restrict__ L, const double *__restrict__ R, double du, double dw,
					  double *__restrict__ F, double &vf)
{
	if (ARITH == 0) {
		unsigned slow = 0;
		f_hllc<DIR, NS, NMS, REINT, true>(c, L, R, du, dw, F, vf, slow);
		if (slow)
			f_hllc<DIR, NS, NMS, REINT, false>(c, L, R, du, dw, F, vf, slow);
	} else {
		r_hllc<DIR, NS, NMS, REINT>(c, L, R, du, dw, F, vf);
	}
}
template <int ARITH> __device__ __forceinline__ double div_dx(const FastConst &c, int d, double a)
{
	if (ARITH == 0) {
		unsigned s3 = 0;
		double dv = div_c<true>(a, c.dx[d], c.y_dx[d], s3);
		if (s3)
			dv = slow_div(a, c.dx[d]);
		return dv;
	}
	return a * c.y_dx[d];
}

// ---------------------------------------------------------------------------------------------------------------
// x sweep: one warp per 30 cells of a row (lane l <-> cell x0 - 1 + l)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <int ARITH, int NS, int NMS, bool REINT, int STAGE, bool DUAL>
__global__ void __launch_bounds__(128) k_sweep_x(FastConst c, const SweepBox *__restrict__ boxes, int tiles_x, int rows_per_box_max)
{
	constexpr int NV = 6 + NS;
	const SweepBox &B = boxes[blockIdx.z];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int ny = B.hi[1] - B.lo[1] + 1, nz = B.hi[2] - B.lo[2] + 1;
	const int row = blockIdx.y * 4 + warp;
	if (row >= ny * nz)
		return; // whole warp
	const int j = B.lo[1] + row % ny, k = B.lo[2] + row / ny;
	const int x0 = B.lo[0] + blockIdx.x * 30;
	if (x0 > B.hi[0])
		return;
	const int i = x0 - 1 + lane;
	const bool cell_ok = (i <= B.hi[0] + 1); // cells lo-1 .. hi+1 carry a parabola
	const int ic = cell_ok ? i : B.hi[0] + 1;
	const A4 &q = B.prim;
	const int64_t o = q.off(ic, j, k);

	// PPM + flattening of the own cell, all variables
	const double chi = q.p[o + NV * q.ns], omchi = 1. - chi;
	double am[NV], ap[NV], q0v[NV];
#pragma unroll
	for (int n = 0; n < NV; ++n) {
		const double *p = q.p + o + n * q.ns;
		const double qm2 = p[-2], qm1 = p[-1], q0 = p[0], qp1 = p[1], qp2 = p[2];
		q0v[n] = q0;
		f_ppm_flat(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), chi, omchi, am[n], ap[n]);
	}
	double mV, mW;
	cell_trans_min<0>(q, o, mV, mW);

	// face between lane-1 and lane: L = ap(lane-1), R = am(lane)
	double Ls[NV];
#pragma unroll
	for (int n = 0; n < NV; ++n)
		Ls[n] = shfl_up1(ap[n]);
	const double mVl = shfl_up1(mV), mWl = shfl_up1(mW);
	const double du = q0v[1] - shfl_up1(q0v[1]);
	double dw = dmin(mVl, mV);
	dw = dmin(dmin(mWl, mW), dw);
	const bool face_ok = (lane >= 1) && (i >= B.lo[0]) && (i <= B.hi[0] + 1);
	double G[NV + 1];
	if (face_ok) {
		double F[NV], vf;
		hllc_face<ARITH, 0, NS, NMS, REINT>(c, Ls, am, du, dw, F, vf);
		const A4 &h = B.hF[0];
		const int64_t oh = h.off(i, j, k);
		if (STAGE == 1) {
#pragma unroll
			for (int n = 0; n < NV; ++n)
				G[n] = F[n];
			G[NV] = vf;
			if (DUAL && (lane <= 30 || i == B.hi[0] + 1)) { // flux_rk2 = 0 + 0.5 F (QuokkaSimulation.hpp:1106-1107)
#pragma unroll
				for (int n = 0; n < NV; ++n)
					h.p[oh + n * h.ns] = 0.0 + 0.5 * F[n];
				h.p[oh + NV * h.ns] = 0.0 + 0.5 * vf;
			}
		} else {
#pragma unroll
			for (int n = 0; n < NV; ++n)
				G[n] = h.p[oh + n * h.ns] + 0.5 * F[n];
			G[NV] = h.p[oh + NV * h.ns] + 0.5 * vf;
		}
	} else {
#pragma unroll
		for (int n = 0; n <= NV; ++n)
			G[n] = 0.0;
	}
	// cell update needs the flux of the next face (lane+1)
	const bool upd = (lane >= 1) && (lane <= 30) && (i <= B.hi[0]);
	const A4 &r = B.rhs;
	const int64_t orr = upd ? r.off(i, j, k) : 0;
#pragma unroll
	for (int n = 0; n < NV; ++n) {
		const double Gn = shfl_dn1(G[n]);
		if (upd)
			r.p[orr + n * r.ns] = c.inv_dx[0] * (G[n] - Gn);
	}
	const double Vn = shfl_dn1(G[NV]);
	if (upd)
	{
			const double dv = div_dx<ARITH>(c, 0, Vn - G[NV]);
			r.p[orr + NV * r.ns] = dv;
		}
}

// ---------------------------------------------------------------------------------------------------------------
// y / z sweeps by marching; LAST carries the stage epilogue
// ---------------------------------------------------------------------------------------------------------------
template <int ARITH, int DIR, int NS, int NMS, bool REINT, int STAGE, bool DUAL, bool LAST>
__global__ void __launch_bounds__(128) k_sweep_m(FastConst c, const SweepBox *__restrict__ boxes, int nseg, unsigned long long *__restrict__ counters)
{
	constexpr int NV = 6 + NS;
	constexpr int TD = (DIR == 1) ? 2 : 1; // the transverse (non-x) axis a warp is pinned to
	const int box = blockIdx.z / nseg, seg = blockIdx.z - box * nseg;
	const SweepBox &B = boxes[box];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int i = B.lo[0] + blockIdx.x * 32 + lane;
	const int t = B.lo[TD] + blockIdx.y * 4 + warp;
	const int s0 = B.lo[DIR] + seg * SEG;
	int bad_cnt = 0, nf_cnt = 0;
	if (i <= B.hi[0] && t <= B.hi[TD] && s0 <= B.hi[DIR]) {
		const int s1 = min(s0 + SEG, B.hi[DIR] + 1); // cells s0 .. s1-1 are updated, faces s0 .. s1 evaluated
		const A4 &q = B.prim;
		const int64_t sN = (DIR == 1) ? q.js : q.ks;
		int idx[3];
		idx[0] = i;
		idx[TD] = t;
		idx[DIR] = s0 - 1;
		int64_t o = q.off(idx[0], idx[1], idx[2]);
		const A4 &h = B.hF[DIR];
		const int64_t shN = (DIR == 1) ? h.js : h.ks;
		int64_t oh = h.off(idx[0], idx[1], idx[2]); // face index == cell index of the cell on its high side
		const A4 &r = B.rhs;
		const int64_t srN = (DIR == 1) ? r.js : r.ks;
		int64_t orr = r.off(idx[0], idx[1], idx[2]);
		const A4 &u0 = B.U0, &uo = B.Uo;
		const int64_t suN = (DIR == 1) ? u0.js : u0.ks, soN = (DIR == 1) ? uo.js : uo.ks;
		int64_t ou = u0.off(idx[0], idx[1], idx[2]), oo = uo.off(idx[0], idx[1], idx[2]);

		double apL[NV], ifl[NV], Gp[NV + 1];
		double mVp = 0, mWp = 0, vNp = 0;
		// unlimited interface value at the low face of the first cell
#pragma unroll
		for (int n = 0; n < NV; ++n) {
			const double *p = q.p + o + n * q.ns;
			ifl[n] = ppm_iface(p[-2 * sN], p[-sN], p[0], p[sN]);
		}
		for (int s = s0 - 1; s <= s1; ++s) {
			const double chi = q.p[o + NV * q.ns], omchi = 1. - chi;
			double am[NV], ap[NV];
			double vN0 = 0;
#pragma unroll
			for (int n = 0; n < NV; ++n) {
				const double *p = q.p + o + n * q.ns;
				const double qm1 = p[-sN], q0 = p[0], qp1 = p[sN], qp2 = p[2 * sN];
				if (n == 1 + DIR)
					vN0 = q0;
				const double ifh = ppm_iface(qm1, q0, qp1, qp2);
				f_ppm_flat(qm1, q0, qp1, ifl[n], ifh, chi, omchi, am[n], ap[n]);
				ifl[n] = ifh;
			}
			double mV, mW;
			cell_trans_min<DIR>(q, o, mV, mW);
			if (s >= s0) {
				const double du = vN0 - vNp;
				double dw = dmin(mVp, mV);
				dw = dmin(dmin(mWp, mW), dw);
				double F[NV], vf, G[NV + 1];
				hllc_face<ARITH, DIR, NS, NMS, REINT>(c, apL, am, du, dw, F, vf);
				if (STAGE == 1) {
#pragma unroll
					for (int n = 0; n < NV; ++n)
						G[n] = F[n];
					G[NV] = vf;
					if (DUAL) {
#pragma unroll
						for (int n = 0; n < NV; ++n)
							h.p[oh + n * h.ns] = 0.0 + 0.5 * F[n];
						h.p[oh + NV * h.ns] = 0.0 + 0.5 * vf;
					}
				} else {
#pragma unroll
					for (int n = 0; n < NV; ++n)
						G[n] = h.p[oh + n * h.ns] + 0.5 * F[n];
					G[NV] = h.p[oh + NV * h.ns] + 0.5 * vf;
				}
				if (s > s0) { // cell s-1: both faces known
					const int64_t orc = orr - srN;
					double rr[NV];
#pragma unroll
					for (int n = 0; n < NV; ++n)
						rr[n] = r.p[orc + n * r.ns] + c.inv_dx[DIR] * (Gp[n] - G[n]);
					const double dv = div_dx<ARITH>(c, DIR, G[NV] - Gp[NV]);
					const double divv = r.p[orc + NV * r.ns] + dv;
					if (!LAST) {
#pragma unroll
						for (int n = 0; n < NV; ++n)
							r.p[orc + n * r.ns] = rr[n];
						r.p[orc + NV * r.ns] = divv;
					} else {
						double U0[NV], Un[NV];
						const int64_t ouc = ou - suN;
#pragma unroll
						for (int n = 0; n < NV; ++n)
							U0[n] = u0.p[ouc + n * u0.ns];
						int bad, nf;
						cell_epilogue<ARITH, NS, NMS>(c, U0, rr, divv, Un, bad, nf);
						bad_cnt += bad;
						nf_cnt += nf;
						const int64_t ooc = oo - soN;
#pragma unroll
						for (int n = 0; n < NV; ++n)
							uo.p[ooc + n * uo.ns] = Un[n];
					}
				}
#pragma unroll
				for (int n = 0; n <= NV; ++n)
					Gp[n] = G[n];
			}
#pragma unroll
			for (int n = 0; n < NV; ++n)
				apL[n] = ap[n];
			mVp = mV;
			mWp = mW;
			vNp = vN0;
			o += sN;
			oh += shN;
			orr += srN;
			ou += suN;
			oo += soN;
		}
	}
	if (LAST) {
		// one atomic per block that saw a flagged / non-finite cell (none in a healthy run)
		const int any = __syncthreads_or(bad_cnt | nf_cnt);
		if (any) {
			if (bad_cnt)
				atomicAdd(counters, (unsigned long long)bad_cnt);
			if (nf_cnt)
				atomicAdd(counters + 1, (unsigned long long)nf_cnt);
		}
	}
}
#include "qk_march.cuh"
} // namespace

#define QK_TRY(x)                                                                                                                                    \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

// descriptors of one launch of a marching sweep along dir (1 | 2) from the host table [box][TM_COUNT]
static void fill_march_maps(MarchMaps &mm, const unsigned char *h_maps, int b0, int nbc, int dir, bool r0)
{
	for (int b = 0; b < nbc; ++b) {
		const unsigned char *src = h_maps + (size_t)(b0 + b) * TM_COUNT * TMAP_BYTES;
		memcpy(mm.m[b][MM_PRIM].b, src + TM_PRIM_M36 * TMAP_BYTES, TMAP_BYTES);
		memcpy(mm.m[b][MM_ROW].b, src + TM_PRIM_R32 * TMAP_BYTES, TMAP_BYTES);
		memcpy(mm.m[b][MM_RHS].b, src + TM_RHS * TMAP_BYTES, TMAP_BYTES);
		memcpy(mm.m[b][MM_STAGE2].b, src + (r0 ? TM_R0 : (TM_HF0 + dir)) * TMAP_BYTES, TMAP_BYTES);
		memcpy(mm.m[b][MM_U0].b, src + TM_U0 * TMAP_BYTES, TMAP_BYTES);
	}
}

// maxn: [0..2] the largest box extents of the launch, [3] the smallest nx, [4] the largest nrows * (nx + 2) (0: does not fit an int)
template <int ARITH, int NS, int NMS, bool REINT, int STAGE, bool DUAL, int ORDER = 3, bool KEEPF = false>
static int launch_stage(int ng, unsigned long long *d_counters, const FastConst &c, const SweepBox *d_tab, const unsigned char *h_maps, int nb, const int maxn[5],
			bool tma, cudaStream_t s)
{
	if (nb == 0)
		return 0; // a rank without boxes (more ranks than boxes) launches nothing and still joins the stage's collectives
	{
		ProfScope p("fused_prim", s);
		const int64_t cells = (int64_t)(maxn[0] + 2 * ng) * (maxn[1] + 2 * ng) * (maxn[2] + 2 * ng);
		dim3 grid((unsigned)std::min<int64_t>((cells + 255) / 256, 4096), nb);
		k_fprim<ARITH, NS, REINT><<<grid, 256, 0, s>>>(c, d_tab, ng);
		QK_KERNEL_CHECK();
	}
	{
		ProfScope p("fused_chi", s);
		const int64_t cells = (int64_t)(maxn[0] + 4) * (maxn[1] + 4) * (maxn[2] + 4);
		dim3 grid((unsigned)std::min<int64_t>((cells + 255) / 256, 4096), nb);
		if constexpr (ARITH == 1) {
			k_fchi9<NS, REINT><<<grid, 256, 0, s>>>(c, d_tab);
			QK_KERNEL_CHECK();
		} else {
			k_fchi<ARITH, NS, REINT><<<grid, 256, 0, s>>>(c, d_tab);
			QK_KERNEL_CHECK();
			k_fchimin<NS><<<grid, 256, 0, s>>>(d_tab);
			QK_KERNEL_CHECK();
		}
	}
	{
		ProfScope p("sweep_x", s);
		const int tiles_x = (maxn[0] + 29) / 30;
		const int rows = maxn[1] * maxn[2];
		// relaxed arithmetic keeps R(U0) instead of 0.5*F(U0) (qk_march.cuh): its x and y sweeps never touch the face arrays
		// and are the same kernels in both stages
		constexpr int XSTAGE = (ARITH == 1) ? 1 : STAGE;
		constexpr bool XDUAL = (ARITH == 1) ? false : DUAL;
		// every box at least 30 cells wide: the rows of a box are laid end to end and cut into 30-slot tiles (k_sweep_xc, qk_march.cuh);
		// QK_XCAT=0 keeps the per-row tiles of k_sweep_xt.  maxn[3] carries the smallest nx of the launch, maxn[4] the largest slot count.
		const char *xe = getenv("QK_XCAT");
		const bool xcat_env = !(xe != nullptr && xe[0] == '0');
		if (tma && xcat_env && maxn[3] >= 30 && maxn[4] > 0) {
			auto kern = k_sweep_xc<ARITH, NS, NMS, REINT, XSTAGE, XDUAL, ORDER, KEEPF>;
			static bool attr_set = false;
			if (!attr_set) {
				QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, XCSmem<6 + NS>::BLOCK_BYTES));
				attr_set = true;
			}
			const int tiles = (maxn[4] + 29) / 30;
			for (int b0 = 0; b0 < nb; b0 += TMAP_MAXB) {
				const int nbc = std::min(TMAP_MAXB, nb - b0);
				XMaps xm;
				for (int b = 0; b < nbc; ++b) {
					const unsigned char *src = h_maps + (size_t)(b0 + b) * TM_COUNT * TMAP_BYTES;
					memcpy(xm.m[b][XM_PRIM].b, src + TM_PRIM_X40 * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_Y3].b, src + TM_PRIM_Y3W * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_Z3].b, src + TM_PRIM_Z3W * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_HF].b, src + TM_HF0 * TMAP_BYTES, TMAP_BYTES);
				}
				dim3 grid((tiles + XC_WARPS * XC_TILES - 1) / (XC_WARPS * XC_TILES), 1, nbc);
				kern<<<grid, 32 * XC_WARPS, XCSmem<6 + NS>::BLOCK_BYTES, s>>>(c, d_tab + b0, xm);
				if (b0 + TMAP_MAXB < nb)
					QK_KERNEL_CHECK();
			}
		} else if (tma) {
			auto kern = k_sweep_xt<ARITH, NS, NMS, REINT, XSTAGE, XDUAL, ORDER, KEEPF>;
			static bool attr_set = false;
			if (!attr_set) {
				QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, XSmem<6 + NS, XSTAGE>::BLOCK_BYTES));
				attr_set = true;
			}
			for (int b0 = 0; b0 < nb; b0 += TMAP_MAXB) { // descriptors are kernel parameters: at most TMAP_MAXB boxes per launch
				const int nbc = std::min(TMAP_MAXB, nb - b0);
				XMaps xm;
				for (int b = 0; b < nbc; ++b) {
					const unsigned char *src = h_maps + (size_t)(b0 + b) * TM_COUNT * TMAP_BYTES;
					memcpy(xm.m[b][XM_PRIM].b, src + TM_PRIM_X38 * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_Y3].b, src + TM_PRIM_Y3 * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_Z3].b, src + TM_PRIM_Z3 * TMAP_BYTES, TMAP_BYTES);
					memcpy(xm.m[b][XM_HF].b, src + TM_HF0 * TMAP_BYTES, TMAP_BYTES);
				}
				dim3 grid(tiles_x, (rows + 4 * XROWS - 1) / (4 * XROWS), nbc);
				kern<<<grid, 128, XSmem<6 + NS, XSTAGE>::BLOCK_BYTES, s>>>(c, d_tab + b0, xm);
				if (b0 + TMAP_MAXB < nb)
					QK_KERNEL_CHECK();
			}
		} else if constexpr (ARITH == 0 && ORDER == 3 && !KEEPF) {
			dim3 grid(tiles_x, (rows + 3) / 4, nb);
			k_sweep_x<ARITH, NS, NMS, REINT, STAGE, DUAL><<<grid, 128, 0, s>>>(c, d_tab, tiles_x, rows);
		} else {
			return QK_ERR_UNSUPPORTED; // the caller runs the exact kernels when the rows cannot be bulk-copied
		}
		QK_KERNEL_CHECK();
	}
	{
		ProfScope p("sweep_y", s);
		const int nseg = (maxn[1] + SEG - 1) / SEG;
		dim3 grid((maxn[0] + 31) / 32, (maxn[2] + 3) / 4, nb * nseg);
		constexpr int YSTAGE = (ARITH == 1) ? 1 : STAGE;
		constexpr bool YDUAL = (ARITH == 1) ? false : DUAL;
		if (tma) {
			auto kern = k_march_t<ARITH, 1, NS, NMS, REINT, YSTAGE, YDUAL, false, ORDER, KEEPF>;
			static bool attr_set = false;
			if (!attr_set) {
				QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MarchSmem<6 + NS, YSTAGE, false, ARITH == 1>::BLOCK_BYTES));
				attr_set = true;
			}
			for (int b0 = 0; b0 < nb; b0 += TMAP_MAXB) {
				const int nbc = std::min(TMAP_MAXB, nb - b0);
				MarchMaps mm;
				fill_march_maps(mm, h_maps, b0, nbc, 1, YSTAGE == 2 && ARITH == 1);
				grid.z = nbc * nseg;
				kern<<<grid, 128, MarchSmem<6 + NS, YSTAGE, false, ARITH == 1>::BLOCK_BYTES, s>>>(c, d_tab + b0, mm, nseg, d_counters);
				if (b0 + TMAP_MAXB < nb)
					QK_KERNEL_CHECK();
			}
		} else if constexpr (ARITH == 0 && ORDER == 3 && !KEEPF) {
			k_sweep_m<ARITH, 1, NS, NMS, REINT, STAGE, DUAL, false><<<grid, 128, 0, s>>>(c, d_tab, nseg, d_counters);
		}
		QK_KERNEL_CHECK();
	}
	{
		ProfScope p("sweep_z", s);
		const int nseg = (maxn[2] + SEG - 1) / SEG;
		dim3 grid((maxn[0] + 31) / 32, (maxn[1] + 3) / 4, nb * nseg);
		if (tma) {
			auto kern = k_march_t<ARITH, 2, NS, NMS, REINT, STAGE, DUAL, true, ORDER, KEEPF>;
			static bool attr_set = false;
			if (!attr_set) {
				QK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MarchSmem<6 + NS, STAGE, true, ARITH == 1>::BLOCK_BYTES));
				attr_set = true;
			}
			for (int b0 = 0; b0 < nb; b0 += TMAP_MAXB) {
				const int nbc = std::min(TMAP_MAXB, nb - b0);
				MarchMaps mm;
				fill_march_maps(mm, h_maps, b0, nbc, 2, ARITH == 1);
				grid.z = nbc * nseg;
				kern<<<grid, 128, MarchSmem<6 + NS, STAGE, true, ARITH == 1>::BLOCK_BYTES, s>>>(c, d_tab + b0, mm, nseg, d_counters);
				if (b0 + TMAP_MAXB < nb)
					QK_KERNEL_CHECK();
			}
		} else if constexpr (ARITH == 0 && ORDER == 3 && !KEEPF) {
			k_sweep_m<ARITH, 2, NS, NMS, REINT, STAGE, DUAL, true><<<grid, 128, 0, s>>>(c, d_tab, nseg, d_counters);
		}
		QK_KERNEL_CHECK();
	}
	return 0;
}

template <int ARITH, int NS, int NMS, bool REINT, int ORDER = 3, bool KEEPF = false>
static int dispatch_stage(int ng, unsigned long long *d_counters, const FastConst &c, const SweepBox *t, const unsigned char *m, int nb, const int maxn[5], int stage,
			  bool dual, bool tma, cudaStream_t s)
{
	if (KEEPF && !tma)
		return QK_ERR_UNSUPPORTED; // the flux-keeping kernels exist in the TMA-staged form only
	if (stage == 1)
		return dual ? launch_stage<ARITH, NS, NMS, REINT, 1, true, ORDER, KEEPF>(ng, d_counters, c, t, m, nb, maxn, tma, s)
			    : launch_stage<ARITH, NS, NMS, REINT, 1, false, ORDER, KEEPF>(ng, d_counters, c, t, m, nb, maxn, tma, s);
	return launch_stage<ARITH, NS, NMS, REINT, 2, true, ORDER, KEEPF>(ng, d_counters, c, t, m, nb, maxn, tma, s);
}


// entry of one translation unit: all instantiated trait sets of one arithmetic mode (KEEPF = false: qk_sweep.cu / qk_sweep_relaxed.cu;
// KEEPF = true, the same kernels also storing the stage's face fluxes for the flux registers: qk_sweep_keepf.cu / qk_sweep_relaxed_keepf.cu)
// PLM (reconstructionOrder_ = 2) is instantiated for the trait set of config C4's hydro (no scalars, reconstruct_eint = false), TMA form only
template <int ARITH, bool KEEPF = false>
static int sweep_stage_dispatch_plm(int ng, unsigned long long *d_counters, const FastConst &c, const SweepBox *db, const unsigned char *dm, int nb, const int maxn[5],
				    int stage, bool dual, cudaStream_t s)
{
	return dispatch_stage<ARITH, 0, 0, false, 2, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, true, s);
}

template <int ARITH, bool KEEPF = false>
static int sweep_stage_dispatch(int ns, bool reint, int ng, unsigned long long *d_counters, const FastConst &c, const SweepBox *db, const unsigned char *dm, int nb,
				const int maxn[5], int stage, bool dual, bool tma, cudaStream_t s)
{
	if (ns == 0)
		return reint ? dispatch_stage<ARITH, 0, 0, true, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s)
			     : dispatch_stage<ARITH, 0, 0, false, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s);
	if (ns == 1)
		return reint ? dispatch_stage<ARITH, 1, 0, true, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s)
			     : dispatch_stage<ARITH, 1, 0, false, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s);
	return reint ? dispatch_stage<ARITH, 3, 2, true, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s)
		     : dispatch_stage<ARITH, 3, 2, false, 3, KEEPF>(ng, d_counters, c, db, dm, nb, maxn, stage, dual, tma, s);
}
