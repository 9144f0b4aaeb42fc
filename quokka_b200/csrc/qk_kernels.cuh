// qk_kernels.cuh -- one CUDA kernel per reference operator (shared by qk_ops.cu, the per-operator C-ABI entry
// points, and qk_level.cu, the faithful level path used for FOFC fallback).  Everything here has internal linkage.
// Compiled with --fmad=false (exact arithmetic contract, see qk_physics.cuh).
#pragma once
#include "qk_physics.cuh"

namespace
{

constexpr int TPB = 256;

struct Iter { // flattened iteration over an index box, x fastest
	int lo[3];
	int n[3];
	int64_t total;
	explicit Iter(const Box3 &b)
	{
		for (int d = 0; d < 3; ++d) {
			lo[d] = b.lo[d];
			n[d] = b.len(d);
		}
		total = (int64_t)n[0] * n[1] * n[2];
	}
	__device__ __forceinline__ bool get(int64_t t, int &i, int &j, int &k) const
	{
		if (t >= total)
			return false;
		const int64_t jk = t / n[0];
		i = lo[0] + (int)(t - jk * n[0]);
		k = lo[2] + (int)(jk / n[1]);
		j = lo[1] + (int)(jk - (jk / n[1]) * n[1]);
		return true;
	}
	unsigned blocks() const { return (unsigned)((total + TPB - 1) / TPB); }
};

// HydroSystem::ConservedToPrimitive  src/hydro/hydro_system.hpp:138-196
__global__ void __launch_bounds__(TPB) k_cons_to_prim(HydroConst c, Iter it, A4 cons, A4 prim)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t o = cons.off(i, j, k), op = prim.off(i, j, k);
	const double rho = cons.p[o], px = cons.p[o + cons.ns], py = cons.p[o + 2 * cons.ns], pz = cons.p[o + 3 * cons.ns];
	const double E = cons.p[o + 4 * cons.ns], Eaux = cons.p[o + 5 * cons.ns];
	const double vx = px / rho, vy = py / rho, vz = pz / rho;
	const double ke = 0.5 * rho * (vx * vx + vy * vy + vz * vz);
	const double Eint_cons = E - ke;
	prim.p[op] = rho;
	prim.p[op + prim.ns] = vx;
	prim.p[op + 2 * prim.ns] = vy;
	prim.p[op + 3 * prim.ns] = vz;
	if (c.reconstruct_eint) {
		prim.p[op + 4 * prim.ns] = Eint_cons / rho;
		prim.p[op + 5 * prim.ns] = Eaux / rho;
	} else {
		prim.p[op + 4 * prim.ns] = c.iso ? rho * c.cs_iso * c.cs_iso : eos_pressure(c, rho, Eint_cons); // ComputePressure, :365-366
		prim.p[op + 5 * prim.ns] = Eaux;
	}
	for (int n = 0; n < c.ns; ++n)
		prim.p[op + (6 + n) * prim.ns] = cons.p[o + (6 + n) * cons.ns];
}

// HydroSystem::ComputeFlatteningCoefficients<DIR>  hydro_system.hpp:531-626
__global__ void __launch_bounds__(TPB) k_flat_coefs(HydroConst c, int dir, Iter it, A4 q, A4 chi)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t s = (dir == 0) ? 1 : (dir == 1) ? q.js : q.ks;
	const int64_t o = q.off(i, j, k);
	double P[5];
#pragma unroll
	for (int m = -2; m <= 2; ++m) {
		double v = q.p[o + m * s + 4 * q.ns];
		if (c.reconstruct_eint) {
			const double r = q.p[o + m * s];
			v = eos_pressure(c, r, r * v);
		}
		if (c.iso) // hydro_system.hpp:579-586
			v = q.p[o + m * s] * (c.cs_iso * c.cs_iso);
		P[m + 2] = v;
	}
	const double rho = q.p[o];
	double KS;
	if (c.iso) { // :604-606
		KS = rho * c.cs_iso * c.cs_iso;
	} else {
		const double cs = eos_sound_speed(c, rho, P[2]);
		KS = (cs * cs) * rho;
	}
	const double vm1 = q.p[o - s + (1 + dir) * q.ns], vp1 = q.p[o + s + (1 + dir) * q.ns];
	chi(i, j, k, 0) = flatten_chi(P[0], P[1], P[3], P[4], KS, vm1, vp1);
}

// HyperbolicSystem::ReconstructStates{Constant,PLM,PPM}<DIR>  src/hyperbolic_system.hpp:129-433
template <int ORDER, int LIMITER> __global__ void __launch_bounds__(TPB) k_reconstruct(int dir, Iter it, int nvars, A4 q, A4 left, A4 right)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t s = (dir == 0) ? 1 : (dir == 1) ? q.js : q.ks;
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	for (int n = 0; n < nvars; ++n) {
		const double *qp = q.p + q.off(i, j, k) + n * q.ns;
		if (ORDER == 3) {
			double am, ap;
			recon_cell<3, 0>(qp[-2 * s], qp[-s], qp[0], qp[s], qp[2 * s], am, ap);
			right(i, j, k, n) = am;
			left(i + e0, j + e1, k + e2, n) = ap;
		} else if (ORDER == 2) {
			// interface-centred: left(i) from cell i-1, right(i) from cell i (:243-246)
			double am, ap, am1, ap1;
			recon_cell<2, LIMITER>(0., qp[-s], qp[0], qp[s], 0., am, ap);
			recon_cell<2, LIMITER>(0., qp[-2 * s], qp[-s], qp[0], 0., am1, ap1);
			left(i, j, k, n) = ap1;
			right(i, j, k, n) = am;
		} else {
			left(i, j, k, n) = qp[-s];
			right(i, j, k, n) = qp[0];
		}
	}
}

// HydroSystem::FlattenShocks<DIR>  hydro_system.hpp:628-694
__global__ void __launch_bounds__(TPB) k_flatten(int dir, Iter it, int nvars, A4 q, A4 c1, A4 c2, A4 c3, A4 left, A4 right)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	double chi = c1(i - 1, j, k, 0);
	chi = dmin(chi, c1(i, j, k, 0));
	chi = dmin(chi, c1(i + 1, j, k, 0));
	chi = dmin(chi, c2(i, j - 1, k, 0));
	chi = dmin(chi, c2(i, j, k, 0));
	chi = dmin(chi, c2(i, j + 1, k, 0));
	chi = dmin(chi, c3(i, j, k - 1, 0));
	chi = dmin(chi, c3(i, j, k, 0));
	chi = dmin(chi, c3(i, j, k + 1, 0));
	for (int n = 0; n < nvars; ++n) {
		const double a_minus = right(i, j, k, n);
		const double a_plus = left(i + e0, j + e1, k + e2, n);
		const double a_mean = q(i, j, k, n);
		right(i, j, k, n) = chi * a_minus + (1. - chi) * a_mean;
		left(i + e0, j + e1, k + e2, n) = chi * a_plus + (1. - chi) * a_mean;
	}
}

// transverse velocity-difference terms of the carbuncle fix at a face (hydro_system.hpp:1018-1034)
__device__ __forceinline__ void face_du_dw(const A4 &q, int dir, int64_t o, int64_t sN, double &du, double &dw, double &div_v)
{
	const int aV = (dir + 1) % 3, aW = (dir + 2) % 3;
	const int64_t sV = (aV == 0) ? 1 : (aV == 1) ? q.js : q.ks;
	const int64_t sW = (aW == 0) ? 1 : (aW == 1) ? q.js : q.ks;
	const double *vN = q.p + o + (1 + dir) * q.ns;
	const double *vV = q.p + o + (1 + aV) * q.ns;
	const double *vW = q.p + o + (1 + aW) * q.ns;
	du = vN[0] - vN[-sN];
	const double dvl = dmin(vV[-sN + sV] - vV[-sN], vV[-sN] - vV[-sN - sV]);
	const double dvr = dmin(vV[sV] - vV[0], vV[0] - vV[-sV]);
	dw = dmin(dvl, dvr);
	const double dwl = dmin(vW[-sN + sW] - vW[-sN], vW[-sN] - vW[-sN - sW]);
	const double dwr = dmin(vW[sW] - vW[0], vW[0] - vW[-sW]);
	dw = dmin(dmin(dwl, dwr), dw);
	div_v = du + 0.5 * (dvl + dvr) + 0.5 * (dwl + dwr); // hydro_system.hpp:1054
}

// HydroSystem::ComputeFluxes<RIEMANN,DIR>  hydro_system.hpp:852-1112
template <int SOLVER> __global__ void __launch_bounds__(TPB) k_compute_fluxes(HydroConst c, int dir, Iter it, A4 flux, A4 fvel, A4 left, A4 right, A4 q)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	double L[QK_MAXV], R[QK_MAXV], F[QK_MAXV];
	for (int n = 0; n < c.nv; ++n) {
		L[n] = left(i, j, k, n);
		R[n] = right(i, j, k, n);
	}
	const int64_t sN = (dir == 0) ? 1 : (dir == 1) ? q.js : q.ks;
	double du, dw, vface, div_v;
	face_du_dw(q, dir, q.off(i, j, k), sN, du, dw, div_v);
	face_flux<SOLVER, true>(c, dir, L, R, du, dw, F, vface, div_v);
	for (int n = 0; n < c.nv; ++n)
		flux(i, j, k, n) = F[n];
	fvel(i, j, k, 0) = vface;
}

// hydroFluxFunction<DIR> / hydroFOFluxFunction<DIR> fused, one thread per face, nothing materialised
// (QuokkaSimulation.hpp:1492-1517, 1559-1568).  Generic-shape version: the tuned pencil kernels are in qk_sweep.cuh.
template <int ORDER, int SOLVER>
__global__ void __launch_bounds__(TPB) k_flux_function(HydroConst c, int dir, Iter it, A4 q, A4 c1, A4 c2, A4 c3, A4 flux, A4 fvel)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const int64_t sN = (dir == 0) ? 1 : (dir == 1) ? q.js : q.ks;
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	const int64_t o = q.off(i, j, k);
	double chiL = 1.0, chiR = 1.0;
	if (ORDER > 1 || SOLVER == QK_HLLC) {
		// min over the 9 neighbours of chi1,chi2,chi3 for cells i-1 (L) and i (R)  hydro_system.hpp:655-669
#pragma unroll
		for (int side = 0; side < 2; ++side) {
			const int ci = i - (1 - side) * e0, cj = j - (1 - side) * e1, ck = k - (1 - side) * e2;
			double chi = c1(ci - 1, cj, ck, 0);
			chi = dmin(chi, c1(ci, cj, ck, 0));
			chi = dmin(chi, c1(ci + 1, cj, ck, 0));
			chi = dmin(chi, c2(ci, cj - 1, ck, 0));
			chi = dmin(chi, c2(ci, cj, ck, 0));
			chi = dmin(chi, c2(ci, cj + 1, ck, 0));
			chi = dmin(chi, c3(ci, cj, ck - 1, 0));
			chi = dmin(chi, c3(ci, cj, ck, 0));
			chi = dmin(chi, c3(ci, cj, ck + 1, 0));
			if (side == 0)
				chiL = chi;
			else
				chiR = chi;
		}
	}
	double L[QK_MAXV], R[QK_MAXV], F[QK_MAXV];
	for (int n = 0; n < c.nv; ++n) {
		const double *qp = q.p + o + n * q.ns;
		if (ORDER == 1 && SOLVER == QK_LLF) { // hydroFOFluxFunction: donor cell, no flattening (QuokkaSimulation.hpp:1559-1568)
			L[n] = qp[-sN];
			R[n] = qp[0];
		} else if (ORDER == 1) { // hydroFluxFunction applies FlattenShocks for every order (:1509)
			const double qm1 = qp[-sN], q0 = qp[0];
			L[n] = chiL * qm1 + (1. - chiL) * qm1;
			R[n] = chiR * q0 + (1. - chiR) * q0;
		} else {
			const double qm3 = qp[-3 * sN], qm2 = qp[-2 * sN], qm1 = qp[-sN], q0 = qp[0], qp1 = qp[sN], qp2 = qp[2 * sN];
			double amL, apL, amR, apR;
			recon_cell<ORDER, QK_MINMOD>(qm3, qm2, qm1, q0, qp1, amL, apL);
			recon_cell<ORDER, QK_MINMOD>(qm2, qm1, q0, qp1, qp2, amR, apR);
			L[n] = chiL * apL + (1. - chiL) * qm1;
			R[n] = chiR * amR + (1. - chiR) * q0;
		}
	}
	double du, dw, vface, div_v;
	face_du_dw(q, dir, o, sN, du, dw, div_v);
	face_flux<SOLVER, true>(c, dir, L, R, du, dw, F, vface, div_v);
	for (int n = 0; n < c.nv; ++n)
		flux(i, j, k, n) = F[n];
	fvel(i, j, k, 0) = vface;
}

__global__ void __launch_bounds__(TPB) k_saxpy(Iter it, int ncomp, A4 dst, double a, A4 src)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	for (int n = 0; n < ncomp; ++n)
		dst(i, j, k, n) = dst(i, j, k, n) + a * src(i, j, k, n);
}

// HydroSystem::ComputeRhsFromFluxes  hydro_system.hpp:448-473
__global__ void __launch_bounds__(TPB) k_rhs(Iter it, int nvars, A4 rhs, A4 fx, A4 fy, A4 fz, double dx0, double dx1, double dx2)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	for (int n = 0; n < nvars; ++n)
		rhs(i, j, k, n) = (1.0 / dx0) * (fx(i, j, k, n) - fx(i + 1, j, k, n)) + (1.0 / dx1) * (fy(i, j, k, n) - fy(i, j + 1, k, n)) +
				  (1.0 / dx2) * (fz(i, j, k, n) - fz(i, j, k + 1, n));
}

// HydroSystem::AddInternalEnergyPdV  hydro_system.hpp:775-814
__global__ void __launch_bounds__(TPB) k_pdv(HydroConst c, Iter it, A4 rhs, A4 u, A4 vx, A4 vy, A4 vz, IA4 redo, double dx0, double dx1, double dx2)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	const double P = cons_pressure(c, u(i, j, k, 0), u(i, j, k, 1), u(i, j, k, 2), u(i, j, k, 3), u(i, j, k, 4));
	double div_v;
	if (redo(i, j, k) == 0) {
		div_v = (vx(i + 1, j, k, 0) - vx(i, j, k, 0)) / dx0 + (vy(i, j + 1, k, 0) - vy(i, j, k, 0)) / dx1 + (vz(i, j, k + 1, 0) - vz(i, j, k, 0)) / dx2;
	} else {
		div_v = 0.5 * ((u(i + 1, j, k, 1) / u(i + 1, j, k, 0) - u(i - 1, j, k, 1) / u(i - 1, j, k, 0)) / dx0 +
			       (u(i, j + 1, k, 2) / u(i, j + 1, k, 0) - u(i, j - 1, k, 2) / u(i, j - 1, k, 0)) / dx1 +
			       (u(i, j, k + 1, 3) / u(i, j, k + 1, 0) - u(i, j, k - 1, 3) / u(i, j, k - 1, 0)) / dx2);
	}
	rhs(i, j, k, 5) = rhs(i, j, k, 5) + (-P * div_v);
}

// HydroSystem::PredictStep  hydro_system.hpp:475-497 (+ redoFlag.sum)
__global__ void __launch_bounds__(TPB) k_predict(HydroConst c, Iter it, int nvars, A4 uo, A4 un, A4 rhs, double dt, IA4 redo, unsigned long long *nbad)
{
	int i, j, k;
	int bad = 0;
	if (it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k)) {
		for (int n = 0; n < nvars; ++n)
			un(i, j, k, n) = uo(i, j, k, n) + dt * rhs(i, j, k, n);
		bad = !(un(i, j, k, 0) > 0.);
		for (int n = 0; n < c.nms; ++n)
			if (un(i, j, k, 6 + n) < 0.0)
				bad = 1;
		redo(i, j, k) = bad;
	}
	const int cnt = __syncthreads_count(bad);
	if (threadIdx.x == 0 && cnt > 0)
		atomicAdd(nbad, (unsigned long long)cnt);
}

// HydroSystem::EnforceLimits on one cell's state held in registers  hydro_system.hpp:698-773
__device__ __forceinline__ void enforce_limits_cell(const HydroConst &c, double *U)
{
	const double rho = U[0];
	double rho_new = rho;
	if (rho < c.dfloor) {
		rho_new = c.dfloor;
		U[0] = rho_new;
		for (int n = 0; n < c.ns; ++n) {
			if (rho_new == 0.0)
				U[6 + n] = 0.0;
			else
				U[6 + n] *= rho / rho_new;
		}
	}
	if (c.nms > 0) {
		double sp_sum = 0.0;
		for (int n = 0; n < c.nms; ++n) {
			if (U[6 + n] < 0.0)
				U[6 + n] = c.small_x * rho_new;
			sp_sum += U[6 + n];
		}
		if ((sp_sum > 2.2250738585072014e-308) && (rho_new > 2.2250738585072014e-308)) {
			sp_sum /= rho_new;
			for (int n = 0; n < c.nms; ++n)
				U[6 + n] /= sp_sum;
		}
	}
	if ((rho_new > 2.2250738585072014e-308) && !c.iso) { // hydro_system.hpp:746
		const double vx1 = U[1] / rho_new, vx2 = U[2] / rho_new, vx3 = U[3] / rho_new;
		const double Ekin = 0.5 * rho_new * (vx1 * vx1 + vx2 * vx2 + vx3 * vx3);
		const double Etot = U[4];
		const double primTemp = eos_tgas_from_eint(c, rho_new, (Etot - Ekin));
		if (primTemp < c.tfloor) {
			U[4] = Ekin + eos_eint_from_tgas(c, rho_new, c.tfloor);
		}
		const double auxTemp = eos_tgas_from_eint(c, rho_new, U[5]);
		if (auxTemp < c.tfloor) {
			U[5] = eos_eint_from_tgas(c, rho_new, c.tfloor);
		}
	}
}

// HydroSystem::SyncDualEnergy on one cell  hydro_system.hpp:816-850; returns false where the reference aborts (rho<=0)
__device__ __forceinline__ bool sync_dual_energy_cell(double *U)
{
	const double eta = 1.0e-3;
	const double rho = U[0];
	if (rho <= 0.)
		return false;
	const double Ekin = (U[1] * U[1] + U[2] * U[2] + U[3] * U[3]) / (2.0 * rho);
	const double Eint_cons = U[4] - Ekin;
	if (Eint_cons > eta * U[4]) {
		U[5] = Eint_cons;
	} else {
		U[4] = U[5] + Ekin;
	}
	return true;
}

__global__ void __launch_bounds__(TPB) k_enforce(HydroConst c, Iter it, A4 s)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	double U[QK_MAXV];
	for (int n = 0; n < c.nv; ++n)
		U[n] = s(i, j, k, n);
	enforce_limits_cell(c, U);
	for (int n = 0; n < c.nv; ++n)
		s(i, j, k, n) = U[n];
}

__global__ void __launch_bounds__(TPB) k_sync(Iter it, A4 s, unsigned long long *nabort)
{
	int i, j, k;
	int bad = 0;
	if (it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k)) {
		double U[6];
		for (int n = 0; n < 6; ++n)
			U[n] = s(i, j, k, n);
		if (sync_dual_energy_cell(U)) {
			s(i, j, k, 4) = U[4];
			s(i, j, k, 5) = U[5];
		} else {
			bad = 1;
		}
	}
	const int cnt = __syncthreads_count(bad);
	if (threadIdx.x == 0 && cnt > 0 && nabort)
		atomicAdd(nabort, (unsigned long long)cnt);
}

// QuokkaSimulation::replaceFluxes, one direction  QuokkaSimulation.hpp:1324-1368 (iterates grow(valid,1))
__global__ void __launch_bounds__(TPB) k_replace(int dir, Iter it, int ncomp, A4 flux, A4 fo, IA4 redo, Box3 fbx)
{
	int i, j, k;
	if (!it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k))
		return;
	if (redo(i, j, k) != 1)
		return;
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	for (int s = 0; s < 2; ++s) {
		const int fi = i + s * e0, fj = j + s * e1, fk = k + s * e2;
		if (fi < fbx.lo[0] || fi > fbx.hi[0] || fj < fbx.lo[1] || fj > fbx.hi[1] || fk < fbx.lo[2] || fk > fbx.hi[2])
			continue;
		for (int n = 0; n < ncomp; ++n)
			flux(fi, fj, fk, n) = fo(fi, fj, fk, n);
	}
}

// order-preserving map of non-negative doubles to unsigned 64-bit for atomicMax
__device__ __forceinline__ unsigned long long d2key(double v)
{
	long long b = __double_as_longlong(v);
	return (b < 0) ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

// ComputeMaxSignalSpeed+norminf (which=0) / maxSignalSpeedLocal (which=1)  hydro_system.hpp:198-252
__global__ void __launch_bounds__(TPB) k_max_signal(HydroConst c, int which, Iter it, A4 u, unsigned long long *out)
{
	int i, j, k;
	double sig = (which == 0) ? 0.0 : -1.7976931348623157e308;
	if (it.get((int64_t)blockIdx.x * TPB + threadIdx.x, i, j, k)) {
		const double rho = u(i, j, k, 0), px = u(i, j, k, 1), py = u(i, j, k, 2), pz = u(i, j, k, 3), E = u(i, j, k, 4);
		const double P = cons_pressure(c, rho, px, py, pz, E);
		const double cs = c.iso ? c.cs_iso : eos_sound_speed(c, rho, P); // hydro_system.hpp:214-218, 242-246
		if (which == 0) {
			const double vx = px / rho, vy = py / rho, vz = pz / rho;
			sig = fabs(cs + sqrt(vx * vx + vy * vy + vz * vz));
		} else {
			const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
			sig = cs + sqrt(2.0 * kinetic_energy / rho);
		}
	}
	// NaN never wins a '<' comparison, as in the reference's max reductions
	__shared__ double sm[TPB / 32];
	for (int o = 16; o > 0; o >>= 1) {
		const double other = __shfl_xor_sync(0xffffffffu, sig, o);
		sig = dmax(sig, other);
	}
	if ((threadIdx.x & 31) == 0)
		sm[threadIdx.x >> 5] = sig;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < TPB / 32; ++w)
			sig = dmax(sig, sm[w]);
		if (!(sig != sig))
			atomicMax(out, d2key(sig));
	}
}


// both maxima in one pass over a box (grid-stride): out[0] = ComputeMaxSignalSpeed + norminf, out[1] = maxSignalSpeedLocal.
// The per-cell values are the ones k_max_signal forms; max is associative, so the result is identical.
__global__ void __launch_bounds__(TPB) k_max_signal2(HydroConst c, Iter it, A4 u, unsigned long long *out)
{
	double s0 = 0.0, s1 = -1.7976931348623157e308;
	for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < it.total; t += (int64_t)gridDim.x * TPB) {
		int i, j, k;
		it.get(t, i, j, k);
		const int64_t o = u.off(i, j, k);
		const double rho = u.p[o], px = u.p[o + u.ns], py = u.p[o + 2 * u.ns], pz = u.p[o + 3 * u.ns], E = u.p[o + 4 * u.ns];
		const double P = cons_pressure(c, rho, px, py, pz, E);
		const double cs = c.iso ? c.cs_iso : eos_sound_speed(c, rho, P);
		const double vx = px / rho, vy = py / rho, vz = pz / rho;
		s0 = dmax(s0, fabs(cs + sqrt(vx * vx + vy * vy + vz * vz)));
		const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
		s1 = dmax(s1, cs + sqrt(2.0 * kinetic_energy / rho));
	}
	__shared__ double sm[2][TPB / 32];
	for (int o = 16; o > 0; o >>= 1) {
		s0 = dmax(s0, __shfl_xor_sync(0xffffffffu, s0, o));
		s1 = dmax(s1, __shfl_xor_sync(0xffffffffu, s1, o));
	}
	if ((threadIdx.x & 31) == 0) {
		sm[0][threadIdx.x >> 5] = s0;
		sm[1][threadIdx.x >> 5] = s1;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < TPB / 32; ++w) {
			s0 = dmax(s0, sm[0][w]);
			s1 = dmax(s1, sm[1][w]);
		}
		if (!(s0 != s0))
			atomicMax(out, d2key(s0));
		if (!(s1 != s1))
			atomicMax(out + 1, d2key(s1));
	}
}

inline double key2d(unsigned long long k)
{
	unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
	double v;
	memcpy(&v, &b, 8);
	return v;
}
inline cudaStream_t S(void *s) { return (cudaStream_t)s; }
} // namespace
