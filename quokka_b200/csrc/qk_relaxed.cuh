// qk_relaxed.cuh -- the RELAXED arithmetic variant of the fused sweep physics (QK_ARITH_FAST).
//
// Same algorithm, same branches, same guards as qk_fast.cuh / the reference, but values are no longer required to be the
// reference's bits: the gamma-law EOS is evaluated in closed form (p = (gamma-1) rho e, c_s^2 = gamma p / rho,
// rho de/dp = 1/(gamma-1), e dr/dp = E/p) instead of through the Microphysics temperature round trip, quotients are a
// product with a reciprocal refined to ~1 ulp (MUFU seed + one cubic Newton step, no correction step, no domain
// checks), the two |v| square roots of the carbuncle switch become one (sqrt is monotonic), and the translation unit
// that instantiates these (qk_sweep_relaxed.cu) is compiled with FMA contraction ON.  Every change perturbs a result by
// O(1 ulp); the drift against the exact path is asserted in tests/test_gpu_relaxed.py (<= 1e-12 of max|U| per component
// after 100 Sedov steps, the tolerance BASELINE.json states).  Non-finite results are caught by the stage epilogue exactly
// as in the exact path and the stage is redone by the faithful (exact) path.
#pragma once
#include "qk_fast.cuh"

// 1/b to about 1 ulp for finite normal b (NaN for b = 0, inf or subnormal: caught downstream as a non-finite cell)
__device__ __forceinline__ double r_rcp(double b)
{
	const double y0 = __hiloint2double(__double2hiint(qk_mufu_rcp64h(b)), 0);
	double e = __fma_rn(-b, y0, 1.0);
	e = __fma_rn(e, e, e);
	return __fma_rn(y0, e, y0);
}

// 1/sqrt(x) to about 1 ulp for finite normal x > 0 (rsqrt seed + two coupled Newton steps; no slow path: x = 0, inf, subnormal or negative
// give inf / NaN, caught downstream as a non-finite cell)
__device__ __forceinline__ double r_rsqrt(double x)
{
	double y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
	const double hx = 0.5 * x;
	double e = __fma_rn(-hx * y0, y0, 0.5); // 0.5 (1 - x y0^2)
	y0 = __fma_rn(y0, e, y0);
	e = __fma_rn(-hx * y0, y0, 0.5);
	return __fma_rn(y0, e, y0);
}
// sqrt(x) = x * rsqrt(x) for x > 0; sqrt(0) = 0 is kept (a cold gas at rest has |v| = 0), NaN for x < 0 as sqrt
__device__ __forceinline__ double r_sqrtx(double x)
{
	const double s = x * r_rsqrt(x);
	return (x == 0.0) ? 0.0 : s;
}

// PPM limiting + FlattenShocks of one cell (ppm_limit / f_ppm_flat of the exact path) with the monotonized-central slope of the extremum
// branch formed without the int -> double sign arithmetic: sign(a) min(...) where a b > 0, else 0 (same value except where a * b underflows)
__device__ __forceinline__ void r_ppm_flat(double qm1, double q0, double qp1, double if_lo, double if_hi, double chi, double omchi, double &am, double &ap)
{
	// bounds = minmax(q0, qm1, qp1).  q0 can be left out: where it is the extremum of the three (a b <= 0) both clamped edges land on the same
	// side of q0, the extremum branch is taken and its slope h is 0 -- the cell is reconstructed as q0 whatever the bounds were
	const bool up = (qm1 < qp1);
	const double lo = up ? qm1 : qp1, hi = up ? qp1 : qm1;
	const double am0 = clampd(if_lo, lo, hi), ap0 = clampd(if_hi, lo, hi);
	const double dq_minus = q0 - am0, dq_plus = ap0 - q0;
	const double a = qp1 - q0, b = q0 - qm1;
	const double m = dmin(0.25 * fabs(a + b), dmin(fabs(a), fabs(b))); // 0.5 * MC
	const double h = (a * b > 0.0) ? copysign(m, a) : 0.0;		    // 0.5 * dq0
	const bool ext = (dq_plus * dq_minus <= 0.0);
	// edge offsets from q0: a_minus = q0 - em, a_plus = q0 + ep; FlattenShocks (hydro_system.hpp:688-691) chi a + (1 - chi) q0 = q0 -+ chi e
	const double tm = 2.0 * dq_plus, tp = 2.0 * dq_minus;
	double em = (fabs(dq_minus) >= fabs(tm)) ? tm : dq_minus;
	double ep = (fabs(dq_plus) >= fabs(tp)) ? tp : dq_plus;
	em = ext ? h : em;
	ep = ext ? h : ep;
	am = __fma_rn(-chi, em, q0);
	ap = __fma_rn(chi, ep, q0);
}

// pressure the EOS returns for an input pressure P: identity inside [1e-200, 1e200], else the eos_reset floor
// rho k_B T_min / (mu m_u)  (extern/Microphysics/interfaces/eos.H:97-139)
__device__ __forceinline__ double r_p_of_p(const FastConst &c, double rho, double P) { return (P < 1.e-200 || P > 1.e200) ? rho * c.pfloor : P; }
// EOS::ComputePressure with e = Eint/rho formed by the caller
__device__ __forceinline__ double r_pressure_from_e(const FastConst &c, double rho, double e)
{
	return (e < 1.e-200 || e > 1.e200) ? rho * c.pfloor : rho * e * c.h.gm1;
}

template <int NS, bool REINT> __device__ __forceinline__ void r_cons_to_prim(const FastConst &c, const double *U, double *q)
{
	const double rho = U[0];
	const double y = r_rcp(rho);
	const double vx = U[1] * y, vy = U[2] * y, vz = U[3] * y;
	const double ke = 0.5 * rho * (vx * vx + vy * vy + vz * vz);
	const double Eint_cons = U[4] - ke;
	q[0] = rho;
	q[1] = vx;
	q[2] = vy;
	q[3] = vz;
	if (REINT) {
		q[4] = Eint_cons * y;
		q[5] = U[5] * y;
	} else {
		q[4] = r_pressure_from_e(c, rho, (rho == 0.0) ? 0.0 : Eint_cons * y);
		q[5] = U[5];
	}
#pragma unroll
	for (int n = 0; n < NS; ++n)
		q[6 + n] = U[6 + n];
}

// Miller-Colella chi along one direction; yKS = 1 / (rho c_s^2) = 1 / (gamma p)
__device__ __forceinline__ double r_flatten_chi(const FastConst &c, double Pm2, double Pm1, double Pp1, double Pp2, double yKS, double vm1, double vp1)
{
	const double beta_max = 0.85, Zmax = 0.75, Zmin = 0.25;
	const double beta_denom = fabs(Pp2 - Pm2);
	const double dP1 = fabs(Pp1 - Pm1);
	const double beta = (beta_denom != 0) ? dP1 * r_rcp(beta_denom) : 0;
	const double chi_min = dmax(0., dmin(1., (beta_max - beta) * c.y_dbeta));
	const double Z = dP1 * yKS;
	double chi = 1.0;
	if (vp1 < vm1)
		chi = dmax(chi_min, dmin(1., (Zmax - Z) / (Zmax - Zmin)));
	return chi;
}

template <int DIR, int NS, int NMS, bool REINT>
__device__ __forceinline__ void r_hllc(const FastConst &c, const double *__restrict__ L, const double *__restrict__ R, double du, double dw,
				       double *__restrict__ F, double &vface)
{
	constexpr int iN = 1 + DIR, iV = 1 + (DIR + 1) % 3, iW = 1 + (DIR + 2) % 3;
	const double rho_L = L[0], rho_R = R[0];
	// one refined 1/sqrt(rho) per state gives both sqrt(rho) (Roe weights) and 1/rho
	const double qL = r_rsqrt(rho_L), qR = r_rsqrt(rho_R);
	const double yL = qL * qL, yR = qR * qR;
	const double vsq_L = L[1] * L[1] + L[2] * L[2] + L[3] * L[3];
	const double vsq_R = R[1] * R[1] + R[2] * R[2] + R[3] * R[3];
	const double ke_L = 0.5 * rho_L * vsq_L, ke_R = 0.5 * rho_R * vsq_R;
	double P_L, P_R, Eint_L, Eint_R;
	if (REINT) {
		P_L = r_pressure_from_e(c, rho_L, (rho_L == 0.0) ? 0.0 : L[4]);
		P_R = r_pressure_from_e(c, rho_R, (rho_R == 0.0) ? 0.0 : R[4]);
		Eint_L = rho_L * L[5];
		Eint_R = rho_R * R[5];
	} else {
		P_L = L[4];
		P_R = R[4];
		Eint_L = L[5];
		Eint_R = R[5];
	}
	// one eos(rp) per state in closed form
	const double p_L = r_p_of_p(c, rho_L, P_L), p_R = r_p_of_p(c, rho_R, P_R);
	const double cs_L = r_sqrtx(c.h.gamma * p_L * yL), cs_R = r_sqrtx(c.h.gamma * p_R * yR);
	const double E_L = p_L * c.inv_gm1 + ke_L;
	const double E_R = p_R * c.inv_gm1 + ke_R;
	const double uL = L[iN], vL = L[iV], wL = L[iW];
	const double uR = R[iN], vR = R[iV], wR = R[iW];

	const double wl = rho_L * qL, wr = rho_R * qR;
	const double norm = r_rcp(wl + wr);
	const double u_tilde = (wl * uL + wr * uR) * norm;
	const double v_tilde = (wl * vL + wr * vR) * norm;
	const double w_tilde = (wl * wL + wr * wR) * norm;
	const double vsq_tilde = u_tilde * u_tilde + v_tilde * v_tilde + w_tilde * w_tilde;
	const double H_L = (E_L + P_L) * yL, H_R = (E_R + P_R) * yR;
	const double H_tilde = (wl * H_L + wr * H_R) * norm;
	const double dU = uL - uR;
	const double C_tilde_rho = 0.5 * (Eint_L * yL + Eint_R * yR);
	const double C_tilde_P = 0.5 * c.bk * (Eint_L * r_rcp(p_L) + Eint_R * r_rcp(p_R)) + c.inv_gm1;
	const double cs_exp = H_tilde - 0.5 * vsq_tilde - C_tilde_rho;
	const double cs_tilde = (cs_exp <= 0) ? 0.5 * (cs_L + cs_R) : r_sqrtx(cs_exp * r_rcp(C_tilde_P));
	const double s_NL = 0.5 * c.h.G * dmax(dU, 0.);
	const double S_L = dmin(uL - (cs_L + s_NL), u_tilde - (cs_tilde + s_NL));
	const double S_R = dmax(uR + (cs_R + s_NL), u_tilde + (cs_tilde + s_NL));
	const double cs_max = dmax(cs_L, cs_R);
	const double tp = dmin(1., (cs_max - dmin(du, 0.)) * r_rcp(cs_max - dmin(dw, 0.)));
	const double theta = (tp * tp) * (tp * tp);
	const double mL = rho_L * (S_L - uL), mR = rho_R * (S_R - uR);
	const double S_star = (theta * (P_R - P_L) + (mL * uL - mR * uR)) * r_rcp(mL - mR);
	const double chi = dmin(1., r_sqrtx(dmax(vsq_L, vsq_R)) * r_rcp(cs_max));
	const double phi = chi * (2. - chi);
	const double P_LR = 0.5 * (P_L + P_R) + 0.5 * phi * (mL * (S_star - uL) + mR * (S_star - uR));

	int region;
	if (S_L > 0.0) {
		region = 0;
	} else if ((S_star > 0.0) && (S_L <= 0.0)) {
		region = 1;
	} else if ((S_star <= 0.0) && (S_R >= 0.0)) {
		region = 2;
	} else {
		region = 3;
	}
	const bool left = (region < 2);
	const bool star = (region == 1) || (region == 2);
	const double rK = left ? rho_L : rho_R, uK = left ? uL : uR, vK = left ? vL : vR, wK = left ? wL : wR;
	const double PK = left ? P_L : P_R, EK = left ? E_L : E_R, EiK = left ? Eint_L : Eint_R, SK = left ? S_L : S_R;
	const double UK[6] = {rK, rK * uK, rK * vK, rK * wK, EK, EiK};
	const double SP = SK * P_LR;
	double Fc[6];
#pragma unroll
	for (int n = 0; n < 6; ++n) {
		double FK = uK * UK[n];
		if (n == 1)
			FK = FK + PK;
		if (n == 4)
			FK = FK + PK * uK;
		Fc[n] = FK;
	}
	double Fsc[NS > 0 ? NS : 1];
#pragma unroll
	for (int n = 0; n < NS; ++n)
		Fsc[n] = uK * (left ? L[6 + n] : R[6 + n]);
	if (star) {
		const double yden = r_rcp(SK - S_star);
		const double sy = S_star * yden;
#pragma unroll
		for (int n = 0; n < 6; ++n) {
			double num = SK * UK[n] - Fc[n];
			double f = sy * num;
			if (n == 1)
				f = f + SP * yden;
			if (n == 4)
				f = f + (SP * S_star) * yden;
			Fc[n] = f;
		}
#pragma unroll
		for (int n = 0; n < NS; ++n) {
			const double Un = left ? L[6 + n] : R[6 + n];
			Fsc[n] = sy * (SK * Un - Fsc[n]);
		}
	}
	F[0] = Fc[0];
	F[iN] = Fc[1];
	F[iV] = Fc[2];
	F[iW] = Fc[3];
	F[4] = Fc[4];
	F[5] = Fc[5];
#pragma unroll
	for (int n = 0; n < NS; ++n)
		F[6 + n] = Fsc[n];
	vface = (F[0] >= 0.) ? F[0] * yR : F[0] * yL;
	if (NMS > 0) {
		double sumL = 0, sumR = 0;
#pragma unroll
		for (int n = 0; n < NMS; ++n) {
			sumL += L[6 + n];
			sumR += R[6 + n];
		}
		const bool pos = (F[0] >= 0.);
		const double ys = F[0] * r_rcp(pos ? sumL : sumR);
#pragma unroll
		for (int n = 0; n < NMS; ++n)
			F[6 + n] = ys * (pos ? L[6 + n] : R[6 + n]);
	}
}
