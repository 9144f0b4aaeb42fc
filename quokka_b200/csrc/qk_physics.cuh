// qk_physics.cuh -- per-cell / per-face device functions of the hydro hot path.
//
// EXACT arithmetic contract: every expression below keeps the reference's IEEE operation order and
// this translation unit is compiled with --fmad=false (as the reference's CUDA build,
// CMakeLists.txt:31, src/CMakeLists.txt:70-88), so results are bit-identical to the reference's
// CPU and GPU paths.  What IS changed, without changing any rounded value:
//   * the three Microphysics eos(rp) calls the reference makes per reconstructed state
//     (ComputeSoundSpeed, ComputeEintFromPres, ComputeOtherDerivatives; hydro_system.hpp:940-946,
//     HLLC.hpp:49-50) are evaluated once and shared;
//   * only the selected branch of the HLLC fan is evaluated (HLLC.hpp:135-150): 6 divides, not 12;
//   * terms that are exactly zero for the gamma-law EOS (rho*dedr, P*D[n] with D[n]=0) are
//     dropped: they can only turn a -0.0 into +0.0, never change a non-zero value;
//   * the unlimited PPM interface value is shared between a_plus(i) and a_minus(i+1), which the
//     reference computes twice with the same grouping (hyperbolic_system.hpp:380-385).
// The dropped zero terms matter only for NON-FINITE intermediates (inf*0 = NaN): face_flux<.., SPECIAL=true>
// keeps every one of them (and the K_visc=0 artificial-viscosity term, hydro_system.hpp:1052-1076) so that
// unphysical states (rho <= 0 surviving a failed FOFC) propagate NaN/Inf exactly as the reference does; the
// faithful path uses SPECIAL=true, the fused sweeps SPECIAL=false and are only run on states whose previous
// stage reported ncells_bad == 0 (qk_sweep.cu).
#pragma once
#include "qk_common.cuh"

// extern/Microphysics/constants/fundamental_constants.H:22,55
#define QK_K_B 1.3806488e-16
#define QK_M_U 1.6605390666e-24

struct HydroConst {
	double gamma, gm1, mu, mumn, boltz, mintemp, mindens, dfloor, tfloor, small_x, G, K_visc;
	int reconstruct_eint, ns, nms, nv;
	// isothermal EOS (gamma == 1: HydroSystem::is_eos_isothermal(), hydro_system.hpp:132-133): P = rho cs_iso^2, c_s = cs_iso, no energy fluxes.
	// Only the one-kernel-per-operator path runs it (qk_sweep.cu sends gamma == 1 there).
	double cs_iso;
	int iso;
};

inline HydroConst make_hydro_const(const qk_hydro_params *p)
{
	HydroConst c;
	c.gamma = p->gamma;
	c.gm1 = p->gamma - 1.0;
	c.mu = p->mean_molecular_weight / QK_M_U; // src/hydro/EOS.hpp:104
	c.mumn = c.mu * QK_M_U;
	c.boltz = p->boltzmann_constant;
	c.mintemp = (1.e-200 < p->small_temp) ? p->small_temp : 1.e-200; // eos_init, interfaces/eos.H:40-44
	c.mindens = (1.e-200 < p->small_dens) ? p->small_dens : 1.e-200;
	c.dfloor = p->density_floor;
	c.tfloor = p->temp_floor;
	c.small_x = p->small_x;
	c.G = 0.5 * (1.0 + p->gamma); // actual_eos.H:267
	c.K_visc = p->K_visc;
	c.reconstruct_eint = p->reconstruct_eint;
	c.ns = p->nscalars;
	c.nms = p->nmscalars;
	c.nv = 6 + p->nscalars;
	c.iso = (p->gamma == 1.0) ? 1 : 0;
	c.cs_iso = p->cs_isothermal;
	return c;
}

// ---- gamma-law EOS through temperature (interfaces/eos.H:395-435,141-205,69-79; gamma_law/actual_eos.H) ----
__device__ __forceinline__ double eos_clamp_rho(const HydroConst &c, double rho) { return dmin(1.e200, dmax(c.mindens, rho)); }

// eos(eos_input_re): returns T; rho is clamped in place
__device__ __forceinline__ double eos_T_from_re(const HydroConst &c, double &rho, double e)
{
	rho = eos_clamp_rho(c, rho);
	if (e < 1.e-200 || e > 1.e200) {
		return c.mintemp; // eos_reset: T = clamp(0) = mintemp
	}
	return e * c.mu * QK_M_U * c.gm1 / QK_K_B; // actual_eos.H:128
}
// eos(eos_input_rp)
__device__ __forceinline__ double eos_T_from_rp(const HydroConst &c, double &rho, double p)
{
	rho = eos_clamp_rho(c, rho);
	if (p < 1.e-200 || p > 1.e200) {
		return c.mintemp;
	}
	return p * c.mu * QK_M_U / (QK_K_B * rho); // actual_eos.H:115
}
// pressure = rho T k_B/(mu m_u)  actual_eos.H:199
__device__ __forceinline__ double eos_p_of_rT(const HydroConst &c, double rho, double T) { return rho * T * QK_K_B / c.mumn; }

// EOS::ComputePressure(rho, Eint)  src/hydro/EOS.hpp:299-340
__device__ __forceinline__ double eos_pressure(const HydroConst &c, double rho, double Eint)
{
	const double e = (rho == 0.0) ? 0.0 : Eint / rho;
	double r = rho;
	const double T = eos_T_from_re(c, r, e);
	return eos_p_of_rT(c, r, T);
}
// EOS::ComputeSoundSpeed(rho, P)  EOS.hpp:342-383
__device__ __forceinline__ double eos_sound_speed(const HydroConst &c, double rho, double P)
{
	double r = rho;
	const double T = eos_T_from_rp(c, r, P);
	const double p = eos_p_of_rT(c, r, T);
	const double rhoinv = 1.0 / r;
	return sqrt(c.gamma * p * rhoinv);
}
// EOS::ComputeTgasFromEint  EOS.hpp:75-114
__device__ __forceinline__ double eos_tgas_from_eint(const HydroConst &c, double rho, double Eint)
{
	double r = rho;
	const double T = eos_T_from_re(c, r, Eint / rho);
	return T * QK_K_B / c.boltz;
}
// EOS::ComputeEintFromTgas  EOS.hpp:116-157 (eos_input_rt: rho and T clamped)
__device__ __forceinline__ double eos_eint_from_tgas(const HydroConst &c, double rho, double Tgas)
{
	const double r = eos_clamp_rho(c, rho);
	const double T = dmin(1.e200, dmax(c.mintemp, Tgas));
	const double p = eos_p_of_rT(c, r, T);
	const double e = p / c.gm1 * (1.0 / r);
	return e * rho * c.boltz / QK_K_B;
}

// everything the flux kernel needs from ONE eos(eos_input_rp) evaluation of a reconstructed state
struct EosRP {
	double cs;   // EOS::ComputeSoundSpeed
	double Eint; // EOS::ComputeEintFromPres = e * rho
	double dedp; // 1/dpde         (ComputeOtherDerivatives, EOS.hpp:288-291)
	double drdp; // 1/(dpdr k_B/boltz)
};
__device__ __forceinline__ EosRP eos_rp_all(const HydroConst &c, double rho, double P)
{
	double r = rho;
	const double T = eos_T_from_rp(c, r, P);
	const double Tinv = 1.0 / T;
	const double rhoinv = 1.0 / r;
	const double p = eos_p_of_rT(c, r, T);
	const double e = p / c.gm1 * rhoinv;
	const double dpdT = p * Tinv;
	const double dpdr = p * rhoinv;
	const double dedT = e * Tinv;
	const double dpde = dpdT * (1.0 / dedT);
	EosRP o;
	o.cs = sqrt(c.gamma * p * rhoinv);
	o.Eint = e * rho;
	o.dedp = 1.0 / dpde;
	o.drdp = 1.0 / (dpdr * QK_K_B / c.boltz);
	return o;
}

// HydroSystem::ComputePressure(cons,i,j,k)  src/hydro/hydro_system.hpp:349-372
__device__ __forceinline__ double cons_pressure(const HydroConst &c, double rho, double px, double py, double pz, double E)
{
	if (c.iso)
		return rho * c.cs_iso * c.cs_iso; // hydro_system.hpp:365-366
	const double vx = px / rho, vy = py / rho, vz = pz / rho;
	const double ke = 0.5 * rho * (vx * vx + vy * vy + vz * vz);
	return eos_pressure(c, rho, E - ke);
}

// ---- limiters + reconstruction (src/hyperbolic_system.hpp:58-66, 164-181, 218-247, 337-433) --------
__device__ __forceinline__ double lim_MC(double a, double b)
{
	return 0.5 * (sgnd(a) + sgnd(b)) * dmin(0.5 * fabs(a + b), dmin(2.0 * fabs(a), 2.0 * fabs(b)));
}
__device__ __forceinline__ double lim_minmod(double a, double b) { return 0.5 * (sgnd(a) + sgnd(b)) * dmin(fabs(a), fabs(b)); }

// unlimited 4th-order interface between cells (m1 | 0): (7/12)(q0+qm1) - (1/12)(qp1+qm2), grouped as :382-385
__device__ __forceinline__ double ppm_iface(double qm2, double qm1, double q0, double qp1)
{
	const double coef_1 = (7. / 12.);
	const double coef_2 = (-1. / 12.);
	return (coef_1 * q0 + coef_2 * qp1) + (coef_1 * qm1 + coef_2 * qm2);
}

// PPM limiting of one cell given its two unlimited interface values; returns a_minus (left edge), a_plus (right edge)
__device__ __forceinline__ void ppm_limit(double qm1, double q0, double qp1, double a_minus, double a_plus, double &am, double &ap)
{
	// bounds = std::minmax({q0, qm1, qp1})
	double lo = q0, hi = q0;
	if (qm1 < lo) lo = qm1;
	if (qp1 < lo) lo = qp1;
	if (!(qm1 < hi)) hi = qm1;
	if (!(qp1 < hi)) hi = qp1;
	double new_a_minus = clampd(a_minus, lo, hi);
	double new_a_plus = clampd(a_plus, lo, hi);
	const double a = q0;
	const double dq_minus = (a - new_a_minus);
	const double dq_plus = (new_a_plus - a);
	const double qa = dq_plus * dq_minus;
	if (qa <= 0.0) {
		const double dq0 = lim_MC(qp1 - q0, q0 - qm1);
		new_a_minus = a - 0.5 * dq0;
		new_a_plus = a + 0.5 * dq0;
	} else {
		if (fabs(dq_minus) >= 2.0 * fabs(dq_plus)) {
			new_a_minus = a - 2.0 * dq_plus;
		}
		if (fabs(dq_plus) >= 2.0 * fabs(dq_minus)) {
			new_a_plus = a + 2.0 * dq_minus;
		}
	}
	am = new_a_minus;
	ap = new_a_plus;
}

// order 1|2|3 reconstruction of one cell: am = state at its left face (rightState(i)), ap = at its right face (leftState(i+1)).
// NOTE PLM in the reference is interface-centred (:243-246): left(i) = q(i-1)+0.25*lim(q(i)-q(i-1), q(i-1)-q(i-2)),
// right(i) = q(i)-0.25*lim(q(i+1)-q(i), q(i)-q(i-1)); expressed per cell: ap(c) = q(c)+0.25*lim(q(c+1)-q(c), q(c)-q(c-1)).
template <int ORDER, int LIMITER> __device__ __forceinline__ void recon_cell(double qm2, double qm1, double q0, double qp1, double qp2, double &am, double &ap)
{
	if (ORDER == 1) {
		am = q0;
		ap = q0;
	} else if (ORDER == 2) {
		const double s = (LIMITER == QK_MC) ? lim_MC(qp1 - q0, q0 - qm1) : lim_minmod(qp1 - q0, q0 - qm1);
		am = q0 - 0.25 * s;
		ap = q0 + 0.25 * s;
	} else {
		ppm_limit(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), am, ap);
	}
}

// ---- Riemann solvers ------------------------------------------------------------------------------
struct FaceState { // quokka::HydroState, src/hydro/HydroState.hpp:10-23 (canonical: u normal)
	double rho, u, v, w, P, Eint;
};

// HydroSystem::ComputeFluxes<HLLC,DIR> body after the L/R gather (hydro_system.hpp:879-1104) + Riemann::HLLC (HLLC.hpp:21-153).
// prims: L[*], R[*] hold the reconstructed primitive components in ARRAY order (rho, vx, vy, vz, P|eint, Eaux|eaux, scalars..).
// Returns F in ARRAY component order and the face velocity.
template <int SOLVER, bool SPECIAL = false>
__device__ __forceinline__ void face_flux(const HydroConst &c, int dir, const double *__restrict__ L, const double *__restrict__ R, double du,
					  double dw, double *__restrict__ F, double &vface, double div_v = 0.0)
{
	const int iN = 1 + dir, iV = 1 + (dir + 1) % 3, iW = 1 + (dir + 2) % 3; // hydro_system.hpp:954-976
	const double rho_L = L[0], rho_R = R[0];
	const double ke_L = 0.5 * rho_L * (L[1] * L[1] + L[2] * L[2] + L[3] * L[3]);
	const double ke_R = 0.5 * rho_R * (R[1] * R[1] + R[2] * R[2] + R[3] * R[3]);
	double P_L, P_R, Eint_L, Eint_R;
	if (c.iso) { // hydro_system.hpp:910-915: E and Eint stay NAN in the reference; their fluxes are set to zero at the end
		P_L = rho_L * (c.cs_iso * c.cs_iso);
		P_R = rho_R * (c.cs_iso * c.cs_iso);
		Eint_L = Eint_R = __longlong_as_double(0x7ff8000000000000LL);
	} else if (c.reconstruct_eint) {
		P_L = eos_pressure(c, rho_L, L[4] * rho_L);
		P_R = eos_pressure(c, rho_R, R[4] * rho_R);
		Eint_L = rho_L * L[5];
		Eint_R = rho_R * R[5];
	} else {
		P_L = L[4];
		P_R = R[4];
		Eint_L = L[5];
		Eint_R = R[5];
	}
	EosRP eL, eR;
	if (c.iso) {
		eL.cs = eR.cs = c.cs_iso;
		eL.Eint = eR.Eint = eL.dedp = eR.dedp = eL.drdp = eR.drdp = Eint_L;
	} else {
		eL = eos_rp_all(c, rho_L, P_L);
		eR = eos_rp_all(c, rho_R, P_R);
	}
	const double cs_L = eL.cs, cs_R = eR.cs;
	const double E_L = eL.Eint + ke_L;
	const double E_R = eR.Eint + ke_R;
	const double uL = L[iN], vL = L[iV], wL = L[iW];
	const double uR = R[iN], vR = R[iV], wR = R[iW];

	if (SOLVER == QK_LLF) { // src/hydro/LLF.hpp:15-43
		const double Sp = dmax(fabs(uL) + cs_L, fabs(uR) + cs_R);
		const double hS = 0.5 * Sp;
		const double UL[6] = {rho_L, rho_L * uL, rho_L * vL, rho_L * wL, E_L, Eint_L};
		const double UR[6] = {rho_R, rho_R * uR, rho_R * vR, rho_R * wR, E_R, Eint_R};
		double Fc[6];
#pragma unroll
		for (int n = 0; n < 6; ++n) {
			double FL = uL * UL[n], FR = uR * UR[n];
			if (n == 1) {
				FL = FL + P_L;
				FR = FR + P_R;
			} else if (n == 4) {
				FL = FL + P_L * uL;
				FR = FR + P_R * uR;
			} else if (SPECIAL) {
				FL = FL + P_L * 0.;
				FR = FR + P_R * 0.;
			}
			Fc[n] = 0.5 * (FL + FR) - hS * (UR[n] - UL[n]);
		}
		F[0] = Fc[0];
		F[iN] = Fc[1];
		F[iV] = Fc[2];
		F[iW] = Fc[3];
		F[4] = Fc[4];
		F[5] = Fc[5];
		for (int n = 0; n < c.ns; ++n) {
			if (SPECIAL)
				F[6 + n] = 0.5 * ((uL * L[6 + n] + P_L * 0.) + (uR * R[6 + n] + P_R * 0.)) - hS * (R[6 + n] - L[6 + n]);
			else
				F[6 + n] = 0.5 * (uL * L[6 + n] + uR * R[6 + n]) - hS * (R[6 + n] - L[6 + n]);
		}
	} else { // HLLC
		const double wl = sqrt(rho_L);
		const double wr = sqrt(rho_R);
		const double norm = 1. / (wl + wr);
		const double u_tilde = (wl * uL + wr * uR) * norm;
		const double v_tilde = (wl * vL + wr * vR) * norm;
		const double w_tilde = (wl * wL + wr * wR) * norm;
		const double vsq_tilde = u_tilde * u_tilde + v_tilde * v_tilde + w_tilde * w_tilde;
		const double H_L = (E_L + P_L) / rho_L;
		const double H_R = (E_R + P_R) / rho_R;
		const double H_tilde = (wl * H_L + wr * H_R) * norm;
		const double dU = uL - uR;
		// gamma != 1 branch (HLLC.hpp:47-73); dedr = 0 for the gamma law (actual_eos.H:230)
		const double eiL = Eint_L / rho_L, eiR = Eint_R / rho_R;
		const double C_tilde_rho = SPECIAL ? 0.5 * (eiL + eiR + rho_L * 0. + rho_R * 0.) : 0.5 * (eiL + eiR);
		const double C_tilde_P = 0.5 * (eiL * eL.drdp + eiR * eR.drdp + rho_L * eL.dedp + rho_R * eR.dedp);
		const double cs_exp = H_tilde - 0.5 * vsq_tilde - C_tilde_rho;
		double cs_tilde;
		if (c.iso || cs_exp <= 0) { // gamma == 1: HLLC.hpp:76-88 (G_L = G_R = 1 = c.G)
			cs_tilde = 0.5 * (cs_L + cs_R);
		} else {
			cs_tilde = sqrt(cs_exp / C_tilde_P);
		}
		const double s_NL = 0.5 * c.G * dmax(dU, 0.);
		const double s_NR = s_NL; // G_L == G_R == (1+gamma)/2
		const double S_L = dmin(uL - (cs_L + s_NL), u_tilde - (cs_tilde + s_NL));
		const double S_R = dmax(uR + (cs_R + s_NR), u_tilde + (cs_tilde + s_NR));
		const double cs_max = dmax(cs_L, cs_R);
		const double tp = dmin(1., (cs_max - dmin(du, 0.)) / (cs_max - dmin(dw, 0.)));
		const double theta = tp * tp * tp * tp;
		const double S_star = (theta * (P_R - P_L) + (rho_L * uL * (S_L - uL) - rho_R * uR * (S_R - uR))) / (rho_L * (S_L - uL) - rho_R * (S_R - uR));
		const double vmag_L = sqrt(uL * uL + vL * vL + wL * wL);
		const double vmag_R = sqrt(uR * uR + vR * vR + wR * wR);
		const double chi = dmin(1., dmax(vmag_L, vmag_R) / cs_max);
		const double phi = chi * (2. - chi);
		const double P_LR = 0.5 * (P_L + P_R) + 0.5 * phi * (rho_L * (S_L - uL) * (S_star - uL) + rho_R * (S_R - uR) * (S_star - uR));

		// fan selection (HLLC.hpp:142-150): region 0: F_L, 1: F*_L, 2: F*_R, 3: F_R
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
		const double den = SK - S_star;
		double Fc[6];
#pragma unroll
		for (int n = 0; n < 6; ++n) {
			double FK = uK * UK[n];
			if (n == 1) {
				FK = FK + PK;
			} else if (n == 4) {
				FK = FK + PK * uK;
			} else if (SPECIAL) {
				FK = FK + PK * 0.;
			}
			double Fs = FK;
			if (star) {
				double num = S_star * (SK * UK[n] - FK);
				if (n == 1) {
					num = num + SP;
				} else if (n == 4) {
					num = num + SP * S_star;
				} else if (SPECIAL) {
					num = num + SP * 0.;
				}
				Fs = num / den;
			}
			Fc[n] = Fs;
		}
		F[0] = Fc[0];
		F[iN] = Fc[1];
		F[iV] = Fc[2];
		F[iW] = Fc[3];
		F[4] = Fc[4];
		F[5] = Fc[5];
		for (int n = 0; n < c.ns; ++n) {
			const double Un = left ? L[6 + n] : R[6 + n];
			const double FK = SPECIAL ? (uK * Un + PK * 0.) : (uK * Un);
			if (SPECIAL)
				F[6 + n] = star ? (S_star * (SK * Un - FK) + SP * 0.) / den : FK;
			else
				F[6 + n] = star ? (S_star * (SK * Un - FK)) / den : FK;
		}
	}
	if (SPECIAL) {
		// artificial viscosity with K_visc = 0 (hydro_system.hpp:1052-1076): adds +0 to every non-momentum component
		// unless the velocity divergence or a state difference is not finite
		const double viscosity = c.K_visc * dmax(-div_v, 0.);
		const double E_Lv = eL.Eint + ke_L, E_Rv = eR.Eint + ke_R;
		F[0] = F[0] + viscosity * (rho_L - rho_R);
		F[4] = F[4] + viscosity * (E_Lv - E_Rv);
		F[5] = F[5] + viscosity * (Eint_L - Eint_R);
		for (int n = 0; n < c.ns; ++n)
			F[6 + n] = F[6 + n] + viscosity * (L[6 + n] - R[6 + n]);
	}
	if (c.iso) { // hydro_system.hpp:1083-1087
		F[4] = 0;
		F[5] = 0;
	}
	// face-centred normal velocity (hydro_system.hpp:1089-1091)
	vface = (F[0] >= 0.) ? (F[0] / rho_R) : (F[0] / rho_L);
	// mass-scalar flux renormalisation (:1060-1074, 1093-1104)
	if (c.nms > 0) {
		double sumL = 0, sumR = 0;
		for (int n = 0; n < c.nms; ++n) {
			sumL += L[6 + n];
			sumR += R[6 + n];
		}
		if (F[0] >= 0.) {
			for (int n = 0; n < c.nms; ++n)
				F[6 + n] = F[0] * L[6 + n] / sumL;
		} else {
			for (int n = 0; n < c.nms; ++n)
				F[6 + n] = F[0] * R[6 + n] / sumR;
		}
	}
}

// Miller-Colella flattening coefficient of one cell along one direction (hydro_system.hpp:588-624).
// Pm2..Pp2 = pressure at i-2..i+2, KS = rho*cs^2 of the cell, vm1/vp1 = normal velocity at i-1/i+1.
__device__ __forceinline__ double flatten_chi(double Pm2, double Pm1, double Pp1, double Pp2, double KS, double vm1, double vp1)
{
	const double beta_max = 0.85, beta_min = 0.75, Zmax = 0.75, Zmin = 0.25;
	const double beta_denom = fabs(Pp2 - Pm2);
	const double dP1 = fabs(Pp1 - Pm1);
	const double beta = (beta_denom != 0) ? (dP1 / beta_denom) : 0;
	const double chi_min = dmax(0., dmin(1., (beta_max - beta) / (beta_max - beta_min)));
	const double Z = dP1 / KS;
	double chi = 1.0;
	if (vp1 < vm1) {
		chi = dmax(chi_min, dmin(1., (Zmax - Z) / (Zmax - Zmin)));
	}
	return chi;
}
