// qk_rad_source.cuh -- one cell of the matter-radiation coupling source terms, single photon group:
// RadSystem<problem_t>::AddSourceTermsSingleGroup, src/radiation/source_terms_single_group.hpp:9-565 (the lambda body :28-564),
// with the reference's compile-time hyper-parameters (src/radiation/radiation_system.hpp:34-52: include_work_term_in_source,
// enable_dE_constrain, !force_rad_floor_in_iteration, !add_line_cooling_to_radiation_in_jac, IMEX_a32 = 0.5), no dust model
// (ISM_Traits default :85-89), zero DefineNetCoolingRate / DefineCosmicRayHeatingRate (:522-545) and constant opacities.
//
// Everything here is QK_HD (host + device) and free of array accesses so that tests/host_src/rad_source_host.cpp can run the
// very same arithmetic on the CPU against the oracle (the product library only ever calls it from k_rad_source).
//
// Arithmetic: the reference's operation order, no FMA contraction (the library is built with --fmad=false; explicit fma()
// calls below are deliberate).  The reference takes T^4, T^3 and lorentz^3 from std::pow; here they are formed in
// double-double and rounded once (pow_dd<N>), i.e. correctly rounded except when the exact value lies within 2^-105 of a
// rounding boundary.  libm's and CUDA's pow are not correctly rounded (0.52 / 2 ulp), so the CPU reference, the reference's
// own CUDA build and this kernel may differ in the last bit of a thermal emission term; after the Newton-Raphson solve
// (residual tolerance 1e-11 of the total energy, :158) the results agree to that tolerance.
//
// RELAXED = true (qk_hydro_params::arith == QK_ARITH_FAST) is the same algorithm -- same iteration, same branches, guards and
// convergence tests -- with the gamma-law EOS in closed form (T = E_gas / (rho c_v'), c_v = rho c_v', c_v' = k_B / (mu m_u (gamma-1))
// instead of the Microphysics round trip: no division where the exact form has eight), quotients over call or cell constants
// as products with a reciprocal formed once, T^4 = (T^2)^2, and the identically-zero cooling terms dropped: one division per
// Newton-Raphson iteration instead of thirteen.  Each operation differs from the exact form by O(1e-16); the converged state
// agrees to the solver's own tolerance (1e-11 E_tot), inside the stated 1e-10 bar.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/quokka_b200.h"

#ifndef QK_HD
#ifdef __CUDACC__
#define QK_HD __host__ __device__ __forceinline__
#define QK_HDM __host__ __device__ __forceinline__ // for static member functions
#else
#define QK_HD static inline
#define QK_HDM inline
#endif
#endif

namespace qk_rsrc
{
// Microphysics constants/fundamental_constants.H:22,55
constexpr double K_B = 1.3806488e-16;
constexpr double M_U = 1.6605390666e-24;

QK_HD double mn(double a, double b) { return (b < a) ? b : a; } // std::min(a, b)
QK_HD double mx(double a, double b) { return (a < b) ? b : a; } // std::max(a, b)

struct Const {
	// RadSystem_Traits / problem opacities
	double c, chat, cscale /* c / chat */, inv_cscale /* 1 / cscale */, cc /* c * c */, c_chat /* c * chat */;
	double a_rad, floor_g /* Erad_floor / nGroups */;
	double kP, kE, kF, kPoE /* kappaP / kappaE or 1 */, two_kE_m_kF /* 2 kappaE - kappaF */;
	int beta_order;
	// EOS_Traits + eos_init limits
	double gamma, gm1, mu /* mean_molecular_weight / m_u */, mumn /* mu * m_u */, kB /* EOS_Traits::boltzmann_constant */;
	double mindens, mintemp;
	// per call
	double dt /* stage 2: (1 - IMEX_a32) dt_radiation */, chat_dt /* chat * dt */, gas_update_factor;
	// relaxed arithmetic (QK_ARITH_FAST): closed forms of the gamma-law EOS and reciprocals of the call constants
	double cvc /* c_v / rho = k_B / (mu m_u (gamma - 1)) */, inv_cvc, inv_arad, inv_kPoE, inv_cc, inv_cchat, Tmin /* mintemp * K_B / k_B */;
};

struct CellIn {
	double rho, mom[3], Egastot, Erad, F[3], src /* radEnergySource(i,j,k,0) */;
};
struct CellOut {
	double mom[3], Egastot, Eint, Erad, F[3];
	int solves, nr_iters, nr_max, fail_nr, fail_outer; // iteration_counter[0..2], iteration_failure_counter[0], [2]
};

// x^N (N = 3, 4) in double-double, rounded once
QK_HD void two_prod(double a, double b, double &p, double &e)
{
	p = a * b;
	e = fma(a, b, -p);
}
template <int N> QK_HD double pow_dd(double x)
{
	double h, l;
	two_prod(x, x, h, l); // x^2 = h + l exactly
	double p, e;
	if (N == 4) { // (h + l)^2 = h^2 + 2 h l + l^2
		two_prod(h, h, p, e);
		e = fma(2.0 * h, l, e);
		e = fma(l, l, e);
	} else { // (h + l) x
		two_prod(h, x, p, e);
		e = fma(l, x, e);
	}
	const double s = p + e;
	// non-finite or underflowing intermediates: fall back to the plain product (same value wherever it is finite and normal)
	return (s == s && fabs(s) <= 1.79769313486231570815e308 && fabs(p) >= 1e-280) ? s : ((N == 4) ? (x * x) * (x * x) : (x * x) * x);
}

// ---- division policies.  Every `a / b` of the reference is IEEE division; D::div(a, D::rcp(b)) returns exactly that
// quotient.  DivPlain is the literal form (host mirror, device fallback).  The device's DivShared (qk_rad_source.cu, built
// on qk_div.cuh) refines the reciprocal of a denominator once with the compiler's own instruction sequence and reuses it for
// every numerator over that denominator -- bit-identical, 3 FP64 instructions per quotient instead of 8 + MUFU.  Denominators
// that are constant over a call (k_B, mu m_u, gamma - 1, a_rad, c c_hat, ...) or over a cell (rho, 2 rho, tau) get their
// reciprocal once per thread, those of one Newton-Raphson iteration (c_v, det) once per iteration.
enum { C_KB = 0, C_kB, C_mumn, C_gm1, C_arad, C_kPoE, C_cc, C_cchat, C_N }; // the call-constant denominators
struct DivPlain {
	struct R {
		double b;
	};
	struct CTab {
		double b[C_N];
	};
	QK_HDM static double divc(double a, const CTab &t, int idx) { return a / t.b[idx]; }
	QK_HDM static R rcp(double b)
	{
		R r;
		r.b = b;
		return r;
	}
	QK_HDM static double div(double a, const R &r) { return a / r.b; }
};

template <class D> struct Recips { // per thread
	typename D::R rho, rho2; // cell constants: rho, 2 rho
	double r, rhoinv;	 // eos: clamped density and 1 / it
};
QK_HD void const_denoms(const Const &k, double b[C_N])
{
	b[C_KB] = K_B;
	b[C_kB] = k.kB;
	b[C_mumn] = k.mumn;
	b[C_gm1] = k.gm1;
	b[C_arad] = k.a_rad;
	b[C_kPoE] = k.kPoE;
	b[C_cc] = k.cc;
	b[C_cchat] = k.c_chat;
}

// ---- gamma-law EOS through Microphysics (interfaces/eos.H:395-435,141-205,69-79; EOS/gamma_law/actual_eos.H:47-294) as
// called by EOS<problem_t>::ComputeTgasFromEint / ComputeEintFromTgas / ComputeEintTempDerivative (src/hydro/EOS.hpp:75-157,200-240)
QK_HD double eos_clamp_rho(const Const &k, double rho) { return mn(1.e200, mx(k.mindens, rho)); }
QK_HD double eos_clamp_T(const Const &k, double T) { return mn(1.e200, mx(k.mintemp, T)); }
template <class D> QK_HD double tgas_from_eint(const Const &k, const typename D::CTab &ct, const Recips<D> &q, double Eint)
{
	const double e = D::div(Eint, q.rho);
	double T;
	if (e < 1.e-200 || e > 1.e200)
		T = k.mintemp; // eos_reset: T = clamp(0)
	else
		T = D::divc(e * k.mu * M_U * k.gm1, ct, C_KB); // actual_eos.H:128
	return D::divc(T * K_B, ct, C_kB);
}
// e(rho, T) of eos_input_rt and the clamped temperature it was evaluated at
template <class D> QK_HD double eos_e_of_rT(const Const &k, const typename D::CTab &ct, const Recips<D> &q, double Tgas, double &T)
{
	T = eos_clamp_T(k, Tgas);
	const double pressure = D::divc(q.r * T * K_B, ct, C_mumn); // actual_eos.H:199
	return D::divc(pressure, ct, C_gm1) * q.rhoinv;	       // :200
}
template <class D> QK_HD double eint_from_tgas(const Const &k, const typename D::CTab &ct, const Recips<D> &q, double rho, double Tgas)
{
	double T;
	const double e = eos_e_of_rT<D>(k, ct, q, Tgas, T);
	return D::divc(e * rho * k.kB, ct, C_KB);
}
template <class D> QK_HD double eint_temp_derivative(const Const &k, const typename D::CTab &ct, const Recips<D> &q, double rho, double Tgas)
{
	double T;
	const double e = eos_e_of_rT<D>(k, ct, q, Tgas, T);
	const double dedT = e * (1.0 / T); // actual_eos.H:194,227
	return D::divc(dedT * rho * k.kB, ct, C_KB);
}

// RadSystem::ComputeEgasFromEint - Eint part (radiation_system.hpp:1289-1309): the kinetic energy p^2 / (2 rho)
template <class D> QK_HD double ekin_of(const Recips<D> &q, double px, double py, double pz)
{
	const double p_sq = px * px + py * py + pz * pz;
	return D::div(p_sq, q.rho2);
}

// RadSystem::ComputeEddingtonFactor :773-790 + ComputeEddingtonTensor :873-916
template <class D> QK_HD void eddington_tensor(double fx, double fy, double fz, double T[3][3])
{
	const double f = sqrt(fx * fx + fy * fy + fz * fz);
	const double fvec[3] = {fx, fy, fz};
	double n[3];
	const bool fpos = (f > 0.);
	const typename D::R Rf = D::rcp(fpos ? f : 1.0);
	for (int ii = 0; ii < 3; ++ii)
		n[ii] = fpos ? D::div(fvec[ii], Rf) : 0.;
	const double fc = (f < 0.) ? 0. : (1. < f) ? 1. : f; // std::clamp(f, 0., 1.)
	const double f_fac = sqrt(4.0 - 3.0 * (fc * fc));
	const double chi = (3.0 + 4.0 * (fc * fc)) / (5.0 + 2.0 * f_fac);
	const double Tdiag = (1.0 - chi) / 2.0;
	const double Tf = (3.0 * chi - 1.0) / 2.0;
	for (int ii = 0; ii < 3; ++ii)
		for (int jj = 0; jj < 3; ++jj) {
			const double delta_ij = (ii == jj) ? 1 : 0;
			T[ii][jj] = Tdiag * delta_ij + Tf * (n[ii] * n[jj]);
		}
}

// RadSystem::Solve3x3matrix :560-580
QK_HD void solve3x3(double C00, double C01, double C02, double C10, double C11, double C12, double C20, double C21, double C22, double Y0, double Y1,
		    double Y2, double X[3])
{
	const double E11 = C11 - C01 * C10 / C00;
	const double E12 = C12 - C02 * C10 / C00;
	const double E21 = C21 - C01 * C20 / C00;
	const double E22 = C22 - C02 * C20 / C00;
	const double Z1 = Y1 - Y0 * C10 / C00;
	const double Z2 = Y2 - Y0 * C20 / C00;
	const double X2 = (Z2 - Z1 * E21 / E11) / (E22 - E12 * E21 / E11);
	const double X1 = (Z1 - E12 * X2) / E11;
	const double X0 = (Y0 - C01 * X1 - C02 * X2) / C00;
	X[0] = X0;
	X[1] = X1;
	X[2] = X2;
}

template <class D, bool RELAXED> QK_HD void source_cell(const Const &k, const typename D::CTab &ct, const CellIn &in, CellOut &out)
{
	typedef typename D::R Rc;
	const double c = k.c, chat = k.chat, cscale = k.cscale, dt = k.dt;
	const double rho = in.rho;
	const double Egastot0 = in.Egastot, Erad0 = in.Erad;
	const double Src = in.src * dt * chat; // :54
	const bool gas = (k.gamma != 1.0);
	const int beta_order = k.beta_order;
	const double nan = NAN;

	Recips<D> q;
	q.rho = D::rcp(rho);
	q.rho2 = D::rcp(2.0 * rho);
	q.r = eos_clamp_rho(k, rho);
	q.rhoinv = (q.r == rho) ? D::div(1.0, q.rho) : (1.0 / q.r);
	// relaxed-mode cell constants
	const double inv_rho = RELAXED ? (1.0 / rho) : 0.0;
	const double inv_2rho = 0.5 * inv_rho;
	const double cv_cell = rho * k.cvc;	   // c_v
	const double inv_cv = inv_rho * k.inv_cvc; // 1 / c_v

	double Egas0 = nan, Ekin0 = nan, Etot0 = nan, Egas_guess = nan;
	double lorentz_factor = nan, lorentz_factor_v = nan, lorentz_factor_v_v = nan;
	double fourPiBoverC = nan, Erad_guess = nan;
	const double kappaP = k.kP, kappaE = k.kE, kappaF = k.kF, kappaPoverE = k.kPoE;
	double work = 0.0, work_prev = 0.0;
	double dMomentum[3] = {0., 0., 0.}, Frad_t1[3] = {0., 0., 0.};
	out.solves = out.nr_iters = out.nr_max = out.fail_nr = out.fail_outer = 0;

	if (gas) { // :82-86
		Egas0 = Egastot0 - (RELAXED ? (in.mom[0] * in.mom[0] + in.mom[1] * in.mom[1] + in.mom[2] * in.mom[2]) * inv_2rho
				    : ekin_of<D>(q, in.mom[0], in.mom[1], in.mom[2]));
		Etot0 = Egas0 + cscale * (Erad0 + Src);
		Ekin0 = Egastot0 - Egas0;
		if (beta_order == 0 || beta_order == 1) { // :115-131
			lorentz_factor = 1.0;
			lorentz_factor_v = 1.0;
		} else {
			const double betaSqr = (in.mom[0] * in.mom[0] + in.mom[1] * in.mom[1] + in.mom[2] * in.mom[2]) / (rho * rho * c * c);
			if (beta_order == 2) {
				lorentz_factor = 1.0 + 0.5 * betaSqr;
				lorentz_factor_v = 1.0;
				lorentz_factor_v_v = 1.0;
			} else { // beta_order == 3 (static_assert(beta_order_ <= 3) :113)
				lorentz_factor = 1.0 + 0.5 * betaSqr;
				lorentz_factor_v = 1.0 + 0.5 * betaSqr;
				lorentz_factor_v_v = 1.0;
			}
		}
	}
	const double resid_limit = 1.0e-11 * Etot0; // resid_tol * Etot0 :158,254
	// tau = dt rho kappaP chat lorentz_factor (:210,217) and J11 (:296-300) do not change within a cell
	const double tau = dt * rho * kappaP * chat * lorentz_factor;
	const Rc Rtau = D::rcp(tau);
	const double inv_tau = (RELAXED && tau > 0.0) ? (1.0 / tau) : 0.0;
	const double J11 = (tau <= 0.0) ? -INFINITY : (RELAXED ? (-kappaPoverE * inv_tau - 1.0) : (D::div(-1.0 * kappaPoverE, Rtau) - 1.0));

	const int max_ite = 5;
	int ite = 0;
	for (; ite < max_ite; ++ite) {
		double R = nan;
		Erad_guess = Erad0;
		if (gas) {
			Egas_guess = Egas0;
			const int maxIter = 100;
			int n = 0;
			for (; n < maxIter; ++n) { // Newton-Raphson :161-352
				double T_gas;
				if (RELAXED) {
					const double e = Egas_guess * inv_rho;
					T_gas = (e < 1.e-200 || e > 1.e200) ? k.Tmin : Egas_guess * inv_cv;
				} else {
					T_gas = tgas_from_eint<D>(k, ct, q, Egas_guess);
				}
				const double T_d = T_gas;
				const double T2 = T_d * T_d;
				fourPiBoverC = k.a_rad * (RELAXED ? (T2 * T2) : pow_dd<4>(T_d)); // ComputeThermalRadiationSingleGroup :471-479
				if (fourPiBoverC < k.floor_g)
					fourPiBoverC = k.floor_g;
				if (n == 0) { // :192-215
					if (beta_order != 0 && ite == 0)
						work = (RELAXED ? (in.mom[0] * in.F[0] + in.mom[1] * in.F[1] + in.mom[2] * in.F[2]) * k.two_kE_m_kF * chat * k.inv_cc
								: D::divc((in.mom[0] * in.F[0] + in.mom[1] * in.F[1] + in.mom[2] * in.F[2]) * k.two_kE_m_kF * chat, ct, C_cc)) *
						       lorentz_factor_v * dt;
					R = (fourPiBoverC - (RELAXED ? Erad_guess * k.inv_kPoE : D::divc(Erad_guess, ct, C_kPoE))) * tau + work;
				} else if (tau > 0.0) { // :216-232
					Erad_guess = kappaPoverE * (fourPiBoverC - (RELAXED ? (R - work) * inv_tau : D::div(R - work, Rtau)));
				}
				// cooling = cooling_derivative = 0, CR_heating = 0 * dt: the terms are kept so that signed zeros and
				// non-finite dt propagate as in the reference (:234-245)
				const double cooling = 0.0, cooling_derivative = 0.0;
				const double CR_heating = 0.0 * dt;
				const double F_G = RELAXED ? (Egas_guess - Egas0 + cscale * R) : (Egas_guess - Egas0 + cscale * R + cooling * dt - CR_heating);
				const double F_D = Erad_guess - Erad0 - (R + Src);
				const double F_D_abs = (tau > 0.0) ? fabs(F_D) : fabs(F_D + R);
				if ((fabs(F_G) < resid_limit) && (cscale * F_D_abs < resid_limit))
					break;
				const double c_v = RELAXED ? cv_cell : eint_temp_derivative<D>(k, ct, q, rho, T_gas);
				const Rc Rcv = D::rcp(c_v);
				const double d_fourpiboverc_d_t = 4. * k.a_rad * (RELAXED ? (T2 * T_d) : pow_dd<3>(T_d)); // :499-503
				const double dEg_dT = kappaPoverE * d_fourpiboverc_d_t;
				// 0 * dt / c_v is +-0 for finite dt and finite non-zero c_v, so J00 is exactly 1; otherwise the quotient is formed
				const double zdt = cooling_derivative * dt;
				const double J00 = (RELAXED || (zdt == 0.0 && c_v != 0.0 && fabs(c_v) <= 1.79769313486231570815e308)) ? 1.0 : (1.0 + D::div(zdt, Rcv));
				const double J01 = cscale;
				const double J10 = RELAXED ? (inv_cv * dEg_dT) : (D::div(1.0, Rcv) * dEg_dT - k.inv_cscale * cooling_derivative * dt);
				const double y0 = -F_G;
				const double y1 = -1. * F_D;
				const double det = J00 * J11 - J01 * J10;
				const Rc Rdet = D::rcp(det);
				const double inv_det = RELAXED ? (1.0 / det) : 0.0;
				const double deltaEgas = RELAXED ? ((J11 * y0 - J01 * y1) * inv_det) : D::div(J11 * y0 - J01 * y1, Rdet);
				const double deltaR = RELAXED ? ((J00 * y1 - J10 * y0) * inv_det) : D::div(J00 * y1 - J10 * y0, Rdet);
				// enable_dE_constrain :330-342: deltaEgas / c_v > std::max(T_gas, T_rad), T_rad = (E_rad / a_rad)^(1/4).  The
				// comparison is false whenever the quotient does not exceed T_gas (std::max(T_gas, NaN) is T_gas), so T_rad
				// (a division and two square roots) is only formed for steps that jump by more than T_gas
				const double dT = RELAXED ? (deltaEgas * inv_cv) : D::div(deltaEgas, Rcv);
				double T_rad = 0.0;
				bool jump = (dT > T_gas);
				if (jump) {
					T_rad = sqrt(sqrt(RELAXED ? (Erad_guess * k.inv_arad) : D::divc(Erad_guess, ct, C_arad)));
					jump = (dT > mx(T_gas, T_rad));
				}
				if (jump) {
					Egas_guess = RELAXED ? (eos_clamp_T(k, T_rad) * cv_cell) : eint_from_tgas<D>(k, ct, q, rho, T_rad);
				} else {
					Egas_guess += deltaEgas;
					R += deltaR;
				}
			}
			if (n >= maxIter) // :354-362
				out.fail_nr += 1;
			out.solves += 1;
			out.nr_iters += n + 1;
			out.nr_max = (out.nr_max < n + 1) ? (n + 1) : out.nr_max;
			if (!RELAXED)
				Erad_guess += k.inv_cscale * (0.0 * dt); // cooling_tend :367-373
		}

		// 2. radiation flux update :396-490
		dMomentum[0] = dMomentum[1] = dMomentum[2] = 0.;
		if (gas && (beta_order != 0)) {
			const double erad = Erad_guess;
			double v_terms[3];
			const Rc Rce = D::rcp(c * erad);
			const double inv_ce = RELAXED ? (1.0 / (c * erad)) : 0.0;
			const double fx = RELAXED ? (in.F[0] * inv_ce) : D::div(in.F[0], Rce);
			const double fy = RELAXED ? (in.F[1] * inv_ce) : D::div(in.F[1], Rce);
			const double fz = RELAXED ? (in.F[2] * inv_ce) : D::div(in.F[2], Rce);
			const double F_coeff = chat * rho * kappaF * dt * lorentz_factor;
			double Tedd[3][3];
			eddington_tensor<D>(fx, fy, fz, Tedd);
			const double lfv3 = (kappaF != kappaE) ? pow_dd<3>(lorentz_factor_v) : 0.;
			for (int n = 0; n < 3; ++n) {
				double Planck_term = kappaP * fourPiBoverC * lorentz_factor_v;
				if (kappaF != kappaE)
					Planck_term += (kappaF - kappaE) * erad * lfv3;
				Planck_term *= k.chat_dt * in.mom[n];
				double pressure_term = 0.0;
				for (int z = 0; z < 3; ++z)
					pressure_term += in.mom[z] * Tedd[n][z] * erad;
				pressure_term *= k.chat_dt * kappaF * lorentz_factor_v;
				v_terms[n] = Planck_term + pressure_term;
			}
			if (beta_order == 1 || kappaF == kappaE) {
				const Rc Rfc = D::rcp(1.0 + F_coeff);
				const double inv_fc = RELAXED ? (1.0 / (1.0 + F_coeff)) : 0.0;
				for (int n = 0; n < 3; ++n) {
					Frad_t1[n] = RELAXED ? ((in.F[n] + v_terms[n]) * inv_fc) : D::div(in.F[n] + v_terms[n], Rfc);
					dMomentum[n] += RELAXED ? (-(Frad_t1[n] - in.F[n]) * k.inv_cchat) : D::divc(-(Frad_t1[n] - in.F[n]), ct, C_cchat);
				}
			} else {
				// gasVel is declared and never assigned in the reference (:407), so the K0 v_i v_j terms are K0 * 0 * 0; they
				// are formed all the same (a non-finite K0 must poison the result as it does there)
				const double gasVel[3] = {0., 0., 0.};
				const double K0 = 2.0 * rho * chat * dt * (kappaF - kappaE) / c / c * pow_dd<3>(lorentz_factor_v_v);
				const double A00 = 1.0 + F_coeff + K0 * gasVel[0] * gasVel[0];
				const double A01 = K0 * gasVel[0] * gasVel[1];
				const double A02 = K0 * gasVel[0] * gasVel[2];
				const double A10 = K0 * gasVel[1] * gasVel[0];
				const double A11 = 1.0 + F_coeff + K0 * gasVel[1] * gasVel[1];
				const double A12 = K0 * gasVel[1] * gasVel[2];
				const double A20 = K0 * gasVel[2] * gasVel[0];
				const double A21 = K0 * gasVel[2] * gasVel[1];
				const double A22 = 1.0 + F_coeff + K0 * gasVel[2] * gasVel[2];
				solve3x3(A00, A01, A02, A10, A11, A12, A20, A21, A22, v_terms[0] + in.F[0], v_terms[1] + in.F[1], v_terms[2] + in.F[2],
					 Frad_t1);
				for (int n = 0; n < 3; ++n)
					dMomentum[n] += D::divc(-(Frad_t1[n] - in.F[n]), ct, C_cchat);
			}
		} else { // :484-490
			const Rc Rfc = D::rcp(1.0 + rho * kappaF * chat * dt);
			const double inv_fc = RELAXED ? (1.0 / (1.0 + rho * kappaF * chat * dt)) : 0.0;
			for (int n = 0; n < 3; ++n) {
				Frad_t1[n] = RELAXED ? (in.F[n] * inv_fc) : D::div(in.F[n], Rfc);
				dMomentum[n] += RELAXED ? (-(Frad_t1[n] - in.F[n]) * k.inv_cchat) : D::divc(-(Frad_t1[n] - in.F[n]), ct, C_cchat);
			}
		}
		const double x1GasMom1 = in.mom[0] + dMomentum[0];
		const double x2GasMom1 = in.mom[1] + dMomentum[1];
		const double x3GasMom1 = in.mom[2] + dMomentum[2];

		// 3. work term :496-541
		if (!gas || beta_order == 0)
			break;
		{
			const double Egastot1 = Egas_guess + (RELAXED ? (x1GasMom1 * x1GasMom1 + x2GasMom1 * x2GasMom1 + x3GasMom1 * x3GasMom1) * inv_2rho
								      : ekin_of<D>(q, x1GasMom1, x2GasMom1, x3GasMom1));
			const double Ekin1 = Egastot1 - Egas_guess;
			const double dEkin_work = Ekin1 - Ekin0;
			Egas_guess -= dEkin_work;
		}
		work_prev = work;
		work = (RELAXED ? (x1GasMom1 * Frad_t1[0] + x2GasMom1 * Frad_t1[1] + x3GasMom1 * Frad_t1[2]) * chat * k.inv_cc
				: D::divc((x1GasMom1 * Frad_t1[0] + x2GasMom1 * Frad_t1[1] + x3GasMom1 * Frad_t1[2]) * chat, ct, C_cc)) *
		       lorentz_factor_v * k.two_kE_m_kF * dt;
		const double lag_tol = 1.0e-13;
		const double dwork = fabs(work - work_prev);
		if ((fabs(work) == 0.0) || (cscale * dwork < lag_tol * Etot0) || (dwork <= lag_tol * R) || (dwork <= 1.0e-8 * fabs(work)))
			break;
	}
	if (ite >= max_ite) // :544-547
		out.fail_outer = 1;

	// 4b. store :549-564
	for (int n = 0; n < 3; ++n) {
		out.mom[n] = in.mom[n] + dMomentum[n] * k.gas_update_factor;
		out.F[n] = Frad_t1[n];
	}
	if (gas) {
		Egas_guess = Egas0 + (Egas_guess - Egas0) * k.gas_update_factor;
		out.Eint = Egas_guess;
		out.Egastot = Egas_guess + (RELAXED ? (out.mom[0] * out.mom[0] + out.mom[1] * out.mom[1] + out.mom[2] * out.mom[2]) * inv_2rho
						    : ekin_of<D>(q, out.mom[0], out.mom[1], out.mom[2]));
		out.Erad = Erad_guess;
	} else {
		out.Eint = nan; // not written by the reference (:558-571); the caller skips these three
		out.Egastot = nan;
		out.Erad = nan;
	}
}

// host: the per-call constants from the three parameter blocks (:13-19,88-91 for dt and gas_update_factor)
static inline Const make_const(const qk_hydro_params *hp, const qk_rad_params *rp, const qk_rad_source_params *sp, double dt_radiation, int stage)
{
	const double IMEX_a32 = 0.5; // radiation_system.hpp:52
	Const k;
	k.c = rp->c_light;
	k.chat = rp->c_hat;
	k.cscale = k.c / k.chat;
	k.inv_cscale = 1 / k.cscale;
	k.cc = k.c * k.c;
	k.c_chat = k.c * k.chat;
	k.a_rad = sp->radiation_constant;
	k.floor_g = rp->Erad_floor / rp->ngroups;
	k.kP = sp->kappa_P;
	k.kE = sp->kappa_E;
	k.kF = sp->kappa_F;
	k.kPoE = (k.kE > 0.0) ? (k.kP / k.kE) : 1.0; // :183-187
	k.two_kE_m_kF = 2.0 * k.kE - k.kF;
	k.beta_order = sp->beta_order;
	k.gamma = hp->gamma;
	k.gm1 = hp->gamma - 1.0;
	k.mu = hp->mean_molecular_weight / M_U; // src/hydro/EOS.hpp:104
	k.mumn = k.mu * M_U;
	k.kB = hp->boltzmann_constant;
	k.mindens = (1.e-200 < hp->small_dens) ? hp->small_dens : 1.e-200; // eos_init, interfaces/eos.H:40-44
	k.mintemp = (1.e-200 < hp->small_temp) ? hp->small_temp : 1.e-200;
	k.dt = (stage == 2) ? (1.0 - IMEX_a32) * dt_radiation : dt_radiation;
	k.chat_dt = k.chat * k.dt;
	k.gas_update_factor = (stage == 1) ? IMEX_a32 : 1.0;
	k.cvc = k.kB / (k.mumn * k.gm1);
	k.inv_cvc = 1.0 / k.cvc;
	k.inv_arad = 1.0 / k.a_rad;
	k.inv_kPoE = 1.0 / k.kPoE;
	k.inv_cc = 1.0 / k.cc;
	k.inv_cchat = 1.0 / k.c_chat;
	k.Tmin = k.mintemp * K_B / k.kB;
	return k;
}
} // namespace qk_rsrc
