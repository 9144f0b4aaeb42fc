/* oracle/quokka_oracle.c -- CPU restatement (plain C) of Quokka's hydro hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see quokka_oracle.h).  Parity status: PINNED against the reference's own
 * code compiled here (oracle/_ref), see tests/test_oracle_vs_ref.py and tests/golden/.
 *
 * Each function cites the reference file:line it follows.  Paths are relative to /root/reference.
 * Compile with -O2 -ffp-contract=off (no FMA contraction; matches the reference's --fmad=false).
 */
#include "quokka_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A4(a, i, j, k, n)                                                                                                                            \
	((a)->p[((int64_t)(i) - (a)->begin[0]) + ((int64_t)(j) - (a)->begin[1]) * (a)->jstride + ((int64_t)(k) - (a)->begin[2]) * (a)->kstride +      \
		(int64_t)(n) * (a)->nstride])

/* std::min / std::max semantics (first argument returned on ties / unordered) */
static inline double dmin(double a, double b) { return (b < a) ? b : a; }
static inline double dmax(double a, double b) { return (a < b) ? b : a; }
/* src/math/math_impl.hpp:17,20 */
static inline double clampd(double v, double lo, double hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
static inline int sgnd(double v) { return (0.0 < v) - (v < 0.0); }

/* extern/Microphysics/constants/fundamental_constants.H:22,55 */
static const double C_k_B = 1.3806488e-16;
static const double C_m_u = 1.6605390666e-24;

enum { RHO = 0, MX = 1, MY = 2, MZ = 3, EN = 4, EI = 5, SC0 = 6 }; /* src/hydro/hydro_system.hpp:54-72 */
#define MAXV (6 + QK_MAX_SCALARS)

/* ------------------------------------------------------------------------------------------------
 * gamma-law EOS through Microphysics: extern/Microphysics/interfaces/eos.H:395-435 (eos),
 * :141-205 (reset_inputs), :69-79 (eos_reset); EOS/gamma_law/actual_eos.H:47-294; chem_eos_t
 * interfaces/eos_type.H:144-164.  Limits from eos_init (eos.H:16-49) with small_temp/small_dens
 * from src/QuokkaSimulation.hpp:165-167.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
	double rho, T, p, e, dpdT, dpdr, dpde, dpdr_e, dedT, dedr, G, mu, cs, cv, cp, gam1;
} eos_state;
enum { EOS_RT, EOS_RP, EOS_RE };

static void actual_eos(const qk_hydro_params *prm, int input, eos_state *s)
{
	const double gam = prm->gamma;
	const double m_nucleon = C_m_u;
	switch (input) {
	case EOS_RT:
		break;
	case EOS_RP: /* actual_eos.H:115 */
		s->T = s->p * s->mu * m_nucleon / (C_k_B * s->rho);
		break;
	case EOS_RE: /* actual_eos.H:128 */
		s->T = s->e * s->mu * m_nucleon * (gam - 1.0) / C_k_B;
		break;
	}
	/* actual_eos.H:194-206 */
	const double Tinv = 1.0 / s->T;
	const double rhoinv = 1.0 / s->rho;
	const double pressure = s->rho * s->T * C_k_B / (s->mu * m_nucleon);
	const double energy = pressure / (gam - 1.0) * rhoinv;
	s->p = pressure;
	s->e = energy;
	/* :225-231 */
	s->dpdT = s->p * Tinv;
	s->dpdr = s->p * rhoinv;
	s->dedT = s->e * Tinv;
	s->dedr = 0.0;
	/* :252-269 */
	s->cv = s->dedT;
	s->cp = gam * s->cv;
	s->gam1 = gam;
	s->dpdr_e = s->dpdr - s->dpdT * s->dedr * (1.0 / s->dedT);
	s->dpde = s->dpdT * (1.0 / s->dedT);
	s->cs = sqrt(gam * s->p * rhoinv);
	s->G = 0.5 * (1.0 + gam);
}

static void eos_call(const qk_hydro_params *prm, int input, eos_state *s)
{
	/* eos_init: mintemp = max(1e-200, small_temp), mindens = max(1e-200, small_dens) */
	const double mintemp = dmax(1.e-200, prm->small_temp), maxtemp = 1.e200;
	const double mindens = dmax(1.e-200, prm->small_dens), maxdens = 1.e200;
	const double mine = 1.e-200, maxe = 1.e200, minp = 1.e-200, maxp = 1.e200;
	int has_been_reset = 0;
	/* reset_inputs, eos.H:141-205 */
	if (input == EOS_RT) {
		s->rho = dmin(maxdens, dmax(mindens, s->rho));
		s->T = dmin(maxtemp, dmax(mintemp, s->T));
	} else {
		s->rho = dmin(maxdens, dmax(mindens, s->rho));
		int bad = (input == EOS_RE) ? (s->e < mine || s->e > maxe) : (s->p < minp || s->p > maxp);
		if (bad) { /* eos_reset, eos.H:69-79 */
			s->T = dmin(maxtemp, dmax(mintemp, s->T));
			s->rho = dmin(maxdens, dmax(mindens, s->rho));
			actual_eos(prm, EOS_RT, s);
			has_been_reset = 1;
		}
	}
	if (!has_been_reset) {
		actual_eos(prm, input, s);
	}
}

static inline eos_state eos_new(const qk_hydro_params *prm)
{
	eos_state s;
	memset(&s, 0, sizeof(s));
	s.mu = prm->mean_molecular_weight / C_m_u; /* src/hydro/EOS.hpp:104,336 */
	return s;
}

/* src/hydro/EOS.hpp:299-340 */
static double eos_pressure(const qk_hydro_params *prm, double rho, double Eint)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.e = (rho == 0.0) ? 0.0 : Eint / rho;
	eos_call(prm, EOS_RE, &s);
	return s.p;
}
/* EOS.hpp:342-383 */
static double eos_sound_speed(const qk_hydro_params *prm, double rho, double P)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.p = P;
	eos_call(prm, EOS_RP, &s);
	return s.cs;
}
/* EOS.hpp:159-198 */
static double eos_eint_from_pres(const qk_hydro_params *prm, double rho, double P)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.p = P;
	eos_call(prm, EOS_RP, &s);
	return s.e * rho;
}
/* EOS.hpp:75-114 */
static double eos_tgas_from_eint(const qk_hydro_params *prm, double rho, double Eint)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.e = Eint / rho;
	eos_call(prm, EOS_RE, &s);
	return s.T * C_k_B / prm->boltzmann_constant;
}
/* EOS.hpp:116-157 */
static double eos_eint_from_tgas(const qk_hydro_params *prm, double rho, double Tgas)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.T = Tgas;
	eos_call(prm, EOS_RT, &s);
	return s.e * rho * prm->boltzmann_constant / C_k_B;
}
/* EOS.hpp:242-297 */
static void eos_other_derivatives(const qk_hydro_params *prm, double rho, double P, double *dedr, double *dedp, double *drdp, double *dpdr_s,
				  double *G)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.p = P;
	eos_call(prm, EOS_RP, &s);
	*dedr = s.dedr;
	*dedp = 1.0 / s.dpde;
	*drdp = 1.0 / (s.dpdr * C_k_B / prm->boltzmann_constant);
	*dpdr_s = s.cs * s.cs;
	*G = s.G;
}

double orc_eos_pressure(const qk_hydro_params *prm, double rho, double Eint) { return eos_pressure(prm, rho, Eint); }
double orc_eos_sound_speed(const qk_hydro_params *prm, double rho, double P) { return eos_sound_speed(prm, rho, P); }
double orc_eos_eint_from_pres(const qk_hydro_params *prm, double rho, double P) { return eos_eint_from_pres(prm, rho, P); }
double orc_eos_tgas_from_eint(const qk_hydro_params *prm, double rho, double Eint) { return eos_tgas_from_eint(prm, rho, Eint); }
double orc_eos_eint_from_tgas(const qk_hydro_params *prm, double rho, double T) { return eos_eint_from_tgas(prm, rho, T); }

/* HydroSystem::ComputePressure(cons,i,j,k)  src/hydro/hydro_system.hpp:349-372 */
static double cons_pressure(const qk_hydro_params *prm, const qk_array4 *c, int i, int j, int k)
{
	const double rho = A4(c, i, j, k, RHO), px = A4(c, i, j, k, MX), py = A4(c, i, j, k, MY), pz = A4(c, i, j, k, MZ);
	const double E = A4(c, i, j, k, EN);
	const double vx = px / rho, vy = py / rho, vz = pz / rho;
	const double ke = 0.5 * rho * (vx * vx + vy * vy + vz * vz);
	const double thermal = E - ke;
	if (prm->gamma == 1.0) /* is_eos_isothermal(), hydro_system.hpp:365-366 */
		return rho * prm->cs_isothermal * prm->cs_isothermal;
	return eos_pressure(prm, rho, thermal);
}
/* HydroSystem::ComputeSoundSpeed(cons,i,j,k)  hydro_system.hpp:374-394 */
static double cons_sound_speed(const qk_hydro_params *prm, const qk_array4 *c, int i, int j, int k)
{
	const double rho = A4(c, i, j, k, RHO);
	const double P = cons_pressure(prm, c, i, j, k);
	return eos_sound_speed(prm, rho, P);
}

/* HydroSystem::ConservedToPrimitive  hydro_system.hpp:138-196 */
void orc_conserved_to_primitive(const qk_hydro_params *prm, const qk_array4 *cons, const qk_array4 *prim, const qk_box *bx)
{
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(cons, i, j, k, RHO), px = A4(cons, i, j, k, MX), py = A4(cons, i, j, k, MY);
				const double pz = A4(cons, i, j, k, MZ), E = A4(cons, i, j, k, EN), Eint_aux = A4(cons, i, j, k, EI);
				const double vx = px / rho, vy = py / rho, vz = pz / rho;
				const double ke = 0.5 * rho * (vx * vx + vy * vy + vz * vz);
				const double Eint_cons = E - ke;
				const double Pgas = cons_pressure(prm, cons, i, j, k);
				const double eint_cons = Eint_cons / rho;
				const double eint_aux = Eint_aux / rho;
				A4(prim, i, j, k, 0) = rho;
				A4(prim, i, j, k, 1) = vx;
				A4(prim, i, j, k, 2) = vy;
				A4(prim, i, j, k, 3) = vz;
				if (prm->reconstruct_eint) {
					A4(prim, i, j, k, 4) = eint_cons;
					A4(prim, i, j, k, 5) = eint_aux;
				} else {
					A4(prim, i, j, k, 4) = Pgas;
					A4(prim, i, j, k, 5) = Eint_aux;
				}
				for (int n = 0; n < prm->nscalars; ++n)
					A4(prim, i, j, k, SC0 + n) = A4(cons, i, j, k, SC0 + n);
			}
}

/* HydroSystem::ComputeFlatteningCoefficients<DIR>  hydro_system.hpp:531-626 */
void orc_flattening_coefficients(const qk_hydro_params *prm, int dir, const qk_array4 *q, const qk_array4 *chi_out, const qk_box *bx)
{
	const double beta_max = 0.85, beta_min = 0.75, Zmax = 0.75, Zmin = 0.25;
	const int e[3] = {dir == 0, dir == 1, dir == 2};
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				double P[5]; /* P(i-2..i+2) along dir */
				for (int s = -2; s <= 2; ++s) {
					const int ii = i + s * e[0], jj = j + s * e[1], kk = k + s * e[2];
					double v = A4(q, ii, jj, kk, 4);
					if (prm->reconstruct_eint) {
						const double r = A4(q, ii, jj, kk, 0);
						v = eos_pressure(prm, r, r * v);
					}
					if (prm->gamma == 1.0) /* :579-586 */
						v = A4(q, ii, jj, kk, 0) * (prm->cs_isothermal * prm->cs_isothermal);
					P[s + 2] = v;
				}
				const double Pplus2 = P[4], Pplus1 = P[3], Pc = P[2], Pminus1 = P[1], Pminus2 = P[0];
				const double beta_denom = fabs(Pplus2 - Pminus2);
				const double beta = (beta_denom != 0) ? (fabs(Pplus1 - Pminus1) / beta_denom) : 0;
				const double chi_min = dmax(0., dmin(1., (beta_max - beta) / (beta_max - beta_min)));
				const double rho = A4(q, i, j, k, 0);
				double K_S;
				if (prm->gamma == 1.0) { /* :604-606 */
					K_S = rho * prm->cs_isothermal * prm->cs_isothermal;
				} else {
					const double cs = eos_sound_speed(prm, rho, Pc);
					K_S = (cs * cs) * rho; /* std::pow(cs,2)*rho */
				}
				const double Z = fabs(Pplus1 - Pminus1) / K_S;
				const int vn = 1 + dir;
				double chi = 1.0;
				if (A4(q, i + e[0], j + e[1], k + e[2], vn) < A4(q, i - e[0], j - e[1], k - e[2], vn)) {
					chi = dmax(chi_min, dmin(1., (Zmax - Z) / (Zmax - Zmin)));
				}
				A4(chi_out, i, j, k, 0) = chi;
			}
}

/* HyperbolicSystem::MC / minmod  src/hyperbolic_system.hpp:58-66 */
static inline double lim_MC(double a, double b) { return 0.5 * (sgnd(a) + sgnd(b)) * dmin(0.5 * fabs(a + b), dmin(2.0 * fabs(a), 2.0 * fabs(b))); }
static inline double lim_minmod(double a, double b) { return 0.5 * (sgnd(a) + sgnd(b)) * dmin(fabs(a), fabs(b)); }

/* HyperbolicSystem::ReconstructStates{Constant :164-181, PLM :218-247, PPM :337-433} */
void orc_reconstruct_states(int order, int limiter, int dir, const qk_array4 *q, const qk_array4 *left, const qk_array4 *right, const qk_box *bx,
			    int nvars)
{
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	for (int n = 0; n < nvars; ++n)
		for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
			for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
				for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
#define Q(s) A4(q, i + (s) * e0, j + (s) * e1, k + (s) * e2, n)
					if (order == 1) {
						A4(left, i, j, k, n) = Q(-1);
						A4(right, i, j, k, n) = Q(0);
					} else if (order == 2) {
						double lslope, rslope;
						if (limiter == QK_MC) {
							lslope = lim_MC(Q(0) - Q(-1), Q(-1) - Q(-2));
							rslope = lim_MC(Q(1) - Q(0), Q(0) - Q(-1));
						} else {
							lslope = lim_minmod(Q(0) - Q(-1), Q(-1) - Q(-2));
							rslope = lim_minmod(Q(1) - Q(0), Q(0) - Q(-1));
						}
						A4(left, i, j, k, n) = Q(-1) + 0.25 * lslope;
						A4(right, i, j, k, n) = Q(0) - 0.25 * rslope;
					} else {
						/* bounds = std::minmax({q(i), q(i-1), q(i+1)}) :365 */
						double lo = Q(0), hi = Q(0);
						if (Q(-1) < lo) lo = Q(-1);
						if (Q(1) < lo) lo = Q(1);
						if (!(Q(-1) < hi)) hi = Q(-1);
						if (!(Q(1) < hi)) hi = Q(1);
						const double coef_1 = (7. / 12.);
						const double coef_2 = (-1. / 12.);
						const double a_minus = (coef_1 * Q(0) + coef_2 * Q(1)) + (coef_1 * Q(-1) + coef_2 * Q(-2));
						const double a_plus = (coef_1 * Q(1) + coef_2 * Q(2)) + (coef_1 * Q(0) + coef_2 * Q(-1));
						double new_a_minus = clampd(a_minus, lo, hi);
						double new_a_plus = clampd(a_plus, lo, hi);
						const double a = Q(0);
						const double dq_minus = (a - new_a_minus);
						const double dq_plus = (new_a_plus - a);
						const double qa = dq_plus * dq_minus;
						if (qa <= 0.0) {
							const double dq0 = lim_MC(Q(1) - Q(0), Q(0) - Q(-1));
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
						A4(right, i, j, k, n) = new_a_minus;
						A4(left, i + e0, j + e1, k + e2, n) = new_a_plus;
					}
#undef Q
				}
}

/* HydroSystem::FlattenShocks<DIR>  hydro_system.hpp:628-694 */
void orc_flatten_shocks(int dir, const qk_array4 *q, const qk_array4 *c1, const qk_array4 *c2, const qk_array4 *c3, const qk_array4 *left,
			const qk_array4 *right, const qk_box *bx, int nvars)
{
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	for (int n = 0; n < nvars; ++n)
		for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
			for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
				for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
					double chi = A4(c1, i - 1, j, k, 0);
					chi = dmin(chi, A4(c1, i, j, k, 0));
					chi = dmin(chi, A4(c1, i + 1, j, k, 0));
					chi = dmin(chi, A4(c2, i, j - 1, k, 0));
					chi = dmin(chi, A4(c2, i, j, k, 0));
					chi = dmin(chi, A4(c2, i, j + 1, k, 0));
					chi = dmin(chi, A4(c3, i, j, k - 1, 0));
					chi = dmin(chi, A4(c3, i, j, k, 0));
					chi = dmin(chi, A4(c3, i, j, k + 1, 0));
					const double a_minus = A4(right, i, j, k, n);
					const double a_plus = A4(left, i + e0, j + e1, k + e2, n);
					const double a_mean = A4(q, i, j, k, n);
					A4(right, i, j, k, n) = chi * a_minus + (1. - chi) * a_mean;
					A4(left, i + e0, j + e1, k + e2, n) = chi * a_plus + (1. - chi) * a_mean;
				}
}

/* quokka::HydroState  src/hydro/HydroState.hpp:10-23 */
typedef struct {
	double rho, u, v, w, P, cs, E, Eint;
	double scalar[QK_MAX_SCALARS];
} hstate;

/* quokka::Riemann::HLLC  src/hydro/HLLC.hpp:21-153 */
static void riemann_hllc(const qk_hydro_params *prm, const hstate *sL, const hstate *sR, double du, double dw, int ns, double *F)
{
	const int nv = 6 + ns;
	const double wl = sqrt(sL->rho);
	const double wr = sqrt(sR->rho);
	const double norm = 1. / (wl + wr);
	const double u_tilde = (wl * sL->u + wr * sR->u) * norm;
	const double v_tilde = (wl * sL->v + wr * sR->v) * norm;
	const double w_tilde = (wl * sL->w + wr * sR->w) * norm;
	const double vsq_tilde = u_tilde * u_tilde + v_tilde * v_tilde + w_tilde * w_tilde;
	const double H_L = (sL->E + sL->P) / sL->rho;
	const double H_R = (sR->E + sR->P) / sR->rho;
	const double H_tilde = (wl * H_L + wr * H_R) * norm;
	double cs_tilde;
	const double dU = sL->u - sR->u;
	double S_L, S_R;
	if (prm->gamma != 1.0) {
		double dedr_L, dedp_L, drdp_L, dpdr_s_L, G_L, dedr_R, dedp_R, drdp_R, dpdr_s_R, G_R;
		eos_other_derivatives(prm, sL->rho, sL->P, &dedr_L, &dedp_L, &drdp_L, &dpdr_s_L, &G_L);
		eos_other_derivatives(prm, sR->rho, sR->P, &dedr_R, &dedp_R, &drdp_R, &dpdr_s_R, &G_R);
		const double C_tilde_rho = 0.5 * ((sL->Eint / sL->rho) + (sR->Eint / sR->rho) + sL->rho * dedr_L + sR->rho * dedr_R);
		const double C_tilde_P = 0.5 * ((sL->Eint / sL->rho) * drdp_L + (sR->Eint / sR->rho) * drdp_R + sL->rho * dedp_L + sR->rho * dedp_R);
		const double cs_exp = H_tilde - 0.5 * vsq_tilde - C_tilde_rho;
		if (cs_exp <= 0) {
			cs_tilde = 0.5 * (sL->cs + sR->cs);
		} else {
			cs_tilde = sqrt(cs_exp / C_tilde_P);
		}
		const double s_NL = 0.5 * G_L * dmax(dU, 0.);
		const double s_NR = 0.5 * G_R * dmax(dU, 0.);
		S_L = dmin(sL->u - (sL->cs + s_NL), u_tilde - (cs_tilde + s_NL));
		S_R = dmax(sR->u + (sR->cs + s_NR), u_tilde + (cs_tilde + s_NR));
	} else {
		cs_tilde = 0.5 * (sL->cs + sR->cs);
		const double G_L = 0.5 * (1.0 + 1.), G_R = 0.5 * (1.0 + 1.);
		const double s_NL = 0.5 * G_L * dmax(dU, 0.);
		const double s_NR = 0.5 * G_R * dmax(dU, 0.);
		S_L = dmin(sL->u - (sL->cs + s_NL), u_tilde - (cs_tilde + s_NL));
		S_R = dmax(sR->u + (sR->cs + s_NR), u_tilde + (cs_tilde + s_NR));
	}
	const double cs_max = dmax(sL->cs, sR->cs);
	const double tp = dmin(1., (cs_max - dmin(du, 0.)) / (cs_max - dmin(dw, 0.)));
	const double theta = tp * tp * tp * tp;
	const double S_star = (theta * (sR->P - sL->P) + (sL->rho * sL->u * (S_L - sL->u) - sR->rho * sR->u * (S_R - sR->u))) /
			      (sL->rho * (S_L - sL->u) - sR->rho * (S_R - sR->u));
	const double vmag_L = sqrt(sL->u * sL->u + sL->v * sL->v + sL->w * sL->w);
	const double vmag_R = sqrt(sR->u * sR->u + sR->v * sR->v + sR->w * sR->w);
	const double chi = dmin(1., dmax(vmag_L, vmag_R) / cs_max);
	const double phi = chi * (2. - chi);
	const double P_LR = 0.5 * (sL->P + sR->P) + 0.5 * phi * (sL->rho * (S_L - sL->u) * (S_star - sL->u) + sR->rho * (S_R - sR->u) * (S_star - sR->u));

	double D_L[MAXV] = {0., 1., 0., 0., sL->u, 0.};
	double D_R[MAXV] = {0., 1., 0., 0., sR->u, 0.};
	double D_star[MAXV] = {0., 1., 0., 0., S_star, 0.};
	double U_L[MAXV] = {sL->rho, sL->rho * sL->u, sL->rho * sL->v, sL->rho * sL->w, sL->E, sL->Eint};
	double U_R[MAXV] = {sR->rho, sR->rho * sR->u, sR->rho * sR->v, sR->rho * sR->w, sR->E, sR->Eint};
	for (int n = 0; n < ns; ++n) {
		U_L[6 + n] = sL->scalar[n];
		U_R[6 + n] = sR->scalar[n];
	}
	for (int n = 0; n < nv; ++n) {
		const double F_L = sL->u * U_L[n] + sL->P * D_L[n];
		const double F_R = sR->u * U_R[n] + sR->P * D_R[n];
		const double F_starL = (S_star * (S_L * U_L[n] - F_L) + S_L * P_LR * D_star[n]) / (S_L - S_star);
		const double F_starR = (S_star * (S_R * U_R[n] - F_R) + S_R * P_LR * D_star[n]) / (S_R - S_star);
		if (S_L > 0.0) {
			F[n] = F_L;
		} else if ((S_star > 0.0) && (S_L <= 0.0)) {
			F[n] = F_starL;
		} else if ((S_star <= 0.0) && (S_R >= 0.0)) {
			F[n] = F_starR;
		} else {
			F[n] = F_R;
		}
	}
}

/* quokka::Riemann::LLF  src/hydro/LLF.hpp:15-43 */
static void riemann_llf(const hstate *sL, const hstate *sR, int ns, double *F)
{
	const int nv = 6 + ns;
	const double Sp = dmax(fabs(sL->u) + sL->cs, fabs(sR->u) + sR->cs);
	double U_L[MAXV] = {sL->rho, sL->rho * sL->u, sL->rho * sL->v, sL->rho * sL->w, sL->E, sL->Eint};
	double U_R[MAXV] = {sR->rho, sR->rho * sR->u, sR->rho * sR->v, sR->rho * sR->w, sR->E, sR->Eint};
	for (int n = 0; n < ns; ++n) {
		U_L[6 + n] = sL->scalar[n];
		U_R[6 + n] = sR->scalar[n];
	}
	double D_L[MAXV] = {0., 1., 0., 0., sL->u, 0.};
	double D_R[MAXV] = {0., 1., 0., 0., sR->u, 0.};
	for (int n = 0; n < nv; ++n) {
		const double F_L = sL->u * U_L[n] + sL->P * D_L[n];
		const double F_R = sR->u * U_R[n] + sR->P * D_R[n];
		F[n] = 0.5 * (F_L + F_R) - 0.5 * Sp * (U_R[n] - U_L[n]);
	}
}

/* HydroSystem::ComputeFluxes<RIEMANN,DIR>  hydro_system.hpp:852-1112 */
void orc_compute_fluxes(const qk_hydro_params *prm, int solver, int dir, const qk_array4 *flux, const qk_array4 *facevel, const qk_array4 *L,
			const qk_array4 *R, const qk_array4 *q, const qk_box *fbx)
{
	const int ns = prm->nscalars, nms = prm->nmscalars, nv = 6 + ns;
	const int aN = dir, aV = (dir + 1) % 3, aW = (dir + 2) % 3; /* array axes of the permuted (i,j,k), ArrayView_3d.hpp:18-113 */
	const int velN = 1 + aN, velV = 1 + aV, velW = 1 + aW;	    /* hydro_system.hpp:954-976 */
	int eN[3] = {0, 0, 0}, eV[3] = {0, 0, 0}, eW[3] = {0, 0, 0};
	eN[aN] = 1;
	eV[aV] = 1;
	eW[aW] = 1;
	for (int k = fbx->lo[2]; k <= fbx->hi[2]; ++k)
		for (int j = fbx->lo[1]; j <= fbx->hi[1]; ++j)
			for (int i = fbx->lo[0]; i <= fbx->hi[0]; ++i) {
				const double rho_L = A4(L, i, j, k, 0), rho_R = A4(R, i, j, k, 0);
				const double vx_L = A4(L, i, j, k, 1), vx_R = A4(R, i, j, k, 1);
				const double vy_L = A4(L, i, j, k, 2), vy_R = A4(R, i, j, k, 2);
				const double vz_L = A4(L, i, j, k, 3), vz_R = A4(R, i, j, k, 3);
				const double ke_L = 0.5 * rho_L * (vx_L * vx_L + vy_L * vy_L + vz_L * vz_L);
				const double ke_R = 0.5 * rho_R * (vx_R * vx_R + vy_R * vy_R + vz_R * vz_R);
				double Eint_L = NAN, Eint_R = NAN, P_L, P_R;
				const int iso = (prm->gamma == 1.0);
				if (iso) { /* :910-915: E and Eint stay NAN, their fluxes are set to zero below */
					P_L = rho_L * (prm->cs_isothermal * prm->cs_isothermal);
					P_R = rho_R * (prm->cs_isothermal * prm->cs_isothermal);
				} else if (prm->reconstruct_eint) {
					const double eint_L = A4(L, i, j, k, 4), eint_R = A4(R, i, j, k, 4);
					P_L = eos_pressure(prm, rho_L, eint_L * rho_L);
					P_R = eos_pressure(prm, rho_R, eint_R * rho_R);
					Eint_L = rho_L * A4(L, i, j, k, 5);
					Eint_R = rho_R * A4(R, i, j, k, 5);
				} else {
					P_L = A4(L, i, j, k, 4);
					P_R = A4(R, i, j, k, 4);
					Eint_L = A4(L, i, j, k, 5);
					Eint_R = A4(R, i, j, k, 5);
				}
				const double cs_L = iso ? prm->cs_isothermal : eos_sound_speed(prm, rho_L, P_L);
				const double E_L = iso ? NAN : eos_eint_from_pres(prm, rho_L, P_L) + ke_L;
				const double cs_R = iso ? prm->cs_isothermal : eos_sound_speed(prm, rho_R, P_R);
				const double E_R = iso ? NAN : eos_eint_from_pres(prm, rho_R, P_R) + ke_R;
				hstate sL, sR;
				sL.rho = rho_L;
				sL.u = A4(L, i, j, k, velN);
				sL.v = A4(L, i, j, k, velV);
				sL.w = A4(L, i, j, k, velW);
				sL.P = P_L;
				sL.cs = cs_L;
				sL.E = E_L;
				sL.Eint = Eint_L;
				sR.rho = rho_R;
				sR.u = A4(R, i, j, k, velN);
				sR.v = A4(R, i, j, k, velV);
				sR.w = A4(R, i, j, k, velW);
				sR.P = P_R;
				sR.cs = cs_R;
				sR.E = E_R;
				sR.Eint = Eint_R;
				for (int n = 0; n < ns; ++n) {
					sL.scalar[n] = A4(L, i, j, k, SC0 + n);
					sR.scalar[n] = A4(R, i, j, k, SC0 + n);
				}
#define QC(di, dj, dk, n) A4(q, i + (di) * eN[0] + (dj) * eV[0] + (dk) * eW[0], j + (di) * eN[1] + (dj) * eV[1] + (dk) * eW[1], k + (di) * eN[2] + (dj) * eV[2] + (dk) * eW[2], n)
				const double du = QC(0, 0, 0, velN) - QC(-1, 0, 0, velN);
				const double dvl = dmin(QC(-1, 1, 0, velV) - QC(-1, 0, 0, velV), QC(-1, 0, 0, velV) - QC(-1, -1, 0, velV));
				const double dvr = dmin(QC(0, 1, 0, velV) - QC(0, 0, 0, velV), QC(0, 0, 0, velV) - QC(0, -1, 0, velV));
				double dw = dmin(dvl, dvr);
				const double dwl = dmin(QC(-1, 0, 1, velW) - QC(-1, 0, 0, velW), QC(-1, 0, 0, velW) - QC(-1, 0, -1, velW));
				const double dwr = dmin(QC(0, 0, 1, velW) - QC(0, 0, 0, velW), QC(0, 0, 0, velW) - QC(0, 0, -1, velW));
				dw = dmin(dmin(dwl, dwr), dw);
#undef QC
				double Fc[MAXV], F[MAXV];
				if (solver == QK_HLLC) {
					riemann_hllc(prm, &sL, &sR, du, dw, ns, Fc);
				} else {
					riemann_llf(&sL, &sR, ns, Fc);
				}
				/* artificial viscosity :1052-1076 */
				const double div_v = du + 0.5 * (dvl + dvr) + 0.5 * (dwl + dwr);
				const double viscosity = prm->K_visc * dmax(-div_v, 0.);
				double U_L[MAXV] = {sL.rho, sL.rho * sL.u, sL.rho * sL.v, sL.rho * sL.w, sL.E, sL.Eint};
				double U_R[MAXV] = {sR.rho, sR.rho * sR.u, sR.rho * sR.v, sR.rho * sR.w, sR.E, sR.Eint};
				double fluxSum_U_L = 0, fluxSum_U_R = 0;
				for (int n = 0; n < ns; ++n) {
					U_L[6 + n] = sL.scalar[n];
					U_R[6 + n] = sR.scalar[n];
					if (n < nms) {
						fluxSum_U_L += U_L[6 + n];
						fluxSum_U_R += U_R[6 + n];
					}
				}
				for (int n = 0; n < nv; ++n)
					F[n] = Fc[n] + viscosity * (U_L[n] - U_R[n]);
				F[velN] = Fc[1];
				F[velV] = Fc[2];
				F[velW] = Fc[3];
				if (iso) { /* :1083-1087 */
					F[EN] = 0;
					F[EI] = 0;
				}
				const double v_norm = (F[0] >= 0.) ? (F[0] / rho_R) : (F[0] / rho_L);
				A4(facevel, i, j, k, 0) = v_norm;
				if (F[0] >= 0.) {
					for (int n = 0; n < nms; ++n)
						F[6 + n] = F[0] * U_L[6 + n] / fluxSum_U_L;
				} else {
					for (int n = 0; n < nms; ++n)
						F[6 + n] = F[0] * U_R[6 + n] / fluxSum_U_R;
				}
				for (int n = 0; n < nv; ++n)
					A4(flux, i, j, k, n) = F[n];
			}
}

/* MultiFab::Saxpy: dst += a*src (extern/amrex/Src/Base/AMReX_MultiFab.cpp Saxpy -> FabArray::Saxpy) */
void orc_saxpy(const qk_array4 *dst, double a, const qk_array4 *src, const qk_box *bx, int ncomp)
{
	for (int n = 0; n < ncomp; ++n)
		for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
			for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
				for (int i = bx->lo[0]; i <= bx->hi[0]; ++i)
					A4(dst, i, j, k, n) += a * A4(src, i, j, k, n);
}

/* HydroSystem::ComputeRhsFromFluxes  hydro_system.hpp:448-473 */
void orc_rhs_from_fluxes(const qk_array4 *rhs, const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz, const double dx[3], const qk_box *bx,
			 int nvars)
{
	for (int n = 0; n < nvars; ++n)
		for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
			for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
				for (int i = bx->lo[0]; i <= bx->hi[0]; ++i)
					A4(rhs, i, j, k, n) = (1.0 / dx[0]) * (A4(fx, i, j, k, n) - A4(fx, i + 1, j, k, n)) +
							      (1.0 / dx[1]) * (A4(fy, i, j, k, n) - A4(fy, i, j + 1, k, n)) +
							      (1.0 / dx[2]) * (A4(fz, i, j, k, n) - A4(fz, i, j, k + 1, n));
}

#define IA4(a, i, j, k) ((a)->p[((int64_t)(i) - (a)->begin[0]) + ((int64_t)(j) - (a)->begin[1]) * (a)->jstride + ((int64_t)(k) - (a)->begin[2]) * (a)->kstride])

/* HydroSystem::AddInternalEnergyPdV  hydro_system.hpp:775-814 */
void orc_add_internal_energy_pdv(const qk_hydro_params *prm, const qk_array4 *rhs, const qk_array4 *c, const double dx[3], const qk_array4 *vx,
				 const qk_array4 *vy, const qk_array4 *vz, const qk_iarray4 *redo, const qk_box *bx)
{
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double Pgas = cons_pressure(prm, c, i, j, k);
				double div_v;
				if (IA4(redo, i, j, k) == 0) {
					div_v = (A4(vx, i + 1, j, k, 0) - A4(vx, i, j, k, 0)) / dx[0] + (A4(vy, i, j + 1, k, 0) - A4(vy, i, j, k, 0)) / dx[1] +
						(A4(vz, i, j, k + 1, 0) - A4(vz, i, j, k, 0)) / dx[2];
				} else {
					div_v = 0.5 * ((A4(c, i + 1, j, k, MX) / A4(c, i + 1, j, k, RHO) - A4(c, i - 1, j, k, MX) / A4(c, i - 1, j, k, RHO)) / dx[0] +
						       (A4(c, i, j + 1, k, MY) / A4(c, i, j + 1, k, RHO) - A4(c, i, j - 1, k, MY) / A4(c, i, j - 1, k, RHO)) / dx[1] +
						       (A4(c, i, j, k + 1, MZ) / A4(c, i, j, k + 1, RHO) - A4(c, i, j, k - 1, MZ) / A4(c, i, j, k - 1, RHO)) / dx[2]);
				}
				A4(rhs, i, j, k, EI) += -Pgas * div_v;
			}
}

/* HydroSystem::isStateValid  hydro_system.hpp:423-446 */
static int state_valid(const qk_hydro_params *prm, const qk_array4 *c, int i, int j, int k)
{
	int ok = (A4(c, i, j, k, RHO) > 0.);
	for (int n = 0; n < prm->nmscalars; ++n)
		if (A4(c, i, j, k, SC0 + n) < 0.0) {
			ok = 0;
			break;
		}
	return ok;
}

/* HydroSystem::PredictStep  hydro_system.hpp:475-497; returns redoFlag.sum() over bx */
int64_t orc_predict_step(const qk_hydro_params *prm, const qk_array4 *uo, const qk_array4 *un, const qk_array4 *rhs, double dt, int nvars,
			 const qk_iarray4 *redo, const qk_box *bx)
{
	int64_t nbad = 0;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				for (int n = 0; n < nvars; ++n)
					A4(un, i, j, k, n) = A4(uo, i, j, k, n) + dt * A4(rhs, i, j, k, n);
				const int bad = !state_valid(prm, un, i, j, k);
				IA4(redo, i, j, k) = bad;
				nbad += bad;
			}
	return nbad;
}

/* HydroSystem::EnforceLimits  hydro_system.hpp:698-773 */
void orc_enforce_limits(const qk_hydro_params *prm, const qk_array4 *s, const qk_box *bx)
{
	const int ns = prm->nscalars, nms = prm->nmscalars;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(s, i, j, k, RHO);
				double rho_new = rho;
				if (rho < prm->density_floor) {
					rho_new = prm->density_floor;
					A4(s, i, j, k, RHO) = rho_new;
					for (int n = 0; n < ns; ++n) {
						if (rho_new == 0.0)
							A4(s, i, j, k, SC0 + n) = 0.0;
						else
							A4(s, i, j, k, SC0 + n) *= rho / rho_new;
					}
				}
				if (nms > 0) {
					double sp_sum = 0.0;
					for (int n = 0; n < nms; ++n) {
						if (A4(s, i, j, k, SC0 + n) < 0.0)
							A4(s, i, j, k, SC0 + n) = prm->small_x * rho_new;
						sp_sum += A4(s, i, j, k, SC0 + n);
					}
					if ((sp_sum > DBL_MIN) && (rho_new > DBL_MIN)) {
						sp_sum /= rho_new;
						for (int n = 0; n < nms; ++n)
							A4(s, i, j, k, SC0 + n) /= sp_sum;
					}
				}
				if ((rho_new > DBL_MIN) && prm->gamma != 1.0) {
					const double vx1 = A4(s, i, j, k, MX) / rho_new;
					const double vx2 = A4(s, i, j, k, MY) / rho_new;
					const double vx3 = A4(s, i, j, k, MZ) / rho_new;
					const double Ekin = 0.5 * rho_new * (vx1 * vx1 + vx2 * vx2 + vx3 * vx3);
					const double Etot = A4(s, i, j, k, EN);
					const double primTemp = eos_tgas_from_eint(prm, rho_new, (Etot - Ekin));
					if (primTemp < prm->temp_floor) {
						const double prim_eint = eos_eint_from_tgas(prm, rho_new, prm->temp_floor);
						A4(s, i, j, k, EN) = Ekin + prim_eint;
					}
					const double auxEint = A4(s, i, j, k, EI);
					const double auxTemp = eos_tgas_from_eint(prm, rho_new, auxEint);
					if (auxTemp < prm->temp_floor) {
						A4(s, i, j, k, EI) = eos_eint_from_tgas(prm, rho_new, prm->temp_floor);
					}
				}
			}
}

/* HydroSystem::SyncDualEnergy  hydro_system.hpp:816-850; returns the number of cells with rho<=0
 * (where the reference aborts; those cells are left untouched here) */
int64_t orc_sync_dual_energy(const qk_hydro_params *prm, const qk_array4 *s, const qk_box *bx)
{
	(void)prm;
	const double eta = 1.0e-3;
	int64_t nabort = 0;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(s, i, j, k, RHO), px = A4(s, i, j, k, MX), py = A4(s, i, j, k, MY), pz = A4(s, i, j, k, MZ);
				const double Etot = A4(s, i, j, k, EN), Eint_aux = A4(s, i, j, k, EI);
				if (rho <= 0.) {
					++nabort;
					continue;
				}
				const double Ekin = (px * px + py * py + pz * pz) / (2.0 * rho);
				const double Eint_cons = Etot - Ekin;
				if (Eint_cons > eta * Etot) {
					A4(s, i, j, k, EI) = Eint_cons;
				} else {
					A4(s, i, j, k, EI) = Eint_aux;
					A4(s, i, j, k, EN) = Eint_aux + Ekin;
				}
			}
	return nabort;
}

static inline int a4_contains(const qk_array4 *a, int i, int j, int k)
{
	return i >= a->begin[0] && i < a->end[0] && j >= a->begin[1] && j < a->end[1] && k >= a->begin[2] && k < a->end[2];
}

/* QuokkaSimulation::replaceFluxes (one direction)  src/QuokkaSimulation.hpp:1324-1368 */
void orc_replace_fluxes(int dir, const qk_array4 *flux, const qk_array4 *fo, const qk_iarray4 *redo, const qk_box *valid, int ncomp)
{
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2);
	for (int n = 0; n < ncomp; ++n)
		for (int k = valid->lo[2] - 1; k <= valid->hi[2] + 1; ++k)
			for (int j = valid->lo[1] - 1; j <= valid->hi[1] + 1; ++j)
				for (int i = valid->lo[0] - 1; i <= valid->hi[0] + 1; ++i) {
					if (IA4(redo, i, j, k) == 1) {
						if (a4_contains(flux, i, j, k))
							A4(flux, i, j, k, n) = A4(fo, i, j, k, n);
						if (a4_contains(flux, i + e0, j + e1, k + e2))
							A4(flux, i + e0, j + e1, k + e2, n) = A4(fo, i + e0, j + e1, k + e2, n);
					}
				}
}

/* which=0: HydroSystem::ComputeMaxSignalSpeed + norminf (hydro_system.hpp:223-252);
 * which=1: HydroSystem::maxSignalSpeedLocal (hydro_system.hpp:198-221) */
double orc_max_signal_speed(const qk_hydro_params *prm, int which, const qk_array4 *c, const qk_box *bx)
{
	double m = (which == 0) ? 0.0 : -DBL_MAX; /* norminf starts from 0 (abs), ParReduce max from lowest */
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(c, i, j, k, RHO), px = A4(c, i, j, k, MX), py = A4(c, i, j, k, MY), pz = A4(c, i, j, k, MZ);
				const double cs = (prm->gamma == 1.0) ? prm->cs_isothermal : cons_sound_speed(prm, c, i, j, k); /* :214-218, 242-246 */
				double sig;
				if (which == 0) {
					const double vx = px / rho, vy = py / rho, vz = pz / rho;
					const double vel_mag = sqrt(vx * vx + vy * vy + vz * vz);
					sig = fabs(cs + vel_mag);
				} else {
					const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
					const double abs_vel = sqrt(2.0 * kinetic_energy / rho);
					sig = cs + abs_vel;
				}
				m = dmax(m, sig);
			}
	return m;
}

/* ================================================================================================
 * Two-moment radiation transport  src/radiation/radiation_system.hpp
 * ============================================================================================== */
/* RadSystem::ConservedToPrimitive  :589-614 */
void orc_rad_conserved_to_primitive(const qk_rad_params *prm, const qk_array4 *cons, const qk_array4 *prim, const qk_box *bx)
{
	const int ns = prm->nstart;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i)
				for (int g = 0; g < prm->ngroups; ++g) {
					const double E_r = A4(cons, i, j, k, ns + 4 * g);
					const double Fx = A4(cons, i, j, k, ns + 4 * g + 1);
					const double Fy = A4(cons, i, j, k, ns + 4 * g + 2);
					const double Fz = A4(cons, i, j, k, ns + 4 * g + 3);
					A4(prim, i, j, k, 4 * g) = E_r;
					A4(prim, i, j, k, 4 * g + 1) = Fx / (prm->c_light * E_r);
					A4(prim, i, j, k, 4 * g + 2) = Fy / (prm->c_light * E_r);
					A4(prim, i, j, k, 4 * g + 3) = Fz / (prm->c_light * E_r);
				}
}

/* RadSystem::ComputeEddingtonFactor  :773-790 (Levermore 1984 closure) */
static double rad_eddington_factor(double f_in)
{
	const double f = clampd(f_in, 0., 1.);
	const double f_fac = sqrt(4.0 - 3.0 * (f * f));
	return (3.0 + 4.0 * (f * f)) / (5.0 + 2.0 * f_fac);
}

/* RadSystem::ComputeEddingtonTensor :873-916 + ComputeRadPressure<DIR> :918-983: F[4] = (F_n, T_nx E, T_ny E, T_nz E), S = max(0.1, sqrt(T_nn)) */
static void rad_pressure(int dir, double erad, const double Fv[3], const double fv[3], double F[4], double *S)
{
	const double f = sqrt(fv[0] * fv[0] + fv[1] * fv[1] + fv[2] * fv[2]);
	double n[3];
	for (int ii = 0; ii < 3; ++ii)
		n[ii] = (f > 0.) ? (fv[ii] / f) : 0.;
	const double chi = rad_eddington_factor(f);
	const double Tdiag = (1.0 - chi) / 2.0;
	const double Tf = (3.0 * chi - 1.0) / 2.0;
	double T[3][3];
	for (int ii = 0; ii < 3; ++ii)
		for (int jj = 0; jj < 3; ++jj) {
			const double delta_ij = (ii == jj) ? 1 : 0;
			T[ii][jj] = Tdiag * delta_ij + Tf * (n[ii] * n[jj]);
		}
	const double Tnormal = T[dir][dir];
	F[0] = Fv[dir];
	F[1] = T[dir][0] * erad;
	F[2] = T[dir][1] * erad;
	F[3] = T[dir][2] * erad;
	const double sq = sqrt(Tnormal);
	*S = (0.1 < sq) ? sq : 0.1; /* std::max(0.1, sqrt(Tnormal)) */
}

/* RadSystem::ComputeFluxes<DIR>  :985-1139; with prm->use_wavespeed_correction the energy component's diffusive term is scaled by
 * epsilon = min(1, 1 / tau_cell) on faces with even i+j+k (:1018-1022,1100-1109; ComputeCellOpticalDepth :803-871 for one group and a constant
 * flux-mean opacity; the gas temperature the reference evaluates there does not enter a constant opacity).  fdiff may be NULL. */
void orc_rad_compute_fluxes(const qk_rad_params *prm, int dir, const qk_array4 *flux, const qk_array4 *fdiff, const qk_array4 *left,
			    const qk_array4 *right, const qk_array4 *cons, const qk_box *facebx)
{
	const int e0 = (dir == 0), e1 = (dir == 1), e2 = (dir == 2), ns = prm->nstart;
	const double c = prm->c_light, chat = prm->c_hat;
	for (int k = facebx->lo[2]; k <= facebx->hi[2]; ++k)
		for (int j = facebx->lo[1]; j <= facebx->hi[1]; ++j)
			for (int i = facebx->lo[0]; i <= facebx->hi[0]; ++i)
				for (int g = 0; g < prm->ngroups; ++g) {
					double erad_L = A4(left, i, j, k, 4 * g), erad_R = A4(right, i, j, k, 4 * g);
					double fL[3], fR[3], FL[3], FR[3];
					for (int m = 0; m < 3; ++m) {
						fL[m] = A4(left, i, j, k, 4 * g + 1 + m);
						fR[m] = A4(right, i, j, k, 4 * g + 1 + m);
					}
					double f_L = sqrt(fL[0] * fL[0] + fL[1] * fL[1] + fL[2] * fL[2]);
					double f_R = sqrt(fR[0] * fR[0] + fR[1] * fR[1] + fR[2] * fR[2]);
					for (int m = 0; m < 3; ++m) {
						FL[m] = fL[m] * (c * erad_L);
						FR[m] = fR[m] * (c * erad_R);
					}
					if ((erad_L <= 0.) || (erad_R <= 0.) || (f_L >= 1.) || (f_R >= 1.)) { /* first-order fallback :1054-1079 */
						erad_L = A4(cons, i - e0, j - e1, k - e2, ns + 4 * g);
						erad_R = A4(cons, i, j, k, ns + 4 * g);
						for (int m = 0; m < 3; ++m) {
							FL[m] = A4(cons, i - e0, j - e1, k - e2, ns + 4 * g + 1 + m);
							FR[m] = A4(cons, i, j, k, ns + 4 * g + 1 + m);
							fL[m] = FL[m] / (c * erad_L);
							fR[m] = FR[m] / (c * erad_R);
						}
						f_L = sqrt(fL[0] * fL[0] + fL[1] * fL[1] + fL[2] * fL[2]);
						f_R = sqrt(fR[0] * fR[0] + fR[1] * fR[1] + fR[2] * fR[2]);
					}
					double F_L[4], F_R[4], S_L, S_R;
					rad_pressure(dir, erad_L, FL, fL, F_L, &S_L);
					S_L *= -1.;
					rad_pressure(dir, erad_R, FR, fR, F_R, &S_R);
					F_L[0] *= chat / c;
					F_R[0] *= chat / c;
					for (int n = 1; n < 4; ++n) {
						F_L[n] *= chat * c;
						F_R[n] *= chat * c;
					}
					S_L *= chat;
					S_R *= chat;
					const double U_L[4] = {erad_L, FL[0], FL[1], FL[2]};
					const double U_R[4] = {erad_R, FR[0], FR[1], FR[2]};
					const double a = S_R / (S_R - S_L), b = S_L / (S_R - S_L), d = S_R * S_L / (S_R - S_L);
					double eps0 = 1.0;
					if (prm->use_wavespeed_correction && ((i + j + k) % 2 == 0)) { /* no correction for odd zones :1105 */
						const double dl = prm->cell_dx[dir];
						const double tau_L = dl * A4(cons, i - e0, j - e1, k - e2, 0) * prm->kappa_F;
						const double tau_R = dl * A4(cons, i, j, k, 0) * prm->kappa_F;
						const double tau = (tau_L * tau_R * 2.) / (tau_L + tau_R); /* harmonic mean :864 */
						const double inv = 1.0 / tau;
						eps0 = (inv < 1.0) ? inv : 1.0; /* std::min(1.0, 1.0 / tau_cell) */
					}
					for (int n = 0; n < 4; ++n) {
						const double eps = (n == 0) ? eps0 : 1.0;
						A4(flux, i, j, k, 4 * g + n) = a * F_L[n] - b * F_R[n] + (eps * d) * (U_R[n] - U_L[n]);
						if (fdiff)
							A4(fdiff, i, j, k, 4 * g + n) = a * F_L[n] - b * F_R[n] + d * (U_R[n] - U_L[n]);
					}
				}
}

/* RadSystem::isStateValid :624-643 / amendRadState :645-665 on the 4*ngroups hyperbolic variables of one cell */
static void rad_validate(const qk_rad_params *prm, double *cons)
{
	const double c = prm->c_light, floor_g = prm->Erad_floor / prm->ngroups;
	int valid = 1;
	for (int g = 0; g < prm->ngroups; ++g) {
		const double E_r = cons[4 * g], Fx = cons[4 * g + 1], Fy = cons[4 * g + 2], Fz = cons[4 * g + 3];
		const double Fnorm = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
		const double f = Fnorm / (c * E_r);
		valid = (valid && (E_r > 0.) && (f <= 1.));
	}
	if (valid)
		return;
	for (int g = 0; g < prm->ngroups; ++g) {
		double E_r = cons[4 * g];
		if (E_r < floor_g) {
			E_r = floor_g;
			cons[4 * g] = floor_g;
		}
		const double Fx = cons[4 * g + 1], Fy = cons[4 * g + 2], Fz = cons[4 * g + 3];
		if (Fx * Fx + Fy * Fy + Fz * Fz > c * c * E_r * E_r) {
			const double Fnorm = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
			cons[4 * g + 1] = Fx / Fnorm * c * E_r;
			cons[4 * g + 2] = Fy / Fnorm * c * E_r;
			cons[4 * g + 3] = Fz / Fnorm * c * E_r;
		}
	}
}

/* RadSystem::PredictStep  :667-710 */
void orc_rad_predict_step(const qk_rad_params *prm, const qk_array4 *uo, const qk_array4 *un, const qk_array4 *fx, const qk_array4 *fy,
			  const qk_array4 *fz, double dt, const double dx[3], const qk_box *bx)
{
	const int nh = 4 * prm->ngroups, ns = prm->nstart;
	double cons[4 * QK_MAX_GROUPS];
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				for (int n = 0; n < nh; ++n)
					cons[n] = A4(uo, i, j, k, ns + n) + ((dt / dx[0]) * (A4(fx, i, j, k, n) - A4(fx, i + 1, j, k, n)) +
									     (dt / dx[1]) * (A4(fy, i, j, k, n) - A4(fy, i, j + 1, k, n)) +
									     (dt / dx[2]) * (A4(fz, i, j, k, n) - A4(fz, i, j, k + 1, n)));
				rad_validate(prm, cons);
				for (int n = 0; n < nh; ++n)
					A4(un, i, j, k, ns + n) = cons[n];
			}
}

/* RadSystem::AddFluxesRK2  :712-771 (IMEX_a32 = 0.5 :52) */
void orc_rad_add_fluxes_rk2(const qk_rad_params *prm, const qk_array4 *unew, const qk_array4 *u0, const qk_array4 *u1, const qk_array4 *fxo,
			    const qk_array4 *fyo, const qk_array4 *fzo, const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz, double dt,
			    const double dx[3], const qk_box *bx)
{
	const int nh = 4 * prm->ngroups, ns = prm->nstart;
	const double IMEX_a32 = 0.5;
	double cons[4 * QK_MAX_GROUPS];
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				for (int n = 0; n < nh; ++n) {
					const double U_0 = A4(u0, i, j, k, ns + n);
					const double U_1 = A4(u1, i, j, k, ns + n);
					const double FxU_0 = (dt / dx[0]) * (A4(fxo, i, j, k, n) - A4(fxo, i + 1, j, k, n));
					const double FxU_1 = (dt / dx[0]) * (A4(fx, i, j, k, n) - A4(fx, i + 1, j, k, n));
					const double FyU_0 = (dt / dx[1]) * (A4(fyo, i, j, k, n) - A4(fyo, i, j + 1, k, n));
					const double FyU_1 = (dt / dx[1]) * (A4(fy, i, j, k, n) - A4(fy, i, j + 1, k, n));
					const double FzU_0 = (dt / dx[2]) * (A4(fzo, i, j, k, n) - A4(fzo, i, j, k + 1, n));
					const double FzU_1 = (dt / dx[2]) * (A4(fz, i, j, k, n) - A4(fz, i, j, k + 1, n));
					cons[n] = (1.0 - IMEX_a32) * U_0 + IMEX_a32 * U_1 + ((0.5 - IMEX_a32) * (FxU_0 + FyU_0 + FzU_0)) +
						  (0.5 * (FxU_1 + FyU_1 + FzU_1));
				}
				rad_validate(prm, cons);
				for (int n = 0; n < nh; ++n)
					A4(unew, i, j, k, ns + n) = cons[n];
			}
}

/* ================================================================================================
 * Matter-radiation coupling, single photon group: RadSystem::AddSourceTermsSingleGroup
 * src/radiation/source_terms_single_group.hpp:9-565 (hyper-parameters src/radiation/radiation_system.hpp:34-52:
 * include_work_term_in_source = true, enable_dE_constrain = true, force_rad_floor_in_iteration = false,
 * add_line_cooling_to_radiation_in_jac = false, IMEX_a32 = 0.5; no dust model (ISM_Traits default :85-89), zero
 * DefineNetCoolingRate / DefineCosmicRayHeatingRate (:522-545)).
 * ============================================================================================== */
/* EOS::ComputeEintTempDerivative  src/hydro/EOS.hpp:200-240 */
static double eos_eint_temp_derivative(const qk_hydro_params *prm, double rho, double Tgas)
{
	eos_state s = eos_new(prm);
	s.rho = rho;
	s.T = Tgas;
	eos_call(prm, EOS_RT, &s);
	return s.dedT * rho * prm->boltzmann_constant / C_k_B;
}
/* RadSystem::ComputeEintFromEgas / ComputeEgasFromEint  radiation_system.hpp:1289-1309 */
static double rad_eint_from_egas(double rho, double px, double py, double pz, double Etot)
{
	const double p_sq = px * px + py * py + pz * pz;
	const double Ekin = p_sq / (2.0 * rho);
	return Etot - Ekin;
}
static double rad_egas_from_eint(double rho, double px, double py, double pz, double Eint)
{
	const double p_sq = px * px + py * py + pz * pz;
	const double Ekin = p_sq / (2.0 * rho);
	return Eint + Ekin;
}
/* RadSystem::ComputeEddingtonTensor  radiation_system.hpp:873-916 (all nine entries) */
static void rad_eddington_tensor(double fx, double fy, double fz, double T[3][3])
{
	const double f = sqrt(fx * fx + fy * fy + fz * fz);
	const double fvec[3] = {fx, fy, fz};
	double n[3];
	for (int ii = 0; ii < 3; ++ii)
		n[ii] = (f > 0.) ? (fvec[ii] / f) : 0.;
	const double chi = rad_eddington_factor(f);
	const double Tdiag = (1.0 - chi) / 2.0;
	const double Tf = (3.0 * chi - 1.0) / 2.0;
	for (int ii = 0; ii < 3; ++ii)
		for (int jj = 0; jj < 3; ++jj) {
			const double delta_ij = (ii == jj) ? 1 : 0;
			T[ii][jj] = Tdiag * delta_ij + Tf * (n[ii] * n[jj]);
		}
}
/* RadSystem::Solve3x3matrix  radiation_system.hpp:560-580 */
static void rad_solve3x3(double C00, double C01, double C02, double C10, double C11, double C12, double C20, double C21, double C22, double Y0,
			 double Y1, double Y2, double X[3])
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

void orc_rad_add_source_terms(const qk_hydro_params *hp, const qk_rad_params *prm, const qk_rad_source_params *sp, const qk_array4 *cons,
			      const qk_array4 *rad_energy_source, const qk_box *bx, double dt_radiation, int stage, int64_t *counters)
{
	const double IMEX_a32 = 0.5;
	const int iE = prm->nstart, iFx = prm->nstart + 1, iFy = prm->nstart + 2, iFz = prm->nstart + 3;
	const double gamma_ = hp->gamma;
	const int beta_order_ = sp->beta_order;
	const double Erad_floor_ = prm->Erad_floor / prm->ngroups; /* radiation_system.hpp:211 */
	const double a_rad = sp->radiation_constant;
	double dt = dt_radiation;
	if (stage == 2) /* :17-19 */
		dt = (1.0 - IMEX_a32) * dt_radiation;

	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double c = prm->c_light;
				const double chat = prm->c_hat;
				/* :40-55 */
				const double rho = A4(cons, i, j, k, RHO);
				const double x1GasMom0 = A4(cons, i, j, k, MX);
				const double x2GasMom0 = A4(cons, i, j, k, MY);
				const double x3GasMom0 = A4(cons, i, j, k, MZ);
				const double gasMtm0[3] = {x1GasMom0, x2GasMom0, x3GasMom0};
				const double Egastot0 = A4(cons, i, j, k, EN);
				const double Erad0 = A4(cons, i, j, k, iE);
				const double Src = (rad_energy_source ? A4(rad_energy_source, i, j, k, 0) : 0.0) * dt * chat;

				double Egas0 = NAN, Ekin0 = NAN, Etot0 = NAN, Egas_guess = NAN, T_gas = NAN, T_d = NAN;
				double lorentz_factor = NAN, lorentz_factor_v = NAN, lorentz_factor_v_v = NAN;
				double fourPiBoverC = NAN, Erad_guess = NAN, kappaP = NAN, kappaE = NAN, kappaF = NAN, kappaPoverE = NAN;
				double work = 0.0, work_prev = 0.0;
				double dMomentum[3] = {0., 0., 0.}, Frad_t1[3] = {0., 0., 0.};
				const double cscale = c / chat;

				if (gamma_ != 1.0) { /* :82-86 */
					Egas0 = rad_eint_from_egas(rho, x1GasMom0, x2GasMom0, x3GasMom0, Egastot0);
					Etot0 = Egas0 + cscale * (Erad0 + Src);
				}
				double gas_update_factor = 1.0;
				if (stage == 1)
					gas_update_factor = IMEX_a32;

				const int max_ite = 5;
				int ite = 0;
				for (; ite < max_ite; ++ite) {
					double R = NAN;
					Erad_guess = Erad0;
					if (gamma_ != 1.0) {
						double tau0 = NAN, tau = NAN;
						Egas_guess = Egas0;
						Ekin0 = Egastot0 - Egas0;
						const double betaSqr =
						    (x1GasMom0 * x1GasMom0 + x2GasMom0 * x2GasMom0 + x3GasMom0 * x3GasMom0) / (rho * rho * c * c);
						if (beta_order_ == 0 || beta_order_ == 1) { /* :115-131 */
							lorentz_factor = 1.0;
							lorentz_factor_v = 1.0;
						} else if (beta_order_ == 2) {
							lorentz_factor = 1.0 + 0.5 * betaSqr;
							lorentz_factor_v = 1.0;
							lorentz_factor_v_v = 1.0;
						} else if (beta_order_ == 3) {
							lorentz_factor = 1.0 + 0.5 * betaSqr;
							lorentz_factor_v = 1.0 + 0.5 * betaSqr;
							lorentz_factor_v_v = 1.0;
						} else {
							lorentz_factor = 1.0 / sqrt(1.0 - betaSqr);
							lorentz_factor_v = lorentz_factor;
							lorentz_factor_v_v = lorentz_factor;
						}
						double F_G = NAN, deltaEgas = NAN, deltaR = NAN, F_D = NAN;
						const double resid_tol = 1.0e-11;
						const int maxIter = 100;
						int n = 0;
						for (; n < maxIter; ++n) { /* Newton-Raphson :161-352 */
							T_gas = eos_tgas_from_eint(hp, rho, Egas_guess);
							T_d = T_gas; /* no dust model :166-167 */
							/* ComputeThermalRadiationSingleGroup  radiation_system.hpp:471-479 */
							fourPiBoverC = a_rad * pow(T_d, 4);
							if (fourPiBoverC < Erad_floor_)
								fourPiBoverC = Erad_floor_;
							kappaP = sp->kappa_P;
							kappaE = sp->kappa_E;
							if (kappaE > 0.0)
								kappaPoverE = kappaP / kappaE;
							else
								kappaPoverE = 1.0;
							if (n == 0) { /* :192-215 */
								kappaF = sp->kappa_F;
								if (beta_order_ != 0) {
									if (ite == 0) {
										const double frad0 = A4(cons, i, j, k, iFx);
										const double frad1 = A4(cons, i, j, k, iFy);
										const double frad2 = A4(cons, i, j, k, iFz);
										work = (x1GasMom0 * frad0 + x2GasMom0 * frad1 + x3GasMom0 * frad2) *
										       (2.0 * kappaE - kappaF) * chat / (c * c) * lorentz_factor_v * dt;
									}
								}
								tau0 = dt * rho * kappaP * chat * lorentz_factor;
								tau = tau0;
								R = (fourPiBoverC - Erad_guess / kappaPoverE) * tau0 + work;
								tau0 = dmax(tau0, 1.0);
							} else { /* :216-232 */
								tau = dt * rho * kappaP * chat * lorentz_factor;
								if (tau > 0.0)
									Erad_guess = kappaPoverE * (fourPiBoverC - (R - work) / tau);
							}
							const double cooling = 0.0, cooling_derivative = 0.0;
							const double CR_heating = 0.0 * dt;
							F_G = Egas_guess - Egas0 + cscale * R + cooling * dt - CR_heating;
							F_D = Erad_guess - Erad0 - (R + Src);
							double F_D_abs = 0.0;
							if (tau > 0.0)
								F_D_abs = fabs(F_D);
							else
								F_D_abs = fabs(F_D + R);
							if ((fabs(F_G) < resid_tol * Etot0) && (cscale * F_D_abs < resid_tol * Etot0))
								break;
							const double c_v = eos_eint_temp_derivative(hp, rho, T_gas);
							/* ComputeThermalRadiationTempDerivativeSingleGroup  radiation_system.hpp:499-503 */
							const double d_fourpiboverc_d_t = 4. * a_rad * pow(T_d, 3);
							const double dEg_dT = kappaPoverE * d_fourpiboverc_d_t;
							double J00, J01, J10, J11;
							J00 = 1.0 + cooling_derivative * dt / c_v;
							J01 = cscale;
							J10 = 1.0 / c_v * dEg_dT - (1 / cscale) * cooling_derivative * dt;
							if (tau <= 0.0)
								J11 = -INFINITY;
							else
								J11 = -1.0 * kappaPoverE / tau - 1.0;
							const double y0 = -F_G;
							const double y1 = -1. * F_D;
							const double det = J00 * J11 - J01 * J10;
							deltaEgas = (J11 * y0 - J01 * y1) / det;
							deltaR = (J00 * y1 - J10 * y0) / det;
							/* enable_dE_constrain :330-342 */
							const double T_rad = sqrt(sqrt(Erad_guess / a_rad));
							if (deltaEgas / c_v > dmax(T_gas, T_rad)) {
								Egas_guess = eos_eint_from_tgas(hp, rho, T_rad);
							} else {
								Egas_guess += deltaEgas;
								R += deltaR;
							}
						}
						if (counters) { /* :354-362 */
							if (n >= maxIter)
								counters[4] += 1;
							counters[0] += 1;
							counters[1] += n + 1;
							if (counters[2] < n + 1)
								counters[2] = n + 1;
						}
						{ /* !add_line_cooling_to_radiation_in_jac :367-373 */
							const double cooling_tend = 0.0 * dt;
							Erad_guess += (1 / cscale) * cooling_tend;
						}
						if (n > 0)
							kappaF = sp->kappa_F;
					} else { /* gamma == 1 :379-391 */
						T_d = T_gas;
						kappaF = sp->kappa_F;
					}

					/* 2. radiation flux update :396-490 */
					double Frad_t0[3];
					dMomentum[0] = dMomentum[1] = dMomentum[2] = 0.;
					Frad_t0[0] = A4(cons, i, j, k, iFx);
					Frad_t0[1] = A4(cons, i, j, k, iFy);
					Frad_t0[2] = A4(cons, i, j, k, iFz);
					if ((gamma_ != 1.0) && (beta_order_ != 0)) {
						const double erad = Erad_guess;
						const double gasVel[3] = {0., 0., 0.}; /* declared and never assigned in the reference (:407) */
						double v_terms[3];
						const double fx = Frad_t0[0] / (c * erad);
						const double fy = Frad_t0[1] / (c * erad);
						const double fz = Frad_t0[2] / (c * erad);
						const double F_coeff = chat * rho * kappaF * dt * lorentz_factor;
						double Tedd[3][3];
						rad_eddington_tensor(fx, fy, fz, Tedd);
						for (int n = 0; n < 3; ++n) {
							double Planck_term = kappaP * fourPiBoverC * lorentz_factor_v;
							if (kappaF != kappaE)
								Planck_term += (kappaF - kappaE) * erad * pow(lorentz_factor_v, 3);
							Planck_term *= chat * dt * gasMtm0[n];
							double pressure_term = 0.0;
							for (int z = 0; z < 3; ++z)
								pressure_term += gasMtm0[z] * Tedd[n][z] * erad;
							pressure_term *= chat * dt * kappaF * lorentz_factor_v;
							v_terms[n] = Planck_term + pressure_term;
						}
						if (beta_order_ == 1 || kappaF == kappaE) {
							for (int n = 0; n < 3; ++n) {
								Frad_t1[n] = (Frad_t0[n] + v_terms[n]) / (1.0 + F_coeff);
								dMomentum[n] += -(Frad_t1[n] - Frad_t0[n]) / (c * chat);
							}
						} else {
							const double K0 = 2.0 * rho * chat * dt * (kappaF - kappaE) / c / c * pow(lorentz_factor_v_v, 3);
							const double A00 = 1.0 + F_coeff + K0 * gasVel[0] * gasVel[0];
							const double A01 = K0 * gasVel[0] * gasVel[1];
							const double A02 = K0 * gasVel[0] * gasVel[2];
							const double A10 = K0 * gasVel[1] * gasVel[0];
							const double A11 = 1.0 + F_coeff + K0 * gasVel[1] * gasVel[1];
							const double A12 = K0 * gasVel[1] * gasVel[2];
							const double A20 = K0 * gasVel[2] * gasVel[0];
							const double A21 = K0 * gasVel[2] * gasVel[1];
							const double A22 = 1.0 + F_coeff + K0 * gasVel[2] * gasVel[2];
							const double B0 = v_terms[0] + Frad_t0[0];
							const double B1 = v_terms[1] + Frad_t0[1];
							const double B2 = v_terms[2] + Frad_t0[2];
							rad_solve3x3(A00, A01, A02, A10, A11, A12, A20, A21, A22, B0, B1, B2, Frad_t1);
							for (int n = 0; n < 3; ++n)
								dMomentum[n] += -(Frad_t1[n] - Frad_t0[n]) / (c * chat);
						}
					} else { /* :484-490 */
						for (int n = 0; n < 3; ++n) {
							Frad_t1[n] = Frad_t0[n] / (1.0 + rho * kappaF * chat * dt);
							dMomentum[n] += -(Frad_t1[n] - Frad_t0[n]) / (c * chat);
						}
					}
					const double x1GasMom1 = A4(cons, i, j, k, MX) + dMomentum[0];
					const double x2GasMom1 = A4(cons, i, j, k, MY) + dMomentum[1];
					const double x3GasMom1 = A4(cons, i, j, k, MZ) + dMomentum[2];

					/* 3. work term :496-524 */
					if ((gamma_ != 1.0) && (beta_order_ != 0)) {
						const double Egastot1 = rad_egas_from_eint(rho, x1GasMom1, x2GasMom1, x3GasMom1, Egas_guess);
						const double Ekin1 = Egastot1 - Egas_guess;
						const double dEkin_work = Ekin1 - Ekin0;
						Egas_guess -= dEkin_work;
					}
					if ((beta_order_ == 0) || (gamma_ == 1.0)) {
						break;
					} else { /* lagged work term :526-541 */
						work_prev = work;
						work = (x1GasMom1 * Frad_t1[0] + x2GasMom1 * Frad_t1[1] + x3GasMom1 * Frad_t1[2]) * chat / (c * c) *
						       lorentz_factor_v * (2.0 * kappaE - kappaF) * dt;
						const double lag_tol = 1.0e-13;
						if ((fabs(work) == 0.0) || (cscale * fabs(work - work_prev) < lag_tol * Etot0) ||
						    (fabs(work - work_prev) <= lag_tol * R) || (fabs(work - work_prev) <= 1.0e-8 * fabs(work)))
							break;
					}
				}
				if (ite >= max_ite && counters) /* :544-547 */
					counters[6] += 1;

				/* 4b. store :549-564 */
				const double x1GasMom1 = A4(cons, i, j, k, MX) + dMomentum[0] * gas_update_factor;
				const double x2GasMom1 = A4(cons, i, j, k, MY) + dMomentum[1] * gas_update_factor;
				const double x3GasMom1 = A4(cons, i, j, k, MZ) + dMomentum[2] * gas_update_factor;
				A4(cons, i, j, k, MX) = x1GasMom1;
				A4(cons, i, j, k, MY) = x2GasMom1;
				A4(cons, i, j, k, MZ) = x3GasMom1;
				if (gamma_ != 1.0) {
					Egas_guess = Egas0 + (Egas_guess - Egas0) * gas_update_factor;
					A4(cons, i, j, k, EI) = Egas_guess;
					A4(cons, i, j, k, EN) = rad_egas_from_eint(rho, x1GasMom1, x2GasMom1, x3GasMom1, Egas_guess);
					A4(cons, i, j, k, iE) = Erad_guess;
				}
				A4(cons, i, j, k, iFx) = Frad_t1[0];
				A4(cons, i, j, k, iFy) = Frad_t1[1];
				A4(cons, i, j, k, iFz) = Frad_t1[2];
			}
}

/* ================================================================================================
 * Coarse <-> fine transfer operators of the AMR ghost fill: the cell-centred interpolater Quokka selects with
 * amr_interpolation_method = 1 (getAmrInterpolaterCellCentered, src/simulation.hpp:1389-1407) =
 * amrex::mf_linear_slope_minmax_interp = MFCellConsLinMinmaxLimitInterp::interp (extern/amrex/Src/AmrCore/AMReX_MFInterpolater.cpp:
 * 332-418) with mf_cell_cons_lin_interp_limit_minmax_llslope + mf_cell_cons_lin_interp (AMReX_MFInterp_3D_C.H:7-109,246-262) and
 * mf_compute_slopes_{x,y,z} (AMReX_MFInterp_C.H:10-90); and amrex::average_down = amrex_avgdown (Base/AMReX_MultiFabUtil_3D_C.H:345-375).
 * ============================================================================================== */
static qk_array4 alloc_a4(qk_box b, int ncomp);
static void free_a4(qk_array4 *a);
static int coarsen_i(int i, int r) { return (i < 0) ? -((-i + r - 1) / r) : i / r; } /* amrex::coarsen(int, int): floor division */

/* mf_compute_slopes_<dir>: centred slope, one-sided at a domain face whose BC is ext_dir or hoextrap */
static double interp_slope(const qk_array4 *u, int i, int j, int k, int nu, int dir, const qk_box *domain, int bclo, int bchi)
{
	const int e[3] = {dir == 0, dir == 1, dir == 2};
	const int idx[3] = {i, j, k};
#define U_(o) A4(u, i + (o)*e[0], j + (o)*e[1], k + (o)*e[2], nu)
	double dc = 0.5 * (U_(1) - U_(-1));
	if (idx[dir] == domain->lo[dir] && (bclo == QK_BC_EXT_DIR || bclo == 4 /* hoextrap */)) {
		if (idx[dir] + 2 < u->end[dir])
			dc = -(16. / 15.) * U_(-1) + 0.5 * U_(0) + (2. / 3.) * U_(1) - 0.1 * U_(2);
		else
			dc = 0.25 * (U_(1) + 5. * U_(0) - 6. * U_(-1));
	}
	if (idx[dir] == domain->hi[dir] && (bchi == QK_BC_EXT_DIR || bchi == 4)) {
		if (idx[dir] - 2 >= u->begin[dir])
			dc = (16. / 15.) * U_(1) - 0.5 * U_(0) - (2. / 3.) * U_(-1) + 0.1 * U_(-2);
		else
			dc = -0.25 * (U_(-1) + 5. * U_(0) - 6. * U_(1));
	}
#undef U_
	return dc;
}

/* slopes of all ncomp components of coarse cell (i,j,k): slope has 3*ncomp components (x block, y block, z block) */
static void interp_limited_slopes(const qk_array4 *slope, const qk_array4 *u, int i, int j, int k, int scomp, int ncomp, const qk_box *domain,
				  const int ratio[3], const int32_t *bc_lo, const int32_t *bc_hi)
{
	double sf[3] = {1.0, 1.0, 1.0};
	for (int ns = 0; ns < ncomp; ++ns) {
		const int nu = ns + scomp;
		double sl[3] = {0., 0., 0.}, dcv[3] = {0., 0., 0.};
		for (int d = 0; d < 3; ++d) {
			const int e0 = (d == 0), e1 = (d == 1), e2 = (d == 2);
			if (ratio[d] > 1) {
				dcv[d] = interp_slope(u, i, j, k, nu, d, domain, bc_lo[3 * ns + d], bc_hi[3 * ns + d]);
				const double df = 2.0 * (A4(u, i + e0, j + e1, k + e2, nu) - A4(u, i, j, k, nu));
				const double db = 2.0 * (A4(u, i, j, k, nu) - A4(u, i - e0, j - e1, k - e2, nu));
				double sd = (df * db >= 0.0) ? dmin(fabs(df), fabs(db)) : 0.;
				sd = copysign(1., dcv[d]) * dmin(sd, fabs(dcv[d]));
				sl[d] = sd;
				A4(slope, i, j, k, ns + d * ncomp) = dcv[d]; /* unlimited slope */
			} else {
				A4(slope, i, j, k, ns + d * ncomp) = 0.0;
			}
		}
		/* :62-87 no new extrema in this component */
		double alpha = 1.0;
		if (sl[0] != 0.0 || sl[1] != 0.0 || sl[2] != 0.0) {
			const double dumax = fabs(sl[0]) * (double)(ratio[0] - 1) / (double)(2 * ratio[0]) +
					     fabs(sl[1]) * (double)(ratio[1] - 1) / (double)(2 * ratio[1]) +
					     fabs(sl[2]) * (double)(ratio[2] - 1) / (double)(2 * ratio[2]);
			const double uc = A4(u, i, j, k, nu);
			double umax = uc, umin = uc;
			const int il = ratio[0] > 1, jl = ratio[1] > 1, kl = ratio[2] > 1;
			for (int ko = -kl; ko <= kl; ++ko)
				for (int jo = -jl; jo <= jl; ++jo)
					for (int io = -il; io <= il; ++io) {
						umin = dmin(umin, A4(u, i + io, j + jo, k + ko, nu));
						umax = dmax(umax, A4(u, i + io, j + jo, k + ko, nu));
					}
			if (dumax * alpha > (umax - uc))
				alpha = (umax - uc) / dumax;
			if (dumax * alpha > (uc - umin))
				alpha = (uc - umin) / dumax;
		}
		for (int d = 0; d < 3; ++d) {
			sl[d] *= alpha;
			if (dcv[d] != 0.0) /* :90-98 */
				sf[d] = dmin(sf[d], sl[d] / dcv[d]);
		}
	}
	for (int ns = 0; ns < ncomp; ++ns) /* :102-106: one limiter per direction for ALL components (preserves linear combinations) */
		for (int d = 0; d < 3; ++d)
			A4(slope, i, j, k, ns + d * ncomp) *= sf[d];
}

/* MFCellConsLinMinmaxLimitInterp::interp on one box pair.  crse covers CoarseBox(fine_region) = coarsen(fine_region) grown by 1;
 * the fine cells of fine_region that lie inside dest_domain are written.  bc_lo / bc_hi: [3 * comp + dim] for comps ccomp.. */
void orc_interp_cons_lin_minmax(const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *fine_region,
				const qk_box *dest_domain, const qk_box *cdomain, const int ratio[3], const int32_t *bc_lo, const int32_t *bc_hi)
{
	qk_box cb; /* crse box shrunk by 1 where ratio > 1 (:391) */
	for (int d = 0; d < 3; ++d) {
		const int m1 = (ratio[d] > 1) ? 1 : 0;
		cb.lo[d] = crse->begin[d] + m1;
		cb.hi[d] = crse->end[d] - 1 - m1;
	}
	qk_array4 slope = alloc_a4(cb, 3 * ncomp);
	for (int k = cb.lo[2]; k <= cb.hi[2]; ++k)
		for (int j = cb.lo[1]; j <= cb.hi[1]; ++j)
			for (int i = cb.lo[0]; i <= cb.hi[0]; ++i)
				interp_limited_slopes(&slope, crse, i, j, k, ccomp, ncomp, cdomain, ratio, bc_lo, bc_hi);
	for (int n = 0; n < ncomp; ++n)
		for (int k = fine_region->lo[2]; k <= fine_region->hi[2]; ++k)
			for (int j = fine_region->lo[1]; j <= fine_region->hi[1]; ++j)
				for (int i = fine_region->lo[0]; i <= fine_region->hi[0]; ++i) {
					if (i < dest_domain->lo[0] || i > dest_domain->hi[0] || j < dest_domain->lo[1] || j > dest_domain->hi[1] ||
					    k < dest_domain->lo[2] || k > dest_domain->hi[2])
						continue;
					const int ic = coarsen_i(i, ratio[0]), jc = coarsen_i(j, ratio[1]), kc = coarsen_i(k, ratio[2]);
					const double xoff = ((double)(i - ic * ratio[0]) + 0.5) / (double)ratio[0] - 0.5;
					const double yoff = ((double)(j - jc * ratio[1]) + 0.5) / (double)ratio[1] - 0.5;
					const double zoff = ((double)(k - kc * ratio[2]) + 0.5) / (double)ratio[2] - 0.5;
					A4(fine, i, j, k, fcomp + n) = A4(crse, ic, jc, kc, ccomp + n) + xoff * A4(&slope, ic, jc, kc, n) +
								       yoff * A4(&slope, ic, jc, kc, n + ncomp) + zoff * A4(&slope, ic, jc, kc, n + ncomp * 2);
				}
	free_a4(&slope);
}

/* QuokkaSimulation::PreInterpState / PostInterpState  src/QuokkaSimulation.hpp:804-841: the hooks FillPatcher runs on the coarse data
 * before, and on the fine data after, the interpolation: gas total energy <-> specific internal energy */
void orc_pre_interp_state(const qk_array4 *c, const qk_box *bx)
{
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(c, i, j, k, RHO), px = A4(c, i, j, k, MX), py = A4(c, i, j, k, MY), pz = A4(c, i, j, k, MZ);
				const double Etot = A4(c, i, j, k, EN);
				const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
				A4(c, i, j, k, EN) = (Etot - kinetic_energy) / rho;
			}
}
void orc_post_interp_state(const qk_array4 *c, const qk_box *bx)
{
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(c, i, j, k, RHO), px = A4(c, i, j, k, MX), py = A4(c, i, j, k, MY), pz = A4(c, i, j, k, MZ);
				const double e = A4(c, i, j, k, EN);
				const double Eint = rho * e;
				const double kinetic_energy = (px * px + py * py + pz * pz) / (2.0 * rho);
				A4(c, i, j, k, EN) = Eint + kinetic_energy;
			}
}

/* amrex_avgdown: crse(i,j,k) = volfrac * sum of the ratio^3 fine cells, summed x fastest */
void orc_average_down(const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *cbx, const int ratio[3])
{
	const double volfrac = 1.0 / (double)(ratio[0] * ratio[1] * ratio[2]);
	for (int n = 0; n < ncomp; ++n)
		for (int k = cbx->lo[2]; k <= cbx->hi[2]; ++k)
			for (int j = cbx->lo[1]; j <= cbx->hi[1]; ++j)
				for (int i = cbx->lo[0]; i <= cbx->hi[0]; ++i) {
					double c = 0;
					for (int kr = 0; kr < ratio[2]; ++kr)
						for (int jr = 0; jr < ratio[1]; ++jr)
							for (int ir = 0; ir < ratio[0]; ++ir)
								c += A4(fine, i * ratio[0] + ir, j * ratio[1] + jr, k * ratio[2] + kr, n + fcomp);
					A4(crse, i, j, k, n + ccomp) = volfrac * c;
				}
}

/* ================================================================================================
 * Level driver (uniform single level, all boxes in this process)
 * ============================================================================================== */
struct orc_level {
	qk_level_desc d;
	qk_box *boxes;
	int32_t *bc_lo, *bc_hi;
	int nb;
	double **snew, **sold; /* per box storage, ncomp x grown box */
	double dt_prev;
	double t;
};

static int64_t box_len(const qk_box *b, int d) { return (int64_t)b->hi[d] - b->lo[d] + 1; }
static qk_box grow(qk_box b, int n)
{
	for (int d = 0; d < 3; ++d) {
		b.lo[d] -= n;
		b.hi[d] += n;
	}
	return b;
}
static qk_box grow_hi(qk_box b, int dir, int n)
{
	b.hi[dir] += n;
	return b;
}
static qk_array4 mk_a4(double *p, qk_box b, int ncomp)
{
	qk_array4 a;
	a.p = p;
	a.jstride = box_len(&b, 0);
	a.kstride = a.jstride * box_len(&b, 1);
	a.nstride = a.kstride * box_len(&b, 2);
	for (int d = 0; d < 3; ++d) {
		a.begin[d] = b.lo[d];
		a.end[d] = b.hi[d] + 1;
	}
	a.ncomp = ncomp;
	return a;
}
static qk_array4 alloc_a4(qk_box b, int ncomp)
{
	int64_t n = box_len(&b, 0) * box_len(&b, 1) * box_len(&b, 2) * ncomp;
	double *p = (double *)calloc((size_t)n, sizeof(double));
	return mk_a4(p, b, ncomp);
}
static qk_iarray4 alloc_ia4(qk_box b)
{
	qk_iarray4 a;
	int64_t n = box_len(&b, 0) * box_len(&b, 1) * box_len(&b, 2);
	a.p = (int32_t *)calloc((size_t)n, sizeof(int32_t));
	a.jstride = box_len(&b, 0);
	a.kstride = a.jstride * box_len(&b, 1);
	a.nstride = a.kstride * box_len(&b, 2);
	for (int d = 0; d < 3; ++d) {
		a.begin[d] = b.lo[d];
		a.end[d] = b.hi[d] + 1;
	}
	a.ncomp = 1;
	return a;
}

orc_level *orc_level_create(const qk_level_desc *desc)
{
	orc_level *L = (orc_level *)calloc(1, sizeof(orc_level));
	L->d = *desc;
	L->nb = desc->nboxes_global;
	L->boxes = (qk_box *)malloc(sizeof(qk_box) * L->nb);
	memcpy(L->boxes, desc->boxes_global, sizeof(qk_box) * L->nb);
	L->bc_lo = (int32_t *)malloc(sizeof(int32_t) * 3 * desc->ncomp);
	L->bc_hi = (int32_t *)malloc(sizeof(int32_t) * 3 * desc->ncomp);
	memcpy(L->bc_lo, desc->bc_lo, sizeof(int32_t) * 3 * desc->ncomp);
	memcpy(L->bc_hi, desc->bc_hi, sizeof(int32_t) * 3 * desc->ncomp);
	L->snew = (double **)malloc(sizeof(double *) * L->nb);
	L->sold = (double **)malloc(sizeof(double *) * L->nb);
	for (int b = 0; b < L->nb; ++b) {
		qk_box g = grow(L->boxes[b], desc->nghost);
		L->snew[b] = alloc_a4(g, desc->ncomp).p;
		L->sold[b] = alloc_a4(g, desc->ncomp).p;
	}
	L->dt_prev = 1.e100; /* dt_.resize(nlevs_max, 1.e100)  src/simulation.hpp:448 */
	L->t = 0.0;
	return L;
}
void orc_level_destroy(orc_level *L)
{
	if (!L)
		return;
	for (int b = 0; b < L->nb; ++b) {
		free(L->snew[b]);
		free(L->sold[b]);
	}
	free(L->snew);
	free(L->sold);
	free(L->boxes);
	free(L->bc_lo);
	free(L->bc_hi);
	free(L);
}
int orc_level_nboxes(const orc_level *L) { return L->nb; }
qk_box orc_level_box(const orc_level *L, int b) { return L->boxes[b]; }
qk_array4 orc_level_state(orc_level *L, int which, int b)
{
	return mk_a4(which == 0 ? L->snew[b] : L->sold[b], grow(L->boxes[b], L->d.nghost), L->d.ncomp);
}

/* fillBoundaryConditions, level 0  src/simulation.hpp:1752-1765:
 *  (1) FabArray::FillBoundary(periodicity): ghost cells that lie (after a periodic shift) inside
 *      another box's VALID region are copied from it (corners included, cross=false);
 *  (2) PhysBCFunct + amrex::FilccCell (extern/amrex/Src/Base/AMReX_FilCC_3D_C.H:38-41,66-75 ...):
 *      ghost cells outside the domain: per-axis mirror (reflect_even/odd) or clamp (foextrap),
 *      applied x then y then z over faces -> edges -> corners (AMReX_PhysBCFunct.H:406-470), i.e.
 *      the composition of the per-axis maps; sources are cells of the same FAB. */
void orc_fill_boundary(orc_level *L, qk_array4 *arrs, int scomp, int ncomp)
{
	const qk_box dom = L->d.domain;
	const int ng = L->d.nghost;
	/* (1) */
	for (int b = 0; b < L->nb; ++b) {
		const qk_box g = grow(L->boxes[b], ng);
		for (int s = 0; s < L->nb; ++s) {
			int smin[3], smax[3];
			for (int d = 0; d < 3; ++d) {
				smin[d] = L->d.periodic[d] ? -1 : 0;
				smax[d] = L->d.periodic[d] ? 1 : 0;
			}
			for (int sz = smin[2]; sz <= smax[2]; ++sz)
				for (int sy = smin[1]; sy <= smax[1]; ++sy)
					for (int sx = smin[0]; sx <= smax[0]; ++sx) {
						if (s == b && sx == 0 && sy == 0 && sz == 0)
							continue;
						const int sh[3] = {sx * (int)box_len(&dom, 0), sy * (int)box_len(&dom, 1), sz * (int)box_len(&dom, 2)};
						/* region (in dst index space) = g ∩ (box_s + sh) */
						qk_box r;
						int empty = 0;
						for (int d = 0; d < 3; ++d) {
							r.lo[d] = g.lo[d] > L->boxes[s].lo[d] + sh[d] ? g.lo[d] : L->boxes[s].lo[d] + sh[d];
							r.hi[d] = g.hi[d] < L->boxes[s].hi[d] + sh[d] ? g.hi[d] : L->boxes[s].hi[d] + sh[d];
							if (r.lo[d] > r.hi[d])
								empty = 1;
						}
						if (empty)
							continue;
						for (int n = scomp; n < scomp + ncomp; ++n)
							for (int k = r.lo[2]; k <= r.hi[2]; ++k)
								for (int j = r.lo[1]; j <= r.hi[1]; ++j)
									for (int i = r.lo[0]; i <= r.hi[0]; ++i)
										A4(&arrs[b], i, j, k, n) = A4(&arrs[s], i - sh[0], j - sh[1], k - sh[2], n);
					}
		}
	}
	/* (2) */
	for (int b = 0; b < L->nb; ++b) {
		const qk_box g = grow(L->boxes[b], ng);
		for (int n = scomp; n < scomp + ncomp; ++n)
			for (int k = g.lo[2]; k <= g.hi[2]; ++k)
				for (int j = g.lo[1]; j <= g.hi[1]; ++j)
					for (int i = g.lo[0]; i <= g.hi[0]; ++i) {
						int idx[3] = {i, j, k};
						int src[3] = {i, j, k};
						double sign = 1.0;
						int outside = 0, skip = 0;
						for (int d = 0; d < 3; ++d) {
							if (L->d.periodic[d])
								continue;
							if (idx[d] < dom.lo[d]) {
								outside = 1;
								const int bc = L->bc_lo[n * 3 + d];
								if (bc == QK_BC_REFLECT_EVEN || bc == QK_BC_REFLECT_ODD) {
									src[d] = 2 * dom.lo[d] - idx[d] - 1;
									if (bc == QK_BC_REFLECT_ODD)
										sign = -sign;
								} else if (bc == QK_BC_FOEXTRAP) {
									src[d] = dom.lo[d];
								} else {
									skip = 1;
								}
							} else if (idx[d] > dom.hi[d]) {
								outside = 1;
								const int bc = L->bc_hi[n * 3 + d];
								if (bc == QK_BC_REFLECT_EVEN || bc == QK_BC_REFLECT_ODD) {
									src[d] = 2 * dom.hi[d] - idx[d] + 1;
									if (bc == QK_BC_REFLECT_ODD)
										sign = -sign;
								} else if (bc == QK_BC_FOEXTRAP) {
									src[d] = dom.hi[d];
								} else {
									skip = 1;
								}
							}
						}
						if (!outside || skip)
							continue;
						const double v = A4(&arrs[b], src[0], src[1], src[2], n);
						A4(&arrs[b], i, j, k, n) = (sign < 0) ? -v : v;
					}
	}
}

static void free_a4(qk_array4 *a) { free(a->p); }

/* QuokkaSimulation::advanceHydroAtLevel  src/QuokkaSimulation.hpp:1032-1322 (uniform level: no
 * flux registers, no Strang sources, no tracers).  Operates state_old (ghost cells are filled in
 * place, as the reference does on state_old_cc_tmp) -> state_new. */
int orc_advance_hydro_level(orc_level *L, const qk_hydro_params *prm, double dt, double cfl, int64_t *bad1, int64_t *bad2)
{
	const int nb = L->nb, ng = L->d.nghost, nv = 6 + prm->nscalars, nc = L->d.ncomp;
	const double *dx = L->d.dx;
	int success = 1;
	if (bad1)
		*bad1 = 0;
	if (bad2)
		*bad2 = 0;

	qk_array4 *Uold = malloc(sizeof(qk_array4) * nb), *Unew = malloc(sizeof(qk_array4) * nb), *Uint = malloc(sizeof(qk_array4) * nb);
	qk_array4 *prim = malloc(sizeof(qk_array4) * nb), *rhs = malloc(sizeof(qk_array4) * nb);
	qk_array4 *chi[3], *lft[3], *rgt[3], *flx[3], *fvl[3], *fof[3], *fov[3], *frk[3], *avg[3];
	qk_iarray4 *redo = malloc(sizeof(qk_iarray4) * nb);
	for (int d = 0; d < 3; ++d) {
		chi[d] = malloc(sizeof(qk_array4) * nb);
		lft[d] = malloc(sizeof(qk_array4) * nb);
		rgt[d] = malloc(sizeof(qk_array4) * nb);
		flx[d] = malloc(sizeof(qk_array4) * nb);
		fvl[d] = malloc(sizeof(qk_array4) * nb);
		fof[d] = malloc(sizeof(qk_array4) * nb);
		fov[d] = malloc(sizeof(qk_array4) * nb);
		frk[d] = malloc(sizeof(qk_array4) * nb);
		avg[d] = malloc(sizeof(qk_array4) * nb);
	}
	for (int b = 0; b < nb; ++b) {
		const qk_box vb = L->boxes[b];
		Uold[b] = orc_level_state(L, 1, b);
		Unew[b] = orc_level_state(L, 0, b);
		Uint[b] = alloc_a4(grow(vb, ng), nc); /* state_inter_cc_, setVal(0) :1056-1057 */
		prim[b] = alloc_a4(grow(vb, ng), nv);
		rhs[b] = alloc_a4(vb, nv);
		redo[b] = alloc_ia4(grow(vb, 1));
		for (int d = 0; d < 3; ++d) {
			chi[d][b] = alloc_a4(grow(vb, 2), 1);
			lft[d][b] = alloc_a4(grow(grow_hi(vb, d, 1), 1), nv);
			rgt[d][b] = alloc_a4(grow(grow_hi(vb, d, 1), 1), nv);
			flx[d][b] = alloc_a4(grow_hi(vb, d, 1), nv);
			fvl[d][b] = alloc_a4(grow_hi(vb, d, 1), 1);
			fof[d][b] = alloc_a4(grow_hi(vb, d, 1), nv);
			fov[d][b] = alloc_a4(grow_hi(vb, d, 1), 1);
			frk[d][b] = alloc_a4(grow_hi(vb, d, 1), nv); /* flux_rk2, setVal(0) :1067-1068 */
			avg[d][b] = alloc_a4(grow(grow_hi(vb, d, 1), 2), 1); /* avgFaceVel ng=2 :1070-1071 */
		}
	}

	/* :1076 */
	orc_fill_boundary(L, Uold, 0, nc);

	/* computeFOHydroFluxes :1096 -> :1519-1568 */
	for (int b = 0; b < nb; ++b) {
		const qk_box vb = L->boxes[b], g4 = grow(vb, ng), g1 = grow(vb, 1);
		orc_conserved_to_primitive(prm, &Uold[b], &prim[b], &g4);
		for (int d = 0; d < 3; ++d) {
			const qk_box fb = grow_hi(vb, d, 1);
			orc_reconstruct_states(1, 0, d, &prim[b], &lft[d][b], &rgt[d][b], &g1, nv);
			orc_compute_fluxes(prm, QK_LLF, d, &fof[d][b], &fov[d][b], &lft[d][b], &rgt[d][b], &prim[b], &fb);
		}
	}

	for (int stage = 1; stage <= prm->integrator_order && success; ++stage) {
		qk_array4 *Uin = (stage == 1) ? Uold : Uint;
		qk_array4 *Uout = (stage == 1 && prm->integrator_order == 2) ? Uint : Unew;
		if (stage == 1 && prm->integrator_order == 1)
			Uout = Uint; /* forward Euler writes state_inter then copies :1286-1287 */
		if (stage == 2)
			orc_fill_boundary(L, Uint, 0, nc); /* :1204 */
		int64_t nbad_total = 0;
		/* computeHydroFluxes :1403-1490 */
		for (int b = 0; b < nb; ++b) {
			const qk_box vb = L->boxes[b], g4 = grow(vb, ng), g2 = grow(vb, 2), g1 = grow(vb, 1);
			orc_conserved_to_primitive(prm, &Uin[b], &prim[b], &g4);
			for (int d = 0; d < 3; ++d)
				orc_flattening_coefficients(prm, d, &prim[b], &chi[d][b], &g2);
			for (int d = 0; d < 3; ++d) {
				const qk_box fb = grow_hi(vb, d, 1);
				orc_reconstruct_states(prm->reconstruction_order, QK_MINMOD, d, &prim[b], &lft[d][b], &rgt[d][b], &g1, nv);
				orc_flatten_shocks(d, &prim[b], &chi[0][b], &chi[1][b], &chi[2][b], &lft[d][b], &rgt[d][b], &g1, nv);
				orc_compute_fluxes(prm, QK_HLLC, d, &flx[d][b], &fvl[d][b], &lft[d][b], &rgt[d][b], &prim[b], &fb);
				/* Saxpy :1105-1108 / :1219-1222 */
				orc_saxpy(&frk[d][b], 0.5, &flx[d][b], &fb, nv);
				orc_saxpy(&avg[d][b], 0.5, &fvl[d][b], &fb, 1);
			}
		}
		/* stage 1 uses (fluxArrays, faceVel); stage 2 uses (flux_rk2, avgFaceVel) :1114-1116, :1228-1230 */
		qk_array4 **F = (stage == 1) ? flx : frk;
		qk_array4 **V = (stage == 1) ? fvl : avg;
		for (int b = 0; b < nb; ++b) {
			const qk_box vb = L->boxes[b];
			memset(redo[b].p, 0, sizeof(int32_t) * (size_t)redo[b].nstride); /* redoFlag.setVal(none) */
			orc_rhs_from_fluxes(&rhs[b], &F[0][b], &F[1][b], &F[2][b], dx, &vb, nv);
			orc_add_internal_energy_pdv(prm, &rhs[b], &Uold[b], dx, &V[0][b], &V[1][b], &V[2][b], &redo[b], &vb);
			nbad_total += orc_predict_step(prm, &Uold[b], &Uout[b], &rhs[b], dt, nv, &redo[b], &vb);
		}
		if (stage == 1 && bad1)
			*bad1 = nbad_total;
		if (stage == 2 && bad2)
			*bad2 = nbad_total;
		if (nbad_total > 0) {
			/* redoFlag.FillBoundary(periodicity) :1157: exchange the 1-cell ghost layer of the flags */
			for (int b = 0; b < nb; ++b) {
				const qk_box g = grow(L->boxes[b], 1);
				const qk_box dom = L->d.domain;
				for (int s = 0; s < nb; ++s)
					for (int sz = -1; sz <= 1; ++sz)
						for (int sy = -1; sy <= 1; ++sy)
							for (int sx = -1; sx <= 1; ++sx) {
								if ((sx && !L->d.periodic[0]) || (sy && !L->d.periodic[1]) || (sz && !L->d.periodic[2]))
									continue;
								if (s == b && !sx && !sy && !sz)
									continue;
								const int sh[3] = {sx * (int)box_len(&dom, 0), sy * (int)box_len(&dom, 1), sz * (int)box_len(&dom, 2)};
								qk_box r;
								int empty = 0;
								for (int d = 0; d < 3; ++d) {
									r.lo[d] = g.lo[d] > L->boxes[s].lo[d] + sh[d] ? g.lo[d] : L->boxes[s].lo[d] + sh[d];
									r.hi[d] = g.hi[d] < L->boxes[s].hi[d] + sh[d] ? g.hi[d] : L->boxes[s].hi[d] + sh[d];
									if (r.lo[d] > r.hi[d])
										empty = 1;
								}
								if (empty)
									continue;
								for (int k = r.lo[2]; k <= r.hi[2]; ++k)
									for (int j = r.lo[1]; j <= r.hi[1]; ++j)
										for (int i = r.lo[0]; i <= r.hi[0]; ++i)
											IA4(&redo[b], i, j, k) = IA4(&redo[s], i - sh[0], j - sh[1], k - sh[2]);
							}
			}
			nbad_total = 0;
			for (int b = 0; b < nb; ++b) {
				const qk_box vb = L->boxes[b];
				for (int d = 0; d < 3; ++d) {
					orc_replace_fluxes(d, &F[d][b], &fof[d][b], &redo[b], &vb, nv);
					orc_replace_fluxes(d, &V[d][b], &fov[d][b], &redo[b], &vb, 1);
				}
			}
			for (int b = 0; b < nb; ++b) {
				const qk_box vb = L->boxes[b];
				orc_rhs_from_fluxes(&rhs[b], &F[0][b], &F[1][b], &F[2][b], dx, &vb, nv);
				orc_add_internal_energy_pdv(prm, &rhs[b], &Uold[b], dx, &V[0][b], &V[1][b], &V[2][b], &redo[b], &vb);
				nbad_total += orc_predict_step(prm, &Uold[b], &Uout[b], &rhs[b], dt, nv, &redo[b], &vb);
			}
			if (nbad_total > 0 && prm->abort_on_fofc_failure) {
				success = 0;
			}
		}
		if (success) {
			for (int b = 0; b < nb; ++b) {
				const qk_box vb = L->boxes[b];
				orc_enforce_limits(prm, &Uout[b], &vb);
				if (prm->use_dual_energy)
					orc_sync_dual_energy(prm, &Uout[b], &vb);
			}
		}
	}
	if (success && prm->integrator_order == 1) {
		for (int b = 0; b < nb; ++b) {
			const qk_box vb = L->boxes[b];
			for (int n = 0; n < nv; ++n)
				for (int k = vb.lo[2]; k <= vb.hi[2]; ++k)
					for (int j = vb.lo[1]; j <= vb.hi[1]; ++j)
						for (int i = vb.lo[0]; i <= vb.hi[0]; ++i)
							A4(&Unew[b], i, j, k, n) = A4(&Uint[b], i, j, k, n);
		}
	}
	/* isCflViolated :992-1013 */
	if (success) {
		double max_signal = -DBL_MAX;
		for (int b = 0; b < nb; ++b) {
			const qk_box vb = L->boxes[b];
			max_signal = dmax(max_signal, orc_max_signal_speed(prm, 1, &Unew[b], &vb));
		}
		const double dx_min = dmin(dmin(dx[0], dx[1]), dx[2]);
		const double dt_cfl = cfl * (dx_min / max_signal);
		if (dt > (1.1 * dt_cfl))
			success = 0;
	}

	for (int b = 0; b < nb; ++b) {
		free_a4(&Uint[b]);
		free_a4(&prim[b]);
		free_a4(&rhs[b]);
		free(redo[b].p);
		for (int d = 0; d < 3; ++d) {
			free_a4(&chi[d][b]);
			free_a4(&lft[d][b]);
			free_a4(&rgt[d][b]);
			free_a4(&flx[d][b]);
			free_a4(&fvl[d][b]);
			free_a4(&fof[d][b]);
			free_a4(&fov[d][b]);
			free_a4(&frk[d][b]);
			free_a4(&avg[d][b]);
		}
	}
	for (int d = 0; d < 3; ++d) {
		free(chi[d]);
		free(lft[d]);
		free(rgt[d]);
		free(flx[d]);
		free(fvl[d]);
		free(fof[d]);
		free(fov[d]);
		free(frk[d]);
		free(avg[d]);
	}
	free(Uold);
	free(Unew);
	free(Uint);
	free(prim);
	free(rhs);
	free(redo);
	return success;
}

/* AMRSimulation::computeTimestep, single level  src/simulation.hpp:703-818 */
static double compute_timestep_floor(orc_level *L, const qk_hydro_params *prm, double cfl, double t_now, double stop_time, double signal_floor);
double orc_compute_timestep(orc_level *L, const qk_hydro_params *prm, double cfl, double t_now, double stop_time)
{
	return compute_timestep_floor(L, prm, cfl, t_now, stop_time, 0.0);
}
/* radiation hydrodynamics: computeMaxSignalLocal (src/QuokkaSimulation.hpp:408-441) takes, per cell, std::max(c_hat / maxSubsteps_,
 * hydro signal speed) (RadSystem::ComputeMaxSignalSpeed = c_hat, src/radiation/radiation_system.hpp:617-624) */
double orc_compute_timestep_radhydro(orc_level *L, const qk_hydro_params *prm, const qk_rad_params *rp, int max_substeps, double cfl, double t_now,
				     double stop_time)
{
	return compute_timestep_floor(L, prm, cfl, t_now, stop_time, rp->c_hat / (double)max_substeps);
}
static double compute_timestep_floor(orc_level *L, const qk_hydro_params *prm, double cfl, double t_now, double stop_time, double signal_floor)
{
	double smax = signal_floor;
	for (int b = 0; b < L->nb; ++b) {
		qk_array4 s = orc_level_state(L, 0, b);
		smax = dmax(smax, orc_max_signal_speed(prm, 0, &s, &L->boxes[b]));
	}
	const double *dx = L->d.dx;
	const double dx_min = dmin(dmin(dx[0], dx[1]), dx[2]);
	const double hydro_dt = cfl * (dx_min / smax);
	double dt_tmp = dmin(hydro_dt, DBL_MAX); /* computeExtraPhysicsTimestep -> max() */
	dt_tmp = dmin(dt_tmp, 1.1 * L->dt_prev);
	double dt_0 = dt_tmp;
	dt_0 = dmin(dt_0, 1.0 * dt_tmp);
	dt_0 = dmin(dt_0, DBL_MAX); /* maxDt_ */
	if (t_now == 0.0)
		dt_0 = dmin(dt_0, DBL_MAX); /* initDt_ */
	const double eps = 1.e-3 * dt_0;
	if (t_now + dt_0 > stop_time - eps)
		dt_0 = stop_time - t_now;
	L->dt_prev = dt_0;
	return dt_0;
}

/* advanceSingleTimestepAtLevel + advanceHydroAtLevelWithRetries  src/QuokkaSimulation.hpp:653-707,885-990 */
int orc_step_with_retries(orc_level *L, const qk_hydro_params *prm, double dt, double cfl)
{
	const int nb = L->nb, ng = L->d.nghost, nc = L->d.ncomp, nv = 6 + prm->nscalars;
	/* std::swap(state_old, state_new) :671 */
	double **tmp = L->sold;
	L->sold = L->snew;
	L->snew = tmp;
	/* keep a pristine copy of state_old (the reference copies it into state_old_cc_tmp each retry :939-940) */
	double **save = malloc(sizeof(double *) * nb);
	int64_t *len = malloc(sizeof(int64_t) * nb);
	for (int b = 0; b < nb; ++b) {
		qk_box g = grow(L->boxes[b], ng);
		len[b] = box_len(&g, 0) * box_len(&g, 1) * box_len(&g, 2) * nc;
		save[b] = malloc(sizeof(double) * (size_t)len[b]);
		memcpy(save[b], L->sold[b], sizeof(double) * (size_t)len[b]);
	}
	int result = -1;
	for (int retry = 0; retry <= 6; ++retry) {
		const int nsub = 1 << retry;
		const double dt_step = dt / nsub;
		for (int b = 0; b < nb; ++b)
			memcpy(L->sold[b], save[b], sizeof(double) * (size_t)len[b]);
		int ok = 1;
		for (int sub = 0; sub < nsub; ++sub) {
			if (sub > 0) { /* amrex::Copy(state_old_cc_tmp, state_new, 0,0,ncompHydro_, nghost) :948 */
				for (int b = 0; b < nb; ++b) {
					qk_box g = grow(L->boxes[b], ng);
					int64_t per = box_len(&g, 0) * box_len(&g, 1) * box_len(&g, 2);
					memcpy(L->sold[b], L->snew[b], sizeof(double) * (size_t)(per * nv));
				}
			}
			ok = orc_advance_hydro_level(L, prm, dt_step, cfl, NULL, NULL);
			if (!ok)
				break;
		}
		if (ok) {
			result = retry;
			break;
		}
	}
	/* state_old keeps the pre-step state without filled ghosts, as in the reference */
	for (int b = 0; b < nb; ++b) {
		memcpy(L->sold[b], save[b], sizeof(double) * (size_t)len[b]);
		free(save[b]);
	}
	free(save);
	free(len);
	if (result >= 0)
		L->t += dt;
	return result;
}

/* ---- radiation transport of one level: the hyperbolic part of advanceRadiationSubstepAtLevel
 * (src/QuokkaSimulation.hpp:1721-1862: advanceRadiationForwardEuler then advanceRadiationMidpointRK2, no source terms,
 * no flux registers) with computeRadiationFluxes / fluxFunction<DIR> (:1903-1986) per box.  state_old -> state_new. */
static void rad_fluxes_of_box(const qk_rad_params *prm, const qk_array4 *U, const qk_box vb, int ng, qk_array4 flx[3])
{
	const int nh = 4 * prm->ngroups;
	const qk_box g4 = grow(vb, ng), g1 = grow(vb, 1);
	qk_array4 prim = alloc_a4(g4, nh);
	orc_rad_conserved_to_primitive(prm, U, &prim, &g4);
	for (int d = 0; d < 3; ++d) {
		const qk_box rb = grow_hi(g1, d, 1), fb = grow_hi(vb, d, 1);
		qk_array4 l = alloc_a4(rb, nh), r = alloc_a4(rb, nh);
		if (prm->reconstruction_order == 3)
			orc_reconstruct_states(3, 0, d, &prim, &l, &r, &g1, nh);
		else
			orc_reconstruct_states(prm->reconstruction_order, QK_MC, d, &prim, &l, &r, &rb, nh);
		flx[d] = alloc_a4(fb, nh);
		orc_rad_compute_fluxes(prm, d, &flx[d], NULL, &l, &r, U, &fb);
		free_a4(&l);
		free_a4(&r);
	}
	free_a4(&prim);
}

/* advanceRadiationForwardEuler :1791-1822 */
static void rad_stage1_level(orc_level *L, const qk_rad_params *prm, double dt, qk_array4 *Uold, qk_array4 *Unew)
{
	const int nb = L->nb, ng = L->d.nghost, nc = L->d.ncomp;
	const double *dx = L->d.dx;
	orc_fill_boundary(L, Uold, 0, nc); /* :1798 */
	qk_rad_params pl = *prm; /* ComputeCellOpticalDepth uses the level's cell sizes */
	for (int d = 0; d < 3; ++d)
		pl.cell_dx[d] = dx[d];
	for (int b = 0; b < nb; ++b) {
		qk_array4 f[3];
		rad_fluxes_of_box(&pl, &Uold[b], L->boxes[b], ng, f);
		orc_rad_predict_step(prm, &Uold[b], &Unew[b], &f[0], &f[1], &f[2], dt, dx, &L->boxes[b]);
		for (int d = 0; d < 3; ++d)
			free_a4(&f[d]);
	}
}
/* advanceRadiationMidpointRK2 :1824-1862 */
static void rad_stage2_level(orc_level *L, const qk_rad_params *prm, double dt, qk_array4 *Uold, qk_array4 *Unew)
{
	const int nb = L->nb, ng = L->d.nghost, nc = L->d.ncomp;
	const double *dx = L->d.dx;
	orc_fill_boundary(L, Unew, 0, nc); /* :1831 */
	/* stateInter and stateNew alias in the reference (:1840-1841): every box's fluxes are computed from the intermediate
	 * state before that box is overwritten, and ghost cells of other boxes are not touched by the update */
	qk_rad_params pl = *prm;
	for (int d = 0; d < 3; ++d)
		pl.cell_dx[d] = dx[d];
	for (int b = 0; b < nb; ++b) {
		qk_array4 fo[3], f[3];
		rad_fluxes_of_box(&pl, &Uold[b], L->boxes[b], ng, fo);
		rad_fluxes_of_box(&pl, &Unew[b], L->boxes[b], ng, f);
		orc_rad_add_fluxes_rk2(prm, &Unew[b], &Uold[b], &Unew[b], &fo[0], &fo[1], &fo[2], &f[0], &f[1], &f[2], dt, dx, &L->boxes[b]);
		for (int d = 0; d < 3; ++d) {
			free_a4(&fo[d]);
			free_a4(&f[d]);
		}
	}
}

void orc_rad_advance_level(orc_level *L, const qk_rad_params *prm, double dt)
{
	const int nb = L->nb;
	qk_array4 *Uold = malloc(sizeof(qk_array4) * nb), *Unew = malloc(sizeof(qk_array4) * nb);
	for (int b = 0; b < nb; ++b) {
		Uold[b] = orc_level_state(L, 1, b);
		Unew[b] = orc_level_state(L, 0, b);
	}
	rad_stage1_level(L, prm, dt, Uold, Unew);
	if (prm->integrator_order == 2)
		rad_stage2_level(L, prm, dt, Uold, Unew);
	free(Uold);
	free(Unew);
}

/* computeNumberOfRadiationSubsteps  src/QuokkaSimulation.hpp:397-406 */
int orc_rad_num_substeps(const orc_level *L, const qk_rad_params *prm, double rad_cfl, double dt_hydro)
{
	const double *dx = L->d.dx;
	const double dx_min = dmin(dmin(dx[0], dx[1]), dx[2]);
	const double dtrad_tmp = rad_cfl * (dx_min / prm->c_hat);
	return (int)ceil(dt_hydro / dtrad_tmp);
}

/* subcycleRadiationAtLevel  src/QuokkaSimulation.hpp:1577-1700 (hydro enabled, no constant dt, no flux registers): nsub IMEX PD-ARS
 * substeps, each = transport stage 1 (old -> new), source terms of stage 1 on new, transport stage 2, source terms of stage 2.
 * esrc: radEnergySource per box (SetRadEnergySource is evaluated at time + dt but is time-independent for the problems used
 * here) or NULL.  Returns nsub. */
int orc_rad_subcycle_level(orc_level *L, const qk_hydro_params *hp, const qk_rad_params *prm, const qk_rad_source_params *sp, const qk_array4 *esrc,
			   double dt_hydro, double rad_cfl, int64_t *counters)
{
	const int nb = L->nb, ns = prm->nstart, nh = 4 * prm->ngroups;
	const int nsub = orc_rad_num_substeps(L, prm, rad_cfl, dt_hydro);
	const double dt_radiation = dt_hydro / (double)nsub;
	qk_array4 *Uold = malloc(sizeof(qk_array4) * nb), *Unew = malloc(sizeof(qk_array4) * nb);
	for (int b = 0; b < nb; ++b) {
		Uold[b] = orc_level_state(L, 1, b);
		Unew[b] = orc_level_state(L, 0, b);
	}
	for (int i = 0; i < nsub; ++i) {
		if (i > 0) /* swapRadiationState: MultiFab::Copy(old, new, nstart, nstart, ncompHyperbolic, 0)  :1570-1574,1604-1610 */
			for (int b = 0; b < nb; ++b) {
				const qk_box *vb = &L->boxes[b];
				for (int n = ns; n < ns + nh; ++n)
					for (int k = vb->lo[2]; k <= vb->hi[2]; ++k)
						for (int j = vb->lo[1]; j <= vb->hi[1]; ++j)
							for (int ii = vb->lo[0]; ii <= vb->hi[0]; ++ii)
								A4(&Uold[b], ii, j, k, n) = A4(&Unew[b], ii, j, k, n);
			}
		rad_stage1_level(L, prm, dt_radiation, Uold, Unew);
		for (int b = 0; b < nb; ++b) /* :1628-1641 */
			orc_rad_add_source_terms(hp, prm, sp, &Unew[b], esrc ? &esrc[b] : NULL, &L->boxes[b], dt_radiation, 1, counters);
		rad_stage2_level(L, prm, dt_radiation, Uold, Unew);
		for (int b = 0; b < nb; ++b) /* :1649-1658 */
			orc_rad_add_source_terms(hp, prm, sp, &Unew[b], esrc ? &esrc[b] : NULL, &L->boxes[b], dt_radiation, 2, counters);
	}
	free(Uold);
	free(Unew);
	return nsub;
}

/* Problem setup of config C4: RadSystem<ShellProblem>::SetRadEnergySource, src/problems/RadhydroShell/test_radhydro_shell.cpp:96-125
 * (simulate_full_box = true): a Gaussian point-like source at the centre of the box.  Written in C with the same libm calls
 * (std::pow(x, 2), std::exp) so that the array equals the reference's bit for bit.  out: one component on box bx. */
void orc_shell_rad_energy_source(const qk_array4 *out, const qk_box *bx, const double dx[3], const double prob_lo[3], const double prob_hi[3])
{
	const double c = 2.99792458e10, Msun = 2.0e33, parsec_in_cm = 3.086e18;
	const double specific_luminosity = 2000., GMC_mass = 1.0e6 * Msun, epsilon = 0.5;
	const double L_star = (epsilon * GMC_mass) * specific_luminosity;
	const double r_0 = 5.0 * parsec_in_cm, sigma_star = 0.3 * r_0;
	const double x0 = prob_lo[0] + 0.5 * (prob_hi[0] - prob_lo[0]);
	const double y0 = prob_lo[1] + 0.5 * (prob_hi[1] - prob_lo[1]);
	const double z0 = prob_lo[2] + 0.5 * (prob_hi[2] - prob_lo[2]);
	const double pi = 3.14159265358979323846; /* M_PI */
	const double source_norm = (1.0 / c) * L_star / pow(2.0 * pi * sigma_star * sigma_star, 1.5);
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double x = prob_lo[0] + (i + 0.5) * dx[0];
				const double y = prob_lo[1] + (j + 0.5) * dx[1];
				const double z = prob_lo[2] + (k + 0.5) * dx[2];
				const double r = sqrt(pow(x - x0, 2) + pow(y - y0, 2) + pow(z - z0, 2));
				A4(out, i, j, k, 0) = source_norm * exp(-(r * r) / (2.0 * sigma_star * sigma_star));
			}
}

/* swap state_new <-> state_old (the radiation subcycle copies new -> old, :1597-1600; a swap has the same effect for the
 * components the transport overwrites and is what the tests drive) */
void orc_level_swap(orc_level *L)
{
	double **t = L->snew;
	L->snew = L->sold;
	L->sold = t;
}

/* ---- time interpolation, regrid tagging, FixupState ------------------------------------------------------- */
/* amrex::FillPatcher::fill  extern/amrex/Src/AmrCore/AMReX_FillPatcher.H:340-387 */
int orc_time_interp(const qk_array4 *dst, int dcomp, const qk_array4 *src0, const qk_array4 *src1, int scomp, int ncomp, const qk_box *region, double t0,
		    double t1, double time)
{
	int idata = 0;
	if (src1) {
		const double teps = fabs(t1 - t0) * 1.e-3;
		if (time > t0 - teps && time < t0 + teps)
			idata = 0;
		else if (time > t1 - teps && time < t1 + teps)
			idata = 1;
		else
			idata = 2;
	}
	const double alpha = (idata == 2) ? (t1 - time) / (t1 - t0) : 0.0;
	const double beta = (idata == 2) ? (time - t0) / (t1 - t0) : 0.0;
	for (int n = 0; n < ncomp; ++n)
		for (int k = region->lo[2]; k <= region->hi[2]; ++k)
			for (int j = region->lo[1]; j <= region->hi[1]; ++j)
				for (int i = region->lo[0]; i <= region->hi[0]; ++i) {
					if (idata == 0)
						A4(dst, i, j, k, dcomp + n) = A4(src0, i, j, k, scomp + n);
					else if (idata == 1)
						A4(dst, i, j, k, dcomp + n) = A4(src1, i, j, k, scomp + n);
					else
						A4(dst, i, j, k, dcomp + n) = alpha * A4(src0, i, j, k, scomp + n) + beta * A4(src1, i, j, k, scomp + n);
				}
	return idata;
}

static double max2(double a, double b) { return (a < b) ? b : a; } /* std::max */

/* QuokkaSimulation<SedovProblem>::ErrorEst  src/problems/HydroBlast3D/test_hydro3d_blast.cpp:118-151 */
void orc_tag_pressure_gradient(const qk_hydro_params *prm, const qk_array4 *cons, char *tags, const qk_box *bx, double eta_threshold, double P_min)
{
	const int nx = bx->hi[0] - bx->lo[0] + 1, ny = bx->hi[1] - bx->lo[1] + 1;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double P = cons_pressure(prm, cons, i, j, k);
				const double P_xplus = cons_pressure(prm, cons, i + 1, j, k), P_xminus = cons_pressure(prm, cons, i - 1, j, k);
				const double P_yplus = cons_pressure(prm, cons, i, j + 1, k), P_yminus = cons_pressure(prm, cons, i, j - 1, k);
				const double P_zplus = cons_pressure(prm, cons, i, j, k + 1), P_zminus = cons_pressure(prm, cons, i, j, k - 1);
				const double del_x = max2(fabs(P_xplus - P), fabs(P - P_xminus));
				const double del_y = max2(fabs(P_yplus - P), fabs(P - P_yminus));
				const double del_z = max2(fabs(P_zplus - P), fabs(P - P_zminus));
				const double gradient_indicator = max2(max2(del_x, del_y), del_z) / P;
				if ((gradient_indicator > eta_threshold) && (P > P_min))
					tags[(i - bx->lo[0]) + (int64_t)nx * ((j - bx->lo[1]) + (int64_t)ny * (k - bx->lo[2]))] = 2; /* TagBox::SET */
			}
}

/* QuokkaSimulation<ShocktubeProblem>::ErrorEst  src/problems/HydroShocktube/test_hydro_shocktube.cpp:146-171 */
void orc_tag_gradient_x(const qk_array4 *state, int comp, char *tags, const qk_box *bx, double dx, double eta_threshold, double q_min)
{
	const int nx = bx->hi[0] - bx->lo[0] + 1, ny = bx->hi[1] - bx->lo[1] + 1;
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				const double rho = A4(state, i, j, k, comp);
				const double del_x = (A4(state, i + 1, j, k, comp) - A4(state, i - 1, j, k, comp)) / (2.0 * dx);
				const double gradient_indicator = sqrt(del_x * del_x) / rho;
				if (gradient_indicator > eta_threshold && rho >= q_min)
					tags[(i - bx->lo[0]) + (int64_t)nx * ((j - bx->lo[1]) + (int64_t)ny * (k - bx->lo[2]))] = 2; /* TagBox::SET */
			}
}

/* QuokkaSimulation::FixupState  src/QuokkaSimulation.hpp:761-770 */
void orc_fixup_state(const qk_hydro_params *prm, const qk_array4 *state, const qk_box *bx)
{
	orc_enforce_limits(prm, state, bx);
	orc_sync_dual_energy(prm, state, bx);
}
