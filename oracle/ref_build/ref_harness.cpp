// oracle/ref_build/ref_harness.cpp -- TEST INFRASTRUCTURE (written for this repo, not reference source).
//
// extern "C" wrappers that call the REFERENCE's own operator templates (HydroSystem<>,
// HyperbolicSystem<> from /root/reference/src, compiled where they lie) on caller-provided host
// arrays, one box at a time.  Used only to pin oracle/quokka_oracle.c (tests/test_oracle_vs_ref.py)
// and to generate tests/golden/*.  Built into oracle/_ref/libquokka_ref.so by oracle/ref_build/Makefile.
//
// The reference's traits are compile-time, so a small set of problem types is instantiated:
//   problem 0: Sedov-like   gamma=1.4, mu=m_u, reconstruct_eint=false, no scalars   (HydroBlast3D traits)
//   problem 1: Sod-like     gamma=1.4, mu=m_u, reconstruct_eint=true,  no scalars   (HydroShocktube traits)
//   problem 2: scalars      gamma=5/3, mu=m_u, reconstruct_eint=true,  3 passive scalars of which 2 mass scalars
//   problem 3: isothermal   gamma=1, cs_isothermal=1.3, reconstruct_eint=false, 1 passive scalar   (the is_eos_isothermal() branches,
//                           src/hydro/hydro_system.hpp:133; BinaryOrbitCIC / StarCluster / RadForce use this EOS)
//   problem 4: isothermal   gamma=1, cs_isothermal=0.7, reconstruct_eint=true, 3 passive scalars of which 2 mass scalars
#include <cstdint>
#include <cstring>

#include "AMReX.H"
#include "AMReX_BoxArray.H"
#include "AMReX_DistributionMapping.H"
#include "AMReX_MultiFab.H"
#include "AMReX_ParmParse.H"
#include "AMReX_iMultiFab.H"
#include "AMReX_MFInterpolater.H"
#include "AMReX_MultiFabUtil.H"
#include "AMReX_FillPatchUtil.H"
#include "AMReX_PhysBCFunct.H"

#include "hydro/hydro_system.hpp"
#include "hyperbolic_system.hpp"
#include "radiation/radiation_system.hpp"
#include "QuokkaSimulation.hpp" // only QuokkaSimulation<P>::PreInterpState / PostInterpState (static members) are instantiated

#include "../../include/quokka_b200.h"

struct P0 {
};
struct P1 {
};
struct P2 {
};
struct P3 {
};
struct P4 {
};

template <> struct quokka::EOS_Traits<P0> {
	static constexpr double gamma = 1.4;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct HydroSystem_Traits<P0> {
	static constexpr bool reconstruct_eint = false;
};
template <> struct Physics_Traits<P0> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

template <> struct quokka::EOS_Traits<P1> {
	static constexpr double gamma = 1.4;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<P1> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

template <> struct quokka::EOS_Traits<P2> {
	static constexpr double gamma = 5. / 3.;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<P2> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 2;
	static constexpr int numPassiveScalars = numMassScalars + 1;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

template <> struct quokka::EOS_Traits<P3> {
	static constexpr double gamma = 1.0;
	static constexpr double cs_isothermal = 1.3;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct HydroSystem_Traits<P3> {
	static constexpr bool reconstruct_eint = false;
};
template <> struct Physics_Traits<P3> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 1;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

template <> struct quokka::EOS_Traits<P4> {
	static constexpr double gamma = 1.0;
	static constexpr double cs_isothermal = 0.7;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<P4> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 2;
	static constexpr int numPassiveScalars = numMassScalars + 1;
	static constexpr bool is_radiation_enabled = false;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};

// radiation problem types (one photon group; hydro + radiation variables, radFirstIndex = 6)
//   R0: c = c_hat = 1, Erad_floor = 0 (dimensionless, RadStreaming-like traits)
//   R1: c = c_cgs, c_hat = c/30 (reduced speed of light), Erad_floor = 1e-12
struct R0 {
};
struct R1 {
};
template <> struct quokka::EOS_Traits<R0> {
	static constexpr double gamma = 5. / 3.;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<R0> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = true;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};
template <> struct RadSystem_Traits<R0> {
	static constexpr double c_light = 1.0;
	static constexpr double c_hat = 1.0;
	static constexpr double radiation_constant = 1.0;
	static constexpr double Erad_floor = 0.;
	static constexpr int beta_order = 0;
};
template <> struct quokka::EOS_Traits<R1> {
	static constexpr double gamma = 5. / 3.;
	static constexpr double mean_molecular_weight = C::m_u;
	static constexpr double boltzmann_constant = C::k_B;
};
template <> struct Physics_Traits<R1> {
	static constexpr bool is_hydro_enabled = true;
	static constexpr int numMassScalars = 0;
	static constexpr int numPassiveScalars = numMassScalars + 0;
	static constexpr bool is_radiation_enabled = true;
	static constexpr bool is_mhd_enabled = false;
	static constexpr int nGroups = 1;
};
template <> struct RadSystem_Traits<R1> {
	static constexpr double c_light = c_light_cgs_;
	static constexpr double c_hat = c_light_cgs_ / 30.0;
	static constexpr double radiation_constant = radiation_constant_cgs_;
	static constexpr double Erad_floor = 1.0e-12;
	static constexpr int beta_order = 0;
};

// problem types for the matter-radiation source terms (RadSystem<P>::AddSourceTermsSingleGroup), one photon group:
//   R2: RadhydroShell traits (src/problems/RadhydroShell/test_radhydro_shell.cpp:39-93,127-135): cgs, gamma = 5/3,
//       mu = 2.2 m_H, c_hat = 860 * 2e5 cm/s, kappa_P = kappa_E = kappa_F = 20, beta_order = 1
//   R3: dimensionless (k_B = 1, mu = 1, a_rad = 1, c = 10, c_hat = 5), kappa_P = 1, kappa_E = 1.5, kappa_F = 2, beta_order = 2
//       (kappa_F != kappa_E: the 3x3 solve and the (kappa_F - kappa_E) Planck term)
//   R4: as R3 with kappa_E = kappa_F = kappa_P = 3 and beta_order = 0 (no work term, no O(beta) flux terms)
//   R5: as R3 with kappa_P = 0.5, kappa_E = kappa_F = 0.25, beta_order = 3, Erad_floor = 1e-6
//   R6: gamma = 1 (isothermal branch: flux update only), c = c_hat = 1, kappa = 2
//   R7: as R3 with kappa_P = kappa_E = 0 (tau = 0: J11 = -inf, the F_D + R residual, kappaPoverE = 1), kappa_F = 0.3, beta_order 1
#define QK_RS_TRAITS(NAME, GAMMA, MU, KB, CL, CH, AR, FLOOR, BETA)                                                                                   \
	struct NAME {                                                                                                                                \
	};                                                                                                                                           \
	template <> struct quokka::EOS_Traits<NAME> {                                                                                                \
		static constexpr double gamma = GAMMA;                                                                                               \
		static constexpr double mean_molecular_weight = MU;                                                                                  \
		static constexpr double boltzmann_constant = KB;                                                                                     \
		static constexpr double cs_isothermal = 1.0;                                                                                         \
	};                                                                                                                                           \
	template <> struct Physics_Traits<NAME> {                                                                                                    \
		static constexpr bool is_hydro_enabled = true;                                                                                       \
		static constexpr int numMassScalars = 0;                                                                                             \
		static constexpr int numPassiveScalars = numMassScalars + 0;                                                                         \
		static constexpr bool is_radiation_enabled = true;                                                                                   \
		static constexpr bool is_mhd_enabled = false;                                                                                        \
		static constexpr int nGroups = 1;                                                                                                    \
	};                                                                                                                                           \
	template <> struct RadSystem_Traits<NAME> {                                                                                                  \
		static constexpr double c_light = CL;                                                                                                \
		static constexpr double c_hat = CH;                                                                                                  \
		static constexpr double radiation_constant = AR;                                                                                     \
		static constexpr double Erad_floor = FLOOR;                                                                                          \
		static constexpr int beta_order = BETA;                                                                                              \
	};
#define QK_RS_OPACITY(NAME, KP, KE, KF)                                                                                                              \
	template <> AMREX_GPU_HOST_DEVICE auto RadSystem<NAME>::ComputePlanckOpacity(const double /*rho*/, const double /*T*/) -> amrex::Real       \
	{                                                                                                                                            \
		return KP;                                                                                                                           \
	}                                                                                                                                            \
	template <> AMREX_GPU_HOST_DEVICE auto RadSystem<NAME>::ComputeEnergyMeanOpacity(const double /*rho*/, const double /*T*/) -> amrex::Real   \
	{                                                                                                                                            \
		return KE;                                                                                                                           \
	}                                                                                                                                            \
	template <> AMREX_GPU_HOST_DEVICE auto RadSystem<NAME>::ComputeFluxMeanOpacity(const double /*rho*/, const double /*T*/) -> amrex::Real     \
	{                                                                                                                                            \
		return KF;                                                                                                                           \
	}
QK_RS_TRAITS(R2, 5. / 3., 2.2 * C::m_u, C::k_B, c_light_cgs_, 860. * 2.0e5, radiation_constant_cgs_, 0., 1)
QK_RS_TRAITS(R3, 5. / 3., 1.0, 1.0, 10.0, 5.0, 1.0, 0., 2)
QK_RS_TRAITS(R4, 1.4, 1.0, 1.0, 10.0, 5.0, 1.0, 0., 0)
QK_RS_TRAITS(R5, 5. / 3., 1.0, 1.0, 10.0, 5.0, 1.0, 1.0e-6, 3)
QK_RS_TRAITS(R6, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0., 1)
QK_RS_TRAITS(R7, 5. / 3., 1.0, 1.0, 10.0, 5.0, 1.0, 0., 1)
QK_RS_OPACITY(R7, 0.0, 0.0, 0.3)
QK_RS_OPACITY(R2, 20.0, 20.0, 20.0)
QK_RS_OPACITY(R3, 1.0, 1.5, 2.0)
QK_RS_OPACITY(R4, 3.0, 3.0, 3.0)
QK_RS_OPACITY(R5, 0.5, 0.25, 0.25)
QK_RS_OPACITY(R6, 2.0, 2.0, 2.0)

namespace
{
bool g_init = false;

void ensure_init()
{
	if (g_init) {
		return;
	}
	int argc = 1;
	static char arg0[] = "ref_harness";
	static char *argv_s[] = {arg0, nullptr};
	char **argv = argv_s;
	amrex::Initialize(argc, argv, false, MPI_COMM_WORLD, []() {
		amrex::ParmParse pp("amrex");
		pp.add("verbose", 0);
		pp.add("signal_handling", 0);
	});
	g_init = true;
}

template <typename P> void set_eos()
{
	// as QuokkaSimulation's constructor does (src/QuokkaSimulation.hpp:160-167)
	init_extern_parameters();
	eos_rp::eos_gamma = quokka::EOS_Traits<P>::gamma;
	amrex::Real small_temp = 1e-10;
	amrex::Real small_dens = 1e-100;
	eos_init(small_temp, small_dens);
}

amrex::Box to_box(const qk_box *b) { return amrex::Box(amrex::IntVect(b->lo[0], b->lo[1], b->lo[2]), amrex::IntVect(b->hi[0], b->hi[1], b->hi[2])); }

// MultiFab with one box; `nodal` = -1 (cell-centred) or the nodal direction
amrex::MultiFab make_mf(const qk_box *valid, int nodal, int ncomp, int ng)
{
	amrex::Box bx = to_box(valid);
	amrex::BoxArray ba(bx);
	if (nodal >= 0) {
		ba = amrex::convert(ba, amrex::IntVect::TheDimensionVector(nodal));
	}
	amrex::DistributionMapping dm(ba);
	amrex::MultiFab mf(ba, dm, ncomp, ng);
	mf.setVal(0.0);
	return mf;
}

void copy_in(amrex::MultiFab &mf, const qk_array4 *a)
{
	auto arr = mf.array(0);
	for (int n = 0; n < arr.ncomp && n < a->ncomp; ++n)
		for (int k = arr.begin.z; k < arr.end.z; ++k)
			for (int j = arr.begin.y; j < arr.end.y; ++j)
				for (int i = arr.begin.x; i < arr.end.x; ++i) {
					if (i < a->begin[0] || i >= a->end[0] || j < a->begin[1] || j >= a->end[1] || k < a->begin[2] || k >= a->end[2])
						continue;
					arr(i, j, k, n) = a->p[(i - a->begin[0]) + (j - a->begin[1]) * a->jstride + (k - a->begin[2]) * a->kstride + n * a->nstride];
				}
}
void copy_out(amrex::MultiFab const &mf, const qk_array4 *a)
{
	auto arr = mf.const_array(0);
	for (int n = 0; n < arr.ncomp && n < a->ncomp; ++n)
		for (int k = arr.begin.z; k < arr.end.z; ++k)
			for (int j = arr.begin.y; j < arr.end.y; ++j)
				for (int i = arr.begin.x; i < arr.end.x; ++i) {
					if (i < a->begin[0] || i >= a->end[0] || j < a->begin[1] || j >= a->end[1] || k < a->begin[2] || k >= a->end[2])
						continue;
					a->p[(i - a->begin[0]) + (j - a->begin[1]) * a->jstride + (k - a->begin[2]) * a->kstride + n * a->nstride] = arr(i, j, k, n);
				}
}
void icopy_in(amrex::iMultiFab &mf, const qk_iarray4 *a)
{
	auto arr = mf.array(0);
	for (int k = arr.begin.z; k < arr.end.z; ++k)
		for (int j = arr.begin.y; j < arr.end.y; ++j)
			for (int i = arr.begin.x; i < arr.end.x; ++i) {
				if (i < a->begin[0] || i >= a->end[0] || j < a->begin[1] || j >= a->end[1] || k < a->begin[2] || k >= a->end[2])
					continue;
				arr(i, j, k) = a->p[(i - a->begin[0]) + (j - a->begin[1]) * a->jstride + (k - a->begin[2]) * a->kstride];
			}
}
void icopy_out(amrex::iMultiFab const &mf, const qk_iarray4 *a)
{
	auto arr = mf.const_array(0);
	for (int k = arr.begin.z; k < arr.end.z; ++k)
		for (int j = arr.begin.y; j < arr.end.y; ++j)
			for (int i = arr.begin.x; i < arr.end.x; ++i) {
				if (i < a->begin[0] || i >= a->end[0] || j < a->begin[1] || j >= a->end[1] || k < a->begin[2] || k >= a->end[2])
					continue;
				a->p[(i - a->begin[0]) + (j - a->begin[1]) * a->jstride + (k - a->begin[2]) * a->kstride] = arr(i, j, k);
			}
}

template <typename P> void cons_to_prim(const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim, int ng)
{
	set_eos<P>();
	const int nv = HydroSystem<P>::nvar_;
	auto c = make_mf(valid, -1, nv, ng);
	auto q = make_mf(valid, -1, nv, ng);
	copy_in(c, cons);
	HydroSystem<P>::ConservedToPrimitive(c, q, ng);
	copy_out(q, prim);
}

template <typename P, FluxDir DIR> void flat_coefs(const qk_box *valid, const qk_array4 *prim, const qk_array4 *chi, int ngprim, int ng)
{
	set_eos<P>();
	const int nv = HydroSystem<P>::nvar_;
	auto q = make_mf(valid, -1, nv, ngprim);
	// the reference launches over primVar_mf's boxes grown by `ng` and indexes chi with the same
	// box index, so give chi the same ghost width as the launch
	auto x = make_mf(valid, -1, 1, ng);
	copy_in(q, prim);
	HydroSystem<P>::template ComputeFlatteningCoefficients<DIR>(q, x, ng);
	copy_out(x, chi);
}

template <typename P, FluxDir DIR>
void reconstruct(int order, int limiter, const qk_box *valid, const qk_array4 *q_in, const qk_array4 *left, const qk_array4 *right, int ngq, int ng,
		 int nvars)
{
	auto q = make_mf(valid, -1, nvars, ngq);
	auto l = make_mf(valid, static_cast<int>(DIR), nvars, ng);
	auto r = make_mf(valid, static_cast<int>(DIR), nvars, ng);
	copy_in(q, q_in);
	copy_in(l, left);
	copy_in(r, right);
	if (order == 3) {
		HyperbolicSystem<P>::template ReconstructStatesPPM<DIR>(q, l, r, ng, nvars);
	} else if (order == 2 && limiter == QK_MINMOD) {
		HyperbolicSystem<P>::template ReconstructStatesPLM<DIR, SlopeLimiter::minmod>(q, l, r, ng, nvars);
	} else if (order == 2) {
		HyperbolicSystem<P>::template ReconstructStatesPLM<DIR, SlopeLimiter::MC>(q, l, r, ng, nvars);
	} else {
		HyperbolicSystem<P>::template ReconstructStatesConstant<DIR>(q, l, r, ng, nvars);
	}
	copy_out(l, left);
	copy_out(r, right);
}

template <typename P, FluxDir DIR>
void flatten(const qk_box *valid, const qk_array4 *q_in, const qk_array4 *c1, const qk_array4 *c2, const qk_array4 *c3, const qk_array4 *left,
	     const qk_array4 *right, int ngq, int ng, int nvars)
{
	auto q = make_mf(valid, -1, nvars, ngq);
	auto x1 = make_mf(valid, -1, 1, 2);
	auto x2 = make_mf(valid, -1, 1, 2);
	auto x3 = make_mf(valid, -1, 1, 2);
	auto l = make_mf(valid, static_cast<int>(DIR), nvars, ng);
	auto r = make_mf(valid, static_cast<int>(DIR), nvars, ng);
	copy_in(q, q_in);
	copy_in(x1, c1);
	copy_in(x2, c2);
	copy_in(x3, c3);
	copy_in(l, left);
	copy_in(r, right);
	HydroSystem<P>::template FlattenShocks<DIR>(q, x1, x2, x3, l, r, ng, nvars);
	copy_out(l, left);
	copy_out(r, right);
}

template <typename P, RiemannSolver RS, FluxDir DIR>
void fluxes(const qk_box *valid, const qk_array4 *flux, const qk_array4 *fvel, const qk_array4 *left, const qk_array4 *right, const qk_array4 *prim,
	    int ngprim, double K_visc)
{
	set_eos<P>();
	const int nv = HydroSystem<P>::nvar_;
	auto q = make_mf(valid, -1, nv, ngprim);
	auto l = make_mf(valid, static_cast<int>(DIR), nv, 1);
	auto r = make_mf(valid, static_cast<int>(DIR), nv, 1);
	auto f = make_mf(valid, static_cast<int>(DIR), nv, 0);
	auto v = make_mf(valid, static_cast<int>(DIR), 1, 0);
	copy_in(q, prim);
	copy_in(l, left);
	copy_in(r, right);
	HydroSystem<P>::template ComputeFluxes<RS, DIR>(f, v, l, r, q, K_visc);
	copy_out(f, flux);
	copy_out(v, fvel);
}

template <typename P>
void update_ops(int op, const qk_box *valid, const qk_array4 *a0, const qk_array4 *a1, const qk_array4 *a2, const qk_array4 *fx, const qk_array4 *fy,
		const qk_array4 *fz, const qk_iarray4 *redo, const double *dx3, double dt, double dfloor, double tfloor, double *scalar_out)
{
	set_eos<P>();
	const int nv = HydroSystem<P>::nvar_;
	amrex::GpuArray<amrex::Real, 3> dx{dx3 ? dx3[0] : 1.0, dx3 ? dx3[1] : 1.0, dx3 ? dx3[2] : 1.0};
	if (op == 0) { // ComputeRhsFromFluxes: a0=rhs
		auto rhs = make_mf(valid, -1, nv, 0);
		std::array<amrex::MultiFab, 3> F{make_mf(valid, 0, nv, 0), make_mf(valid, 1, nv, 0), make_mf(valid, 2, nv, 0)};
		copy_in(F[0], fx);
		copy_in(F[1], fy);
		copy_in(F[2], fz);
		HydroSystem<P>::ComputeRhsFromFluxes(rhs, F, dx, nv);
		copy_out(rhs, a0);
	} else if (op == 1) { // AddInternalEnergyPdV: a0=rhs (in/out), a1=cons (ng>=1), fx..=facevel
		auto rhs = make_mf(valid, -1, nv, 0);
		auto cons = make_mf(valid, -1, nv, 4);
		std::array<amrex::MultiFab, 3> V{make_mf(valid, 0, 1, 0), make_mf(valid, 1, 1, 0), make_mf(valid, 2, 1, 0)};
		amrex::iMultiFab rf(amrex::BoxArray(to_box(valid)), amrex::DistributionMapping(amrex::BoxArray(to_box(valid))), 1, 1);
		rf.setVal(0);
		copy_in(rhs, a0);
		copy_in(cons, a1);
		copy_in(V[0], fx);
		copy_in(V[1], fy);
		copy_in(V[2], fz);
		icopy_in(rf, redo);
		HydroSystem<P>::AddInternalEnergyPdV(rhs, cons, dx, V, rf);
		copy_out(rhs, a0);
	} else if (op == 2) { // PredictStep: a0=old, a1=new, a2=rhs
		auto uo = make_mf(valid, -1, nv, 0);
		auto un = make_mf(valid, -1, nv, 0);
		auto rhs = make_mf(valid, -1, nv, 0);
		amrex::iMultiFab rf(amrex::BoxArray(to_box(valid)), amrex::DistributionMapping(amrex::BoxArray(to_box(valid))), 1, 1);
		rf.setVal(0);
		copy_in(uo, a0);
		copy_in(rhs, a2);
		HydroSystem<P>::PredictStep(uo, un, rhs, dt, nv, rf);
		copy_out(un, a1);
		icopy_out(rf, redo);
		if (scalar_out) {
			*scalar_out = static_cast<double>(rf.sum(0));
		}
	} else if (op == 3) { // EnforceLimits: a0=state in/out
		auto s = make_mf(valid, -1, nv, 0);
		copy_in(s, a0);
		HydroSystem<P>::EnforceLimits(dfloor, tfloor, s);
		copy_out(s, a0);
	} else if (op == 4) { // SyncDualEnergy
		auto s = make_mf(valid, -1, nv, 0);
		copy_in(s, a0);
		HydroSystem<P>::SyncDualEnergy(s);
		copy_out(s, a0);
	} else if (op == 5) { // ComputeMaxSignalSpeed + norminf
		auto s = make_mf(valid, -1, nv, 0);
		auto m = make_mf(valid, -1, 1, 0);
		copy_in(s, a0);
		HydroSystem<P>::ComputeMaxSignalSpeed(s.const_array(0), m.array(0), to_box(valid));
		*scalar_out = m.norminf();
	} else if (op == 6) { // maxSignalSpeedLocal
		auto s = make_mf(valid, -1, nv, 0);
		copy_in(s, a0);
		*scalar_out = HydroSystem<P>::maxSignalSpeedLocal(s);
	}
}

// ---- radiation: RadSystem<P> static functions take Array4s of one box ----
template <typename P> void rad_cons_to_prim(const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim, int ng)
{
	auto c = make_mf(valid, -1, RadSystem<P>::nvar_, ng);
	auto q = make_mf(valid, -1, RadSystem<P>::nvarHyperbolic_, ng);
	copy_in(c, cons);
	RadSystem<P>::ConservedToPrimitive(c.const_array(0), q.array(0), amrex::grow(to_box(valid), ng));
	copy_out(q, prim);
}
template <typename P, FluxDir DIR>
void rad_fluxes(const qk_box *valid, const qk_array4 *flux, const qk_array4 *fdiff, const qk_array4 *left, const qk_array4 *right, const qk_array4 *cons,
		int ngcons)
{
	const int nh = RadSystem<P>::nvarHyperbolic_;
	auto c = make_mf(valid, -1, RadSystem<P>::nvar_, ngcons);
	auto l = make_mf(valid, static_cast<int>(DIR), nh, 1);
	auto r = make_mf(valid, static_cast<int>(DIR), nh, 1);
	auto f = make_mf(valid, static_cast<int>(DIR), nh, 0);
	auto fd = make_mf(valid, static_cast<int>(DIR), nh, 0);
	copy_in(c, cons);
	copy_in(l, left);
	copy_in(r, right);
	amrex::GpuArray<amrex::Real, 3> dx{1.0, 1.0, 1.0};
	RadSystem<P>::template ComputeFluxes<DIR>(f.array(0), fd.array(0), l.const_array(0), r.const_array(0),
						  amrex::surroundingNodes(to_box(valid), static_cast<int>(DIR)), c.const_array(0), dx, false);
	copy_out(f, flux);
	copy_out(fd, fdiff);
}
// the same with the optical-depth wavespeed correction (radiation.use_wavespeed_correction, :1018-1022,1100-1109): needs the gas state of
// consVar (ComputeCellOpticalDepth :803-871), the cell sizes and a problem type with opacities (R2 .. R7)
template <typename P, FluxDir DIR>
void rad_fluxes_wsc(const qk_box *valid, const qk_array4 *flux, const qk_array4 *fdiff, const qk_array4 *left, const qk_array4 *right, const qk_array4 *cons,
		    int ngcons, const double *dx3)
{
	set_eos<P>();
	const int nh = RadSystem<P>::nvarHyperbolic_;
	auto c = make_mf(valid, -1, RadSystem<P>::nvar_, ngcons);
	auto l = make_mf(valid, static_cast<int>(DIR), nh, 1);
	auto r = make_mf(valid, static_cast<int>(DIR), nh, 1);
	auto f = make_mf(valid, static_cast<int>(DIR), nh, 0);
	auto fd = make_mf(valid, static_cast<int>(DIR), nh, 0);
	copy_in(c, cons);
	copy_in(l, left);
	copy_in(r, right);
	amrex::GpuArray<amrex::Real, 3> dx{dx3[0], dx3[1], dx3[2]};
	RadSystem<P>::template ComputeFluxes<DIR>(f.array(0), fd.array(0), l.const_array(0), r.const_array(0),
						  amrex::surroundingNodes(to_box(valid), static_cast<int>(DIR)), c.const_array(0), dx, true);
	copy_out(f, flux);
	copy_out(fd, fdiff);
}
template <typename P>
void rad_update(int op, const qk_box *valid, const qk_array4 *unew, const qk_array4 *u0, const qk_array4 *u1, const qk_array4 *const *fold,
		const qk_array4 *const *fnew, double dt, const double *dx3)
{
	const int nh = RadSystem<P>::nvarHyperbolic_, nv = RadSystem<P>::nvar_;
	amrex::GpuArray<amrex::Real, 3> dx{dx3[0], dx3[1], dx3[2]};
	auto un = make_mf(valid, -1, nv, 0);
	auto a0 = make_mf(valid, -1, nv, 0);
	auto a1 = make_mf(valid, -1, nv, 0);
	std::array<amrex::MultiFab, 3> FO{make_mf(valid, 0, nh, 0), make_mf(valid, 1, nh, 0), make_mf(valid, 2, nh, 0)};
	std::array<amrex::MultiFab, 3> F{make_mf(valid, 0, nh, 0), make_mf(valid, 1, nh, 0), make_mf(valid, 2, nh, 0)};
	copy_in(un, unew);
	copy_in(a0, u0);
	for (int d = 0; d < 3; ++d) {
		copy_in(F[d], fnew[d]);
	}
	amrex::Box bx = to_box(valid);
	if (op == 0) { // PredictStep
		RadSystem<P>::PredictStep(a0.const_array(0), un.array(0), {F[0].const_array(0), F[1].const_array(0), F[2].const_array(0)},
					  {F[0].const_array(0), F[1].const_array(0), F[2].const_array(0)}, dt, dx, bx, nh);
	} else { // AddFluxesRK2
		copy_in(a1, u1);
		for (int d = 0; d < 3; ++d) {
			copy_in(FO[d], fold[d]);
		}
		RadSystem<P>::AddFluxesRK2(un.array(0), a0.const_array(0), a1.const_array(0), {FO[0].const_array(0), FO[1].const_array(0), FO[2].const_array(0)},
					   {F[0].const_array(0), F[1].const_array(0), F[2].const_array(0)},
					   {FO[0].const_array(0), FO[1].const_array(0), FO[2].const_array(0)},
					   {F[0].const_array(0), F[1].const_array(0), F[2].const_array(0)}, dt, dx, bx, nh);
	}
	copy_out(un, unew);
}
#define DISPATCH_R(problem, CALL)                                                                                                                    \
	switch (problem) {                                                                                                                           \
	case 0: {                                                                                                                                    \
		using P = R0;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 1: {                                                                                                                                    \
		using P = R1;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	default:                                                                                                                                     \
		return -1;                                                                                                                           \
	}

#define DISPATCH_RS(problem, CALL)                                                                                                                   \
	switch (problem) {                                                                                                                           \
	case 2: {                                                                                                                                    \
		using P = R2;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 3: {                                                                                                                                    \
		using P = R3;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 4: {                                                                                                                                    \
		using P = R4;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 5: {                                                                                                                                    \
		using P = R5;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 6: {                                                                                                                                    \
		using P = R6;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 7: {                                                                                                                                    \
		using P = R7;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	default:                                                                                                                                     \
		return -1;                                                                                                                           \
	}

template <typename P>
void rad_source_terms(const qk_box *valid, const qk_array4 *cons, const qk_array4 *src, double dt, int stage, int64_t *counters)
{
	set_eos<P>();
	auto c = make_mf(valid, -1, RadSystem<P>::nvar_, 0);
	auto e = make_mf(valid, -1, 1, 0);
	copy_in(c, cons);
	if (src != nullptr) {
		copy_in(e, src);
	} else {
		e.setVal(0.);
	}
	int it[4] = {0, 0, 0, 0};
	int fail[3] = {0, 0, 0};
	auto arr = c.array(0);
	RadSystem<P>::AddSourceTermsSingleGroup(arr, e.const_array(0), to_box(valid), dt, stage, 0.0, it, fail);
	copy_out(c, cons);
	if (counters != nullptr) {
		counters[0] += it[0];
		counters[1] += it[1];
		counters[2] = std::max<int64_t>(counters[2], it[2]);
		counters[3] += it[3];
		for (int n = 0; n < 3; ++n) {
			counters[4 + n] += fail[n];
		}
	}
}

#define DISPATCH_P(problem, CALL)                                                                                                                    \
	switch (problem) {                                                                                                                           \
	case 0: {                                                                                                                                    \
		using P = P0;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 1: {                                                                                                                                    \
		using P = P1;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 2: {                                                                                                                                    \
		using P = P2;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 3: {                                                                                                                                    \
		using P = P3;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 4: {                                                                                                                                    \
		using P = P4;                                                                                                                        \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	default:                                                                                                                                     \
		return -1;                                                                                                                           \
	}
#define DISPATCH_D(dir, CALL)                                                                                                                        \
	switch (dir) {                                                                                                                               \
	case 0: {                                                                                                                                    \
		constexpr FluxDir D = FluxDir::X1;                                                                                                   \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	case 1: {                                                                                                                                    \
		constexpr FluxDir D = FluxDir::X2;                                                                                                   \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	default: {                                                                                                                                   \
		constexpr FluxDir D = FluxDir::X3;                                                                                                   \
		CALL;                                                                                                                                \
	} break;                                                                                                                                     \
	}
} // namespace

extern "C" {

int ref_nvar(int problem)
{
	DISPATCH_P(problem, return HydroSystem<P>::nvar_);
	return -1;
}

int ref_cons_to_prim(int problem, const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim, int ng)
{
	ensure_init();
	DISPATCH_P(problem, cons_to_prim<P>(valid, cons, prim, ng));
	return 0;
}

int ref_flattening_coefficients(int problem, int dir, const qk_box *valid, const qk_array4 *prim, const qk_array4 *chi, int ngprim, int ng)
{
	ensure_init();
	DISPATCH_P(problem, DISPATCH_D(dir, (flat_coefs<P, D>(valid, prim, chi, ngprim, ng))));
	return 0;
}

int ref_reconstruct(int problem, int order, int limiter, int dir, const qk_box *valid, const qk_array4 *q, const qk_array4 *left,
		    const qk_array4 *right, int ngq, int ng, int nvars)
{
	ensure_init();
	DISPATCH_P(problem, DISPATCH_D(dir, (reconstruct<P, D>(order, limiter, valid, q, left, right, ngq, ng, nvars))));
	return 0;
}

int ref_flatten_shocks(int problem, int dir, const qk_box *valid, const qk_array4 *q, const qk_array4 *c1, const qk_array4 *c2, const qk_array4 *c3,
		       const qk_array4 *left, const qk_array4 *right, int ngq, int ng, int nvars)
{
	ensure_init();
	DISPATCH_P(problem, DISPATCH_D(dir, (flatten<P, D>(valid, q, c1, c2, c3, left, right, ngq, ng, nvars))));
	return 0;
}

int ref_compute_fluxes(int problem, int solver, int dir, const qk_box *valid, const qk_array4 *flux, const qk_array4 *fvel, const qk_array4 *left,
		       const qk_array4 *right, const qk_array4 *prim, int ngprim, double K_visc)
{
	ensure_init();
	if (solver == QK_HLLC) {
		DISPATCH_P(problem, DISPATCH_D(dir, (fluxes<P, RiemannSolver::HLLC, D>(valid, flux, fvel, left, right, prim, ngprim, K_visc))));
	} else {
		DISPATCH_P(problem, DISPATCH_D(dir, (fluxes<P, RiemannSolver::LLF, D>(valid, flux, fvel, left, right, prim, ngprim, K_visc))));
	}
	return 0;
}

int ref_update_op(int problem, int op, const qk_box *valid, const qk_array4 *a0, const qk_array4 *a1, const qk_array4 *a2, const qk_array4 *fx,
		  const qk_array4 *fy, const qk_array4 *fz, const qk_iarray4 *redo, const double *dx3, double dt, double dfloor, double tfloor,
		  double *scalar_out)
{
	ensure_init();
	DISPATCH_P(problem, update_ops<P>(op, valid, a0, a1, a2, fx, fy, fz, redo, dx3, dt, dfloor, tfloor, scalar_out));
	return 0;
}

// ---- radiation (problem = 0: R0, 1: R1) ----
int ref_rad_params(int problem, qk_rad_params *out)
{
	DISPATCH_R(problem, {
		out->c_light = RadSystem<P>::c_light_;
		out->c_hat = RadSystem<P>::c_hat_;
		out->Erad_floor = RadSystem_Traits<P>::Erad_floor;
		out->ngroups = RadSystem<P>::nGroups_;
		out->nstart = RadSystem<P>::nstartHyperbolic_;
		out->reconstruction_order = 3;
		out->integrator_order = 2;
	});
	return 0;
}
int ref_rad_cons_to_prim(int problem, const qk_box *valid, const qk_array4 *cons, const qk_array4 *prim, int ng)
{
	ensure_init();
	DISPATCH_R(problem, rad_cons_to_prim<P>(valid, cons, prim, ng));
	return 0;
}
int ref_rad_compute_fluxes(int problem, int dir, const qk_box *valid, const qk_array4 *flux, const qk_array4 *fdiff, const qk_array4 *left,
			   const qk_array4 *right, const qk_array4 *cons, int ngcons)
{
	ensure_init();
	DISPATCH_R(problem, DISPATCH_D(dir, (rad_fluxes<P, D>(valid, flux, fdiff, left, right, cons, ngcons))));
	return 0;
}
int ref_rad_compute_fluxes_wsc(int problem, int dir, const qk_box *valid, const qk_array4 *flux, const qk_array4 *fdiff, const qk_array4 *left,
			       const qk_array4 *right, const qk_array4 *cons, int ngcons, const double *dx3)
{
	ensure_init();
	DISPATCH_RS(problem, DISPATCH_D(dir, (rad_fluxes_wsc<P, D>(valid, flux, fdiff, left, right, cons, ngcons, dx3))));
	return 0;
}
int ref_rad_update(int problem, int op, const qk_box *valid, const qk_array4 *unew, const qk_array4 *u0, const qk_array4 *u1, const qk_array4 *fxo,
		   const qk_array4 *fyo, const qk_array4 *fzo, const qk_array4 *fx, const qk_array4 *fy, const qk_array4 *fz, double dt, const double *dx3)
{
	ensure_init();
	const qk_array4 *fo[3] = {fxo, fyo, fzo};
	const qk_array4 *fn[3] = {fx, fy, fz};
	DISPATCH_R(problem, rad_update<P>(op, valid, unew, u0, u1, fo, fn, dt, dx3));
	return 0;
}

// ---- matter-radiation source terms (problem = 2..6: R2..R6) ----
int ref_rad_source_params(int problem, qk_hydro_params *hp, qk_rad_params *rp, qk_rad_source_params *sp)
{
	DISPATCH_RS(problem, {
		std::memset(hp, 0, sizeof(*hp));
		hp->gamma = quokka::EOS_Traits<P>::gamma;
		hp->mean_molecular_weight = quokka::EOS_Traits<P>::mean_molecular_weight;
		hp->boltzmann_constant = quokka::EOS_Traits<P>::boltzmann_constant;
		hp->small_temp = 1e-10;
		hp->small_dens = 1e-100;
		rp->c_light = RadSystem<P>::c_light_;
		rp->c_hat = RadSystem<P>::c_hat_;
		rp->Erad_floor = RadSystem_Traits<P>::Erad_floor;
		rp->ngroups = RadSystem<P>::nGroups_;
		rp->nstart = RadSystem<P>::nstartHyperbolic_;
		rp->reconstruction_order = 3;
		rp->integrator_order = 2;
		sp->radiation_constant = RadSystem<P>::radiation_constant_;
		sp->kappa_P = RadSystem<P>::ComputePlanckOpacity(1.0, 1.0);
		sp->kappa_E = RadSystem<P>::ComputeEnergyMeanOpacity(1.0, 1.0);
		sp->kappa_F = RadSystem<P>::ComputeFluxMeanOpacity(1.0, 1.0);
		sp->beta_order = RadSystem<P>::beta_order_;
		sp->opacity_model = QK_OPACITY_CONSTANT;
	});
	return 0;
}
int ref_rad_add_source_terms(int problem, const qk_box *valid, const qk_array4 *cons, const qk_array4 *src, double dt, int stage, int64_t *counters)
{
	ensure_init();
	DISPATCH_RS(problem, rad_source_terms<P>(valid, cons, src, dt, stage, counters));
	return 0;
}

// ---- AMR transfer operators: AMReX's own code (the interpolater Quokka selects, src/simulation.hpp:1389-1407) ----
// crse: FAB on CoarseBox(fine_region); fine: FAB containing fine_region; bc_lo/bc_hi: [3 * comp + dim]
int ref_interp_cons_lin_minmax(const qk_array4 *crse, const qk_array4 *fine, int ncomp, const qk_box *fine_region, const qk_box *dest_domain,
			       const qk_box *cdomain, const int *ratio, const int32_t *bc_lo, const int32_t *bc_hi)
{
	ensure_init();
	const amrex::IntVect rr(ratio[0], ratio[1], ratio[2]);
	const amrex::Box fbx = to_box(fine_region);
	const amrex::Box cbx = amrex::mf_linear_slope_minmax_interp.CoarseBox(fbx, rr);
	qk_box cb{{cbx.smallEnd(0), cbx.smallEnd(1), cbx.smallEnd(2)}, {cbx.bigEnd(0), cbx.bigEnd(1), cbx.bigEnd(2)}};
	auto cmf = make_mf(&cb, -1, ncomp, 0);
	auto fmf = make_mf(fine_region, -1, ncomp, 0);
	copy_in(cmf, crse);
	copy_in(fmf, fine);
	amrex::Vector<amrex::BCRec> bcs(ncomp);
	for (int n = 0; n < ncomp; ++n) {
		for (int d = 0; d < 3; ++d) {
			bcs[n].setLo(d, bc_lo[3 * n + d]);
			bcs[n].setHi(d, bc_hi[3 * n + d]);
		}
	}
	amrex::RealBox rb({0., 0., 0.}, {1., 1., 1.});
	amrex::Geometry cgeom(to_box(cdomain), rb, 0, {0, 0, 0});
	amrex::Geometry fgeom(amrex::refine(to_box(cdomain), rr), rb, 0, {0, 0, 0});
	amrex::mf_linear_slope_minmax_interp.interp(cmf, 0, fmf, 0, ncomp, amrex::IntVect(0), cgeom, fgeom, to_box(dest_domain), rr, bcs, 0);
	copy_out(fmf, fine);
	return 0;
}
int ref_average_down(const qk_array4 *crse, const qk_array4 *fine, int ncomp, const qk_box *cbx, const int *ratio)
{
	ensure_init();
	const amrex::IntVect rr(ratio[0], ratio[1], ratio[2]);
	qk_box fb;
	for (int d = 0; d < 3; ++d) {
		fb.lo[d] = cbx->lo[d] * ratio[d];
		fb.hi[d] = (cbx->hi[d] + 1) * ratio[d] - 1;
	}
	auto cmf = make_mf(cbx, -1, ncomp, 0);
	auto fmf = make_mf(&fb, -1, ncomp, 0);
	copy_in(fmf, fine);
	amrex::average_down(fmf, cmf, 0, ncomp, rr);
	copy_out(cmf, crse);
	return 0;
}

// ---- time interpolation between two states: amrex::FillPatchSingleLevel (AMReX_FillPatchUtil_I.H:140-175), the same
// alpha * old + beta * new expression FillPatcher::fill applies to the coarse data (AMReX_FillPatcher.H:372-383) ----
int ref_time_interp(const qk_box *bx, const qk_array4 *dst, const qk_array4 *src0, const qk_array4 *src1, int ncomp, double t0, double t1, double time)
{
	ensure_init();
	auto d = make_mf(bx, -1, ncomp, 0);
	auto s0 = make_mf(bx, -1, ncomp, 0);
	auto s1 = make_mf(bx, -1, ncomp, 0);
	copy_in(s0, src0);
	copy_in(s1, src1);
	amrex::RealBox rb({0., 0., 0.}, {1., 1., 1.});
	amrex::Geometry geom(to_box(bx), rb, 0, {0, 0, 0});
	amrex::PhysBCFunctNoOp bc;
	amrex::FillPatchSingleLevel(d, time, amrex::Vector<amrex::MultiFab *>{&s0, &s1}, amrex::Vector<amrex::Real>{t0, t1}, 0, 0, ncomp, geom, bc, 0);
	copy_out(d, dst);
	return 0;
}

// ---- QuokkaSimulation<P>::PreInterpState / PostInterpState (src/QuokkaSimulation.hpp:804-841), problem 0 (P0) ----
int ref_pre_post_interp_state(int post, const qk_box *bx, const qk_array4 *state)
{
	ensure_init();
	auto mf = make_mf(bx, -1, 6, 0);
	copy_in(mf, state);
	if (post != 0) {
		QuokkaSimulation<P0>::PostInterpState(mf, 0, 6);
	} else {
		QuokkaSimulation<P0>::PreInterpState(mf, 0, 6);
	}
	copy_out(mf, state);
	return 0;
}

} // extern "C"
