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
#include <cstdint>
#include <cstring>

#include "AMReX.H"
#include "AMReX_BoxArray.H"
#include "AMReX_DistributionMapping.H"
#include "AMReX_MultiFab.H"
#include "AMReX_ParmParse.H"
#include "AMReX_iMultiFab.H"

#include "hydro/hydro_system.hpp"
#include "hyperbolic_system.hpp"
#include "radiation/radiation_system.hpp"

#include "../../include/quokka_b200.h"

struct P0 {
};
struct P1 {
};
struct P2 {
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

} // extern "C"
