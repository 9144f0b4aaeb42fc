// tests/host_src/rad_source_host.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-cell source-term arithmetic
// (quokka_b200/csrc/qk_rad_source.cuh, the body of the CUDA kernel k_rad_source) for the HOST with g++ -ffp-contract=off and
// runs it over one box, so that the kernel's arithmetic can be compared with the oracle on a machine without a GPU
// (tests/test_rad_source_host.py).  Built into tests/host_src/_build/; never linked into libquokka_b200.so.
#include "../../quokka_b200/csrc/qk_rad_source.cuh"

static inline double &at(const qk_array4 *a, int i, int j, int k, int n)
{
	return a->p[(int64_t)(i - a->begin[0]) + (int64_t)(j - a->begin[1]) * a->jstride + (int64_t)(k - a->begin[2]) * a->kstride + n * a->nstride];
}

template <bool RELAXED>
static void run_box(const qk_hydro_params *hp, const qk_rad_params *rp, const qk_rad_source_params *sp, const qk_array4 *cons,
					  const qk_array4 *src, const qk_box *bx, double dt_radiation, int stage, int64_t *counters)
{
	const qk_rsrc::Const k = qk_rsrc::make_const(hp, rp, sp, dt_radiation, stage);
	const int ns = rp->nstart;
	qk_rsrc::DivPlain::CTab ct;
	qk_rsrc::const_denoms(k, ct.b);
	for (int kk = bx->lo[2]; kk <= bx->hi[2]; ++kk)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				qk_rsrc::CellIn in;
				qk_rsrc::CellOut out;
				in.rho = at(cons, i, j, kk, 0);
				for (int m = 0; m < 3; ++m) {
					in.mom[m] = at(cons, i, j, kk, 1 + m);
					in.F[m] = at(cons, i, j, kk, ns + 1 + m);
				}
				in.Egastot = at(cons, i, j, kk, 4);
				in.Erad = at(cons, i, j, kk, ns);
				in.src = src ? at(src, i, j, kk, 0) : 0.0;
				qk_rsrc::source_cell<qk_rsrc::DivPlain, RELAXED>(k, ct, in, out);
				for (int m = 0; m < 3; ++m) {
					at(cons, i, j, kk, 1 + m) = out.mom[m];
					at(cons, i, j, kk, ns + 1 + m) = out.F[m];
				}
				if (k.gamma != 1.0) {
					at(cons, i, j, kk, 4) = out.Egastot;
					at(cons, i, j, kk, 5) = out.Eint;
					at(cons, i, j, kk, ns) = out.Erad;
				}
				if (counters) {
					counters[0] += out.solves;
					counters[1] += out.nr_iters;
					counters[2] = counters[2] < out.nr_max ? out.nr_max : counters[2];
					counters[4] += out.fail_nr;
					counters[6] += out.fail_outer;
				}
			}
}

extern "C" void host_rad_add_source_terms(const qk_hydro_params *hp, const qk_rad_params *rp, const qk_rad_source_params *sp, const qk_array4 *cons,
					  const qk_array4 *src, const qk_box *bx, double dt_radiation, int stage, int64_t *counters)
{
	if (hp->arith == QK_ARITH_FAST)
		run_box<true>(hp, rp, sp, cons, src, bx, dt_radiation, stage, counters);
	else
		run_box<false>(hp, rp, sp, cons, src, bx, dt_radiation, stage, counters);
}
