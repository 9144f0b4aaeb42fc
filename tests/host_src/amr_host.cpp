// tests/host_src/amr_host.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-coarse-cell AMR transfer arithmetic
// (quokka_b200/csrc/qk_amr.cuh, the bodies of k_amr_interp / k_amr_avgdown) for the HOST and runs it over one box pair so that it
// can be compared with the oracle without a GPU (tests/test_amr_host.py).  Never linked into libquokka_b200.so.
#include "../../quokka_b200/csrc/qk_amr.cuh"

extern "C" void host_amr_interp(const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *fine_region,
				const qk_box *dest_domain, const qk_box *cdomain, const int *ratio, const int32_t *bc_lo, const int32_t *bc_hi)
{
	qk_amr::InterpParams P;
	qk_amr::Box region;
	for (int d = 0; d < 3; ++d) {
		P.cdomain.lo[d] = cdomain->lo[d];
		P.cdomain.hi[d] = cdomain->hi[d];
		P.dest.lo[d] = dest_domain->lo[d];
		P.dest.hi[d] = dest_domain->hi[d];
		P.ratio[d] = ratio[d];
		region.lo[d] = fine_region->lo[d];
		region.hi[d] = fine_region->hi[d];
	}
	P.ccomp = ccomp;
	P.fcomp = fcomp;
	P.ncomp = ncomp;
	for (int n = 0; n < 3 * ncomp; ++n) {
		P.bc_lo[n] = bc_lo[n];
		P.bc_hi[n] = bc_hi[n];
	}
	const qk_amr::V4 c = qk_amr::view(*crse), f = qk_amr::view(*fine);
	int clo[3], chi[3];
	for (int d = 0; d < 3; ++d) {
		clo[d] = qk_amr::coarsen(region.lo[d], ratio[d]);
		chi[d] = qk_amr::coarsen(region.hi[d], ratio[d]);
	}
	for (int k = clo[2]; k <= chi[2]; ++k)
		for (int j = clo[1]; j <= chi[1]; ++j)
			for (int i = clo[0]; i <= chi[0]; ++i)
				qk_amr::interp_coarse_cell(c, f, i, j, k, region, P);
}

extern "C" void host_amr_average_down(const qk_array4 *crse, int ccomp, const qk_array4 *fine, int fcomp, int ncomp, const qk_box *cbx, const int *ratio)
{
	const qk_amr::V4 c = qk_amr::view(*crse), f = qk_amr::view(*fine);
	for (int n = 0; n < ncomp; ++n)
		for (int k = cbx->lo[2]; k <= cbx->hi[2]; ++k)
			for (int j = cbx->lo[1]; j <= cbx->hi[1]; ++j)
				for (int i = cbx->lo[0]; i <= cbx->hi[0]; ++i)
					qk_amr::at(c, i, j, k, n + ccomp) = qk_amr::avgdown_cell(f, i, j, k, n + fcomp, ratio);
}

extern "C" void host_amr_prepost(int post, const qk_array4 *state, const qk_box *bx)
{
	const qk_amr::V4 c = qk_amr::view(*state);
	for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
		for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
			for (int i = bx->lo[0]; i <= bx->hi[0]; ++i) {
				if (post)
					qk_amr::prepost_cell<true>(c, i, j, k);
				else
					qk_amr::prepost_cell<false>(c, i, j, k);
			}
}

// time interpolation and regrid tagging: the per-cell functions of k_amr_time_interp / k_tag
extern "C" int host_time_interp(const qk_array4 *dst, const qk_array4 *s0, const qk_array4 *s1, int ncomp, const qk_box *bx, double t0, double t1, double time)
{
	const qk_amr::V4 d = qk_amr::view(*dst), a = qk_amr::view(*s0), b = qk_amr::view(*s1);
	const int which = qk_amr::time_interp_branch(t0, t1, time, true);
	const double alpha = (which == 2) ? (t1 - time) / (t1 - t0) : 0.0, beta = (which == 2) ? (time - t0) / (t1 - t0) : 0.0;
	for (int n = 0; n < ncomp; ++n)
		for (int k = bx->lo[2]; k <= bx->hi[2]; ++k)
			for (int j = bx->lo[1]; j <= bx->hi[1]; ++j)
				for (int i = bx->lo[0]; i <= bx->hi[0]; ++i)
					qk_amr::at(d, i, j, k, n) = qk_amr::time_interp_value(which, alpha, beta, qk_amr::at(a, i, j, k, n), qk_amr::at(b, i, j, k, n));
	return which;
}
extern "C" int host_tag_pressure(const double *P7, double eta, double pmin) { return qk_amr::tag_pressure_gradient(P7, eta, pmin) ? 1 : 0; }
extern "C" int host_tag_gradient_x(double qm, double q0, double qp, double dx, double eta, double qmin) { return qk_amr::tag_gradient_x(qm, q0, qp, dx, eta, qmin) ? 1 : 0; }
