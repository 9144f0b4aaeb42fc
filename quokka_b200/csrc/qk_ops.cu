// qk_ops.cu -- one CUDA kernel per reference operator (the fine-grained C-ABI entry points used for
// kernel-by-kernel parity and as the simplest drop-in for HydroSystem<>/HyperbolicSystem<> calls).
// The fused, tuned path lives in qk_level.cu / qk_sweep.cuh; both share qk_physics.cuh.
// Compiled with --fmad=false (exact arithmetic contract, see qk_physics.cuh).
#include "qk_kernels.cuh"
#include <string.h>

int64_t g_qk_launches = 0;

namespace
{

// tiny device scratch for host-visible scalars (counts, maxima)
unsigned long long *g_scalar_dev = nullptr;
unsigned long long *g_scalar_host = nullptr;
int ensure_scalars()
{
	if (g_scalar_dev)
		return 0;
	QK_CUDA(cudaMalloc(&g_scalar_dev, 64));
	QK_CUDA(cudaMallocHost(&g_scalar_host, 64));
	return 0;
}
} // namespace

// ---- library ----------------------------------------------------------------------------------------
extern "C" int qk_abi_version(void) { return QK_ABI_VERSION; }
extern "C" int qk_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}
int qk_require_device() { return qk_device_count() > 0 ? QK_OK : QK_ERR_NO_DEVICE; }
extern "C" const char *qk_error_string(int code)
{
	switch (code) {
	case QK_OK:
		return "ok";
	case QK_ERR_NO_DEVICE:
		return "no CUDA device: libquokka_b200 has no CPU fallback";
	case QK_ERR_BAD_ARG:
		return "bad argument";
	case QK_ERR_UNSUPPORTED:
		return "unsupported configuration (isothermal EOS, K_visc != 0, > QK_MAX_SCALARS scalars, MHD)";
	case QK_ERR_NOMEM:
		return "out of device memory";
	default:
		return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown";
	}
}
extern "C" int64_t qk_launch_count(void) { return g_qk_launches; }

static int check_params(const qk_hydro_params *p)
{
	if (!p)
		return QK_ERR_BAD_ARG;
	if (p->gamma == 1.0 || p->K_visc != 0.0 || p->nscalars > QK_MAX_SCALARS || p->nscalars < 0 || p->nmscalars > p->nscalars)
		return QK_ERR_UNSUPPORTED;
	return qk_require_device();
}
#define QK_TRY(x)                                                                                                                                    \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

// ---- per-operator entry points ----------------------------------------------------------------------
extern "C" int qk_hydro_conserved_to_primitive(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons,
					       const qk_array4 *prim, int nghost, void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(nghost));
		k_cons_to_prim<<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(cons[b]), A4(prim[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_flattening_coefficients(const qk_hydro_params *prm, int dir, int nboxes, const qk_box *valid, const qk_array4 *prim,
						const qk_array4 *chi, int nghost, void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(nghost));
		k_flat_coefs<<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, A4(prim[b]), A4(chi[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_reconstruct_states(int order, int limiter, int dir, int nboxes, const qk_box *valid, const qk_array4 *q, const qk_array4 *left,
				     const qk_array4 *right, int nghost, int nvars, void *stream)
{
	QK_TRY(qk_require_device());
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(nghost));
		if (order == 3)
			k_reconstruct<3, 0><<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, nvars, A4(q[b]), A4(left[b]), A4(right[b]));
		else if (order == 2 && limiter == QK_MC)
			k_reconstruct<2, QK_MC><<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, nvars, A4(q[b]), A4(left[b]), A4(right[b]));
		else if (order == 2)
			k_reconstruct<2, QK_MINMOD><<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, nvars, A4(q[b]), A4(left[b]), A4(right[b]));
		else if (order == 1)
			k_reconstruct<1, 0><<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, nvars, A4(q[b]), A4(left[b]), A4(right[b]));
		else
			return QK_ERR_BAD_ARG;
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_flatten_shocks(int dir, int nboxes, const qk_box *valid, const qk_array4 *q, const qk_array4 *chi1, const qk_array4 *chi2,
				       const qk_array4 *chi3, const qk_array4 *left, const qk_array4 *right, int nghost, int nvars, void *stream)
{
	QK_TRY(qk_require_device());
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(nghost));
		k_flatten<<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, nvars, A4(q[b]), A4(chi1[b]), A4(chi2[b]), A4(chi3[b]), A4(left[b]), A4(right[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_compute_fluxes(const qk_hydro_params *prm, int solver, int dir, int nboxes, const qk_box *valid, const qk_array4 *flux,
				       const qk_array4 *facevel, const qk_array4 *left, const qk_array4 *right, const qk_array4 *prim, void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).face(dir));
		if (solver == QK_HLLC)
			k_compute_fluxes<QK_HLLC><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, A4(flux[b]), A4(facevel[b]), A4(left[b]), A4(right[b]), A4(prim[b]));
		else
			k_compute_fluxes<QK_LLF><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, A4(flux[b]), A4(facevel[b]), A4(left[b]), A4(right[b]), A4(prim[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_flux_function(const qk_hydro_params *prm, int fo, int dir, int nboxes, const qk_box *valid, const qk_array4 *prim,
				      const qk_array4 *chi1, const qk_array4 *chi2, const qk_array4 *chi3, const qk_array4 *flux, const qk_array4 *facevel,
				      void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).face(dir));
		const A4 q(prim[b]), f(flux[b]), v(facevel[b]);
		if (fo) {
			k_flux_function<1, QK_LLF><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, q, q, q, q, f, v);
		} else {
			const A4 c1(chi1[b]), c2(chi2[b]), c3(chi3[b]);
			if (prm->reconstruction_order == 3)
				k_flux_function<3, QK_HLLC><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, q, c1, c2, c3, f, v);
			else if (prm->reconstruction_order == 2)
				k_flux_function<2, QK_HLLC><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, q, c1, c2, c3, f, v);
			else
				k_flux_function<1, QK_HLLC><<<it.blocks(), TPB, 0, S(stream)>>>(c, dir, it, q, c1, c2, c3, f, v);
		}
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_saxpy(int nboxes, const qk_box *region, const qk_array4 *dst, double a, const qk_array4 *src, int ncomp, void *stream)
{
	QK_TRY(qk_require_device());
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(region[b])));
		k_saxpy<<<it.blocks(), TPB, 0, S(stream)>>>(it, ncomp, A4(dst[b]), a, A4(src[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_rhs_from_fluxes(int nboxes, const qk_box *valid, const qk_array4 *rhs, const qk_array4 *fx, const qk_array4 *fy,
					const qk_array4 *fz, const double dx[3], int nvars, void *stream)
{
	QK_TRY(qk_require_device());
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_rhs<<<it.blocks(), TPB, 0, S(stream)>>>(it, nvars, A4(rhs[b]), A4(fx[b]), A4(fy[b]), A4(fz[b]), dx[0], dx[1], dx[2]);
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_add_internal_energy_pdv(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *rhs,
						const qk_array4 *cons, const double dx[3], const qk_array4 *vx, const qk_array4 *vy, const qk_array4 *vz,
						const qk_iarray4 *redo, void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_pdv<<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(rhs[b]), A4(cons[b]), A4(vx[b]), A4(vy[b]), A4(vz[b]), IA4(redo[b]), dx[0], dx[1], dx[2]);
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_predict_step(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons_old,
				     const qk_array4 *cons_new, const qk_array4 *rhs, double dt, int nvars, const qk_iarray4 *redo, int64_t *ncells_bad,
				     void *stream)
{
	QK_TRY(check_params(prm));
	QK_TRY(ensure_scalars());
	const HydroConst c = make_hydro_const(prm);
	QK_CUDA(cudaMemsetAsync(g_scalar_dev, 0, 8, S(stream)));
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_predict<<<it.blocks(), TPB, 0, S(stream)>>>(c, it, nvars, A4(cons_old[b]), A4(cons_new[b]), A4(rhs[b]), dt, IA4(redo[b]), g_scalar_dev);
		QK_KERNEL_CHECK();
	}
	if (ncells_bad) {
		QK_CUDA(cudaMemcpyAsync(g_scalar_host, g_scalar_dev, 8, cudaMemcpyDeviceToHost, S(stream)));
		QK_CUDA(cudaStreamSynchronize(S(stream)));
		*ncells_bad = (int64_t)g_scalar_host[0];
	}
	return 0;
}

extern "C" int qk_hydro_enforce_limits(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, void *stream)
{
	QK_TRY(check_params(prm));
	const HydroConst c = make_hydro_const(prm);
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_enforce<<<it.blocks(), TPB, 0, S(stream)>>>(c, it, A4(state[b]));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_sync_dual_energy(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *state, int64_t *nabort,
					 void *stream)
{
	QK_TRY(check_params(prm));
	QK_TRY(ensure_scalars());
	QK_CUDA(cudaMemsetAsync(g_scalar_dev, 0, 8, S(stream)));
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_sync<<<it.blocks(), TPB, 0, S(stream)>>>(it, A4(state[b]), g_scalar_dev);
		QK_KERNEL_CHECK();
	}
	if (nabort) {
		QK_CUDA(cudaMemcpyAsync(g_scalar_host, g_scalar_dev, 8, cudaMemcpyDeviceToHost, S(stream)));
		QK_CUDA(cudaStreamSynchronize(S(stream)));
		*nabort = (int64_t)g_scalar_host[0];
	}
	return 0;
}

extern "C" int qk_hydro_replace_fluxes(int dir, int nboxes, const qk_box *valid, const qk_array4 *flux, const qk_array4 *fo_flux,
				       const qk_iarray4 *redo, int ncomp, void *stream)
{
	QK_TRY(qk_require_device());
	for (int b = 0; b < nboxes; ++b) {
		Iter it(Box3(valid[b]).grown(1));
		k_replace<<<it.blocks(), TPB, 0, S(stream)>>>(dir, it, ncomp, A4(flux[b]), A4(fo_flux[b]), IA4(redo[b]), Box3(valid[b]).face(dir));
		QK_KERNEL_CHECK();
	}
	return 0;
}

extern "C" int qk_hydro_max_signal_speed(const qk_hydro_params *prm, int which, int nboxes, const qk_box *valid, const qk_array4 *cons,
					 double *max_out, void *stream)
{
	QK_TRY(check_params(prm));
	QK_TRY(ensure_scalars());
	const HydroConst c = make_hydro_const(prm);
	QK_CUDA(cudaMemsetAsync(g_scalar_dev, 0, 8, S(stream)));
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		k_max_signal<<<it.blocks(), TPB, 0, S(stream)>>>(c, which, it, A4(cons[b]), g_scalar_dev);
		QK_KERNEL_CHECK();
	}
	QK_CUDA(cudaMemcpyAsync(g_scalar_host, g_scalar_dev, 8, cudaMemcpyDeviceToHost, S(stream)));
	QK_CUDA(cudaStreamSynchronize(S(stream)));
	*max_out = (g_scalar_host[0] == 0ull) ? ((which == 0) ? 0.0 : -1.7976931348623157e308) : key2d(g_scalar_host[0]);
	return 0;
}
