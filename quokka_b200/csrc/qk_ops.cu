// qk_ops.cu -- one CUDA kernel per reference operator (the fine-grained C-ABI entry points used for
// kernel-by-kernel parity and as the simplest drop-in for HydroSystem<>/HyperbolicSystem<> calls).
// The fused, tuned path lives in qk_level.cu / qk_sweep.cuh; both share qk_physics.cuh.
// Compiled with --fmad=false (exact arithmetic contract, see qk_physics.cuh).
#include "qk_kernels.cuh"
#include "qk_div.cuh"
#include <string.h>
#include <algorithm>

#include <map>
#include <string>
#include <vector>

int64_t g_qk_launches = 0;
bool g_qk_prof_on = false;

// per-class event pairs; resolved (synchronised + summed) only when the report is read
namespace
{
struct ProfClass {
	std::vector<cudaEvent_t> ev; // begin,end,begin,end...
	size_t used = 0;
	int64_t launches = 0;
};
std::map<std::string, ProfClass> g_prof;
} // namespace

void qk_prof_mark(const char *name, int launches, cudaStream_t s, bool begin)
{
	ProfClass &c = g_prof[name];
	if (c.used == c.ev.size()) {
		cudaEvent_t e;
		if (cudaEventCreate(&e) != cudaSuccess)
			return;
		c.ev.push_back(e);
	}
	cudaEventRecord(c.ev[c.used++], s);
	if (!begin)
		c.launches += launches;
}

extern "C" int qk_prof_enable(int on)
{
	g_qk_prof_on = (on != 0);
	if (on)
		for (auto &kv : g_prof) {
			kv.second.used = 0;
			kv.second.launches = 0;
		}
	return 0;
}

// "name launches total_ms\n" per class into buf; returns the number of bytes needed
extern "C" int qk_prof_report(char *buf, int buflen)
{
	std::string out;
	for (auto &kv : g_prof) {
		ProfClass &c = kv.second;
		if (c.used < 2)
			continue;
		double ms = 0;
		for (size_t i = 0; i + 1 < c.used; i += 2) {
			cudaEventSynchronize(c.ev[i + 1]);
			float t = 0;
			if (cudaEventElapsedTime(&t, c.ev[i], c.ev[i + 1]) == cudaSuccess)
				ms += t;
		}
		char line[256];
		snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(), (long long)c.launches, ms);
		out += line;
	}
	if (buf && buflen > 0) {
		strncpy(buf, out.c_str(), (size_t)buflen - 1);
		buf[buflen - 1] = 0;
	}
	return (int)out.size() + 1;
}

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
		return "unsupported configuration (isothermal EOS, > QK_MAX_SCALARS scalars, MHD)";
	case QK_ERR_NOMEM:
		return "out of device memory";
	case QK_ERR_NOT_CONVERGED:
		return "matter-radiation coupling failed to converge, or the radiation subcycle exceeds maxSubsteps + 1";
	default:
		return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown";
	}
}
extern "C" int64_t qk_launch_count(void) { return g_qk_launches; }

static int check_params(const qk_hydro_params *p)
{
	if (!p)
		return QK_ERR_BAD_ARG;
	if (p->nscalars > QK_MAX_SCALARS || p->nscalars < 0 || p->nmscalars > p->nscalars)
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
	ProfScope prof_("cons_to_prim", S(stream));
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
	ProfScope prof_("flatten_coefs", S(stream));
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
	ProfScope prof_("reconstruct", S(stream));
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
	ProfScope prof_("flatten_shocks", S(stream));
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
	ProfScope prof_("compute_fluxes", S(stream));
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
	ProfScope prof_("flux_function", S(stream));
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
	ProfScope prof_("saxpy", S(stream));
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
	ProfScope prof_("rhs_from_fluxes", S(stream));
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
	ProfScope prof_("pdv", S(stream));
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
	ProfScope prof_("predict_step", S(stream));
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
	ProfScope prof_("enforce_limits", S(stream));
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
	ProfScope prof_("sync_dual_energy", S(stream));
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
	ProfScope prof_("replace_fluxes", S(stream));
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
	ProfScope prof_("max_signal_speed", S(stream));
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


// both reductions of a time step in ONE pass (internal; the driver caches out[0] for the next computeTimestep):
// out[0] = ComputeMaxSignalSpeed + norminf (simulation.hpp:709-710), out[1] = maxSignalSpeedLocal (isCflViolated)
int qk_hydro_max_signal_both(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, double out[2], cudaStream_t s)
{
	QK_TRY(check_params(prm));
	QK_TRY(ensure_scalars());
	const HydroConst c = make_hydro_const(prm);
	QK_CUDA(cudaMemsetAsync(g_scalar_dev, 0, 16, s));
	ProfScope prof_("max_signal_speed", s);
	for (int b = 0; b < nboxes; ++b) {
		Iter it((Box3(valid[b])));
		const unsigned blocks = std::min<unsigned>(it.blocks(), 148u * 8u);
		k_max_signal2<<<blocks, TPB, 0, s>>>(c, it, A4(cons[b]), g_scalar_dev);
		QK_KERNEL_CHECK();
	}
	QK_CUDA(cudaMemcpyAsync(g_scalar_host, g_scalar_dev, 16, cudaMemcpyDeviceToHost, s));
	QK_CUDA(cudaStreamSynchronize(s));
	out[0] = (g_scalar_host[0] == 0ull) ? 0.0 : key2d(g_scalar_host[0]);
	out[1] = (g_scalar_host[1] == 0ull) ? -1.7976931348623157e308 : key2d(g_scalar_host[1]);
	return 0;
}

// ---- self-test of qk_div.cuh on the device (tests/test_gpu_division.py) ------------------------------------------
namespace
{
__device__ __forceinline__ unsigned long long sm64(unsigned long long &s)
{
	s += 0x9E3779B97F4A7C15ull;
	unsigned long long z = s;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
// mode 0: full-range random bit patterns; 1: random mantissas, exponents within +-40 of 1; 2: specials mixed in
__global__ void k_div_selftest(unsigned long long seed, int mode, int per_thread, unsigned long long *out)
{
	unsigned long long s = seed + 0x1234567ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
	unsigned long long bad_div = 0, bad_rcp = 0;
	const double specials[12] = {0.0, -0.0, 1.0, -1.0, 4.9406564584124654e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
				     __longlong_as_double(0x7ff0000000000000ll), __longlong_as_double(0xfff0000000000000ll),
				     __longlong_as_double(0x7ff8000000000000ll), 1e-300, 1e300};
	for (int it = 0; it < per_thread; ++it) {
		unsigned long long ua = sm64(s), ub = sm64(s);
		if (mode == 1) {
			ua = (ua & 0x800fffffffffffffull) | ((0x3ffull - 40 + (sm64(s) % 81)) << 52);
			ub = (ub & 0x800fffffffffffffull) | ((0x3ffull - 40 + (sm64(s) % 81)) << 52);
		}
		double a = __longlong_as_double((long long)ua), b = __longlong_as_double((long long)ub);
		if (mode == 2) {
			const unsigned long long pick = sm64(s);
			if ((pick & 3) == 0)
				a = specials[(pick >> 8) % 12];
			if ((pick & 12) == 0)
				b = specials[(pick >> 16) % 12];
		}
		const QkRcp r = qk_rcp(b);
		const double q1 = qk_div(a, r);
		const double q2 = a / b;
		if (__double_as_longlong(q1) != __double_as_longlong(q2) && !((q1 != q1) && (q2 != q2)))
			++bad_div;
		// is the division's refined reciprocal the correctly rounded 1/b wherever the fast path applies?
		const double y = 1.0 / b;
		if (r.ok && __double_as_longlong(r.y) != __double_as_longlong(y) && !((r.y != r.y) && (y != y)))
			++bad_rcp;
	}
	if (bad_div)
		atomicAdd(out, bad_div);
	if (bad_rcp)
		atomicAdd(out + 1, bad_rcp);
}
} // namespace

extern "C" int qk_selftest_division(uint64_t seed, int mode, int64_t npairs, int64_t *bad_div, int64_t *bad_rcp)
{
	QK_TRY(qk_require_device());
	QK_TRY(ensure_scalars());
	QK_CUDA(cudaMemset(g_scalar_dev, 0, 16));
	const int threads = 256, blocks = 148 * 8, per = (int)((npairs + (int64_t)threads * blocks - 1) / ((int64_t)threads * blocks));
	k_div_selftest<<<blocks, threads>>>(seed, mode, per, g_scalar_dev);
	QK_KERNEL_CHECK();
	QK_CUDA(cudaMemcpy(g_scalar_host, g_scalar_dev, 16, cudaMemcpyDeviceToHost));
	if (bad_div)
		*bad_div = (int64_t)g_scalar_host[0];
	if (bad_rcp)
		*bad_rcp = (int64_t)g_scalar_host[1];
	return 0;
}
