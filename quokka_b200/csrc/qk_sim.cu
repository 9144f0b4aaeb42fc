// qk_sim.cu -- C++ host driver above the operator ABI: the single-level image of
// AMRSimulation<problem_t>::evolve / computeTimestep (src/simulation.hpp:722-977) and
// QuokkaSimulation<problem_t>::advanceSingleTimestepAtLevel / advanceHydroAtLevelWithRetries /
// advanceHydroAtLevel (src/QuokkaSimulation.hpp:653-707, 885-990, 1032-1322) for a uniform level
// (no AMR, no Strang sources, no tracers).  It owns state_new / state_old / state_inter as device
// FABs in AMReX layout and calls only the C ABI of include/quokka_b200.h.
#include "qk_level.h"

#include <chrono>
#include <float.h>
#include <string.h>

extern "C" int qk_fill_boundary(qk_level *L, const qk_array4 *state, int scomp, int ncomp, void *stream);

struct qk_sim {
	qk_level *lev = nullptr;
	qk_comm *comm = nullptr;
	qk_hydro_params prm;
	double cfl = 0.3;
	double t = 0.0;
	double dt_prev = 1.e100; // dt_.resize(nlevs_max, 1.e100)  src/simulation.hpp:448
	int64_t cell_updates = 0;
	int64_t ncells_global = 0;
	int64_t retries = 0;
	int nb = 0;
	std::vector<qk_array4> snew, sold, sint, stmp;
	double *pool[4] = {nullptr, nullptr, nullptr, nullptr};
	std::vector<size_t> off;
	size_t pool_doubles = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	// local ComputeMaxSignalSpeed maximum of state_new, computed in the same pass as isCflViolated's at the end of the last
	// advance (the next computeTimestep reads the same state); invalidated whenever state_new is changed from outside
	bool sig_is_global = false; // sig_local is already the maximum over all ranks
	bool sig_valid = false;
	double sig_local = 0.0;
	// radiation (is_radiation_enabled): subcycleRadiationAtLevel after the hydro advance, c_hat / maxSubsteps_ in the time step
	bool rad_on = false, src_on = false;
	qk_rad_params rprm;
	qk_rad_source_params sprm;
	std::vector<qk_array4> esrc; // radEnergySource per local box (caller-owned device memory) or empty
	double rad_cfl = 0.3;	     // radiationCflNumber_  src/QuokkaSimulation.hpp:125
	int max_substeps = 10;	     // maxSubsteps_  :126
	int64_t rad_failures = 0;    // nf_coupling + nf_dust + nf_outer over the run (:1692-1699)
	int last_nsub = 0;
	int64_t rad_cell_updates = 0; // radiationCellUpdates_
	// host <-> device transfer of the VALID cells only (qk_sim_set_state_valid / qk_sim_get_state_valid): contiguous staging per box, copies on
	// their own stream, (un)packing kernels on the simulation's stream, one event per box
	double *stage = nullptr;
	std::vector<size_t> stage_off;
	cudaStream_t copy_stream = nullptr;
	std::vector<cudaEvent_t> box_ev;
};
int qk_hydro_max_signal_both(const qk_hydro_params *prm, int nboxes, const qk_box *valid, const qk_array4 *cons, double out[2], cudaStream_t s);

static inline int blen(const qk_box &b, int d) { return b.hi[d] - b.lo[d] + 1; }

static int alloc_state(qk_sim *s, int which, std::vector<qk_array4> &out)
{
	if (!s->pool[which]) {
		if (cudaMalloc(&s->pool[which], s->pool_doubles * sizeof(double)) != cudaSuccess) {
			cudaGetLastError();
			return QK_ERR_NOMEM;
		}
		QK_CUDA(cudaMemsetAsync(s->pool[which], 0, s->pool_doubles * sizeof(double), s->stream));
	}
	out.resize(s->nb);
	for (int b = 0; b < s->nb; ++b) {
		qk_box g = s->lev->valid[b];
		for (int d = 0; d < 3; ++d) {
			g.lo[d] -= s->lev->nghost;
			g.hi[d] += s->lev->nghost;
		}
		qk_array4 a;
		a.p = s->pool[which] + s->off[b];
		a.jstride = blen(g, 0);
		a.kstride = a.jstride * blen(g, 1);
		a.nstride = a.kstride * blen(g, 2);
		for (int d = 0; d < 3; ++d) {
			a.begin[d] = g.lo[d];
			a.end[d] = g.hi[d] + 1;
		}
		a.ncomp = s->lev->ncomp;
		out[b] = a;
	}
	return 0;
}

extern "C" int qk_sim_create(const qk_level_desc *desc, const qk_hydro_params *prm, double cfl, qk_comm *comm, qk_sim **out)
{
	if (!desc || !prm || !out)
		return QK_ERR_BAD_ARG;
	int rc = qk_require_device();
	if (rc)
		return rc;
	qk_sim *s = new qk_sim();
	s->prm = *prm;
	s->cfl = cfl;
	s->comm = comm;
	rc = qk_level_create(desc, &s->lev);
	if (rc) {
		delete s;
		return rc;
	}
	qk_level_set_comm(s->lev, comm);
	s->nb = (int)s->lev->valid.size();
	for (int b = 0; b < desc->nboxes_global; ++b)
		s->ncells_global += (int64_t)blen(desc->boxes_global[b], 0) * blen(desc->boxes_global[b], 1) * blen(desc->boxes_global[b], 2);
	s->off.resize(s->nb);
	size_t total = 0;
	for (int b = 0; b < s->nb; ++b) {
		const qk_box &v = s->lev->valid[b];
		const int ng = s->lev->nghost;
		s->off[b] = total;
		size_t n = (size_t)(blen(v, 0) + 2 * ng) * (blen(v, 1) + 2 * ng) * (blen(v, 2) + 2 * ng) * s->lev->ncomp;
		total += (n + 31) & ~(size_t)31;
	}
	s->pool_doubles = total;
	cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
	cudaEventCreate(&s->ev0);
	cudaEventCreate(&s->ev1);
	rc = alloc_state(s, 0, s->snew);
	rc = rc ? rc : alloc_state(s, 1, s->sold);
	rc = rc ? rc : alloc_state(s, 2, s->sint); // state_inter_cc_, setVal(0) (QuokkaSimulation.hpp:1056-1057)
	if (rc) {
		qk_sim_destroy(s);
		return rc;
	}
	*out = s;
	return 0;
}

extern "C" void qk_sim_destroy(qk_sim *s)
{
	if (!s)
		return;
	if (s->stream)
		cudaStreamSynchronize(s->stream);
	qk_level_destroy(s->lev);
	for (int i = 0; i < 4; ++i)
		if (s->pool[i])
			cudaFree(s->pool[i]);
	if (s->copy_stream) {
		cudaStreamSynchronize(s->copy_stream);
		cudaStreamDestroy(s->copy_stream);
	}
	for (cudaEvent_t e : s->box_ev)
		cudaEventDestroy(e);
	if (s->stage)
		cudaFree(s->stage);
	if (s->ev0)
		cudaEventDestroy(s->ev0);
	if (s->ev1)
		cudaEventDestroy(s->ev1);
	if (s->stream)
		cudaStreamDestroy(s->stream);
	delete s;
}

extern "C" int qk_sim_nlocal(const qk_sim *s) { return s ? s->nb : 0; }

// Physics_Traits::is_radiation_enabled for this simulation: every qk_sim_step then runs subcycleRadiationAtLevel after the hydro
// advance (src/QuokkaSimulation.hpp:690-694) and computeTimestep takes std::max(c_hat / maxSubsteps_, hydro signal speed)
// (computeMaxSignalLocal :408-441).  src == NULL: transport only.  esrc: one-component device FABs per local box, or NULL.
extern "C" int qk_sim_enable_radiation(qk_sim *s, const qk_rad_params *rad, const qk_rad_source_params *src, const qk_array4 *esrc, double rad_cfl,
				       int max_substeps)
{
	if (!s || !rad || !(rad_cfl > 0.0) || max_substeps < 1)
		return QK_ERR_BAD_ARG;
	if (rad->nstart + 4 * rad->ngroups > s->lev->ncomp)
		return QK_ERR_BAD_ARG;
	s->rprm = *rad;
	s->src_on = (src != nullptr);
	if (src)
		s->sprm = *src;
	s->esrc.clear();
	if (esrc)
		s->esrc.assign(esrc, esrc + s->nb);
	s->rad_cfl = rad_cfl;
	s->max_substeps = max_substeps;
	s->rad_on = true;
	s->sig_valid = false;
	return 0;
}
extern "C" int qk_sim_last_rad_substeps(const qk_sim *s) { return s ? s->last_nsub : 0; }
extern "C" int64_t qk_sim_rad_cell_updates(const qk_sim *s) { return s ? s->rad_cell_updates : 0; }
extern "C" qk_level *qk_sim_level(qk_sim *s) { return s ? s->lev : nullptr; }
extern "C" void *qk_sim_stream(qk_sim *s) { return s ? (void *)s->stream : nullptr; }
extern "C" double qk_sim_time(const qk_sim *s) { return s ? s->t : 0.0; }
extern "C" int64_t qk_sim_cell_updates(const qk_sim *s) { return s ? s->cell_updates : 0; }
extern "C" int64_t qk_sim_retries(const qk_sim *s) { return s ? s->retries : 0; }
// number of doubles of local box b including ghost cells: ncomp * prod(len+2*nghost)
extern "C" int64_t qk_sim_box_doubles(const qk_sim *s, int b)
{
	if (!s || b < 0 || b >= s->nb)
		return 0;
	return (int64_t)s->snew[b].nstride * s->lev->ncomp;
}
extern "C" int qk_sim_state_desc(qk_sim *s, int which, int b, qk_array4 *out)
{
	if (!s || b < 0 || b >= s->nb || !out)
		return QK_ERR_BAD_ARG;
	*out = (which == 0 ? s->snew : which == 1 ? s->sold : s->sint)[b];
	return 0;
}

// host <-> device transfer of state_new of local box b (whole FAB incl. ghost cells, AMReX layout).
// `host` should be pinned for the copies to be asynchronous; both calls are stream-ordered and qk_sim_sync waits.
extern "C" int qk_sim_set_state(qk_sim *s, int b, const double *host)
{
	if (!s || b < 0 || b >= s->nb || !host)
		return QK_ERR_BAD_ARG;
	s->sig_valid = false;
	QK_CUDA(cudaMemcpyAsync(s->snew[b].p, host, (size_t)qk_sim_box_doubles(s, b) * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	return 0;
}
extern "C" int qk_sim_get_state(qk_sim *s, int b, double *host)
{
	if (!s || b < 0 || b >= s->nb || !host)
		return QK_ERR_BAD_ARG;
	QK_CUDA(cudaMemcpyAsync(host, s->snew[b].p, (size_t)qk_sim_box_doubles(s, b) * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	return 0;
}
// ---- valid cells only: what a host-resident caller actually owns (ghost cells are filled by the step itself) ----
__global__ void __launch_bounds__(256) k_stage_copy(A4 a, Box3 bx, int ncomp, double *__restrict__ stage, int to_state)
{
	const int nx = bx.hi[0] - bx.lo[0] + 1, ny = bx.hi[1] - bx.lo[1] + 1, nz = bx.hi[2] - bx.lo[2] + 1;
	const int64_t total = (int64_t)nx * ny * nz * ncomp;
	for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
		const int64_t row = t / nx;
		const int i = (int)(t - row * nx);
		const int64_t pl = row / ny;
		const int j = (int)(row - pl * ny);
		const int n = (int)(pl / nz), k = (int)(pl - (int64_t)n * nz);
		double *p = a.p + a.off(bx.lo[0] + i, bx.lo[1] + j, bx.lo[2] + k) + (int64_t)n * a.ns;
		if (to_state)
			*p = stage[t];
		else
			stage[t] = *p;
	}
}

static int stage_setup(qk_sim *s)
{
	if (s->stage)
		return 0;
	size_t tot = 0;
	s->stage_off.resize(s->nb);
	for (int b = 0; b < s->nb; ++b) {
		s->stage_off[b] = tot;
		tot += (size_t)Box3(s->lev->valid[b]).ncells() * s->lev->ncomp;
	}
	if (cudaMalloc(&s->stage, std::max<size_t>(tot, 1) * sizeof(double)) != cudaSuccess)
		return QK_ERR_NOMEM;
	QK_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
	s->box_ev.resize(s->nb);
	for (int b = 0; b < s->nb; ++b)
		QK_CUDA(cudaEventCreateWithFlags(&s->box_ev[b], cudaEventDisableTiming));
	return 0;
}

extern "C" int64_t qk_sim_box_valid_doubles(const qk_sim *s, int b)
{
	if (!s || b < 0 || b >= s->nb)
		return 0;
	return Box3(s->lev->valid[b]).ncells() * s->lev->ncomp;
}

// state_new of local box b from / to a contiguous host array (ncomp, nz, ny, nx) of its VALID cells (pinned for asynchronous copies).  The copy
// runs on a separate stream and the (un)packing kernel on the simulation's, so the copy of box b+1 overlaps the kernel of box b.  Ghost cells
// are not transferred: the step fills them (fillBoundaryConditions) before anything reads them.
extern "C" int qk_sim_set_state_valid(qk_sim *s, int b, const double *host)
{
	if (!s || b < 0 || b >= s->nb || !host)
		return QK_ERR_BAD_ARG;
	int rc = stage_setup(s);
	if (rc)
		return rc;
	s->sig_valid = false;
	const int64_t n = qk_sim_box_valid_doubles(s, b);
	double *st = s->stage + s->stage_off[b];
	if (b == 0) { // the staging buffer may still be read by the kernels of an earlier download: order the copy stream behind them
		QK_CUDA(cudaEventRecord(s->ev0, s->stream));
		QK_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev0, 0));
	}
	QK_CUDA(cudaMemcpyAsync(st, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s->copy_stream));
	QK_CUDA(cudaEventRecord(s->box_ev[b], s->copy_stream));
	QK_CUDA(cudaStreamWaitEvent(s->stream, s->box_ev[b], 0));
	k_stage_copy<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s->stream>>>(A4(s->snew[b]), Box3(s->lev->valid[b]), s->lev->ncomp, st, 1);
	QK_KERNEL_CHECK();
	return 0;
}
extern "C" int qk_sim_get_state_valid(qk_sim *s, int b, double *host)
{
	if (!s || b < 0 || b >= s->nb || !host)
		return QK_ERR_BAD_ARG;
	int rc = stage_setup(s);
	if (rc)
		return rc;
	const int64_t n = qk_sim_box_valid_doubles(s, b);
	double *st = s->stage + s->stage_off[b];
	k_stage_copy<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s->stream>>>(A4(s->snew[b]), Box3(s->lev->valid[b]), s->lev->ncomp, st, 0);
	QK_KERNEL_CHECK();
	QK_CUDA(cudaEventRecord(s->box_ev[b], s->stream));
	QK_CUDA(cudaStreamWaitEvent(s->copy_stream, s->box_ev[b], 0));
	QK_CUDA(cudaMemcpyAsync(host, st, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->copy_stream));
	return 0;
}

extern "C" int qk_sim_sync(qk_sim *s)
{
	if (!s)
		return QK_ERR_BAD_ARG;
	QK_CUDA(cudaStreamSynchronize(s->stream));
	if (s->copy_stream)
		QK_CUDA(cudaStreamSynchronize(s->copy_stream));
	return 0;
}
extern "C" void qk_sim_reset_clock(qk_sim *s, double t, double dt_prev)
{
	s->t = t;
	s->dt_prev = dt_prev;
	s->sig_valid = false;
	s->cell_updates = 0;
	s->retries = 0;
}

#define QK_TRY(x)                                                                                                                                    \
	do {                                                                                                                                         \
		int r_ = (x);                                                                                                                        \
		if (r_ != 0)                                                                                                                         \
			return r_;                                                                                                                   \
	} while (0)

static inline double dminh(double a, double b) { return (b < a) ? b : a; }
static inline double dmaxh(double a, double b) { return (a < b) ? b : a; }

// AMRSimulation::computeTimestep for a single level (src/simulation.hpp:703-818)
extern "C" int qk_sim_compute_timestep(qk_sim *s, double stop_time, double *dt_out)
{
	if (!s || !dt_out)
		return QK_ERR_BAD_ARG;
	double smax = 0.0;
	if (s->sig_valid)
		smax = s->sig_local;
	else
		QK_TRY(qk_hydro_max_signal_speed(&s->prm, 0, s->nb, s->lev->valid.data(), s->snew.data(), &smax, s->stream));
	if (!(s->sig_valid && s->sig_is_global)) // the one-pass kernel of the previous step has already reduced its maximum over the ranks on the device
		QK_TRY(qk_comm_allreduce_max_f64(s->comm, &smax, s->stream));
	if (s->rad_on) // std::max(maxSignalRadiation, maxSignalHydro) per cell  src/QuokkaSimulation.hpp:426-433
		smax = dmaxh(s->rprm.c_hat / static_cast<double>(s->max_substeps), smax);
	const double *dx = s->lev->dx;
	const double dx_min = dminh(dminh(dx[0], dx[1]), dx[2]);
	const double hydro_dt = s->cfl * (dx_min / smax);
	double dt_tmp = dminh(hydro_dt, DBL_MAX);
	dt_tmp = dminh(dt_tmp, 1.1 * s->dt_prev);
	double dt_0 = dt_tmp;
	dt_0 = dminh(dt_0, 1.0 * dt_tmp);
	dt_0 = dminh(dt_0, DBL_MAX);
	if (s->t == 0.0)
		dt_0 = dminh(dt_0, DBL_MAX);
	const double eps = 1.e-3 * dt_0;
	if (s->t + dt_0 > stop_time - eps)
		dt_0 = stop_time - s->t;
	s->dt_prev = dt_0;
	*dt_out = dt_0;
	return 0;
}

// QuokkaSimulation::advanceHydroAtLevel (src/QuokkaSimulation.hpp:1032-1322): Uold (ghosts filled in place) -> snew.
// *ok = 0 on FOFC failure or CFL violation.
static int advance_hydro(qk_sim *s, std::vector<qk_array4> &Uold, double dt, int *ok)
{
	qk_level *L = s->lev;
	const int nc = L->ncomp;
	*ok = 1;
	QK_TRY(qk_fill_boundary(L, Uold.data(), 0, nc, s->stream)); // :1076
	int64_t bad = 0;
	if (s->prm.integrator_order == 2) {
		QK_TRY(qk_hydro_advance_stage(L, &s->prm, 1, Uold.data(), Uold.data(), s->sint.data(), dt, &bad, s->stream));
		if (bad > 0 && s->prm.abort_on_fofc_failure) {
			*ok = 0;
			return 0;
		}
		QK_TRY(qk_fill_boundary(L, s->sint.data(), 0, nc, s->stream)); // :1204
		QK_TRY(qk_hydro_advance_stage(L, &s->prm, 2, Uold.data(), s->sint.data(), s->snew.data(), dt, &bad, s->stream));
	} else {
		QK_TRY(qk_hydro_advance_stage(L, &s->prm, 1, Uold.data(), Uold.data(), s->snew.data(), dt, &bad, s->stream));
	}
	if (bad > 0 && s->prm.abort_on_fofc_failure) {
		*ok = 0;
		return 0;
	}
	// isCflViolated (:992-1013)
	// both maxima of state_new (isCflViolated's and the next computeTimestep's) in one pass
	double both[2];
	{
		int rc = qk_fused_max_signal(L, &s->prm, s->snew.data(), both, s->stream); // both maxima come back reduced over all ranks
		s->sig_is_global = true;
		if (rc == QK_ERR_UNSUPPORTED) { // decided by the constants: the same on every rank
			rc = qk_hydro_max_signal_both(&s->prm, s->nb, L->valid.data(), s->snew.data(), both, s->stream);
			s->sig_is_global = false;
		}
		QK_TRY(rc);
	}
	s->sig_local = both[0];
	s->sig_valid = true;
	double smax = both[1];
	if (!s->sig_is_global)
		QK_TRY(qk_comm_allreduce_max_f64(s->comm, &smax, s->stream));
	const double dx_min = dminh(dminh(L->dx[0], L->dx[1]), L->dx[2]);
	const double dt_cfl = s->cfl * (dx_min / smax);
	if (dt > 1.1 * dt_cfl)
		*ok = 0;
	return 0;
}

// advanceSingleTimestepAtLevel + advanceHydroAtLevelWithRetries (:653-707, :885-990).
// The reference copies state_old into state_old_cc_tmp before every attempt (:939-940); the valid cells of
// state_old are never written by the advance, so the first attempt works on state_old directly and the copy
// is only made when a retry with sub-steps needs to overwrite it (:948).
extern "C" int qk_sim_step(qk_sim *s, double dt, int *retries_out)
{
	if (!s)
		return QK_ERR_BAD_ARG;
	s->sig_valid = false;
	std::swap(s->snew, s->sold); // :671
	std::swap(s->pool[0], s->pool[1]);
	int result = -1;
	for (int retry = 0; retry <= 6; ++retry) { // max_retries = 6 (:891)
		const int nsub = 1 << retry;
		const double dt_step = dt / nsub;
		int ok = 1;
		if (nsub == 1) {
			QK_TRY(advance_hydro(s, s->sold, dt_step, &ok));
		} else {
			QK_TRY(alloc_state(s, 3, s->stmp));
			QK_CUDA(cudaMemcpyAsync(s->pool[3], s->pool[1], s->pool_doubles * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
			for (int sub = 0; sub < nsub && ok; ++sub) {
				if (sub > 0) // amrex::Copy(state_old_cc_tmp, state_new, 0, 0, ncompHydro_, nghost) :948
					for (int b = 0; b < s->nb; ++b)
						QK_CUDA(cudaMemcpyAsync(s->stmp[b].p, s->snew[b].p, (size_t)s->snew[b].nstride * (6 + s->prm.nscalars) * sizeof(double),
									cudaMemcpyDeviceToDevice, s->stream));
				QK_TRY(advance_hydro(s, s->stmp, dt_step, &ok));
			}
		}
		if (ok) {
			result = retry;
			break;
		}
	}
	if (retries_out)
		*retries_out = result;
	if (result < 0)
		return 0; // the reference aborts here (:966-989); the caller sees retries = -1
	if (s->rad_on) { // subcycleRadiationAtLevel :690-694; state_inter is free after the hydro advance and serves as U_tmp
		int nsub = 0;
		// the reference reads its failure counters after every substep and aborts (:1692-1711); here they are accumulated over the
		// subcycle and read once per coarse step (one stream synchronisation)
		int64_t counters[QK_RAD_SOURCE_NCOUNTERS] = {};
		QK_TRY(qk_rad_subcycle(s->lev, &s->prm, &s->rprm, s->src_on ? &s->sprm : nullptr, s->sold.data(), s->snew.data(), s->sint.data(),
				       s->esrc.empty() ? nullptr : s->esrc.data(), dt, s->rad_cfl, s->src_on ? counters : nullptr, &nsub, s->stream));
		s->last_nsub = nsub;
		int64_t nfail = counters[4] + counters[5] + counters[6];
		QK_TRY(s->lev->global_sum(&nfail, s->stream));
		s->rad_failures += nfail;
		if (nfail > 0 || nsub > s->max_substeps + 1) { // AMREX_ALWAYS_ASSERT(nsubSteps <= maxSubsteps_ + 1) :1597
			if (retries_out)
				*retries_out = -1;
			return QK_ERR_NOT_CONVERGED;
		}
		s->rad_cell_updates += (int64_t)nsub * s->ncells_global; // radiationCellUpdates_ :1697
		s->sig_valid = false;					 // the source terms changed the gas state
	}
	s->retries += result;
	s->t += dt;
	s->cell_updates += s->ncells_global; // cellUpdates_ += CountCells(lev)  simulation.hpp:1285
	return 0;
}

// AMRSimulation::evolve (simulation.hpp:827-977) without I/O: up to max_steps coarse steps or stop_time.
// elapsed_s is the host wall-clock of the loop (the reference's FOM denominator), device_ms the CUDA-event time.
extern "C" int qk_sim_evolve(qk_sim *s, int max_steps, double stop_time, int *steps_done, double *elapsed_s, double *device_ms)
{
	if (!s)
		return QK_ERR_BAD_ARG;
	QK_CUDA(cudaStreamSynchronize(s->stream));
	const auto t0 = std::chrono::steady_clock::now();
	QK_CUDA(cudaEventRecord(s->ev0, s->stream));
	int n = 0;
	for (; n < max_steps && s->t < stop_time; ++n) {
		double dt;
		QK_TRY(qk_sim_compute_timestep(s, stop_time, &dt));
		int r;
		QK_TRY(qk_sim_step(s, dt, &r));
		if (r < 0) {
			if (steps_done)
				*steps_done = n;
			return QK_ERR_UNSUPPORTED;
		}
	}
	QK_CUDA(cudaEventRecord(s->ev1, s->stream));
	QK_CUDA(cudaStreamSynchronize(s->stream));
	const auto t1 = std::chrono::steady_clock::now();
	float ms = 0;
	QK_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
	if (steps_done)
		*steps_done = n;
	if (elapsed_s)
		*elapsed_s = std::chrono::duration<double>(t1 - t0).count();
	if (device_ms)
		*device_ms = ms;
	return 0;
}
