// qk_sweep.cu -- the FUSED stage of QuokkaSimulation::advanceHydroAtLevel (src/QuokkaSimulation.hpp:1099-1198,
// 1202-1285): per RK stage
//
//   k_fprim    HydroSystem::ConservedToPrimitive (ng = 4)                                     hydro_system.hpp:138-196
//   k_fchi     ComputeFlatteningCoefficients<X1,X2,X3> in ONE pass (ng = 2), sharing rho c_s^2  :531-626
//   k_fchimin  the 9-point min of FlattenShocks, once per cell instead of once per direction    :655-669
//   k_sweep_x  PPM -> flatten -> HLLC -> flux difference along x; lane <-> cell, neighbour states and fluxes
//              exchanged with warp shuffles; nothing but 0.5*F (RK2 average) and the RHS is written
//   k_sweep_m  the same along y / z by MARCHING: lane <-> x (coalesced), each thread walks a pencil segment and
//              keeps the previous face's flux and the previous cell's right state in registers, so every PPM
//              parabola and every Riemann problem is evaluated exactly once; the z sweep carries the epilogue
//              (ComputeRhsFromFluxes sum, AddInternalEnergyPdV, PredictStep, redoFlag count, EnforceLimits,
//              SyncDualEnergy) and writes the new state -- no rhs / flux / face-velocity MultiFabs exist.
//
// All local boxes go into one launch per kernel.  Arithmetic is the reference's (see qk_fast.cuh); the association
// order of the three directional contributions is that of ComputeRhsFromFluxes (x, then +y, then +z) and of
// AddInternalEnergyPdV's div v.  A stage whose result flags a cell (redoFlag) or contains a non-finite value is
// REDONE by the faithful path of qk_level.cu, which owns the first-order flux correction.
#include "qk_sweep_kernels.cuh"

#include <cuda.h>
#include <stdio.h> // CUtensorMap (the encoder lives in qk_level.cu: qk_encode_tile)


// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// qk_sweep_relaxed.cu: the same kernels instantiated with relaxed arithmetic (FMA contraction, closed-form EOS)
int qk_sweep_stage_relaxed(int ns, bool reint, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5], int stage,
			   bool dual, bool tma, cudaStream_t s);

int qk_sweep_stage_relaxed_plm(int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5], int stage, bool dual,
			       cudaStream_t s);
// qk_sweep_keepf.cu / qk_sweep_relaxed_keepf.cu: the TMA-staged kernels instantiated with KEEPF = true (they also store the stage's own face
// fluxes into SweepBox::fo for incrementFluxRegisters); arith selects the translation unit, order 2 = the PLM instantiation
int qk_sweep_stage_keepf(int ns, bool reint, int order, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb, const int maxn[5],
			 int stage, bool dual, cudaStream_t s);
int qk_sweep_stage_relaxed_keepf(int ns, bool reint, int order, int ng, unsigned long long *d_counters, const FastConst &c, const void *boxes, const void *tmaps, int nb,
				 const int maxn[5], int stage, bool dual, cudaStream_t s);

static_assert(sizeof(CUtensorMap) == TMAP_BYTES, "descriptor size");
static bool encode_tile(CUtensorMap *m, const qk_array4 &a, unsigned bx, unsigned by, unsigned bz, unsigned bc) { return qk_encode_tile(m, a, bx, by, bz, bc); }

struct FusedState {
	int nv = 0; // 6 + nscalars the scratch was built for
	std::vector<qk_array4> prim, chi3, rhs, hF[3];
	std::vector<qk_array4> fo[3]; // kept face fluxes of the last stage (tight rows; allocated on the first flux-keeping stage)
	bool fo_valid = false;	      // the last stage of this level ran fused with KEEPF and was not redone: fo[] are its fluxes
	SweepBox *d_boxes = nullptr;
	SweepBox *h_boxes = nullptr; // pinned staging, one table per in-flight stage (ring of 8)
	CUtensorMap *h_maps = nullptr;			   // TM_COUNT descriptors per box of the stage being launched (host memory)
	std::vector<CUtensorMap> static_maps;		   // descriptors of the level's own scratch (encoded once); empty: TMA staging unavailable
	bool maps_ok = false;
	int ring = 0;
	cudaEvent_t ev[8];
	bool ev_used[8];
	bool tainted = false; // a previous stage left flagged cells in the state: stay on the faithful path
	int tma_agreed = -1;  // N ranks: -1 not negotiated yet, else min over ranks of "my rows can be bulk-copied" (the kernel family is a collective choice)
};

void qk_fused_free(qk_level *L)
{
	if (!L->fused)
		return;
	FusedState *F = L->fused;
	if (F->d_boxes)
		cudaFree(F->d_boxes);
	if (F->h_maps)
		free(F->h_maps);
	if (F->h_boxes) {
		cudaFreeHost(F->h_boxes);
		for (int i = 0; i < 8; ++i)
			cudaEventDestroy(F->ev[i]);
	}
	delete F;
	L->fused = nullptr;
}

// one pass over the local boxes: out[0] = ComputeMaxSignalSpeed + norminf (simulation.hpp:709-710), out[1] = maxSignalSpeedLocal
// (isCflViolated), both already reduced over all ranks of the level's communicator; returns QK_ERR_UNSUPPORTED when the constants leave the shared-reciprocal domain (caller uses qk_ops.cu's kernels)
int qk_fused_max_signal(qk_level *L, const qk_hydro_params *prm, const qk_array4 *state, double out[2], cudaStream_t s)
{
	FastConst c;
	if (!make_fast_const(prm, L->dx, 0.0, &c))
		return QK_ERR_UNSUPPORTED;
	QK_TRY(L->ensure_counters());
	QK_CUDA(cudaMemsetAsync(L->d_counters + 4, 0, 16, s));
	{
		ProfScope p("max_signal_speed", s);
		const int nb = (int)L->valid.size();
		for (int b0 = 0; b0 < nb; b0 += SigBoxes::MAXB) {
			SigBoxes sb;
			const int n = std::min(SigBoxes::MAXB, nb - b0);
			int64_t maxcells = 1;
			for (int b = 0; b < n; ++b) {
				sb.u[b] = A4(state[b0 + b]);
				sb.bx[b] = Box3(L->valid[b0 + b]);
				maxcells = std::max(maxcells, sb.bx[b].ncells());
			}
			const unsigned per_box = (unsigned)std::max<int64_t>(1, std::min<int64_t>((maxcells + 255) / 256, (148 * 16 + n - 1) / n));
			if (prm->arith == QK_ARITH_FAST) // closed-form EOS, as the relaxed sweeps (this unit is compiled without FMA contraction)
				k_signal<1><<<dim3(per_box, n), 256, 0, s>>>(c, sb, L->d_counters + 4);
			else
				k_signal<0><<<dim3(per_box, n), 256, 0, s>>>(c, sb, L->d_counters + 4);
			QK_KERNEL_CHECK();
		}
	}
	// N ranks: the two maxima are order-preserving 64-bit keys, so the global maxima are one in-place device all-reduce (no extra host round trip)
	if (L->comm && L->nranks > 1)
		QK_TRY(qk_comm_allreduce_dev_u64(L->comm, L->d_counters + 4, 2, 1, s));
	QK_CUDA(cudaMemcpyAsync(L->h_counters + 4, L->d_counters + 4, 16, cudaMemcpyDeviceToHost, s));
	QK_CUDA(cudaStreamSynchronize(s));
	auto key2d = [](unsigned long long k) {
		unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
		double v;
		memcpy(&v, &b, 8);
		return v;
	};
	out[0] = (L->h_counters[4] == 0ull) ? 0.0 : key2d(L->h_counters[4]);
	out[1] = (L->h_counters[5] == 0ull) ? -1.7976931348623157e308 : key2d(L->h_counters[5]);
	return 0;
}

void qk_fused_untaint(qk_level *L)
{
	if (L->fused)
		L->fused->tainted = false;
}

static int fused_setup(qk_level *L, int nv)
{
	if (L->fused && L->fused->nv == nv)
		return 0;
	if (L->fused)
		return QK_ERR_UNSUPPORTED;
	FusedState *F = new FusedState();
	L->fused = F;
	const int nb = (int)L->valid.size();
	int rc = 0;
	rc = rc ? rc : L->alloc_fabs(F->prim, nv + 1, L->nghost, -1, true);
	rc = rc ? rc : L->alloc_fabs(F->chi3, 3, 2, -1);
	rc = rc ? rc : L->alloc_fabs(F->rhs, nv + 1, 0, -1, true);
	for (int d = 0; d < 3 && !rc; ++d)
		rc = L->alloc_fabs(F->hF[d], nv + 1, 0, d, true);
	if (rc)
		return rc;
	QK_CUDA(cudaMalloc(&F->d_boxes, sizeof(SweepBox) * nb * 8));
	QK_CUDA(cudaMallocHost(&F->h_boxes, sizeof(SweepBox) * nb * 8));
	F->h_maps = static_cast<CUtensorMap *>(aligned_alloc(64, sizeof(CUtensorMap) * TM_COUNT * std::max(nb, 1)));
	if (!F->h_maps)
		return QK_ERR_NOMEM;
	// descriptors of the scratch arrays never change
	F->static_maps.assign((size_t)TM_COUNT * nb, CUtensorMap{});
	F->maps_ok = true;
	for (int b = 0; b < nb && F->maps_ok; ++b) {
		CUtensorMap *m = &F->static_maps[(size_t)TM_COUNT * b];
		const unsigned nc = (unsigned)nv + 1;
		F->maps_ok = encode_tile(&m[TM_PRIM_M36], F->prim[b], 36, 1, 1, nc) && encode_tile(&m[TM_PRIM_R32], F->prim[b], 32, 1, 1, 1) &&
			     encode_tile(&m[TM_PRIM_X38], F->prim[b], 38, 1, 1, nc) && encode_tile(&m[TM_PRIM_Y3], F->prim[b], 34, 3, 1, 1) &&
			     encode_tile(&m[TM_PRIM_Z3], F->prim[b], 34, 1, 3, 1) && encode_tile(&m[TM_RHS], F->rhs[b], 32, 1, 1, nc) &&
			     encode_tile(&m[TM_HF0], F->hF[0][b], 32, 1, 1, nc) && encode_tile(&m[TM_HF1], F->hF[1][b], 32, 1, 1, nc) &&
			     encode_tile(&m[TM_HF2], F->hF[2][b], 32, 1, 1, nc) && encode_tile(&m[TM_R0], F->hF[2][b], 32, 1, 1, (unsigned)nv) &&
			     encode_tile(&m[TM_PRIM_X40], F->prim[b], 40, 1, 1, nc) && encode_tile(&m[TM_PRIM_Y3W], F->prim[b], 36, 3, 1, 1) &&
			     encode_tile(&m[TM_PRIM_Z3W], F->prim[b], 36, 1, 3, 1);
	}
	for (int i = 0; i < 8; ++i) {
		QK_CUDA(cudaEventCreateWithFlags(&F->ev[i], cudaEventDisableTiming));
		F->ev_used[i] = false;
	}
	F->nv = nv;
	return 0;
}

// face-flux arrays of the last stage if it ran on the fused flux-keeping path (else nullptr: the faithful path's scr.flx hold them)
const std::vector<qk_array4> *qk_fused_kept_fluxes(const qk_level *L, int dir)
{
	return (L->fused && L->fused->fo_valid) ? &L->fused->fo[dir] : nullptr;
}

int qk_fused_stage(qk_level *L, const qk_hydro_params *prm, int stage, const qk_array4 *U0, const qk_array4 *Ustage, const qk_array4 *Uout, double dt,
		   int64_t *ncells_bad, cudaStream_t s, bool *handled, bool keepf)
{
	*handled = false;
	if (L->fused)
		L->fused->fo_valid = false;
	// configurations the fused kernels are instantiated for; anything else runs the faithful path
	const int ns = prm->nscalars, nms = prm->nmscalars;
	const bool inst = (ns == 0 && nms == 0) || (ns == 1 && nms == 0) || (ns == 3 && nms == 2);
	const int order = prm->reconstruction_order;
	// PPM for every instantiated trait set; PLM (minmod) for the scalar-free, reconstruct_eint = false set (config C4's hydro)
	const bool can = (order == 3 || (order == 2 && ns == 0 && !prm->reconstruct_eint)) && inst && prm->use_dual_energy && L->nghost >= 4 && prm->K_visc == 0.0 &&
			 prm->gamma != 1.0; // the isothermal EOS (is_eos_isothermal(), hydro_system.hpp:133) is built on the operator path only
	if (!can) {
		// say so once: the one-kernel-per-operator path is 3-4x slower and a maintainer flipping a run-time switch should know
		static bool warned = false;
		if (!warned && getenv("QK_QUIET") == nullptr) {
			warned = true;
			fprintf(stderr,
				"[quokka_b200] this configuration takes the one-kernel-per-operator path (fused sweeps exist for PPM, for PLM without scalars and "
				"reconstruct_eint, with dual energy, 4 ghost cells, K_visc = 0 and gamma != 1): reconstruction_order = %d, nscalars = %d, K_visc = %g, gamma = %g\n",
				order, ns, prm->K_visc, prm->gamma);
		}
		return 0;
	}
	if (L->fused && L->fused->tainted)
		return 0;
	FastConst c;
	if (!make_fast_const(prm, L->dx, dt, &c))
		return 0;
	const int nv = 6 + ns;
	QK_TRY(fused_setup(L, nv));
	QK_TRY(L->ensure_counters());
	FusedState *F = L->fused;
	const int nb = (int)L->valid.size();
	if (keepf && F->fo[0].empty()) {
		for (int d = 0; d < 3; ++d)
			QK_TRY(L->alloc_fabs(F->fo[d], nv, 0, d));
	}
	// box table through the pinned ring
	const int slot = F->ring;
	F->ring = (F->ring + 1) % 8;
	if (F->ev_used[slot])
		QK_CUDA(cudaEventSynchronize(F->ev[slot]));
	SweepBox *hb = F->h_boxes + (size_t)slot * nb;
	int maxn[5] = {1, 1, 1, 1 << 30, 0}; // largest extents; smallest nx; largest slot count of the concatenated x sweep (launch_stage)
	int64_t max_slots = 0, total_slots = 0;
	bool tma = (getenv("QK_NO_TMA") == nullptr) && F->maps_ok;
	CUtensorMap *hm = F->h_maps;
	for (int b = 0; b < nb; ++b) {
		// the tile descriptors: the level's own scratch (encoded once) + the caller's U0 (its arrays rotate from step to step)
		if (tma) {
			memcpy(hm + (size_t)TM_COUNT * b, &F->static_maps[(size_t)TM_COUNT * b], sizeof(CUtensorMap) * TM_COUNT);
			tma = encode_tile(&hm[(size_t)TM_COUNT * b + TM_U0], U0[b], 32, 1, 1, (unsigned)nv);
		}
		SweepBox &B = hb[b];
		B.U0 = A4(U0[b]);
		B.Us = A4(Ustage[b]);
		B.Uo = A4(Uout[b]);
		B.prim = A4(F->prim[b]);
		B.chi3 = A4(F->chi3[b]);
		B.rhs = A4(F->rhs[b]);
		for (int d = 0; d < 3; ++d) {
			B.hF[d] = A4(F->hF[d][b]);
			if (keepf)
				B.fo[d] = A4(F->fo[d][b]);
			B.lo[d] = L->valid[b].lo[d];
			B.hi[d] = L->valid[b].hi[d];
			maxn[d] = std::max(maxn[d], B.hi[d] - B.lo[d] + 1);
		}
		maxn[3] = std::min(maxn[3], B.hi[0] - B.lo[0] + 1);
		const int64_t slots = (int64_t)(B.hi[1] - B.lo[1] + 1) * (B.hi[2] - B.lo[2] + 1) * (B.hi[0] - B.lo[0] + 3);
		max_slots = std::max<int64_t>(max_slots, slots);
		total_slots += slots;
	}
	maxn[4] = (max_slots < (int64_t(1) << 31) - 64 && (L->nghost & 1) == 0) ? (int)max_slots : 0; // (its windows start at even array columns)
	{
		// The concatenated x sweep pays off where the launch fills the GPU several times over (256^3 per GPU: 570 k tiles, x sweep 7 % faster).
		// On the small levels of an AMR hierarchy (a few 10 k tiles: the Sod tube of config C1, a 64^3 Sedov) its tiles, most of which wait for
		// two staged windows, are latency-bound and the per-row tiles are up to 1.5x faster (measured inside the reference's driver,
		// profiles/r02_ref_cuda_amr_final.json vs r02_ref_cuda_amr_xcat0.json): below QK_XCAT_MIN_TILES (default 65536) tiles the level keeps them.
		const char *e = getenv("QK_XCAT_MIN_TILES"); // read per call: the parity tests switch it
		const long long min_tiles = e ? atoll(e) : 65536ll;
		if (total_slots / 30 < min_tiles)
			maxn[4] = 0;
	}
	if (L->comm && L->nranks > 1) {
		// Every rank must run the same kernel family: a rank on the faithful path pairs different collectives than one on the fused path.
		// The alignment of the local rows is negotiated once per level; a rank whose rows later stop qualifying fails loudly instead of hanging.
		if (F->tma_agreed < 0) {
			int64_t v = tma ? 0 : 1; // sum of "cannot" over ranks
			QK_TRY(L->global_sum(&v, s));
			F->tma_agreed = (v == 0) ? 1 : 0;
		}
		if (F->tma_agreed == 1 && !tma)
			return QK_ERR_UNSUPPORTED;
		tma = (F->tma_agreed == 1);
	}
	SweepBox *db = F->d_boxes + (size_t)slot * nb;
	const unsigned char *dm = reinterpret_cast<const unsigned char *>(hm); // HOST table: the launchers copy what a kernel uses into its parameters
	QK_CUDA(cudaMemcpyAsync(db, hb, sizeof(SweepBox) * nb, cudaMemcpyHostToDevice, s));
	QK_CUDA(cudaEventRecord(F->ev[slot], s));
	F->ev_used[slot] = true;
	QK_CUDA(cudaMemsetAsync(L->d_counters, 0, 32, s));

	const bool dual = (prm->integrator_order == 2);
	int rc;
	if ((order == 2 || keepf) && !tma)
		return 0; // the PLM and the flux-keeping kernels exist in the TMA-staged form only: the faithful path takes the stage (*handled stays false)
	if (keepf)
		rc = (prm->arith == QK_ARITH_FAST)
			 ? qk_sweep_stage_relaxed_keepf(ns, prm->reconstruct_eint != 0, order, L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, s)
			 : qk_sweep_stage_keepf(ns, prm->reconstruct_eint != 0, order, L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, s);
	else if (order == 2)
		rc = (prm->arith == QK_ARITH_FAST) ? qk_sweep_stage_relaxed_plm(L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, s)
						   : sweep_stage_dispatch_plm<0>(L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, s);
	else if (prm->arith == QK_ARITH_FAST && tma) // the relaxed kernels exist in the TMA-staged form only
		rc = qk_sweep_stage_relaxed(ns, prm->reconstruct_eint != 0, L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, tma, s);
	else
		rc = sweep_stage_dispatch<0>(ns, prm->reconstruct_eint != 0, L->nghost, L->d_counters, c, db, dm, nb, maxn, stage, dual, tma, s);
	QK_TRY(rc);
	// redoFlag.sum() over all ranks (QuokkaSimulation.hpp:1146): in-place device all-reduce of the two counters, then the one D2H copy
	if (L->comm && L->nranks > 1)
		QK_TRY(qk_comm_allreduce_dev_u64(L->comm, L->d_counters, 2, 0, s));
	QK_CUDA(cudaMemcpyAsync(L->h_counters, L->d_counters, 32, cudaMemcpyDeviceToHost, s));
	QK_CUDA(cudaStreamSynchronize(s));
	const int64_t flagged = (int64_t)(L->h_counters[0] + L->h_counters[1]);
	if (flagged > 0) {
		// redo the whole stage with the faithful path (FOFC lives there); it recomputes F(U0) itself in stage 2
		L->scr.rk_valid = false;
		int64_t bad = 0;
		QK_TRY(L->faithful_stage(prm, stage, U0, Ustage, Uout, dt, &bad, s));
		if (bad > 0)
			F->tainted = true;
		if (ncells_bad)
			*ncells_bad = bad;
	} else {
		F->fo_valid = keepf;
		if (ncells_bad)
			*ncells_bad = 0;
	}
	*handled = true;
	return 0;
}
