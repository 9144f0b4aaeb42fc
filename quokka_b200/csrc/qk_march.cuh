// qk_march.cuh -- the y / z sweeps with TMA-staged pencils (included by qk_sweep.cu, inside its anonymous namespace).
//
// Same mapping and arithmetic as k_sweep_m (lane <-> x, one warp walks a 32-cell segment of a pencil along DIR and
// keeps the previous face's flux and the previous cell's right state in registers), but no thread ever issues a
// global LOAD: every warp runs its own producer/consumer pipeline in shared memory --
//
//   prim ring   5 slots, one per cell row along DIR: rho, vy, vz, P, Eaux [, scalars], chi_min rows of 32 and a 36-wide
//               vx row (x neighbours for the carbuncle term); row r+3 is requested while row r is being computed
//   trans ring  2 slots: the two transverse rows of the other velocity component (z +- 1 for the y sweep, y +- 1 for z)
//   aux slot    what the update of cell r-1 needs at the END of step r: 0.5*F(U0) of face r (stage 2), the partial
//               RHS of the x (+y) sweeps, and U0 (z sweep: epilogue) -- requested at the top of the step
//
// each slot is filled by 1-D bulk copies of the TMA engine (cp.async.bulk -> UBLKCP) issued by lane 0 and completing on
// an mbarrier the whole warp waits on.  Global memory latency is therefore hidden behind a full step (~1.5 k
// instructions) of PPM + HLLC arithmetic instead of being exposed at the top of every step, with no registers spent
// on prefetch buffers.  Requires 16-byte aligned rows: even pitches / component strides and an even ghost offset
// (the level's own scratch is allocated that way; the caller's U0 is checked on the host, else k_sweep_m runs).
#pragma once
#include "qk_tma.cuh"

#ifndef QK_MARCH_Y_BLOCKS
#define QK_MARCH_Y_BLOCKS 4 // resident CTAs per SM of the relaxed y sweep: 5 (96 registers) was measured slower (spills): 2.08 vs 1.84 ms per step
#endif

// R0 (relaxed arithmetic only): instead of the three 0.5*F(U0) face arrays, stage 1 keeps its complete right-hand side
// R(U0) per cell (in the z-face scratch array) and stage 2 forms 0.5*R(U0) + 0.5*R(U1): algebraically the same update
// (the RK2 flux average is linear), 120 B/cell less traffic in each stage, but not the reference's rounding order.
//
// Staging is by TENSOR-MAP TMA (cp.async.bulk.tensor.4d, qk_tma.cuh): one copy brings a whole [36 x (NV+1)] row tile of the primitives
// (cells i0-2 .. i0+33: the x neighbours of the carbuncle term ride in the same tile), one copy each the transverse rows, the partial RHS,
// U0 and the stage-1 data -- 7 copies per marching step instead of one 1-D bulk copy per component row (23-35), which removes most of the
// uniform-datapath address arithmetic from the issue stream.  Descriptors: TM_COUNT CUtensorMaps per box, encoded on the host (qk_sweep.cu).
enum SweepTmap {
	TM_PRIM_M36 = 0, // prim   box {36, 1, 1, NV+1}
	TM_PRIM_R32,	 // prim   box {32, 1, 1, 1}     one component row (transverse velocity rows of the marching sweeps)
	TM_PRIM_X38,	 // prim   box {38, 1, 1, NV+1}  x sweep window
	TM_PRIM_Y3,	 // prim   box {34, 3, 1, 1}     rows j-1 .. j+1 of one component (x sweep)
	TM_PRIM_Z3,	 // prim   box {34, 1, 3, 1}     rows k-1 .. k+1
	TM_RHS,		 // rhs    box {32, 1, 1, NV+1}
	TM_HF0,		 // hF[0]  box {32, 1, 1, NV+1}
	TM_HF1,
	TM_HF2,
	TM_R0,	  // hF[2]  box {32, 1, 1, NV}     R(U0) of the relaxed mode
	TM_U0,	  // U0     box {32, 1, 1, NV}
	// the concatenated x sweep (k_sweep_xc): its windows start at an arbitrary cell and a tensor-map copy of FP64 must start at an EVEN x
	// coordinate (16 bytes; an odd one raises "illegal instruction" on sm_100a -- scripts/diag/tma_coord_test.cu), so they are one pair wider
	TM_PRIM_X40, // prim   box {40, 1, 1, NV+1}
	TM_PRIM_Y3W, // prim   box {36, 3, 1, 1}
	TM_PRIM_Z3W, // prim   box {36, 1, 3, 1}
	TM_COUNT
};
// The descriptors travel as __grid_constant__ kernel PARAMETERS (the one place a TMA descriptor needs no proxy fence and cannot go stale in the
// per-SM descriptor cache): a launch covers at most TMAP_MAXB boxes (qk_tma.cuh), each kernel family carries only the descriptors it uses.
enum { MM_PRIM = 0, MM_ROW, MM_RHS, MM_STAGE2, MM_U0, MM_COUNT }; // marching sweeps: prim tile, one-component row, rhs, hF[DIR] | R(U0), U0
enum { XM_PRIM = 0, XM_Y3, XM_Z3, XM_HF, XM_COUNT };		   // x sweep
struct MarchMaps {
	TmapBytes m[TMAP_MAXB][MM_COUNT];
};
struct XMaps {
	TmapBytes m[TMAP_MAXB][XM_COUNT];
};

template <int NV, int STAGE, bool LAST, bool R0> struct MarchSmem {
	static constexpr int NR = 4;
	static constexpr int XW = 36;					// cells i0-2 .. i0+33 of every component (+ chi)
	static constexpr int PR = (((NV + 1) * XW + 15) / 16) * 16;	// one prim slot, padded to a 128-byte multiple
	static constexpr int TR = 64;
	static constexpr int AUX_HF = 0;
	static constexpr int HF_ROWS = (STAGE == 2) ? (R0 ? (LAST ? NV : 0) : NV + 1) : 0;
	static constexpr int AUX_RHS = HF_ROWS * 32;
	static constexpr int AUX_U0 = AUX_RHS + (NV + 1) * 32;
	static constexpr int AUX = AUX_U0 + (LAST ? NV * 32 : 0);
	static constexpr int WARP_DOUBLES = NR * PR + TR + AUX;
	static constexpr int WARP_BYTES = WARP_DOUBLES * 8 + 128; // + mbarriers: [0..3] prim, [4] trans, [5] aux (every TMA destination stays 128-byte aligned)
	static constexpr int BLOCK_BYTES = 4 * WARP_BYTES;
};

template <int ARITH, int DIR, int NS, int NMS, bool REINT, int STAGE, bool DUAL, bool LAST, int ORDER = 3, bool KEEPF = false>
__global__ void __launch_bounds__(128, (ARITH == 1 && !LAST) ? QK_MARCH_Y_BLOCKS : 4) k_march_t(FastConst c, const SweepBox *__restrict__ boxes, const __grid_constant__ MarchMaps tmaps, int nseg, unsigned long long *__restrict__ counters)
{
	constexpr int NV = 6 + NS;
	constexpr bool R0 = (ARITH == 1);
	using SM = MarchSmem<NV, STAGE, LAST, R0>;
	constexpr int TD = (DIR == 1) ? 2 : 1;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); // warp-uniform in the compiler's eyes
	double *const prim_s = reinterpret_cast<double *>(smem_raw + (size_t)warp * SM::WARP_BYTES);
	double *const trans_s = prim_s + SM::NR * SM::PR;
	double *const aux_s = trans_s + SM::TR;
	uint64_t *const bars = reinterpret_cast<uint64_t *>(aux_s + SM::AUX); // [0..3] prim, [4] trans, [5] aux

	const int box = blockIdx.z / nseg, seg = blockIdx.z - box * nseg;
	const SweepBox &B = boxes[box];
	const int i0 = B.lo[0] + blockIdx.x * 32;
	const int t = B.lo[TD] + blockIdx.y * 4 + warp;
	const int s0 = B.lo[DIR] + seg * SEG;
	int bad_cnt = 0, nf_cnt = 0;
	if (i0 <= B.hi[0] && t <= B.hi[TD] && s0 <= B.hi[DIR]) {
		const int nact = min(32, B.hi[0] - i0 + 1);
		const bool active = lane < nact;
		const int s1 = min(s0 + SEG, B.hi[DIR] + 1);
		const int i = i0 + lane;
		const A4 &q = B.prim;
		const A4 &h = B.hF[DIR];
		const A4 &rh = B.rhs;
		const A4 &u0 = B.U0;
		const A4 &uo = B.Uo;
		const TmapBytes *const M = tmaps.m[box]; // this box's descriptors (parameter space)
		if (lane == 0) {
#pragma unroll
			for (int b = 0; b < 6; ++b)
				mbar_init(&bars[b], 1);
			mbar_init_fence();
		}
		__syncwarp();

		// tile coordinates (warp-uniform): x of the tile, the fixed transverse index and the origin of the marching index in each array
		const int64_t shN = (DIR == 1) ? h.js : h.ks, srN = (DIR == 1) ? rh.js : rh.ks, soN = (DIR == 1) ? uo.js : uo.ks;
		const int qx = i0 - 2 - q.bx, qT = (DIR == 1) ? t - q.bz : t - q.by, qN0 = (DIR == 1) ? q.by : q.bz;
		const int hx = i0 - h.bx, hT = (DIR == 1) ? t - h.bz : t - h.by, hN0 = (DIR == 1) ? h.by : h.bz;
		const int rx = i0 - rh.bx, rT = (DIR == 1) ? t - rh.bz : t - rh.by, rN0 = (DIR == 1) ? rh.by : rh.bz;
		const int ux = i0 - u0.bx, uT = (DIR == 1) ? t - u0.bz : t - u0.by, uN0 = (DIR == 1) ? u0.by : u0.bz;
		int row_q = s0 - 3; // next prim row to stage
		int row_t = s0 - 1; // next cell whose transverse rows are staged
		// one tile copy: marching coordinate nn, transverse coordinate tt
		auto tile = [&](double *dst, int which, int x, int nn, int tt, int comp, uint64_t *bar) {
			tma_tile_g2s(dst, &M[which], x, (DIR == 1) ? nn : tt, (DIR == 1) ? tt : nn, comp, bar);
		};
		auto issue_prim = [&](int slot) { // row row_q, all components, cells i0-2 .. i0+33
			uint64_t *bar = &bars[slot];
			mbar_arrive_expect_tx(bar, (unsigned)(NV + 1) * SM::XW * 8u);
			tile(prim_s + slot * SM::PR, MM_PRIM, qx, row_q - qN0, qT, 0, bar);
		};
		auto issue_trans = [&]() { // the other transverse velocity component either side of cell row_t (y sweep: v_z at z -+ 1; z sweep: v_y at y -+ 1)
			uint64_t *bar = &bars[4];
			mbar_arrive_expect_tx(bar, 2u * 256u);
			tile(trans_s, MM_ROW, qx + 2, row_t - qN0, qT - 1, (DIR == 1) ? 3 : 2, bar);
			tile(trans_s + 32, MM_ROW, qx + 2, row_t - qN0, qT + 1, (DIR == 1) ? 3 : 2, bar);
		};
		// what the end of step r reads: hF of face r (stage 2), rhs (+U0) of cell r-1
		auto aux_bytes = [&](int r) -> unsigned {
			unsigned b = 0;
			if (r >= s0) {
				if (STAGE == 2 && !R0)
					b += (unsigned)(NV + 1) * 256u;
				if (r > s0)
					b += (unsigned)(NV + 1) * 256u + (LAST ? (unsigned)NV * 256u : 0u) + ((STAGE == 2 && R0 && LAST) ? (unsigned)NV * 256u : 0u);
			}
			return b;
		};
		auto issue_aux = [&](int r) {
			uint64_t *bar = &bars[5];
			mbar_arrive_expect_tx(bar, aux_bytes(r));
			if (STAGE == 2 && !R0)
				tile(aux_s + SM::AUX_HF, MM_STAGE2, hx, r - hN0, hT, 0, bar); // 0.5 F(U0) of face r
			if (r > s0) {
				if (STAGE == 2 && R0 && LAST)
					tile(aux_s + SM::AUX_HF, MM_STAGE2, hx, r - 1 - hN0, hT, 0, bar); // R(U0) of cell r-1
				tile(aux_s + SM::AUX_RHS, MM_RHS, rx, r - 1 - rN0, rT, 0, bar);
				if (LAST)
					tile(aux_s + SM::AUX_U0, MM_U0, ux, r - 1 - uN0, uT, 0, bar);
			}
		};
		// component n (n = NV: chi) of this lane's cell in a prim slot (cell i0 + lane sits at entry lane + 2 of a 36-wide row)
		auto PV = [&](const double *sl, int n) -> double { return sl[n * SM::XW + lane + 2]; };

		// prologue: rows s0-3 .. s0 into slots 0 .. 3, transverse rows of cell s0-1
		for (int sl = 0; sl < SM::NR; ++sl) {
			if (elect_one())
				issue_prim(sl);
			++row_q;
		}
		if (elect_one())
			issue_trans();
		++row_t;
		// rows s0-3 .. s0 -> unlimited interface value at the low face of cell s0-1
#pragma unroll
		for (int b = 0; b < 4; ++b)
			mbar_wait(&bars[b], 0);
		double apL[NV], ifl[NV], Gp[NV + 1];
		double mVp = 0, mWp = 0, vNp = 0;
		if (active && ORDER == 3) {
#pragma unroll
			for (int n = 0; n < NV; ++n)
				ifl[n] = ppm_iface(PV(prim_s, n), PV(prim_s + SM::PR, n), PV(prim_s + 2 * SM::PR, n), PV(prim_s + 3 * SM::PR, n));
		}
		__syncwarp();
		if (elect_one())
			issue_prim(0); // row s0+1 replaces row s0-3
		++row_q;
		unsigned aux_phase = 0;
		int64_t o_h = h.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		int64_t o_r = rh.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		int64_t o_o = uo.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		const A4 &fk = B.fo[DIR];
		const int64_t sfN = (DIR == 1) ? fk.js : fk.ks;
		int64_t o_f = KEEPF ? fk.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1) : 0; // face r of the kept-flux array
		// k4: slot of row r+2 (the newest row a step reads), par4 its phase parity; rows r+1, r, r-1 sit in the slots before it
		int k4 = 0;
		unsigned par4 = 1, tpar = 0;

		for (int r = s0 - 1; r <= s1; ++r) {
			__syncwarp(); // the aux slot is free again
			const bool have_aux = aux_bytes(r) != 0;
			if (have_aux && elect_one())
				issue_aux(r);
			const double *s_p2 = prim_s + k4 * SM::PR;
			const double *s_p1 = prim_s + ((k4 + 3) & 3) * SM::PR;
			const double *s_0 = prim_s + ((k4 + 2) & 3) * SM::PR;
			const double *s_m1 = prim_s + ((k4 + 1) & 3) * SM::PR;
			mbar_wait(&bars[k4], par4);
			double am[NV], ap[NV];
			double vN0 = 0, mV = 0, mW = 0;
			if (active) {
				const double chi = s_0[NV * SM::XW + lane + 2], omchi = 1. - chi;
#pragma unroll
				for (int n = 0; n < NV; ++n) {
					const double qm1 = PV(s_m1, n), q0 = PV(s_0, n), qp1 = PV(s_p1, n), qp2 = PV(s_p2, n);
					if (n == 1 + DIR)
						vN0 = q0;
					if (ORDER == 3) {
						const double ifh = ppm_iface(qm1, q0, qp1, qp2);
						if (ARITH == 1)
							r_ppm_flat(qm1, q0, qp1, ifl[n], ifh, chi, omchi, am[n], ap[n]);
						else
							f_ppm_flat(qm1, q0, qp1, ifl[n], ifh, chi, omchi, am[n], ap[n]);
						ifl[n] = ifh;
					} else {
						f_plm_flat(qm1, q0, qp1, chi, omchi, am[n], ap[n]);
					}
				}
			}
			mbar_wait(&bars[4], tpar);
			if (active) { // transverse velocity-difference minima of cell r (hydro_system.hpp:1022-1033)
				const double *tr = trans_s;
				const double x0 = s_0[SM::XW + lane + 2], xm = s_0[SM::XW + lane + 1], xp = s_0[SM::XW + lane + 3];
				const double mx = dmin(xp - x0, x0 - xm); // along x
				const double t0 = s_0[((DIR == 1) ? 3 : 2) * SM::XW + lane + 2];
				const double mt = dmin(tr[32 + lane] - t0, t0 - tr[lane]); // along the other transverse axis
				if (DIR == 1) { // V = z, W = x
					mV = mt;
					mW = mx;
				} else { // V = x, W = y
					mV = mx;
					mW = mt;
				}
			}
			__syncwarp(); // row r-1 and the transverse rows have been consumed by every lane: refill them for step r+1
			if (elect_one()) {
				if (r + 3 <= s1 + 2)
					issue_prim((k4 + 1) & 3);
				if (r + 1 <= s1)
					issue_trans();
			}
			++row_q;
			++row_t;
			if (r >= s0) {
				double G[NV + 1];
				if (active) {
					const double du = vN0 - vNp;
					double dw = dmin(mVp, mV);
					dw = dmin(dmin(mWp, mW), dw);
					double F[NV], vf;
					hllc_face<ARITH, DIR, NS, NMS, REINT>(c, apL, am, du, dw, F, vf);
#pragma unroll
					for (int n = 0; n < NV; ++n)
						G[n] = F[n];
					G[NV] = vf;
					if (KEEPF) { // the stage's own flux of face r, for the flux registers
#pragma unroll
						for (int n = 0; n < NV; ++n)
							fk.p[o_f + n * fk.ns] = F[n];
					}
				}
				if (have_aux) {
					mbar_wait(&bars[5], aux_phase);
					aux_phase ^= 1u;
				}
				if (active) {
					if (STAGE == 1) {
						if (DUAL && !R0) { // flux_rk2 = 0 + 0.5 F (QuokkaSimulation.hpp:1106-1107)
#pragma unroll
							for (int n = 0; n <= NV; ++n)
								h.p[o_h + n * h.ns] = 0.0 + 0.5 * G[n];
						}
					} else if (!R0) {
#pragma unroll
						for (int n = 0; n <= NV; ++n)
							G[n] = aux_s[SM::AUX_HF + n * 32 + lane] + 0.5 * G[n];
					}
					if (r > s0) { // cell r-1: both faces known
						const int64_t orc = o_r - srN;
						double rr[NV];
#pragma unroll
						for (int n = 0; n < NV; ++n)
							rr[n] = aux_s[SM::AUX_RHS + n * 32 + lane] + c.inv_dx[DIR] * (Gp[n] - G[n]);
						const double dv = div_dx<ARITH>(c, DIR, G[NV] - Gp[NV]);
						const double divv = aux_s[SM::AUX_RHS + NV * 32 + lane] + dv;
						if (!LAST) {
#pragma unroll
							for (int n = 0; n < NV; ++n)
								rh.p[orc + n * rh.ns] = rr[n];
							rh.p[orc + NV * rh.ns] = divv;
						} else {
							double U0[NV], Un[NV];
#pragma unroll
							for (int n = 0; n < NV; ++n)
								U0[n] = aux_s[SM::AUX_U0 + n * 32 + lane];
							int bad, nf;
							cell_epilogue<ARITH, NS, NMS, (STAGE == 2 && R0)>(c, U0, rr, divv, Un, bad, nf, aux_s + SM::AUX_HF + lane);
							if (STAGE == 1 && DUAL && R0) { // keep R(U0) of cell r-1 for stage 2
								const int64_t ohc = o_h - shN;
#pragma unroll
								for (int n = 0; n < NV; ++n)
									h.p[ohc + n * h.ns] = rr[n];
							}
							bad_cnt += bad;
							nf_cnt += nf;
							const int64_t ooc = o_o - soN;
#pragma unroll
							for (int n = 0; n < NV; ++n)
								uo.p[ooc + n * uo.ns] = Un[n];
						}
					}
#pragma unroll
					for (int n = 0; n <= NV; ++n)
						Gp[n] = G[n];
				}
			}
			if (active) {
#pragma unroll
				for (int n = 0; n < NV; ++n)
					apL[n] = ap[n];
				mVp = mV;
				mWp = mW;
				vNp = vN0;
			}
			o_h += shN;
			o_r += srN;
			o_o += soN;
			if (KEEPF)
				o_f += sfN;
			// rotate the ring
			k4 = (k4 + 1) & 3;
			if (k4 == 0)
				par4 ^= 1u;
			tpar ^= 1u;
		}
	}
	if (LAST) {
		const int any = __syncthreads_or(bad_cnt | nf_cnt);
		if (any) {
			if (bad_cnt)
				atomicAdd(counters, (unsigned long long)bad_cnt);
			if (nf_cnt)
				atomicAdd(counters + 1, (unsigned long long)nf_cnt);
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// x sweep with TMA-staged rows: one warp owns a 30-cell tile in x (lane l <-> cell x0-1+l, parabola for every lane,
// flux for lanes 1..31, update for lanes 1..30; neighbour states and fluxes by warp shuffle, as in k_sweep_x) and
// walks XROWS consecutive rows; the rows of row m+1 are bulk-copied into the other stage while row m is computed.
// ---------------------------------------------------------------------------------------------------------------
constexpr int XROWS = 8;
template <int NV, int STAGE> struct XSmem {
	static constexpr int PW = 38;				     // cells x0-4 .. x0+33 of every component (+ chi)
	static constexpr int PRIM = (((NV + 1) * PW + 15) / 16) * 16; // one [38 x (NV+1)] tile, padded to a 128-byte multiple
	static constexpr int TW = 34;				     // cells x0-2 .. x0+31 of the transverse velocity rows
	static constexpr int T3 = ((3 * TW + 15) / 16) * 16;	     // rows -1, 0, +1 of one component (one [34 x 3] tile)
	static constexpr int TRANS = 2 * T3;			     // vy at j-1 .. j+1, vz at k-1 .. k+1
	static constexpr int AUX = (STAGE == 2) ? (NV + 1) * 32 : 0;  // 0.5 F(U0) at faces x0 .. x0+31 (stage 2 only; the relaxed mode never runs STAGE 2 here)
	static constexpr int STAGE_DOUBLES = PRIM + TRANS + AUX;
	static constexpr int WARP_BYTES = 2 * STAGE_DOUBLES * 8 + 128; // two stages + two mbarriers
	static constexpr int BLOCK_BYTES = 4 * WARP_BYTES;
};

template <int ARITH, int NS, int NMS, bool REINT, int STAGE, bool DUAL, int ORDER = 3, bool KEEPF = false>
__global__ void __launch_bounds__(128) k_sweep_xt(FastConst c, const SweepBox *__restrict__ boxes, const __grid_constant__ XMaps tmaps)
{
	constexpr int NV = 6 + NS;
	using SM = XSmem<NV, STAGE>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
	double *const st0 = reinterpret_cast<double *>(smem_raw + (size_t)warp * SM::WARP_BYTES);
	uint64_t *const bars = reinterpret_cast<uint64_t *>(st0 + 2 * SM::STAGE_DOUBLES);
	const SweepBox &B = boxes[blockIdx.z];
	const int ny = B.hi[1] - B.lo[1] + 1, nz = B.hi[2] - B.lo[2] + 1;
	const int nrows = ny * nz;
	const int row0 = (blockIdx.y * 4 + warp) * XROWS;
	const int x0 = B.lo[0] + blockIdx.x * 30;
	if (row0 >= nrows || x0 > B.hi[0])
		return; // whole warp
	const int rows = min(XROWS, nrows - row0);
	const int i = x0 - 1 + lane;
	const A4 &q = B.prim;
	const A4 &h = B.hF[0];
	const A4 &r = B.rhs;
	const TmapBytes *const M = tmaps.m[blockIdx.z];
	if (lane == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		mbar_init_fence();
	}
	__syncwarp();
	// (j, k) of the row being staged / computed, advanced incrementally (one integer division per warp instead of two per row)
	int jn = B.lo[1] + row0 % ny, kn = B.lo[2] + row0 / ny; // next row to stage
	int j = jn, k = kn;					  // row being computed
	const int qx = x0 - 4 - q.bx, hx = x0 - h.bx;
	auto issue = [&](int m) { // stage row row0+m = (jn, kn): 3 tile copies (4 in stage 2)
		double *dst = st0 + (m & 1) * SM::STAGE_DOUBLES;
		uint64_t *bar = &bars[m & 1];
		mbar_arrive_expect_tx(bar, (unsigned)((NV + 1) * SM::PW + 6 * SM::TW + ((STAGE == 2) ? (NV + 1) * 32 : 0)) * 8u);
		const int jy = jn - q.by, kz = kn - q.bz;
		tma_tile_g2s(dst, &M[XM_PRIM], qx, jy, kz, 0, bar);
		tma_tile_g2s(dst + SM::PRIM, &M[XM_Y3], qx + 2, jy - 1, kz, 2, bar);	     // vy at j-1, j, j+1
		tma_tile_g2s(dst + SM::PRIM + SM::T3, &M[XM_Z3], qx + 2, jy, kz - 1, 3, bar); // vz at k-1, k, k+1
		if (STAGE == 2)
			tma_tile_g2s(dst + SM::PRIM + SM::TRANS, &M[XM_HF], hx, jn - h.by, kn - h.bz, 0, bar);
	};
	if (elect_one())
		issue(0);
	if (++jn > B.hi[1]) {
		jn = B.lo[1];
		++kn;
	}
	const bool face_ok = (lane >= 1) && (i >= B.lo[0]) && (i <= B.hi[0] + 1);
	const bool upd = (lane >= 1) && (lane <= 30) && (i <= B.hi[0]);
	for (int m = 0; m < rows; ++m) {
		__syncwarp();
		if (m + 1 < rows && elect_one())
			issue(m + 1);
		if (++jn > B.hi[1]) {
			jn = B.lo[1];
			++kn;
		}
		const double *sp = st0 + (m & 1) * SM::STAGE_DOUBLES;
		mbar_wait(&bars[m & 1], (unsigned)(m >> 1) & 1u);
		// PPM + flattening of the own cell (x0-1+lane sits at index lane+3 of a prim row)
		const double chi = sp[NV * SM::PW + lane + 3], omchi = 1. - chi;
		double am[NV], ap[NV], q0v1 = 0;
#pragma unroll
		for (int n = 0; n < NV; ++n) {
			const double *p = sp + n * SM::PW + lane + 3;
			const double qm2 = p[-2], qm1 = p[-1], q0 = p[0], qp1 = p[1], qp2 = p[2];
			if (n == 1)
				q0v1 = q0;
			if (ORDER == 3)
				if (ARITH == 1)
					r_ppm_flat(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), chi, omchi, am[n], ap[n]);
				else
					f_ppm_flat(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), chi, omchi, am[n], ap[n]);
			else
				f_plm_flat(qm1, q0, qp1, chi, omchi, am[n], ap[n]);
		}
		// transverse minima: V = y, W = z (cell x0-1+lane sits at index lane+1 of a transverse row)
		const double *tr = sp + SM::PRIM;
		const double vy0 = sp[2 * SM::PW + lane + 3], vz0 = sp[3 * SM::PW + lane + 3];
		const double mV = dmin(tr[2 * SM::TW + lane + 1] - vy0, vy0 - tr[lane + 1]);				  // rows j+1, j-1
		const double mW = dmin(tr[SM::T3 + 2 * SM::TW + lane + 1] - vz0, vz0 - tr[SM::T3 + lane + 1]); // rows k+1, k-1
		double Ls[NV];
#pragma unroll
		for (int n = 0; n < NV; ++n)
			Ls[n] = shfl_up1(ap[n]);
		const double mVl = shfl_up1(mV), mWl = shfl_up1(mW);
		const double du = q0v1 - shfl_up1(q0v1);
		double dw = dmin(mVl, mV);
		dw = dmin(dmin(mWl, mW), dw);
		double G[NV + 1];
		if (face_ok) {
			double F[NV], vf;
			hllc_face<ARITH, 0, NS, NMS, REINT>(c, Ls, am, du, dw, F, vf);
			if (KEEPF && (lane <= 30 || i == B.hi[0] + 1)) { // the stage's own flux of face i, for the flux registers
				const A4 &fk = B.fo[0];
				const int64_t of = fk.off(i, j, k);
#pragma unroll
				for (int n = 0; n < NV; ++n)
					fk.p[of + n * fk.ns] = F[n];
			}
			if (STAGE == 1) {
#pragma unroll
				for (int n = 0; n < NV; ++n)
					G[n] = F[n];
				G[NV] = vf;
				if (DUAL && (lane <= 30 || i == B.hi[0] + 1)) { // flux_rk2 = 0 + 0.5 F (QuokkaSimulation.hpp:1106-1107)
					const int64_t oh = h.off(i, j, k);
#pragma unroll
					for (int n = 0; n <= NV; ++n)
						h.p[oh + n * h.ns] = 0.0 + 0.5 * G[n];
				}
			} else {
				const double *ax = sp + SM::PRIM + SM::TRANS + lane - 1; // face i = x0-1+lane sits at index lane-1
#pragma unroll
				for (int n = 0; n < NV; ++n)
					G[n] = ax[n * 32] + 0.5 * F[n];
				G[NV] = ax[NV * 32] + 0.5 * vf;
			}
		} else {
#pragma unroll
			for (int n = 0; n <= NV; ++n)
				G[n] = 0.0;
		}
		const int64_t orr = upd ? r.off(i, j, k) : 0;
#pragma unroll
		for (int n = 0; n < NV; ++n) {
			const double Gn = shfl_dn1(G[n]);
			if (upd)
				r.p[orr + n * r.ns] = c.inv_dx[0] * (G[n] - Gn);
		}
		const double Vn = shfl_dn1(G[NV]);
		if (upd) {
			const double dv = div_dx<ARITH>(c, 0, Vn - G[NV]);
			r.p[orr + NV * r.ns] = dv;
		}
		if (++j > B.hi[1]) {
			j = B.lo[1];
			++k;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// x sweep over the CONCATENATED rows of a box (round 2; replaces k_sweep_xt wherever every box is at least 30 cells wide).
// k_sweep_xt cuts every row into 30-cell tiles of its own, so a 128-cell row costs five warp passes of which the last one
// updates 8 cells (17 % of the lanes of the sweep idle; 43 % for the 32-cell boxes of an AMR level).  Here the rows of a box
// are laid end to end as SLOTS -- row r contributes the nx + 2 cells lo-1 .. hi+1 (the two end cells carry a parabola only) --
// and tile t owns slots 30 t .. 30 t + 29 of that sequence, whatever row they fall in: lane l of the warp holds slot
// 30 t - 1 + l, lanes 1 .. 30 are the owned ones (parabola, face flux against lane l-1, update against lane l+1 exactly as in
// k_sweep_xt), lanes 0 and 31 the halo.  A tile that runs over the end of a row continues with the first slots of the next
// row: its lanes then read from two staged windows, A (the row the tile starts in; double-buffered across tiles) and B (the
// next row, cells lo-4 .. lo+35; one buffer, requested as soon as the previous tile has read its own).  The first slot of a
// row is never a face and the last never an update, so no shuffle crosses a row boundary with a value that is used.
// 130 slots per 128-cell row -> 4.33 warp passes instead of 5.  Stage 2 of the exact mode reads 0.5 F(U0) with plain
// loads after the Riemann solve (a staged copy would have to live in the single B buffer until then).
// ---------------------------------------------------------------------------------------------------------------
constexpr int XC_TILES = 8; // consecutive tiles per warp
constexpr int XC_WARPS = 3; // warps per CTA: 3 windows per warp -> 36 KB per CTA at NV = 6, six CTAs (18 warps) per SM
template <int NV> struct XCSmem {
	static constexpr int PW = 40;				     // cells c0-3-sh .. c0+36-sh of every component (+ chi), c0 = the cell of lane 0, sh = 0 | 1 so that the window starts at an even x
	static constexpr int PRIM = (((NV + 1) * PW + 15) / 16) * 16; // one [40 x (NV+1)] tile, padded to a 128-byte multiple
	static constexpr int TW = 36;				     // cells c0-1-sh .. c0+34-sh of the transverse velocity rows
	static constexpr int T3 = ((3 * TW + 15) / 16) * 16;	     // rows -1, 0, +1 of one component (one [36 x 3] tile)
	static constexpr int BUF = PRIM + 2 * T3;		     // one window: primitives, vy at j-1 .. j+1, vz at k-1 .. k+1
	static constexpr int WARP_BYTES = 3 * BUF * 8 + 128;	     // windows A (two) and B + three mbarriers
	static constexpr int BLOCK_BYTES = XC_WARPS * WARP_BYTES;
};

template <int ARITH, int NS, int NMS, bool REINT, int STAGE, bool DUAL, int ORDER = 3, bool KEEPF = false>
__global__ void __launch_bounds__(32 * XC_WARPS) k_sweep_xc(FastConst c, const SweepBox *__restrict__ boxes, const __grid_constant__ XMaps tmaps)
{
	constexpr int NV = 6 + NS;
	using SM = XCSmem<NV>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31;
	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
	double *const st0 = reinterpret_cast<double *>(smem_raw + (size_t)warp * SM::WARP_BYTES);
	uint64_t *const bars = reinterpret_cast<uint64_t *>(st0 + 3 * SM::BUF); // [0], [1]: windows A; [2]: window B
	const SweepBox &B = boxes[blockIdx.z];
	const int nx = B.hi[0] - B.lo[0] + 1, ny = B.hi[1] - B.lo[1] + 1, nz = B.hi[2] - B.lo[2] + 1;
	const int nslot = nx + 2, nrows = ny * nz;
	const int ntiles = (nrows * nslot + 29) / 30; // the host checked that the slot count fits an int and nx >= 30
	const int t0 = (blockIdx.x * XC_WARPS + warp) * XC_TILES;
	if (t0 >= ntiles)
		return; // whole warp
	const int nt = min(XC_TILES, ntiles - t0);
	const A4 &q = B.prim;
	const A4 &h = B.hF[0];
	const A4 &r = B.rhs;
	const TmapBytes *const M = tmaps.m[blockIdx.z];
	if (lane == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		mbar_init(&bars[2], 1);
		mbar_init_fence();
	}
	__syncwarp();
	// one window: the 40 cells from array column qx (EVEN) of row (j, k) of every primitive + chi, and the transverse velocity rows around them
	auto issue_win = [&](double *dst, uint64_t *bar, int qx, int j, int k) {
		mbar_arrive_expect_tx(bar, (unsigned)((NV + 1) * SM::PW + 6 * SM::TW) * 8u);
		const int jy = j - q.by, kz = k - q.bz;
		tma_tile_g2s(dst, &M[XM_PRIM], qx, jy, kz, 0, bar);
		tma_tile_g2s(dst + SM::PRIM, &M[XM_Y3], qx + 2, jy - 1, kz, 2, bar);	     // vy at j-1, j, j+1
		tma_tile_g2s(dst + SM::PRIM + SM::T3, &M[XM_Z3], qx + 2, jy, kz - 1, 3, bar); // vz at k-1, k, k+1
	};
	// the tile being computed (warp-uniform): slot of lane 0 within its row (pos0, -1 for the first tile of a box), that row (rowA -> jA, kA),
	// how many lanes stay in it (nA) and whether the remaining lanes continue in the next row (two -> jB, kB)
	int pos0, rowA, jA, kA;
	{
		const int g0 = 30 * t0 - 1;
		rowA = (g0 < 0) ? 0 : (int)((unsigned)g0 / (unsigned)nslot);
		pos0 = g0 - rowA * nslot;
		const int kq = (int)((unsigned)rowA / (unsigned)ny);
		jA = B.lo[1] + (rowA - kq * ny);
		kA = B.lo[2] + kq;
	}
	auto next_row = [&](int j, int k, int &jn, int &kn) {
		jn = j + 1;
		kn = k;
		if (jn > B.hi[1]) {
			jn = B.lo[1];
			++kn;
		}
	};
	int nA = min(32, nslot - pos0);
	bool two = (nA < 32) && (rowA + 1 < nrows);
	int jB, kB;
	next_row(jA, kA, jB, kB);
	double *const winB = st0 + 2 * SM::BUF;
	// array column of cell lo - 4 (the first cell window B needs; 0 for the level's own scratch, whose ghost width is 4): even, or the host
	// would not have chosen this kernel.  Window A wants to start at column qx0 + pos0: it starts one column earlier when that is odd.
	const int qx0 = B.lo[0] - 4 - q.bx;
	if (elect_one()) {
		issue_win(st0, &bars[0], qx0 + (pos0 & ~1), jA, kA);
		if (two)
			issue_win(winB, &bars[2], qx0, jB, kB);
	}
	unsigned bpar = 0;
	for (int m = 0; m < nt; ++m) {
		__syncwarp();
		// the next tile: 30 slots further on, at most one row further (nslot >= 32)
		const bool more = (m + 1 < nt);
		int pos0n = pos0 + 30, rowAn = rowA, jAn = jA, kAn = kA;
		if (pos0n >= nslot) {
			pos0n -= nslot;
			++rowAn;
			jAn = jB;
			kAn = kB;
		}
		const int nAn = min(32, nslot - pos0n);
		const bool twon = (nAn < 32) && (rowAn + 1 < nrows);
		int jBn, kBn;
		next_row(jAn, kAn, jBn, kBn);
		if (more && elect_one())
			issue_win(st0 + ((m + 1) & 1) * SM::BUF, &bars[(m + 1) & 1], qx0 + (pos0n & ~1), jAn, kAn);
		// this lane's slot
		const bool inB = (lane >= nA) && two;
		const int lq = inB ? lane - nA : lane + (pos0 & 1); // position within its window: the own cell sits at index lq + 3 of a prim row
		const int pos = inB ? lane - nA : pos0 + lane;	    // slot within its row: cell lo - 1 + pos
		const bool valid = (lane < nA) || inB;	  // lanes past the last slot of the box compute on whatever window A holds and store nothing
		const int i = B.lo[0] - 1 + pos;
		const int j = inB ? jB : jA, k = inB ? kB : kA;
		const double *sp = inB ? winB : st0 + (m & 1) * SM::BUF;
		mbar_wait(&bars[m & 1], (unsigned)(m >> 1) & 1u);
		if (two) {
			mbar_wait(&bars[2], bpar);
			bpar ^= 1u;
		}
		// PPM + flattening of the own cell (it sits at index lq + 3 of a prim row)
		const double chi = sp[NV * SM::PW + lq + 3], omchi = 1. - chi;
		double am[NV], ap[NV], q0v1 = 0;
#pragma unroll
		for (int n = 0; n < NV; ++n) {
			const double *p = sp + n * SM::PW + lq + 3;
			const double qm2 = p[-2], qm1 = p[-1], q0 = p[0], qp1 = p[1], qp2 = p[2];
			if (n == 1)
				q0v1 = q0;
			if (ORDER == 3)
				if (ARITH == 1)
					r_ppm_flat(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), chi, omchi, am[n], ap[n]);
				else
					f_ppm_flat(qm1, q0, qp1, ppm_iface(qm2, qm1, q0, qp1), ppm_iface(qm1, q0, qp1, qp2), chi, omchi, am[n], ap[n]);
			else
				f_plm_flat(qm1, q0, qp1, chi, omchi, am[n], ap[n]);
		}
		// transverse minima: V = y, W = z (the own cell sits at index lq + 1 of a transverse row)
		const double *tr = sp + SM::PRIM;
		const double vy0 = sp[2 * SM::PW + lq + 3], vz0 = sp[3 * SM::PW + lq + 3];
		const double mV = dmin(tr[2 * SM::TW + lq + 1] - vy0, vy0 - tr[lq + 1]);			  // rows j+1, j-1
		const double mW = dmin(tr[SM::T3 + 2 * SM::TW + lq + 1] - vz0, vz0 - tr[SM::T3 + lq + 1]); // rows k+1, k-1
		__syncwarp(); // window B has been read by every lane: the next tile's may land in it
		if (more && twon && elect_one())
			issue_win(winB, &bars[2], qx0, jBn, kBn);
		double Ls[NV];
#pragma unroll
		for (int n = 0; n < NV; ++n)
			Ls[n] = shfl_up1(ap[n]);
		const double mVl = shfl_up1(mV), mWl = shfl_up1(mW);
		const double du = q0v1 - shfl_up1(q0v1);
		double dw = dmin(mVl, mV);
		dw = dmin(dmin(mWl, mW), dw);
		// a face needs the cell on its low side in the same row (pos >= 1 puts it in lane - 1), an update the face on its high side (lane + 1)
		const bool face_ok = valid && (lane >= 1) && (pos >= 1);
		const bool upd = valid && (lane >= 1) && (lane <= 30) && (pos >= 1) && (pos <= nx);
		const bool own_face = face_ok && (lane <= 30); // lane 31's face is lane 1's of the next tile
		double G[NV + 1];
		if (face_ok) {
			double F[NV], vf;
			hllc_face<ARITH, 0, NS, NMS, REINT>(c, Ls, am, du, dw, F, vf);
			if (KEEPF && own_face) { // the stage's own flux of face i, for the flux registers
				const A4 &fk = B.fo[0];
				const int64_t of = fk.off(i, j, k);
#pragma unroll
				for (int n = 0; n < NV; ++n)
					fk.p[of + n * fk.ns] = F[n];
			}
			if (STAGE == 1) {
#pragma unroll
				for (int n = 0; n < NV; ++n)
					G[n] = F[n];
				G[NV] = vf;
				if (DUAL && own_face) { // flux_rk2 = 0 + 0.5 F (QuokkaSimulation.hpp:1106-1107)
					const int64_t oh = h.off(i, j, k);
#pragma unroll
					for (int n = 0; n <= NV; ++n)
						h.p[oh + n * h.ns] = 0.0 + 0.5 * G[n];
				}
			} else {
				const int64_t oh = h.off(i, j, k); // 0.5 F(U0) of face i
#pragma unroll
				for (int n = 0; n < NV; ++n)
					G[n] = h.p[oh + n * h.ns] + 0.5 * F[n];
				G[NV] = h.p[oh + NV * h.ns] + 0.5 * vf;
			}
		} else {
#pragma unroll
			for (int n = 0; n <= NV; ++n)
				G[n] = 0.0;
		}
		const int64_t orr = upd ? r.off(i, j, k) : 0;
#pragma unroll
		for (int n = 0; n < NV; ++n) {
			const double Gn = shfl_dn1(G[n]);
			if (upd)
				r.p[orr + n * r.ns] = c.inv_dx[0] * (G[n] - Gn);
		}
		const double Vn = shfl_dn1(G[NV]);
		if (upd) {
			const double dv = div_dx<ARITH>(c, 0, Vn - G[NV]);
			r.p[orr + NV * r.ns] = dv;
		}
		pos0 = pos0n;
		rowA = rowAn;
		jA = jAn;
		kA = kAn;
		nA = nAn;
		two = twon;
		jB = jBn;
		kB = kBn;
	}
}
