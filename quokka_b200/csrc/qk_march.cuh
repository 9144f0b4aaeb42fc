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

template <int NV> struct MarchSmem {
	static constexpr int NR = 5;
	static constexpr int PR = (NV + 1) * 32 + 36; // rows n*32 for n = 0..NV (n = NV: chi; n = 1 unused) + wide vx row
	static constexpr int WIDE = (NV + 1) * 32;
	static constexpr int TR = 64;
	static constexpr int AUX_HF = 0, AUX_RHS = (NV + 1) * 32, AUX_U0 = 2 * (NV + 1) * 32;
	static constexpr int AUX = 2 * (NV + 1) * 32 + NV * 32;
	static constexpr int WARP_DOUBLES = NR * PR + 2 * TR + AUX;
	static constexpr int WARP_BYTES = WARP_DOUBLES * 8 + 64; // + 8 mbarriers
	static constexpr int BLOCK_BYTES = 4 * WARP_BYTES;
};

template <int DIR, int NS, int NMS, bool REINT, int STAGE, bool DUAL, bool LAST>
__global__ void __launch_bounds__(128) k_march_t(FastConst c, const SweepBox *__restrict__ boxes, int nseg, unsigned long long *__restrict__ counters)
{
	constexpr int NV = 6 + NS;
	using SM = MarchSmem<NV>;
	constexpr int TD = (DIR == 1) ? 2 : 1;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double *const prim_s = reinterpret_cast<double *>(smem_raw + (size_t)warp * SM::WARP_BYTES);
	double *const trans_s = prim_s + SM::NR * SM::PR;
	double *const aux_s = trans_s + 2 * SM::TR;
	uint64_t *const bars = reinterpret_cast<uint64_t *>(aux_s + SM::AUX); // [0..4] prim, [5..6] trans, [7] aux

	const int box = blockIdx.z / nseg, seg = blockIdx.z - box * nseg;
	const SweepBox &B = boxes[box];
	const int i0 = B.lo[0] + blockIdx.x * 32;
	const int t = B.lo[TD] + blockIdx.y * 4 + warp;
	const int s0 = B.lo[DIR] + seg * SEG;
	int bad_cnt = 0, nf_cnt = 0;
	if (i0 <= B.hi[0] && t <= B.hi[TD] && s0 <= B.hi[DIR]) {
		const int nact = min(32, B.hi[0] - i0 + 1);
		const bool active = lane < nact;
		const unsigned rowb = (unsigned)((nact + 1) & ~1) * 8u, wideb = rowb + 32u;
		const int s1 = min(s0 + SEG, B.hi[DIR] + 1);
		const int i = i0 + lane;
		const A4 &q = B.prim;
		const A4 &h = B.hF[DIR];
		const A4 &rh = B.rhs;
		const A4 &u0 = B.U0;
		const A4 &uo = B.Uo;
		// offset of (i0, row) in an array, the transverse index fixed at t
		auto off_row = [&](const A4 &a, int row) -> int64_t {
			return (DIR == 1) ? a.off(i0, row, t) : a.off(i0, t, row);
		};
		if (lane == 0) {
#pragma unroll
			for (int b = 0; b < 8; ++b)
				mbar_init(&bars[b], 1);
			mbar_init_fence();
		}
		__syncwarp();

		auto issue_prim = [&](int row) {
			const int slot = (row - (s0 - 3)) % SM::NR;
			double *dst = prim_s + slot * SM::PR;
			uint64_t *bar = &bars[slot];
			const double *src = q.p + off_row(q, row);
			mbar_arrive_expect_tx(bar, (unsigned)NV * rowb + wideb);
#pragma unroll
			for (int n = 0; n <= NV; ++n) {
				if (n == 1)
					continue;
				bulk_g2s(dst + n * 32, src + n * q.ns, rowb, bar);
			}
			bulk_g2s(dst + SM::WIDE, src - 2 + q.ns, wideb, bar);
		};
		auto issue_trans = [&](int row) {
			const int slot = (row - (s0 - 1)) & 1;
			double *dst = trans_s + slot * SM::TR;
			uint64_t *bar = &bars[5 + slot];
			const int64_t sT = (DIR == 1) ? q.ks : q.js;	      // y sweep: V = z; z sweep: W = y
			const double *src = q.p + off_row(q, row) + ((DIR == 1) ? 3 : 2) * q.ns; // vz | vy
			mbar_arrive_expect_tx(bar, 2u * rowb);
			bulk_g2s(dst, src - sT, rowb, bar);
			bulk_g2s(dst + 32, src + sT, rowb, bar);
		};
		// what the end of step r reads: hF of face r (stage 2), rhs (+U0) of cell r-1
		auto aux_bytes = [&](int r) -> unsigned {
			unsigned b = 0;
			if (r >= s0) {
				if (STAGE == 2)
					b += (unsigned)(NV + 1) * rowb;
				if (r > s0)
					b += (unsigned)(NV + 1) * rowb + (LAST ? (unsigned)NV * rowb : 0u);
			}
			return b;
		};
		auto issue_aux = [&](int r) {
			uint64_t *bar = &bars[7];
			mbar_arrive_expect_tx(bar, aux_bytes(r));
			if (STAGE == 2) {
				const double *src = h.p + off_row(h, r);
#pragma unroll
				for (int n = 0; n <= NV; ++n)
					bulk_g2s(aux_s + SM::AUX_HF + n * 32, src + n * h.ns, rowb, bar);
			}
			if (r > s0) {
				const double *src = rh.p + off_row(rh, r - 1);
#pragma unroll
				for (int n = 0; n <= NV; ++n)
					bulk_g2s(aux_s + SM::AUX_RHS + n * 32, src + n * rh.ns, rowb, bar);
				if (LAST) {
					const double *su = u0.p + off_row(u0, r - 1);
#pragma unroll
					for (int n = 0; n < NV; ++n)
						bulk_g2s(aux_s + SM::AUX_U0 + n * 32, su + n * u0.ns, rowb, bar);
				}
			}
		};
		// value of component n (n = NV: chi) of the cell (lane, row)
		auto P = [&](int row, int n) -> double {
			const double *sl = prim_s + ((row - (s0 - 3)) % SM::NR) * SM::PR;
			return (n == 1) ? sl[SM::WIDE + lane + 2] : sl[n * 32 + lane];
		};

		if (lane == 0) {
			for (int row = s0 - 3; row <= s0 + 1; ++row)
				issue_prim(row);
			issue_trans(s0 - 1);
		}
		// rows s0-3 .. s0 -> unlimited interface value at the low face of cell s0-1
#pragma unroll
		for (int b = 0; b < 4; ++b)
			mbar_wait(&bars[b], 0);
		double apL[NV], ifl[NV], Gp[NV + 1];
		double mVp = 0, mWp = 0, vNp = 0;
		if (active) {
#pragma unroll
			for (int n = 0; n < NV; ++n)
				ifl[n] = ppm_iface(P(s0 - 3, n), P(s0 - 2, n), P(s0 - 1, n), P(s0, n));
		}
		unsigned aux_phase = 0;
		int64_t o_h = h.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		int64_t o_r = rh.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		int64_t o_o = uo.off(i, (DIR == 1) ? s0 - 1 : t, (DIR == 1) ? t : s0 - 1);
		const int64_t shN = (DIR == 1) ? h.js : h.ks, srN = (DIR == 1) ? rh.js : rh.ks, soN = (DIR == 1) ? uo.js : uo.ks;

		for (int r = s0 - 1; r <= s1; ++r) {
			__syncwarp(); // every lane is done with the slots about to be refilled
			const bool have_aux = aux_bytes(r) != 0;
			if (lane == 0) {
				if (r + 3 <= s1 + 2)
					issue_prim(r + 3);
				if (r + 1 <= s1)
					issue_trans(r + 1);
				if (have_aux)
					issue_aux(r);
			}
			{ // row r+2 is the newest one this step reads
				const int k = r + 2 - (s0 - 3);
				mbar_wait(&bars[k % SM::NR], (unsigned)(k / SM::NR) & 1u);
			}
			double am[NV], ap[NV];
			double vN0 = 0, mV = 0, mW = 0;
			if (active) {
				const double chi = P(r, NV), omchi = 1. - chi;
#pragma unroll
				for (int n = 0; n < NV; ++n) {
					const double qm1 = P(r - 1, n), q0 = P(r, n), qp1 = P(r + 1, n), qp2 = P(r + 2, n);
					if (n == 1 + DIR)
						vN0 = q0;
					const double ifh = ppm_iface(qm1, q0, qp1, qp2);
					f_ppm_flat(qm1, q0, qp1, ifl[n], ifh, chi, omchi, am[n], ap[n]);
					ifl[n] = ifh;
				}
			}
			{
				const int k = r - (s0 - 1);
				mbar_wait(&bars[5 + (k & 1)], (unsigned)(k >> 1) & 1u);
			}
			if (active) { // transverse velocity-difference minima of cell r (hydro_system.hpp:1022-1033)
				const double *sl = prim_s + ((r - (s0 - 3)) % SM::NR) * SM::PR;
				const double *tr = trans_s + ((r - (s0 - 1)) & 1) * SM::TR;
				const double x0 = sl[SM::WIDE + lane + 2], xm = sl[SM::WIDE + lane + 1], xp = sl[SM::WIDE + lane + 3];
				const double mx = dmin(xp - x0, x0 - xm); // along x
				const double t0 = sl[((DIR == 1) ? 3 : 2) * 32 + lane];
				const double mt = dmin(tr[32 + lane] - t0, t0 - tr[lane]); // along the other transverse axis
				if (DIR == 1) { // V = z, W = x
					mV = mt;
					mW = mx;
				} else { // V = x, W = y
					mV = mx;
					mW = mt;
				}
			}
			if (r >= s0) {
				double G[NV + 1];
				if (active) {
					const double du = vN0 - vNp;
					double dw = dmin(mVp, mV);
					dw = dmin(dmin(mWp, mW), dw);
					double F[NV], vf;
					unsigned slow = 0;
					f_hllc<DIR, NS, NMS, REINT, true>(c, apL, am, du, dw, F, vf, slow);
					if (slow)
						f_hllc<DIR, NS, NMS, REINT, false>(c, apL, am, du, dw, F, vf, slow);
#pragma unroll
					for (int n = 0; n < NV; ++n)
						G[n] = F[n];
					G[NV] = vf;
				}
				if (have_aux) {
					mbar_wait(&bars[7], aux_phase);
					aux_phase ^= 1u;
				}
				if (active) {
					if (STAGE == 1) {
						if (DUAL) { // flux_rk2 = 0 + 0.5 F (QuokkaSimulation.hpp:1106-1107)
#pragma unroll
							for (int n = 0; n <= NV; ++n)
								h.p[o_h + n * h.ns] = 0.0 + 0.5 * G[n];
						}
					} else {
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
						unsigned s3 = 0;
						double dv = div_c<true>(G[NV] - Gp[NV], c.dx[DIR], c.y_dx[DIR], s3);
						if (s3)
							dv = slow_div(G[NV] - Gp[NV], c.dx[DIR]);
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
							cell_epilogue<NS, NMS>(c, U0, rr, divv, Un, bad, nf);
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
