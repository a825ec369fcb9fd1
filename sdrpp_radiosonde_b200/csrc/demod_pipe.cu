/*
 * demod_pipe.cu — K1 (production): IQ / FM samples -> hard bits for the GFSK sondes, as a warp-specialised software
 * pipeline inside each CTA.  Reference chain: SD/demod/gfsk.c:56-131 = agc_apply (dsp/agc.c:19-34) ->
 * filter_fwd_sample / filter_get (dsp/filter.c:41-64) -> advance_timeslot / retime (dsp/timing.c:28-76) -> slicer,
 * behind the FM discriminator of the plugin (dsp::demod::FM, src/main.cpp:57).
 *
 * One CTA owns up to PIPE_G = 8 channels of one sonde type and streams the buffer through tiles of T = 256 samples.
 * What is serial in time per channel (the two AGC recurrences and the timing loop) runs on two warps whose lanes are
 * channels; everything that is a pure function of position runs on a pool of parallel-work warps:
 *
 *   LD  1 warp    TMA producer: one bulk copy per channel row segment and tile into a 3-deep ring (cp.async.bulk)
 *   PW  n warps   pull (tile, channel) work items from a shared-memory queue, in dependency order:
 *                   item A  S1  FM discriminator of one channel's tile                              -> x
 *                   item B  S3  gain apply a = s * (5 / avg)      (agc.c:27,31)                     -> q
 *                           S4  49-tap FIR at every sample / polyphase branch (filter.c:49-64)      -> y (slot ring)
 *   AG  1 warp    both AGC recurrences fused (agc.c:24-28), lane = channel                          x -> s, v
 *   TM  1 warp    Gardner NCO + retime + slicer, lane = channel (timing_exact.cuh)                  y -> bit ring (HBM)
 *
 *        HBM --LD--> raw[3] --A--> x[3] --AG--> s,v[2] --B--> q[2] --B--> y[3 tiles] --TM--> bits
 *
 * Scheduling facts this layout is built on (tools/ubench, measured on the B200): the SMSP arbiter is round-robin
 * (no warp-id priority), every fp32 instruction occupies the FMA pipe of its SMSP for one cycle (packed f32x2: two),
 * and a dependent chain is slowed by ~1.5 cycles per operation for every FMA-bound warp sharing its SMSP.  So the two
 * serial warps sit on SMSPs of their own (warp id % 4 = 0: TM + the mostly sleeping LD, 1: AG), the PW warps on the
 * other two (plus optionally a few beside TM), and because that makes the SMSPs unequal the PW work is distributed
 * dynamically.  Items only ever wait for earlier items or for the other roles, so the queue cannot deadlock.
 *
 * Everything is the reference's fp32 operation order (strict_math.cuh), so soft symbols stay bit-identical.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cuda.h>

#include "pipe_common.cuh"
#include "timing_exact.cuh"

static __constant__ sonde_modem c_modem[SONDE_NTYPES_];
/* every FIR tap twice, as the (c, c) operand of the packed fp32x2 multiply */
static __constant__ float2 c_taps2[SONDE_NTYPES_][SONDE_MAX_PHASES * SONDE_FIR_TAPS];

extern "C" cudaError_t sonde_upload_modems_pipe(const sonde_modem *m)
{
	static float2 t2[SONDE_NTYPES_][SONDE_MAX_PHASES * SONDE_FIR_TAPS];
	for (int t = 0; t < SONDE_NTYPES_; t++)
		for (int i = 0; i < SONDE_MAX_PHASES * SONDE_FIR_TAPS; i++) t2[t][i] = make_float2(m[t].taps[i], m[t].taps[i]);
	const cudaError_t e = cudaMemcpyToSymbol(c_taps2, t2, sizeof(t2));
	if (e != cudaSuccess) return e;
	return cudaMemcpyToSymbol(c_modem, m, sizeof(sonde_modem) * SONDE_NTYPES_);
}

namespace {

using namespace pipe;

/* tensor maps of the input [rows][len] for the tensor staging of a tile (encoded per call on the host, encode_k1_maps): the
 * rows are split as row = quotient * step + residue, so that channels `step` rows apart are consecutive in the third
 * dimension; `tile` has a box of T columns x 1 x tma_box_rows, `lookback` one of 2 columns x 1 x the same rows */
struct k1_maps {
	CUtensorMap tile, lookback;
};

/* AGC_SPLIT 1 puts the bias and the level recurrence on two warps (A1, A2).  Measured and NOT adopted (exp47, tools/ubench/
 * ubench4/5): alone on an SM the bias loop runs at 13.5 cycles/sample (its dependent chain is 3 x 4.04 = 12.1), inside the
 * kernel at 19 whether fused with the level recurrence or not, because any second warp on its SMSP — the level warp as
 * much as a parallel-work warp — costs every dependent operation about 1.5 cycles; giving A1 an SMSP of its own would
 * take issue capacity from the parallel warps, which are the bound.  RS41 0.644 ms split vs 0.631 fused, M10 0.98 vs 0.85. */
#ifndef AGC_SPLIT
#define AGC_SPLIT 0
#endif
constexpr int W_TM = 0, W_AG = 1, W_LD = 4, W_A2 = 5; /* warp ids of the single-warp roles (SMSP = id % 4)       */
constexpr uint32_t ROLE_MASK = (1u << W_TM) | (1u << W_AG) | (1u << W_LD) | (AGC_SPLIT ? (1u << W_A2) : 0u);
constexpr int NRAW = 3, NSV = 2, NY = 3;             /* ring depths in tiles (NX = 3 from pipe_common)            */
constexpr int XS = T + 12;                           /* x row: 8 floats of slack for the AGC prefetch; 268 % 32 = 12 keeps the
                                                        8 channel lanes' LDS.128 on distinct banks                 */
constexpr int YM = 32;                               /* slots mirrored after the end of the y ring                */
constexpr int QN = 244;                              /* paired FIR input, see q_phys(): 242 pairs, rows 16 B aligned */
#ifndef DISC_ROLLED
#define DISC_ROLLED 1
#endif
constexpr uint32_t ZSENT = 0x7fc5a5a5u;              /* s value of a sample that bypassed the AGC (agc.c:23)      */

template <int P>
struct smem_t {
	float2 rawt[NRAW][G][T];             /* TMA landing zone, one tile: IQ rows of T float2 (FM: dense rows of T floats
	                                        from the start of the slot) — the shape a 2-D box of G rows lands in            */
	float2 rawlb[NRAW][G][2];            /* IQ: the two samples before the tile (the discriminator's look-back)         */
	float x[NX][G][XS];                  /* discriminator output / FM input                                          */
	float s[NSV][G][RS];                 /* bias-removed samples                                                     */
	float v[NSV][G][RS];                 /* moving_avg before each sample's update                                   */
	float2 q[2][G][QN];                  /* AGC output, 48 samples of history first, stored as pairs (j, j + 64)     */
	float y[G][NY * T * P + YM + 4];     /* FIR output per NCO slot, ring of NY tiles + mirror of the first slots    */
	int zflag[NX];                       /* tile contains exact-zero samples                                         */
	int svz[NSV];                        /* the same flag travelling with s, v                                       */
	int qnext;                           /* work queue head                                                          */
	int chan[G], row[G];
	unsigned long long negzero2;         /* (-0.0f, -0.0f), read at run time so that the packed product stays an FFMA2 */
	/* mbarriers: only where the waiter is one sequential warp (LD: rawfree, AG: xfull / svfree, TM: yfull) or the
	 * wait is guarded (rawfull).  A parity wait passes falsely when the barrier is a whole phase behind, and the PW
	 * items are not sequential waiters, so what they wait for is published as monotonic counters instead. */
	unsigned long long rawfull[NRAW], rawfree[NRAW], xfull[NX], svfree[NSV], vfree[NSV], yfull[NY];
	int ag_done;                         /* tiles whose s AND v are written (the fused warp, or A2 when the AGC is split) */
	int a1_done;                         /* tiles the bias warp has finished (x slot consumed, s written)              */
	int tm_rel;                          /* y tiles the timing warp has released                                     */
	int a_done[G];                       /* per channel: tiles whose gain stage + head copy are done                  */
	int s4_done[G][2];                   /* per channel and tile parity: tiles whose FIR is done                     */
};

/* counter hand-off: the writer's lanes have synchronised (__syncwarp) before one lane publishes */
__device__ __forceinline__ void flag_publish(int *f, int v)
{
	__threadfence_block();
	*reinterpret_cast<volatile int *>(f) = v;
}
__device__ __forceinline__ void flag_wait(const int *f, int need, long long &acc, bool on)
{
	if (*reinterpret_cast<const volatile int *>(f) < need) {
		const long long t0 = on ? clock64() : 0;
		while (*reinterpret_cast<const volatile int *>(f) < need) __nanosleep(40);
		if (on) acc += clock64() - t0;
	}
	__threadfence_block();
}

/* Warp barrier for shared-memory hand-offs between lanes.  An explicit bar.warp.sync: after a loop whose trip count
 * differs between lanes ptxas turned __syncwarp() into a NOP (it took the warp for converged), and lanes that had left
 * the loop early read entries their neighbours had not written yet. */
__device__ __forceinline__ void warp_sync_hard()
{
	asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
}

/* ---- work queue: blocks of `gact` items: A(0), A(1), then for k = 0 .. ntiles-1: A(k+2), B(k) ------------------------ */
__device__ __forceinline__ int queue_pull(int *qnext, int lane)
{
	int v = 0;
	if (lane == 0) v = atomicAdd(qnext, 1);
	return __shfl_sync(FULL, v, 0);
}

/* ---- item A: FM discriminator of tile k, channel g ---------------------------------------------------------------------
 * Lane l owns the four sample pairs (2l + 64i, 2l + 64i + 1), i = 0..3: consecutive lanes read consecutive 16-byte
 * pieces of the row (conflict-free LDS.128 / fully coalesced LDG.128) and write consecutive 8-byte pieces of x. */
template <int P, bool IQ, bool TMA>
__device__ __forceinline__ void item_disc(smem_t<P> &sm, const demod_params &p, const int k, const int g, const int lane,
                                          const int ntiles, long long &wacc, const bool prof_on)
{
	const int L = p.len;
	const int n = min(T, L - k * T);
	const int xs = k % NX, rs = k % NRAW;
	const int ch = sm.chan[g];
	const size_t rowoff = (size_t)sm.row[g] * p.row_stride + (size_t)k * T;
	bool zero = false;
	/* x slot free = the AGC warp has consumed tile k-3.  This also guards the parity wait below: tile k-3 then has
	 * landed, so rawfull[rs] is in tile k's phase or past it. */
	if (k >= NX) flag_wait(AGC_SPLIT ? &sm.a1_done : &sm.ag_done, k - NX + 1, wacc, prof_on);
	if (TMA) mbar_wait_t(&sm.rawfull[rs], (k / NRAW) & 1, wacc, prof_on);
	if (IQ) {
		/* DISC_ROLLED: the four 64-sample pieces one after the other in a rolled loop (a quarter of the code: the parallel
		 * warps stream through their straight-line bodies once per item and the two SMs of a TPC share instruction fetch)
		 * instead of eight samples as one straight-line block */
		float carry = 0.0f;
		if (lane == 0) {
			/* phase of the sample before the tile: the previous call's last sample, or the look-back sample */
			if (k == 0) {
				carry = p.st[ch].disc_prev;
			} else {
				const float2 pr = TMA ? sm.rawlb[rs][g][1] : __ldg(static_cast<const float2 *>(p.in) + rowoff - 1);
				carry = det_phase(pr.x, pr.y);
			}
		}
		float last = 0.0f;
#if DISC_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
		for (int i = 0; i < 4; i++) {
			const int t = 2 * lane + 64 * i;
			float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			if (TMA) {
				f = *reinterpret_cast<const float4 *>(&sm.rawt[rs][g][t]);
			} else {
				const float2 *src = static_cast<const float2 *>(p.in) + rowoff + t;
				/* rows need not be 16-byte aligned on this path */
				if (t < n) { const float2 a = __ldg(src); f.x = a.x; f.y = a.y; }
				if (t + 1 < n) { const float2 a = __ldg(src + 1); f.z = a.x; f.w = a.y; }
			}
			/* samples past the end of the buffer: (0, 0), phase 0 */
			const float re[2] = {(t < n) ? f.x : 0.0f, (t + 1 < n) ? f.z : 0.0f};
			const float im[2] = {(t < n) ? f.y : 0.0f, (t + 1 < n) ? f.w : 0.0f};
			float ph[2];
			det_phase_n<2>(re, im, ph);
			/* previous sample's phase: the neighbouring lane's second sample; lane 0 takes lane 31's of the piece before */
			float up = __shfl_up_sync(FULL, ph[1], 1);
			if (lane == 0) up = carry;
			carry = __shfl_sync(FULL, ph[1], 31);
			const float x0 = disc_step(ph[0], up, p.fm_gain), x1 = disc_step(ph[1], ph[0], p.fm_gain);
			zero |= (t < n && x0 == 0.0f) || (t + 1 < n && x1 == 0.0f);
			*reinterpret_cast<float2 *>(&sm.x[xs][g][t]) = make_float2(x0, x1);
			/* the phase of the buffer's last sample is the next call's `previous phase` */
			if (((n - 1) >> 6) == i) last = ((n - 1) & 1) ? ph[1] : ph[0];
		}
		if (k == ntiles - 1 && (((n - 1) & 63) >> 1) == lane) p.st[ch].disc_prev = last;
	} else {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int t = 2 * lane + 64 * i;
			float x0, x1;
			if (TMA) {
				const float2 f = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(&sm.rawt[rs][0][0]) + g * T + t);
				x0 = f.x; x1 = f.y;
			} else {
				const float *src = static_cast<const float *>(p.in) + rowoff + t;
				x0 = (t < n) ? __ldg(src) : 0.0f;
				x1 = (t + 1 < n) ? __ldg(src + 1) : 0.0f;
			}
			zero |= (t < n && x0 == 0.0f) || (t + 1 < n && x1 == 0.0f);
			*reinterpret_cast<float2 *>(&sm.x[xs][g][t]) = make_float2(x0, x1);
		}
	}
	if (__any_sync(FULL, zero) && lane == 0) atomicOr(&sm.zflag[xs], 1);
	warp_sync_hard();
	if (lane == 0) {
		mbar_arrive(&sm.xfull[xs]);
		if (TMA) mbar_arrive(&sm.rawfree[rs]);
	}
}

/* ---- the FIR input of one tile -------------------------------------------------------------------------------------------
 * a_h[m], m < 304: 48 samples of history, then the tile.  The FIR wants fp32x2 operands, so the array is stored as
 * pairs  pair[j] = (a_h[j], a_h[j + 64])  for j in [0, 112) and [128, 240): pair j yields the outputs of samples j
 * and j + 64 with one packed multiply and one packed add per tap.  A lane's window is 52 consecutive pairs read with
 * LDS.128; the second region is stored 2 pairs (one 16-byte piece) further so that the eight lanes of a quarter warp
 * (four windows of each region, starts 4 pairs apart) hit eight different 16-byte bank groups. */
__device__ __forceinline__ int q_phys(const int j) { return j + ((j >> 7) << 1); }
__device__ __forceinline__ bool q_has_x(const int m) { return m < 112 || (m >= 128 && m < 240); }
__device__ __forceinline__ bool q_has_y(const int m) { return (m >= 64 && m < 176) || m >= 192; }
__device__ __forceinline__ void q_store(float2 *q, const int m, const float val)
{
	if (q_has_x(m)) q[q_phys(m)].x = val;
	if (q_has_y(m)) q[q_phys(m - 64)].y = val;
}
__device__ __forceinline__ float q_load(const float2 *q, const int m)
{
	return q_has_x(m) ? q[q_phys(m)].x : q[q_phys(m - 64)].y;
}

/* float offsets inside a q row of the (up to) two slots of sample lane + 32 i; QN - 1 is a spare pair nobody reads */
struct q_offsets {
	int x[8], y[8];
};
__device__ __forceinline__ q_offsets q_offsets_of(const int lane)
{
	q_offsets o;
#pragma unroll
	for (int i = 0; i < 8; i++) {
		const int m = SONDE_FIR_HIST + lane + 32 * i;
		o.x[i] = q_has_x(m) ? 2 * q_phys(m) : 2 * (QN - 1);
		o.y[i] = q_has_y(m) ? 2 * q_phys(m - 64) + 1 : 2 * (QN - 1) + 1;
	}
	return o;
}

/* ---- item B: gain apply + FIR of tile k, channel g -------------------------------------------------------------------- */
template <int P>
__device__ __forceinline__ void item_fir(smem_t<P> &sm, const demod_params &p, const int k, const int g, const int lane,
                                         const int ntiles, const int type, const unsigned long long negzero2, const q_offsets &qoff,
                                         long long &wacc, const bool prof_on)
{
	const int L = p.len;
	const int n = min(T, L - k * T);
	const int ss = k % NSV, ab = k & 1;
	float2 *q = sm.q[ab][g];
	const float2 *qprev = sm.q[ab ^ 1][g];

	flag_wait(&sm.ag_done, k + 1, wacc, prof_on);
	/* q[ab] is free once the FIR of tile k-2 has read it and the item of tile k-1 has copied its head out of it (the
	 * latter also means the tail this item's head needs has been written) */
	if (k >= 2) flag_wait(&sm.s4_done[g][k & 1], k / 2, wacc, prof_on);
	if (k >= 1) flag_wait(&sm.a_done[g], k, wacc, prof_on);
	/* ---- S3: a = s * (5 / avg_before)  (agc.c:27,31); a bypassed (exact zero) sample passes as 0.
	 * Lane l owns samples l + 32 i (conflict-free scalar accesses). ---- */
	{
		const bool zs = sm.svz[ss] != 0;
		const float *sp = sm.s[ss][g], *vp = sm.v[ss][g];
		float sv[8], vv[8], gn[8];
		bool redo = false;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			sv[i] = sp[lane + 32 * i];
			vv[i] = vp[lane + 32 * i];
		}
		/* the eight divisions as one branch-free block (strict_math.cuh), the out-of-range ones redone exactly */
#pragma unroll
		for (int i = 0; i < 8; i++) {
			gn[i] = fdiv_inrange(5.0f, vv[i]);
			redo |= !fdiv_inrange_ok(5.0f, vv[i]);
		}
		if (__builtin_expect(redo, 0)) {
#pragma unroll
			for (int i = 0; i < 8; i++)
				if (!fdiv_inrange_ok(5.0f, vv[i])) gn[i] = fdiv(5.0f, vv[i]);
		}
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const int t = lane + 32 * i;
			float o = 0.0f;
			if (t < n && !(zs && __float_as_uint(sv[i]) == ZSENT)) o = fmul(sv[i], gn[i]);
			/* q_store(q, 48 + t, o) with the two target slots precomputed per lane (qoff): a slot the sample does
			 * not have is the row's spare pair, so both stores are unconditional */
			reinterpret_cast<float *>(q)[qoff.x[i]] = o;
			reinterpret_cast<float *>(q)[qoff.y[i]] = o;
		}
	}
	warp_sync_hard();
	if (lane == 0) {
		mbar_arrive(&sm.svfree[ss]);
		if (AGC_SPLIT) mbar_arrive(&sm.vfree[ss]);
	}
	/* head = the last 48 inputs before this tile: a_h_prev[256 + m], kept as the .y of pair 192 + m (straight-line on
	 * purpose, see warp_sync_hard) */
	static_assert(SONDE_FIR_HIST == 48, "two entries for the first 16 lanes, one for the others");
	q[lane].x = qprev[q_phys(192 + lane)].y;
	if (lane < 16) q[32 + lane].x = qprev[q_phys(192 + 32 + lane)].y;
	warp_sync_hard();
	if (lane == 0) flag_publish(&sm.a_done[g], k + 1);
	if (k == ntiles - 1) {
		/* filter memory for the next call: the last 48 inputs */
		float *hist = p.st[sm.chan[g]].hist;
		hist[lane] = q_load(q, n + lane);
		if (lane < 16) hist[32 + lane] = q_load(q, n + 32 + lane);
	}

	/* ---- S4: FIR, reference summation order (filter.c:59-61).  Lane -> window: quarter warp w = lane / 8, region
	 * rho = (lane / 4) % 2, t = lane % 4: pairs j0 .. j0 + 3 with j0 = 128 rho + 16 w + 4 t, i.e. the outputs of samples
	 * j0 + r (low halves) and j0 + 64 + r (high halves).  Output pair r shares one packed multiply and one packed add
	 * per tap (sm_100 FFMA2 / FADD2, round-to-nearest per half, never fused: a*c + (-0) == fl(a*c)). ---- */
	const int ys = k % NY;
	if (k >= NY) flag_wait(&sm.tm_rel, k - NY + 1, wacc, prof_on);
	{
		const int j0 = 128 * ((lane >> 2) & 1) + 16 * (lane >> 3) + 4 * (lane & 3);
		const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(q + q_phys(j0));
		float *yrow = sm.y[g] + ys * (T * P);
#pragma unroll 1
		for (int br = 0; br < P; br++) {
			/* the window is (re)loaded per polyphase branch: 52 packed operands are 104 registers, which do not
			 * survive two unrolled 49-tap passes without spilling */
			unsigned long long w[SONDE_FIR_HIST + 4];
#pragma unroll
			for (int i = 0; i < (SONDE_FIR_HIST + 4) / 2; i++) {
				const ulonglong2 t = src[i];
				w[2 * i] = t.x;
				w[2 * i + 1] = t.y;
			}
			unsigned long long acc[4] = {0ull, 0ull, 0ull, 0ull};                  /* (+0, +0) */
			const unsigned long long *tp = reinterpret_cast<const unsigned long long *>(c_taps2[type] + br * SONDE_FIR_TAPS);
#pragma unroll
			for (int i = 0; i < SONDE_FIR_TAPS; i++) {
				const unsigned long long c = tp[i];
#pragma unroll
				for (int r = 0; r < 4; r++) acc[r] = add2(acc[r], fma2(w[r + i], c, negzero2));
			}
			/* unpack: low halves are outputs j0 + r, high halves j0 + 64 + r */
			float lo[4], hi[4];
#pragma unroll
			for (int r = 0; r < 4; r++) {
				lo[r] = __uint_as_float((uint32_t)acc[r]);
				hi[r] = __uint_as_float((uint32_t)(acc[r] >> 32));
			}
			if (P == 1) {
				*reinterpret_cast<float4 *>(yrow + j0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
				*reinterpret_cast<float4 *>(yrow + j0 + 64) = make_float4(hi[0], hi[1], hi[2], hi[3]);
				if (ys == 0 && j0 < YM)
					*reinterpret_cast<float4 *>(sm.y[g] + NY * T + j0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
			} else {
				/* slot of (sample t, branch br) = P t + (P - 1 - br): filter_get(phase) uses branch P-1-phase (filter.c:54) */
				const int o = P - 1 - br;
#pragma unroll
				for (int r = 0; r < 4; r++) {
					yrow[P * (j0 + r) + o] = lo[r];
					yrow[P * (j0 + 64 + r) + o] = hi[r];
				}
				if (ys == 0 && j0 < YM / P) {
#pragma unroll
					for (int r = 0; r < 4; r++) sm.y[g][NY * T * P + P * (j0 + r) + o] = lo[r];
				}
			}
		}
	}
	warp_sync_hard();
	if (lane == 0) {
		mbar_arrive(&sm.yfull[ys]);
		flag_publish(&sm.s4_done[g][k & 1], k / 2 + 1);
	}
}

/* ---- AG: both AGC recurrences of one channel over one tile (agc.c:23-28) ------------------------------------------------
 * bias:   s = x - bias ; bias = bias * (1 - 0.01) + s * 0.01
 * level:  v = avg      ; avg  = avg * (1 - 0.001) + |s| * 0.001
 * Two independent dependency chains (3 and 2 operations per sample) in one instruction stream; the chain of the bias
 * (12 cycles per sample) is the floor.  ncu on the first version of this loop (4 samples per iteration, results stored
 * at its end) showed a third of the time in short-scoreboard stalls: the instruction after an STS.128 that overwrites
 * one of its source registers waits until the store has read them, which under the shared-memory traffic of the other
 * warps takes ~20 cycles.  So blocks of 8 samples rotate through three register sets and a block is stored only after
 * the NEXT block has been computed: no register is rewritten within a block's time of the store that reads it. */
struct agc_block {
	float4 s0, s1, v0, v1;
};

__device__ __forceinline__ void agc_block8(const float4 xa, const float4 xb, float &bias, float &avg, agc_block &o)
{
	const float b1 = fsub(1.0f, 0.01f), b0 = 0.01f, g1 = fsub(1.0f, 0.001f), g0 = 0.001f;
	o.s0.x = fsub(xa.x, bias); bias = fadd(fmul(bias, b1), fmul(o.s0.x, b0));
	o.s0.y = fsub(xa.y, bias); bias = fadd(fmul(bias, b1), fmul(o.s0.y, b0));
	o.s0.z = fsub(xa.z, bias); bias = fadd(fmul(bias, b1), fmul(o.s0.z, b0));
	o.s0.w = fsub(xa.w, bias); bias = fadd(fmul(bias, b1), fmul(o.s0.w, b0));
	o.s1.x = fsub(xb.x, bias); bias = fadd(fmul(bias, b1), fmul(o.s1.x, b0));
	o.s1.y = fsub(xb.y, bias); bias = fadd(fmul(bias, b1), fmul(o.s1.y, b0));
	o.s1.z = fsub(xb.z, bias); bias = fadd(fmul(bias, b1), fmul(o.s1.z, b0));
	o.s1.w = fsub(xb.w, bias); bias = fadd(fmul(bias, b1), fmul(o.s1.w, b0));
	o.v0.x = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s0.x), g0));
	o.v0.y = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s0.y), g0));
	o.v0.z = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s0.z), g0));
	o.v0.w = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s0.w), g0));
	o.v1.x = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s1.x), g0));
	o.v1.y = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s1.y), g0));
	o.v1.z = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s1.z), g0));
	o.v1.w = avg; avg = fadd(fmul(avg, g1), fmul(fabsf(o.s1.w), g0));
}

__device__ __forceinline__ void agc_store8(float *__restrict__ s, float *__restrict__ v, const int i, const agc_block &o)
{
	*reinterpret_cast<float4 *>(s + i) = o.s0;
	*reinterpret_cast<float4 *>(s + i + 4) = o.s1;
	*reinterpret_cast<float4 *>(v + i) = o.v0;
	*reinterpret_cast<float4 *>(v + i + 4) = o.v1;
}

/* `check_zero`: the tile contains exact-zero samples, which bypass the AGC and do not update its state (agc.c:23);
 * their s is the sentinel ZSENT so that S3 can pass them as 0.  x rows have 8 floats of slack for the prefetch. */
__device__ __forceinline__ void agc_tile(const float *__restrict__ x, float *__restrict__ s, float *__restrict__ v, const int n,
                                         float &bias, float &avg, const bool check_zero)
{
	const float b1 = fsub(1.0f, 0.01f), b0 = 0.01f, g1 = fsub(1.0f, 0.001f), g0 = 0.001f;
	int i = 0;
	if (!check_zero) {
		const int nb = n >> 3;                   /* whole blocks of 8 */
		if (nb > 0) {
			agc_block A, B, C;
			const float4 *xp = reinterpret_cast<const float4 *>(x);
			float4 xa = xp[0], xb = xp[1];
			int b = 0;
			/* prime: block 0 -> A */
			{
				const float4 na = xp[2], nbv = xp[3];
				agc_block8(xa, xb, bias, avg, A);
				xa = na; xb = nbv;
			}
			/* steady state, three blocks per trip: compute the next block, then store the one before it */
			for (b = 1; b + 2 < nb; b += 3) {
				{
					const float4 na = xp[2 * b + 2], nbv = xp[2 * b + 3];
					agc_block8(xa, xb, bias, avg, B);
					agc_store8(s, v, 8 * (b - 1), A);
					xa = na; xb = nbv;
				}
				{
					const float4 na = xp[2 * b + 4], nbv = xp[2 * b + 5];
					agc_block8(xa, xb, bias, avg, C);
					agc_store8(s, v, 8 * b, B);
					xa = na; xb = nbv;
				}
				{
					const float4 na = xp[2 * b + 6], nbv = xp[2 * b + 7];
					agc_block8(xa, xb, bias, avg, A);
					agc_store8(s, v, 8 * (b + 1), C);
					xa = na; xb = nbv;
				}
			}
			/* block b-1 is in A and not stored yet; up to two more whole blocks */
			if (b < nb) {
				const float4 na = xp[2 * b + 2], nbv = xp[2 * b + 3];
				agc_block8(xa, xb, bias, avg, B);
				agc_store8(s, v, 8 * (b - 1), A);
				xa = na; xb = nbv;
				if (b + 1 < nb) {
					agc_block8(xa, xb, bias, avg, C);
					agc_store8(s, v, 8 * b, B);
					agc_store8(s, v, 8 * (b + 1), C);
				} else {
					agc_store8(s, v, 8 * b, B);
				}
			} else {
				agc_store8(s, v, 8 * (b - 1), A);
			}
			i = nb << 3;
		}
		for (; i < n; i++) {
			const float o = fsub(x[i], bias);
			bias = fadd(fmul(bias, b1), fmul(o, b0));
			s[i] = o;
			v[i] = avg;
			avg = fadd(fmul(avg, g1), fmul(fabsf(o), g0));
		}
	} else {
		for (; i < n; i++) {
			const float xi = x[i];
			v[i] = avg;
			if (xi == 0.0f) { s[i] = __uint_as_float(ZSENT); continue; }
			const float o = fsub(xi, bias);
			bias = fadd(fmul(bias, b1), fmul(o, b0));
			s[i] = o;
			avg = fadd(fmul(avg, g1), fmul(fabsf(o), g0));
		}
	}
}

/* ---- split AGC (AGC_SPLIT): one recurrence per warp ------------------------------------------------------------------
 * A1  s = x - bias ; bias = bias * 0.99 + s * 0.01        (3 dependent operations per sample)
 * A2  v = avg      ; avg  = avg * 0.999 + |s| * 0.001     (2), one tile behind A1, reading s from shared memory
 * In one warp the two chains share an in-order instruction stream and neither runs at its own latency (18.6 busy
 * cycles/sample with nothing else on the SM); on two warps each does.  Same register rotation as agc_tile: a block of 8 is
 * stored only after the next one has been computed. */
__device__ __forceinline__ void bias_tile(const float *__restrict__ x, float *__restrict__ s, const int n, float &bias, const bool check_zero)
{
	const float b1 = fsub(1.0f, 0.01f), b0 = 0.01f;
	if (!check_zero) {
		agc_bias_tile_fast(x, s, n, bias);
	} else {
		for (int i = 0; i < n; i++) {                 /* exact zeros bypass the AGC (agc.c:23) */
			const float xi = x[i];
			if (xi == 0.0f) { s[i] = __uint_as_float(ZSENT); continue; }
			const float o = fsub(xi, bias);
			bias = fadd(fmul(bias, b1), fmul(o, b0));
			s[i] = o;
		}
	}
}
__device__ __forceinline__ void level_tile(const float *__restrict__ s, float *__restrict__ v, const int n, float &avg, const bool check_zero)
{
	const float g1 = fsub(1.0f, 0.001f), g0 = 0.001f;
	if (!check_zero) {
		agc_level_tile_fast(s, v, n, avg);
	} else {
		for (int i = 0; i < n; i++) {
			const float si = s[i];
			v[i] = avg;
			if (__float_as_uint(si) == ZSENT) continue;
			avg = fadd(fmul(avg, g1), fmul(fabsf(si), g0));
		}
	}
}

template <int P, int KM0, int NM, int KS0, int NS, bool IQ, bool SOFT, bool TMA>
__global__ void __maxnreg__(64)
demod_pipe_kernel(const demod_params p, const int group_base, const int n_here, const __grid_constant__ k1_maps maps)
{
	if ((int)blockIdx.x >= n_here) return;            /* the padding CTA of an odd group count (launch_tpc_pairs) */
	extern __shared__ __align__(128) unsigned char smem_raw[];        /* 2-D TMA destinations want 128 bytes */
	smem_t<P> &sm = *reinterpret_cast<smem_t<P> *>(smem_raw);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nthreads = blockDim.x;
	const int grp = group_base + blockIdx.x;
	const int type = p.group_type[grp];
	const sonde_modem &md = c_modem[type];
	const int *chans = p.group_chan + (size_t)grp * G;
	const int L = p.len;
	const int ntiles = (L + T - 1) / T;
	const bool prof_on = p.prof != nullptr;
	long long wacc[2] = {0, 0};
	long long n_rounds = 0, n_slow = 0;
	const long long t_start = prof_on ? clock64() : 0;
	/* the channels of a group are packed at its front (build_groups) */
	int gact = 0;
#pragma unroll
	for (int g = 0; g < G; g++) gact += chans[g] >= 0;

	/* ---- prologue ------------------------------------------------------------------------- */
	if (tid == 0) {
		sm.negzero2 = 0x8000000080000000ull;
		sm.qnext = 0;
		for (int i = 0; i < NRAW; i++) { mbar_init(&sm.rawfull[i], 1); mbar_init(&sm.rawfree[i], gact); }
		for (int i = 0; i < NX; i++) { mbar_init(&sm.xfull[i], gact); sm.zflag[i] = 0; }
		/* split AGC: s is read by the gain stage of every channel AND by the level warp, v by the gain stage only */
		for (int i = 0; i < NSV; i++) { mbar_init(&sm.svfree[i], gact + (AGC_SPLIT ? 1 : 0)); mbar_init(&sm.vfree[i], gact); sm.svz[i] = 0; }
		for (int i = 0; i < NY; i++) mbar_init(&sm.yfull[i], gact);
		sm.ag_done = 0;
		sm.a1_done = 0;
		sm.tm_rel = 0;
		for (int g = 0; g < G; g++) { sm.a_done[g] = 0; sm.s4_done[g][0] = 0; sm.s4_done[g][1] = 0; }
	}
	for (int i = tid; i < G * SONDE_FIR_HIST; i += nthreads) {
		const int g = i / SONDE_FIR_HIST, j = i % SONDE_FIR_HIST;
		const int ch = chans[g];
		/* tile 0 reads its head from "the tail of tile -1" = buffer 1: a_h[256 + j] is the .y of pair 192 + j */
		sm.q[1][g][q_phys(192 + j)].y = (ch >= 0) ? p.st[ch].hist[j] : 0.0f;
	}
	if (tid < G) { sm.chan[tid] = chans[tid]; sm.row[tid] = chans[tid] >= 0 ? p.in_row[chans[tid]] : 0; }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();

	if (warp == W_LD) {
		/* =============================== LD: TMA producer ===================================
		 * one bulk copy per channel and tile (a 2 KB row segment, plus 2 samples of look-back so that the
		 * discriminator of a tile does not depend on the previous tile's item) into sm.raw, up to NRAW tiles ahead. */
		if (TMA && lane == 0) {
			constexpr uint32_t ESZ = IQ ? 8u : 4u;
			/* the tensor copy fetches tma_box_rows rows: a group that would reach past the input (the last, partial one of an
			 * interleaved batch) takes the row copies instead */
			const int step = p.tma_row_step;
			const int box = (p.tma_box_rows > 0 && sm.row[0] + (p.tma_box_rows - 1) * step < p.n_rows) ? p.tma_box_rows : 0;
			const int rres = box > 0 ? sm.row[0] % step : 0, rquo = box > 0 ? sm.row[0] / step : 0;
			for (int tile = 0; tile < ntiles; tile++) {
				const int slot = tile % NRAW;
				if (tile >= NRAW) mbar_wait(&sm.rawfree[slot], ((tile / NRAW) - 1) & 1);
				const bool lb = IQ && tile > 0;
				if (box > 0) {
					/* ONE tensor copy per tile: T columns x `box` consecutive rows (columns past the buffer are zero-filled
					 * and count towards the transaction), plus the 2-column look-back box */
					mbar_expect_tx(&sm.rawfull[slot], (uint32_t)box * (T * ESZ + (lb ? 16u : 0u)));
					tma_load_3d(&sm.rawt[slot][0][0], &maps.tile, tile * T, rres, rquo, &sm.rawfull[slot]);
					if (lb) tma_load_3d(&sm.rawlb[slot][0][0], &maps.lookback, tile * T - 2, rres, rquo, &sm.rawfull[slot]);
				} else {
					const int n = min(T, L - tile * T);
					const uint32_t bytes = (uint32_t)n * ESZ;
					mbar_expect_tx(&sm.rawfull[slot], (bytes + (lb ? 16u : 0u)) * (uint32_t)gact);
					for (int g = 0; g < gact; g++) {
						const char *src = static_cast<const char *>(p.in) + ((size_t)sm.row[g] * p.row_stride + (size_t)tile * T) * ESZ;
						void *dst = IQ ? static_cast<void *>(&sm.rawt[slot][g][0])
						               : static_cast<void *>(reinterpret_cast<float *>(&sm.rawt[slot][0][0]) + g * T);
						tma_load_1d(dst, src, bytes, &sm.rawfull[slot]);
						if (lb) tma_load_1d(&sm.rawlb[slot][g][0], src - 16, 16u, &sm.rawfull[slot]);
					}
				}
			}
		}
		return;
	} else if (AGC_SPLIT && warp == W_AG) {
		/* =============================== A1: the bias recurrence ============================ */
		const int g = lane & (G - 1);
		const bool own = lane < gact;
		float bias = own ? p.st[chans[g]].agc_bias : 0.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int xs = k % NX, ss = k % NSV;
			mbar_wait_t(&sm.xfull[xs], (k / NX) & 1, wacc[0], prof_on);
			if (k >= NSV) mbar_wait_t(&sm.svfree[ss], ((k / NSV) - 1) & 1, wacc[1], prof_on);
			const int z = sm.zflag[xs];
			if (own) bias_tile(sm.x[xs][g], sm.s[ss][g], n, bias, z != 0);
			warp_sync_hard();
			if (lane == 0) {
				sm.svz[ss] = z;
				sm.zflag[xs] = 0;
				flag_publish(&sm.a1_done, k + 1);
			}
		}
		if (own) p.st[chans[g]].agc_bias = bias;
	} else if (AGC_SPLIT && warp == W_A2) {
		/* =============================== A2: the level recurrence, one tile behind A1 ====== */
		const int g = lane & (G - 1);
		const bool own = lane < gact;
		float avg = own ? p.st[chans[g]].agc_avg : 5.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int ss = k % NSV;
			flag_wait(&sm.a1_done, k + 1, wacc[0], prof_on);
			if (k >= NSV) mbar_wait_t(&sm.vfree[ss], ((k / NSV) - 1) & 1, wacc[1], prof_on);
			const int z = sm.svz[ss];
			if (own) level_tile(sm.s[ss][g], sm.v[ss][g], n, avg, z != 0);
			warp_sync_hard();
			if (lane == 0) {
				mbar_arrive(&sm.svfree[ss]);                 /* this warp is done reading s */
				flag_publish(&sm.ag_done, k + 1);
			}
		}
		if (own) p.st[chans[g]].agc_avg = avg;
	} else if (!AGC_SPLIT && warp == W_AG) {
		/* =============================== AG: the two AGC recurrences ======================== */
		const int g = lane & (G - 1);
		const bool own = lane < gact;
		float bias = own ? p.st[chans[g]].agc_bias : 0.0f;
		float avg = own ? p.st[chans[g]].agc_avg : 5.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int xs = k % NX, ss = k % NSV;
			mbar_wait_t(&sm.xfull[xs], (k / NX) & 1, wacc[0], prof_on);
			if (k >= NSV) mbar_wait_t(&sm.svfree[ss], ((k / NSV) - 1) & 1, wacc[1], prof_on);
			const int z = sm.zflag[xs];
			if (own) agc_tile(sm.x[xs][g], sm.s[ss][g], sm.v[ss][g], n, bias, avg, z != 0);
			warp_sync_hard();
			if (lane == 0) {
				sm.svz[ss] = z;
				sm.zflag[xs] = 0;
				flag_publish(&sm.ag_done, k + 1);
			}
		}
		if (own) { p.st[chans[g]].agc_bias = bias; p.st[chans[g]].agc_avg = avg; }
	} else if (warp == W_TM) {
		/* =============================== TM: timing + slicer ================================
		 * lane = channel; see timing_exact.cuh.  Tile k's epoch runs every round whose candidate slots lie inside
		 * tiles <= k; a round that would reach into tile k+1 waits for the next epoch, so tile k-1 is released at
		 * the end of epoch k (the y ring holds NY = 3 tiles: PW can be a full tile ahead). */
		constexpr int TP = T * P, RING = NY * TP, W = KS0 + NS - 1;
		static_assert(W <= YM && W <= TP, "round window vs mirror");
		const int g = lane & (G - 1);
		const bool own = lane < gact;
		const int ch = own ? chans[g] : 0;
		tmx_regs tr = {};
		const tmx_consts tc = {md.freq0, md.alpha, md.beta, md.max_fdev};
		uint8_t *ring = p.ring + (size_t)ch * p.ring_bytes;
		float *soft = (own && p.soft) ? p.soft + (size_t)ch * p.soft_stride : nullptr;
		/* the ring is a power of two and far smaller than 2^32 bits, so the low 32 bits of the
		 * stream position address it */
		const uint32_t ring_mask = p.ring_bytes - 1;
		uint64_t nbits0 = 0;
		if (own) {
			const demod_state &st = p.st[ch];
			tr.prev = st.t_prev; tr.phase = st.t_phase; tr.freq = st.t_freq;
			tr.target = (float)st.t_state;
			tr.interm = 0.0f;                               /* gfsk.c:73 */
			nbits0 = st.nbits; tr.nb = (uint32_t)nbits0; tr.nsoft = 0;
			/* rebuild the 32-bit word in progress: its complete bytes are in the ring, the last < 8 bits in the state */
			const uint32_t cnt = tr.nb & 31u, wordoff = ((tr.nb >> 5) << 2) & ring_mask;
			uint32_t acc = 0;
			for (uint32_t b = 0; b < (cnt >> 3); b++) acc = (acc << 8) | ring[wordoff + b];
			tr.acc = (acc << (cnt & 7u)) | (st.bit_acc & ((1u << (cnt & 7u)) - 1u));
		} else {
			tr.freq = tc.center; tr.target = 1.0f;
		}
		const int ns_total = L * P;
		int sabs = 0, sring = 0;
		const float *yrow = sm.y[g];
		for (int k = 0; k < ntiles; k++) {
			mbar_wait_t(&sm.yfull[k % NY], (k / NY) & 1, wacc[0], prof_on);
			const int end = (k == ntiles - 1) ? ns_total : (k + 1) * TP;
			if (own) {
				tmx_run<KM0, NM, KS0, NS, RING, SOFT>(tr, yrow, sabs, sring, end, tc, ring, ring_mask, soft, p.soft_stride, n_slow);
				if (k == ntiles - 1) {
					while (sabs < ns_total) {
						const int used = tmx_literal_round<SOFT>(tr, yrow + sring, min(W, ns_total - sabs), tc, ring, ring_mask, soft, p.soft_stride);
						sabs += used;
						sring += used;
						sring = (sring >= RING) ? sring - RING : sring;
					}
				}
			}
			warp_sync_hard();
			if (k >= 1 && lane == 0) flag_publish(&sm.tm_rel, k);      /* tiles 0 .. k-1 are consumed */
		}
		/* without the soft-symbol tap the fast rounds do not count symbols: one per bit of this call */
		if (!SOFT) tr.nsoft = (int)(tr.nb - (uint32_t)nbits0);
		n_rounds = tr.nsoft;
		if (own) {
			demod_state &st = p.st[ch];
			const uint64_t nbits = nbits0 + (uint64_t)(tr.nb - (uint32_t)nbits0);
			const uint32_t cnt = tr.nb & 31u, wordoff = ((tr.nb >> 5) << 2) & ring_mask, rem = cnt & 7u;
			for (uint32_t b = 0; b < (cnt >> 3); b++) ring[wordoff + b] = (uint8_t)(tr.acc >> (cnt - 8u * (b + 1u)));
			if (rem) ring[wordoff + (cnt >> 3)] = (uint8_t)(tr.acc << (8u - rem));
			st.t_prev = tr.prev; st.t_phase = tr.phase; st.t_freq = tr.freq; st.t_state = (int)tr.target;
			st.bit_acc = tr.acc & ((1u << rem) - 1u); st.bit_cnt = (int)rem; st.nbits = nbits; st.nsoft = tr.nsoft;
			p.nbits_out[ch] = nbits;
		}
	} else if ((p.pw_mask >> warp) & 1u) {
		/* =============================== PW: work queue ===================================== */
		const unsigned long long negzero2 = *reinterpret_cast<volatile unsigned long long *>(&sm.negzero2);
		const int total = (2 + 2 * ntiles) * gact;
		const q_offsets qoff = q_offsets_of(lane);
		for (;;) {
			const int idx = queue_pull(&sm.qnext, lane);
			if (idx >= total) break;
			const int b = idx / gact, g = idx - b * gact;
			if (b < 2 || !(b & 1)) {
				const int k = (b < 2) ? b : (b - 2) / 2 + 2;
				if (k < ntiles) item_disc<P, IQ, TMA>(sm, p, k, g, lane, ntiles, wacc[0], prof_on);
			} else {
				item_fir<P>(sm, p, (b - 3) / 2, g, lane, ntiles, type, negzero2, qoff, wacc[1], prof_on);
			}
			if (prof_on) n_rounds++;
		}
	} else {
		return;
	}
	if (prof_on && lane == 0) {
		/* per CTA: [role*4 + {wait0, wait1, total, counts}] ; roles: 0 = first PW warp, 1 = AG, 2 = last PW warp, 3 = TM */
		const int first_pw = __ffs(p.pw_mask) - 1, last_pw = 31 - __clz(p.pw_mask);
		const int role = warp == W_AG ? 1 : warp == W_TM ? 3 : warp == first_pw ? 0 : warp == last_pw ? 2 : -1;
		if (role >= 0) {
			long long *o = p.prof + (size_t)blockIdx.x * 16 + role * 4;
			o[0] = wacc[0]; o[1] = wacc[1]; o[2] = clock64() - t_start;
			o[3] = (n_rounds << 32) | n_slow;
		}
	}
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

/* The input of this call as a 2-D tensor [n_rows][len] (row pitch row_stride): columns past `len` and before 0 read as
 * zero, which is what the tile tail and the look-back of tile 0 want.  Driver entry point through the runtime (libcuda is
 * not linked); encoding is host arithmetic only. */
cudaError_t encode_k1_maps(k1_maps *m, const demod_params *p, bool iq)
{
	static encode_tiled_fn encode = [] {
		encode_tiled_fn f = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
		return f;
	}();
	if (!encode) return cudaErrorNotSupported;
	const size_t esz = iq ? 8 : 4;
	const int step = p->tma_row_step > 0 ? p->tma_row_step : 1;
	const cuuint64_t gdim[3] = {(cuuint64_t)p->len, (cuuint64_t)step, (cuuint64_t)((p->n_rows + step - 1) / step)};
	const cuuint64_t gstr[2] = {(cuuint64_t)p->row_stride * esz, (cuuint64_t)p->row_stride * esz * step};
	const cuuint32_t estr[3] = {1, 1, 1};
	const CUtensorMapDataType dt = iq ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
	const cuuint32_t box[3] = {(cuuint32_t)T, 1, (cuuint32_t)p->tma_box_rows};
	if (encode(&m->tile, dt, 3, const_cast<void *>(p->in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
		return cudaErrorInvalidValue;
	m->lookback = m->tile;
	if (iq) {
		const cuuint32_t boxl[3] = {2, 1, (cuuint32_t)p->tma_box_rows};
		if (encode(&m->lookback, dt, 3, const_cast<void *>(p->in), gdim, gstr, boxl, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
			return cudaErrorInvalidValue;
	}
	return cudaSuccess;
}

template <int P, int KM0, int NM, int KS0, int NS, bool IQ, bool SOFT, bool TMA>
cudaError_t launch3(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	static std::atomic<unsigned long long> attr_done{0};
	auto kern = demod_pipe_kernel<P, KM0, NM, KS0, NS, IQ, SOFT, TMA>;
	const cudaError_t ea = sonde_ensure_dynamic_smem(kern, (int)sizeof(smem_t<P>), attr_done);
	if (ea != cudaSuccess) return ea;
	const int nwarps = 32 - __builtin_clz(p->pw_mask | ROLE_MASK);
	k1_maps maps;
	memset(&maps, 0, sizeof(maps));
	if (TMA && p->tma_box_rows > 0) {
		const cudaError_t em = encode_k1_maps(&maps, p, IQ);
		if (em != cudaSuccess) return em;
	}
	return launch_tpc_pairs(kern, n_groups, nwarps * 32, sizeof(smem_t<P>), stream, p->tpc_pairs != 0, *p, group_base, n_groups, maps);
}

template <int P, int KM0, int NM, int KS0, int NS, bool IQ>
cudaError_t launch(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	if (p->use_tma)
		return p->soft ? launch3<P, KM0, NM, KS0, NS, IQ, true, true>(p, group_base, n_groups, stream)
		               : launch3<P, KM0, NM, KS0, NS, IQ, false, true>(p, group_base, n_groups, stream);
	return p->soft ? launch3<P, KM0, NM, KS0, NS, IQ, true, false>(p, group_base, n_groups, stream)
	               : launch3<P, KM0, NM, KS0, NS, IQ, false, false>(p, group_base, n_groups, stream);
}

}  // namespace

/* variant 0: 1 polyphase branch, ~10 NCO slots per symbol (RS41)          mid-symbol hit on slot 4..6, symbol on 9..11
 * variant 1: 1 branch, 19.2 / 20 slots per symbol (DFM, iMS-100, MRZ-N1)   mid-symbol 8..11, symbol 18..21
 * variant 2: 2 branches, 10 slots per symbol (M10/M20)                     as variant 0
 * (windows for the 48 kS/s modems; at other rates the rounds that miss them take the literal path) */
extern "C" cudaError_t sonde_launch_demod_pipe(const demod_params *p, int group_base, int n_groups, int variant,
                                               cudaStream_t stream)
{
	if (n_groups <= 0) return cudaSuccess;
	if (!(p->pw_mask & ~ROLE_MASK) || (p->pw_mask & ROLE_MASK)) return cudaErrorInvalidValue;
	switch (variant) {
	case 0:  return p->is_iq ? launch<1, 4, 3, 9, 3, true>(p, group_base, n_groups, stream) : launch<1, 4, 3, 9, 3, false>(p, group_base, n_groups, stream);
	case 1:  return p->is_iq ? launch<1, 8, 4, 18, 4, true>(p, group_base, n_groups, stream) : launch<1, 8, 4, 18, 4, false>(p, group_base, n_groups, stream);
	case 2:  return p->is_iq ? launch<2, 4, 3, 9, 3, true>(p, group_base, n_groups, stream) : launch<2, 4, 3, 9, 3, false>(p, group_base, n_groups, stream);
	default: return cudaErrorInvalidValue;
	}
}
