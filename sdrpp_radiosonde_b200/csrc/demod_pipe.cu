/*
 * demod_pipe.cu — K1 (production): IQ / FM samples -> hard bits, as a warp-specialised
 * software pipeline inside each CTA.
 *
 * One CTA owns PIPE_G = 8 channels of one sonde type and streams the chunk through tiles of
 * PIPE_T = 256 samples.  Eleven warps, four roles, tiles handed over through mbarrier-guarded
 * shared-memory rings, so that the three serial recurrences of the reference chain run
 * concurrently with each other and with the data-parallel work:
 *
 *   PW  warps 0-7   S1 load + FM discriminator (tile k+1, inputs prefetched one tile earlier)
 *                   S3 gain apply  a = s * (5 / avg)                       (tile k)
 *                   S4 49-tap FIR at every position / polyphase branch     (tile k)
 *   A1  warp 8      AGC bias recurrence      s = x - bias ; bias = .99 bias + .01 s   (agc.c:24-25)
 *   A2  warp 9      AGC level recurrence     v = avg ; avg = .999 avg + .001 |s|      (agc.c:27-28)
 *   TM  warp 10     Gardner NCO + retime + slicer, event driven: each lane (= channel) jumps from
 *                   timing hit to timing hit and only *selects* precomputed FIR outputs
 *                                                                         (timing.c:28-76, gfsk.c:75-125)
 *
 *        HBM --S1--> x[3] --A1--> s[2] --A2--> v[2] --S3--> a[2] --S4--> y[2] --TM--> bit ring (HBM)
 *
 * Lanes of A1/A2/TM are channels, so a serial step costs one warp instruction for all 8
 * channels; the per-sample critical paths are 3 dependent fp32 ops (A1), 2 (A2) and one add
 * per NCO slot plus ~20 ops per symbol (TM).  Everything is the reference's fp32 operation
 * order (strict_math.cuh), so soft symbols stay bit-identical.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "pipe_common.cuh"
#include "timing_round.cuh"

static __constant__ sonde_modem c_modem[SONDE_NTYPES_];

extern "C" cudaError_t sonde_upload_modems_pipe(const sonde_modem *m)
{
	return cudaMemcpyToSymbol(c_modem, m, sizeof(sonde_modem) * SONDE_NTYPES_);
}

namespace {

using namespace pipe;

/* Warp roles by scheduler (warp id % 4).  The serial lanes are latency-bound (IPC ~0.3) and lose issue slots
 * to any PW warp on their scheduler.  Two placements, picked per kernel variant from measurements
 * (tools/stalls.py):
 *   layout 0 (timing lane is the critical stage: RS41, M10): TM alone on SMSP3
 *        SMSP0: 6 PW, LD | SMSP1: 6 PW | SMSP2: A1, A2, 4 PW | SMSP3: TM (+ 5 warps that exit at once); the producer's
 *        barrier polling next to the timing lane cost it ~2 cycles/sample
 *   layout 1 (AGC lanes are the critical stage: DFM, iMS-100, MRZ-N1): all three serial lanes on SMSP3
 *        SMSP0: 6 PW | SMSP1: 5 PW | SMSP2: 5 PW | SMSP3: TM, A1, A2, LD
 * LD is the TMA producer warp: one lane issues the bulk copies of the input tiles (it exits at once when the
 * input rows are not 16-byte aligned and the PW threads load their samples themselves). */
template <int LAYOUT>
struct roles;
template <>
struct roles<0> {
	static constexpr int NWARPS = 25, W_TM = 3, W_A1 = 2, W_A2 = 6, W_LD = 24;
	static __device__ __forceinline__ bool idle(int warp) { return (warp & 3) == 3 && warp != W_TM; }
	static __device__ __forceinline__ int pw_index(int warp)
	{
		const int q = warp >> 2, r = warp & 3;        /* ids 0,4,..,20 -> 0..5 ; 1,5,..,21 -> 6..11 ; 10,14,18,22 -> 12..15 */
		return r == 0 ? q : r == 1 ? 6 + q : 12 + (q - 2);
	}
};
template <>
struct roles<1> {
	static constexpr int NWARPS = 21, W_TM = 3, W_A1 = 7, W_A2 = 11, W_LD = 15;
	static __device__ __forceinline__ bool idle(int warp) { return (warp & 3) == 3 && warp > W_A2 && warp != W_LD; }
	static __device__ __forceinline__ int pw_index(int warp)
	{
		const int q = warp >> 2, r = warp & 3;        /* ids 0,4,..,20 -> 0..5 ; 1,5,..,17 -> 6..10 ; 2,6,..,18 -> 11..15 */
		return r == 0 ? q : r == 1 ? 6 + q : 11 + q;
	}
};
constexpr int NRAW = 3;                  /* bulk-copy ring: S1 runs two tiles ahead of S3, the copies one more */

template <int P>
struct smem_t {
	float x[NX][G][RS];                  /* discriminator output / FM input          */
	float s[NS2][G][RS];                 /* bias-removed samples                     */
	float v[NS2][G][RS];                 /* moving_avg before each sample's update   */
	float a[NS2][G][AS];                 /* AGC output, [0,48) = previous tile tail  */
	float y[NS2][P][G][RS];              /* FIR output per polyphase branch          */
	float2 raw[NRAW][G][T];              /* TMA landing zone: raw IQ (or FM in .x-packed form), NRAW tiles in flight */
	float ph[G][RS];                     /* S1 scratch: phases, [g][0] = previous    */
	float carry[2][G];                   /* last phase of the previous tile          */
	float2 taps[P * SONDE_FIR_TAPS];     /* each tap duplicated for the packed fp32x2 FIR */
	int   zflag[NX];                     /* tile contains exact-zero samples         */
	unsigned long long negzero2;         /* (-0.0f, -0.0f), read at run time so that the packed product stays an FFMA2 */
	unsigned long long rawfull[NRAW], rawfree[NRAW];   /* complete_tx barriers of the bulk copies / slot released by PW */
	int chan[G], row[G];
	unsigned long long xfull[NX], sfull[NS2], sfree[NS2], vfull[NS2], vfree[NS2], yfull[NS2], yfree[NS2];
};

template <int P, int N, bool IQ, bool SOFT, int LAYOUT, bool TMA>
__global__ void __launch_bounds__(roles<LAYOUT>::NWARPS * 32, 1)
demod_pipe_kernel(const demod_params p, const int group_base)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	smem_t<P> &sm = *reinterpret_cast<smem_t<P> *>(smem_raw);

	using RL = roles<LAYOUT>;
	constexpr int NTHREADS = RL::NWARPS * 32;
	constexpr int W_A1 = RL::W_A1, W_A2 = RL::W_A2, W_TM = RL::W_TM, W_LD = RL::W_LD;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int grp = group_base + blockIdx.x;
	const sonde_modem &md = c_modem[p.group_type[grp]];
	const int *chans = p.group_chan + (size_t)grp * G;
	const int L = p.len;
	const int ntiles = (L + T - 1) / T;
	const bool prof_on = p.prof != nullptr;
	long long wacc[2] = {0, 0};
	long long n_rounds = 0, n_slow = 0;
	const long long t_start = prof_on ? clock64() : 0;

	/* ---- prologue ------------------------------------------------------------------------- */
	if (tid == 0) {
		sm.negzero2 = 0x8000000080000000ull;
		for (int i = 0; i < NRAW; i++) { mbar_init(&sm.rawfull[i], 1); mbar_init(&sm.rawfree[i], NPW); }
		for (int i = 0; i < NX; i++) { mbar_init(&sm.xfull[i], NPW); sm.zflag[i] = 0; }
		for (int i = 0; i < NS2; i++) {
			mbar_init(&sm.sfull[i], 1); mbar_init(&sm.sfree[i], 1 + NPW);
			mbar_init(&sm.vfull[i], 1); mbar_init(&sm.vfree[i], NPW);
			mbar_init(&sm.yfull[i], NPW); mbar_init(&sm.yfree[i], 1);
		}
	}
	for (int i = tid; i < P * SONDE_FIR_TAPS; i += NTHREADS) sm.taps[i] = make_float2(md.taps[i], md.taps[i]);
	for (int i = tid; i < G * SONDE_FIR_HIST; i += NTHREADS) {
		const int g = i / SONDE_FIR_HIST, k = i % SONDE_FIR_HIST;
		const int ch = chans[g];
		/* tile 0 reads its head from "slot 1's tail" */
		sm.a[1][g][T + k] = (ch >= 0) ? p.st[ch].hist[k] : 0.0f;
	}
	if (tid < G) sm.carry[0][tid] = (chans[tid] >= 0) ? p.st[chans[tid]].disc_prev : 0.0f;
	if (tid < G) { sm.chan[tid] = chans[tid]; sm.row[tid] = chans[tid] >= 0 ? p.in_row[chans[tid]] : 0; }
	__syncthreads();

	if (RL::idle(warp)) return;
	if (warp == W_LD) {
		/* =============================== LD: TMA producer ===================================
		 * 8 bulk copies per tile (one 2 KB row segment per channel) into sm.raw, up to NRAW tiles ahead. */
		if (TMA && lane == 0) {
			constexpr uint32_t ESZ = IQ ? 8u : 4u;
			for (int tile = 0; tile < ntiles; tile++) {
				const int slot = tile % NRAW;
				if (tile >= NRAW) mbar_wait(&sm.rawfree[slot], ((tile / NRAW) - 1) & 1);
				const int n = min(T, L - tile * T);
				const uint32_t bytes = (uint32_t)n * ESZ;
				uint32_t total = 0;
#pragma unroll
				for (int g = 0; g < G; g++) total += (sm.chan[g] >= 0) ? bytes : 0u;
				mbar_expect_tx(&sm.rawfull[slot], total);
#pragma unroll
				for (int g = 0; g < G; g++)
					if (sm.chan[g] >= 0)
						tma_load_1d(&sm.raw[slot][g][0],
						            static_cast<const char *>(p.in) + ((size_t)sm.row[g] * p.row_stride + (size_t)tile * T) * ESZ,
						            bytes, &sm.rawfull[slot]);
			}
		}
		return;
	}
	if (warp != W_A1 && warp != W_A2 && warp != W_TM) {
		/* =============================== PW: S1 / S3 / S4 =================================== */
		const int pw = RL::pw_index(warp);           /* 0..NPW-1                                     */
		const int pt = pw * 32 + lane;           /* PW thread index                              */
		const int t = pt % T;                    /* sample column owned in S1 / S3               */
		const int g0 = (pt / T) * CPT;           /* first of the CPT channels owned in S1 / S3   */
		const int fir_g = pw % G;                /* channel row owned in S4                      */
		const int fir_seg = (pw / G) * 32 + lane;/* segment of R outputs owned in S4             */
		int ch_of[CPT], row_of[CPT];
#pragma unroll
		for (int c = 0; c < CPT; c++) {
			ch_of[c] = chans[g0 + c];
			row_of[c] = ch_of[c] >= 0 ? p.in_row[ch_of[c]] : 0;
		}

		const unsigned long long negzero2 = *reinterpret_cast<volatile unsigned long long *>(&sm.negzero2);
		/* Input staging.  When the rows are 16-byte aligned (p.use_tma, decided on the host) one elected thread
		 * streams each tile with 8 bulk copies (one 2 KB row segment per channel) into sm.raw two tiles ahead
		 * and the PW threads read their samples from shared memory; otherwise every thread prefetches its
		 * own samples into registers one tile ahead with coalesced loads. */
		constexpr bool tma = TMA;
		float2 q[CPT];                           /* register prefetch (non-TMA path)             */
		auto prefetch = [&](int tile) {
			if (tma || tile >= ntiles) return;
			const int i = tile * T + t;
#pragma unroll
			for (int c = 0; c < CPT; c++) {
				q[c] = make_float2(0.0f, 0.0f);
				if (i < L && ch_of[c] >= 0) {
					if (IQ) q[c] = __ldg(static_cast<const float2 *>(p.in) + (size_t)row_of[c] * p.row_stride + i);
					else    q[c].x = __ldg(static_cast<const float *>(p.in) + (size_t)row_of[c] * p.row_stride + i);
				}
			}
		};
		/* S1 of `tile`; starts the loads of a later tile */
		auto stage1 = [&](int tile) {
			const int slot = tile % NX;
			const int n = min(T, L - tile * T);
			float cur[CPT];
			if (pt == 0) sm.zflag[slot] = 0;
			if (tma) {
				mbar_wait_t(&sm.rawfull[tile % NRAW], (tile / NRAW) & 1, wacc[1], prof_on);
#pragma unroll
				for (int c = 0; c < CPT; c++) {
					q[c] = make_float2(0.0f, 0.0f);
					if (t < n && ch_of[c] >= 0) {
						if (IQ) q[c] = sm.raw[tile % NRAW][g0 + c][t];
						else    q[c].x = reinterpret_cast<const float *>(&sm.raw[tile % NRAW][g0 + c][0])[t];
					}
				}
				warp_arrive(&sm.rawfree[tile % NRAW], lane);   /* values are in registers: the slot may be refilled */
			}
#pragma unroll
			for (int c = 0; c < CPT; c++) {
				const int g = g0 + c;
				if (IQ) {
					const float phv = (t < n) ? det_phase(q[c].x, q[c].y) : 0.0f;
					sm.ph[g][t + 1] = phv;
					if (t == n - 1) sm.carry[(tile + 1) & 1][g] = phv;
				} else {
					cur[c] = q[c].x;
				}
			}
			if (IQ && pt < G) sm.ph[pt][0] = sm.carry[tile & 1][pt];
			if (!tma) prefetch(tile + 1);
			pw_barrier();
			bool zero = false;
#pragma unroll
			for (int c = 0; c < CPT; c++) {
				const int g = g0 + c;
				const float xv = IQ ? disc_step(sm.ph[g][t + 1], sm.ph[g][t], p.fm_gain) : cur[c];
				sm.x[slot][g][t] = xv;
				zero |= (t < n && ch_of[c] >= 0 && xv == 0.0f);
			}
			if (zero) atomicOr(&sm.zflag[slot], 1);
			warp_arrive(&sm.xfull[slot], lane);
		};

		prefetch(0);
		stage1(0);
		if (ntiles > 1) stage1(1);

		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int xs = k % NX, ss = k % NS2;
			const uint32_t par = (k / NS2) & 1;

			/* ---- S1(k+2) ---- */
			/* two tiles ahead: the AGC lanes need about one PW iteration for a tile (A1 then A2), so with S1 only one
			 * tile ahead S3(k) would wait for them */
			if (k + 2 < ntiles) stage1(k + 2);

			/* ---- S3(k): a = s * (5 / avg_before)  (agc.c:27,31); zero samples pass as 0 ---- */
			mbar_wait_t(&sm.vfull[ss], par, wacc[0], prof_on);            /* implies sfull[ss] (A2 consumed it first) */
			const bool zslow = sm.zflag[xs] != 0;
#pragma unroll
			for (int c = 0; c < CPT; c++) {
				const int g = g0 + c;
				float o = 0.0f;
				if (t < n && ch_of[c] >= 0 && !(zslow && sm.x[xs][g][t] == 0.0f))
					o = fmul(sm.s[ss][g][t], fdiv(5.0f, sm.v[ss][g][t]));
				sm.a[ss][g][SONDE_FIR_HIST + t] = o;
			}
			/* head = tail of the previous tile */
			for (int i = pt; i < G * SONDE_FIR_HIST; i += NPWT) {
				const int g = i / SONDE_FIR_HIST, j = i % SONDE_FIR_HIST;
				sm.a[ss][g][j] = sm.a[ss ^ 1][g][T + j];
			}
			warp_arrive(&sm.sfree[ss], lane);
			if (lane == 0) mbar_arrive(&sm.vfree[ss]);
			pw_barrier();

			/* ---- S4(k) ---- */
			mbar_wait_t(&sm.yfree[ss], par ^ 1, wacc[1], prof_on);
			fir_segment<P>(sm.a[ss][fir_g], sm.y[ss], sm.taps, negzero2, fir_g, fir_seg);
			warp_arrive(&sm.yfull[ss], lane);
		}

		/* save the filter memory (last 48 inputs) and the discriminator phase */
		pw_barrier();
		{
			const int ls = (ntiles - 1) % NS2;
			const int nl = L - (ntiles - 1) * T;
			for (int i = pt; i < G * SONDE_FIR_HIST; i += NPWT) {
				const int g = i / SONDE_FIR_HIST, j = i % SONDE_FIR_HIST;
				if (chans[g] >= 0) p.st[chans[g]].hist[j] = sm.a[ls][g][nl + j];
			}
			if (IQ && pt < G && chans[pt] >= 0) p.st[chans[pt]].disc_prev = sm.carry[ntiles & 1][pt];
		}
	} else if (warp == W_A1) {
		/* =============================== A1: bias recurrence ================================ */
		const int g = lane & (G - 1);
		const bool own = lane < G && chans[g] >= 0;
		float bias = own ? p.st[chans[g]].agc_bias : 0.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int xs = k % NX, ss = k % NS2;
			const uint32_t par = (k / NS2) & 1;
			mbar_wait_t(&sm.xfull[xs], (k / NX) & 1, wacc[0], prof_on);
			mbar_wait_t(&sm.sfree[ss], par ^ 1, wacc[1], prof_on);
			const float *__restrict__ x = sm.x[xs][g];
			float *__restrict__ s = sm.s[ss][g];
			if (lane < G) agc_bias_tile(x, s, n, bias, sm.zflag[xs] != 0);
			warp_arrive(&sm.sfull[ss], lane);
		}
		if (own) p.st[chans[g]].agc_bias = bias;
	} else if (warp == W_A2) {
		/* =============================== A2: level recurrence =============================== */
		const int g = lane & (G - 1);
		const bool own = lane < G && chans[g] >= 0;
		float avg = own ? p.st[chans[g]].agc_avg : 5.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int xs = k % NX, ss = k % NS2;
			const uint32_t par = (k / NS2) & 1;
			mbar_wait_t(&sm.sfull[ss], par, wacc[0], prof_on);
			mbar_wait_t(&sm.vfree[ss], par ^ 1, wacc[1], prof_on);
			const float *__restrict__ s = sm.s[ss][g];
			const float *__restrict__ x = sm.x[xs][g];
			float *__restrict__ v = sm.v[ss][g];
			if (lane < G) agc_level_tile(s, x, v, n, avg, sm.zflag[xs] != 0);
			warp_arrive(&sm.vfull[ss], lane);
			if (lane == 0) mbar_arrive(&sm.sfree[ss]);
		}
		if (own) p.st[chans[g]].agc_avg = avg;
	} else {
		/* =============================== TM: timing + slicer ================================
		 * lane = channel.  Per "round" a lane advances its NCO by up to N slots, i.e. to its next symbol
		 * instant (timing.c:28-43), then retimes and slices (timing.c:45-76, gfsk.c:99-115).
		 *
		 * The reference finds the hit slots by comparing after every add.  Here the slot numbers are
		 * PREDICTED arithmetically (c = ceil((threshold - phase) / freq)), the reference's chain of adds
		 * is then run for exactly that many slots as predicated FADDs (same operations, same order, so
		 * the phase value is the reference's), and the prediction is VERIFIED on the chain values
		 * (p[c-1] < threshold <= p[c]; the chain is monotone because freq > 0).  The critical path per
		 * symbol is therefore the add chain itself.  If any lane's check fails (rounding put a slot on
		 * the other side of a threshold, or a pathological state), the round is redone for the warp with
		 * the literal slot-by-slot loop, so the result is exact in every case. */
		const int g = lane & (G - 1);
		const bool own = lane < G && chans[g] >= 0;
		const int ch = own ? chans[g] : 0;
		tm_regs tr = {};
		const float center = md.freq0, alpha = md.alpha, beta = md.beta, max_fdev = md.max_fdev;
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
			tr.acc = st.bit_acc; nbits0 = st.nbits; tr.nb = (uint32_t)nbits0; tr.nsoft = 0;
		} else {
			tr.freq = center; tr.target = 1.0f;
		}
		float rf = rcp_approx(tr.freq);                      /* only steers the prediction */
		/* A predicted mid-symbol slot is trusted only when (1 - phase)/freq is at least DELTA away from
		 * an integer.  |phase| < 4, so each of the <= N adds of the chain rounds by <= 2^-23 and the chain
		 * deviates from phase + k*freq by < N * 1.2e-7, i.e. N * 1.2e-7 / freq slots; the approximate
		 * quotient adds < 4 * 2^-23 * (4 / freq) slots.  DELTA is 8x that sum. */
		const float DELTA = 8.0f * ((float)(N + 16) * 1.2e-7f) / center;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int ss = k % NS2;
			mbar_wait_t(&sm.yfull[ss], (k / NS2) & 1, wacc[0], prof_on);
			const float (*y)[G][RS] = sm.y[ss];
			const int ns = own ? n * P : 0;
			tm_tile<P, N, SOFT, G, RS, (P == 2 ? 8 : 0)>(tr, rf, y, g, ns, center, alpha, beta, max_fdev, DELTA, ring, ring_mask, soft,
			                           p.soft_stride, n_rounds, n_slow, prof_on);
			__syncwarp();
			warp_arrive(&sm.yfree[ss], lane);
		}
		if (own) {
			demod_state &st = p.st[ch];
			const uint64_t nbits = nbits0 + (uint64_t)(tr.nb - (uint32_t)nbits0);
			const int cnt = (int)(tr.nb & 7u);
			st.t_prev = tr.prev; st.t_phase = tr.phase; st.t_freq = tr.freq; st.t_state = (int)tr.target;
			st.bit_acc = tr.acc & ((1u << cnt) - 1u); st.bit_cnt = cnt; st.nbits = nbits; st.nsoft = tr.nsoft;
			p.nbits_out[ch] = nbits;
			if (cnt) ring[(uint32_t)(nbits >> 3) & ring_mask] = (uint8_t)(tr.acc << (8 - cnt));
		}
	}
	if (prof_on && lane == 0) {
		/* per CTA: [role*4 + {wait0, wait1, total}] ; roles: 0 = PW (warp 0), 1 = A1, 2 = A2, 3 = TM */
		const int role = warp == W_A1 ? 1 : warp == W_A2 ? 2 : warp == W_TM ? 3 : (warp == 0 ? 0 : -1);
		if (role >= 0) {
			long long *o = p.prof + (size_t)blockIdx.x * 16 + role * 4;
			o[0] = wacc[0]; o[1] = wacc[1]; o[2] = clock64() - t_start;
			o[3] = (n_rounds << 32) | n_slow;
		}
	}
}

template <int P, int N, bool IQ, bool SOFT, bool TMA>
cudaError_t launch2(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	constexpr int LAYOUT = (N == 24) ? 1 : 0;
	static std::atomic<unsigned long long> attr_done{0};
	const cudaError_t ea = sonde_ensure_dynamic_smem(demod_pipe_kernel<P, N, IQ, SOFT, LAYOUT, TMA>, (int)sizeof(smem_t<P>), attr_done);
	if (ea != cudaSuccess) return ea;
	demod_pipe_kernel<P, N, IQ, SOFT, LAYOUT, TMA><<<n_groups, roles<LAYOUT>::NWARPS * 32, sizeof(smem_t<P>), stream>>>(*p, group_base);
	return cudaGetLastError();
}

template <int P, int N, bool IQ>
cudaError_t launch(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	if (p->use_tma)
		return p->soft ? launch2<P, N, IQ, true, true>(p, group_base, n_groups, stream)
		               : launch2<P, N, IQ, false, true>(p, group_base, n_groups, stream);
	return p->soft ? launch2<P, N, IQ, true, false>(p, group_base, n_groups, stream)
	               : launch2<P, N, IQ, false, false>(p, group_base, n_groups, stream);
}

}  // namespace

/* variant 0: 1 polyphase branch, <= 10 NCO slots per symbol (RS41)        -> 12-slot rounds
 * variant 1: 1 branch, ~20 slots per symbol (DFM, iMS-100, MRZ-N1)         -> 24-slot rounds
 * variant 2: 2 branches, 10 slots per symbol (M10/M20)                     -> 12-slot rounds */
extern "C" cudaError_t sonde_launch_demod_pipe(const demod_params *p, int group_base, int n_groups, int variant,
                                               cudaStream_t stream)
{
	if (n_groups <= 0) return cudaSuccess;
	switch (variant) {
	case 0:  return p->is_iq ? launch<1, 12, true>(p, group_base, n_groups, stream) : launch<1, 12, false>(p, group_base, n_groups, stream);
	case 1:  return p->is_iq ? launch<1, 24, true>(p, group_base, n_groups, stream) : launch<1, 24, false>(p, group_base, n_groups, stream);
	case 2:  return p->is_iq ? launch<2, 12, true>(p, group_base, n_groups, stream) : launch<2, 12, false>(p, group_base, n_groups, stream);
	default: return cudaErrorInvalidValue;
	}
}
