/*
 * demod_pipe_afsk.cu — K1 for the two AFSK sondes (iMet-1/4, SRS-C50): FM / IQ samples -> hard bits as a
 * warp-specialised software pipeline, the AFSK counterpart of demod_pipe.cu.
 *
 * Reference chain per sample (SD/demod/afsk.c:104-141):
 *      s   = agc_apply(x) / len * 2
 *      out = s * cexpf(-j p_mark)  ;  mark_sum  += out - mark_history[idx]     (boxcar over one symbol)
 *      out = s * cexpf(-j p_space) ;  space_sum += out - space_history[idx]
 *      f   = cabsf(mark_sum) - cabsf(space_sum)        -> 49-tap FIR -> Gardner loop -> slicer
 *      p   = fmod(p + f_tone, 2 pi)                    (double fmod, stored as float)
 *
 * Serial in time per channel: the two AGC recurrences, the two NCO phases (data independent), the four
 * running sums and the timing loop.  Each gets a warp whose lanes are channels (or channel x component);
 * everything else — discriminator, gain, sincos, the four mixer products, the boxcar differences
 * out[n] - out[n-len], the two square roots, the FIR at every position — is done by 16 parallel-work
 * warps.  Tiles are handed over through mbarrier-guarded shared-memory buffers:
 *
 *   HBM -S1-> x[3] -A1-> s[2] -A2-> v[2] -\
 *                              NC -> pn[2] -+-S3m-> o -> d -BX-> (sums, in place) -S3c-> a[2] -S4-> y[2] -TM-> bits
 *
 *   PW, iteration k:  S1(k+1)  S3m(k)  S4(k-1)  S3c(k)      (the FIR of the previous tile hides BX(k))
 *
 * The NCO phase sequence depends only on the number of samples a decoder has consumed, so all channels of
 * a handle normally share it; the prologue checks that bit for bit and, if so, sincos is evaluated once
 * per sample column instead of once per channel (8x fewer double-precision sincos).
 *
 * libm: as in demod.cu (the phase-by-phase kernel this one is checked against bit for bit) sincos is
 * evaluated in double and rounded to float, cabsf is (float)sqrt((double)re*re + (double)im*im) like glibc,
 * and fmod(x, 2 pi) for 0 <= x < 4 pi is one exact conditional subtraction.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "pipe_common.cuh"
#include "timing_round.cuh"
#include "timing_exact.cuh"

static __constant__ sonde_modem c_modem[SONDE_NTYPES_];

extern "C" cudaError_t sonde_upload_modems_afsk_pipe(const sonde_modem *m)
{
	return cudaMemcpyToSymbol(c_modem, m, sizeof(sonde_modem) * SONDE_NTYPES_);
}

namespace {

using namespace pipe;

/* Warp placement (scheduler = warp id % 4).  PW is the throughput-bound role of this kernel (two double
 * square roots per sample), the five serial lanes have slack:
 *   layout 0: all serial lanes on SMSP3                       SMSP0: 6 PW | SMSP1: 5 PW | SMSP2: 5 PW | SMSP3: TM A1 A2 NC BX
 *   layout 1: PW spread evenly, serial lanes on top           SMSP0: 4 PW, A1, NC | SMSP1: 4 PW, A2 | SMSP2: 4 PW, BX | SMSP3: 4 PW, TM */
template <int LAYOUT>
struct aroles;
template <>
struct aroles<0> {
	static constexpr int NWARPS = 21, W_TM = 3, W_A1 = 7, W_A2 = 11, W_NC = 15, W_BX = 19;
	static __device__ __forceinline__ int pw_index(int warp)
	{
		const int q = warp >> 2, r = warp & 3;        /* ids 0,4,..,20 -> 0..5 ; 1,5,..,17 -> 6..10 ; 2,6,..,18 -> 11..15 */
		return r == 0 ? q : r == 1 ? 6 + q : 11 + q;
	}
};
template <>
struct aroles<1> {
	static constexpr int NWARPS = 21, W_TM = 19, W_A1 = 16, W_A2 = 17, W_NC = 20, W_BX = 18;
	static __device__ __forceinline__ int pw_index(int warp) { return warp; }
};

constexpr int OS = SONDE_AFSK_MAXLEN + T;  /* mixer outputs kept per channel: `len` history + one tile */
constexpr int DS = T + 1;                  /* row stride (float4) of the difference / sum buffer: the 32 BX lanes hit 32 banks */

struct asmem_t {
	float x[NX][G][RS];                  /* discriminator output / FM input                       */
	float s[NS2][G][RS];                 /* bias-removed samples                                  */
	float v[NS2][G][RS];                 /* moving_avg before each sample's update                */
	float a[NS2][G][AS];                 /* filter input |mark| - |space|, [0,48) = previous tile tail */
	float y[NS2][1][G][RS];              /* FIR output                                            */
	float ph[G][RS];                     /* S1 scratch: phases, [g][0] = previous                 */
	float pn[2][2][G][RS];               /* NCO phase used for each sample: [slot][mark/space][row][t] */
	float4 o[G][OS];                     /* mixer outputs (mark re, im, space re, im); [0,len) = the previous `len` */
	float4 d[G][DS];                     /* out[n] - out[n-len]; BX turns it into the running sums in place */
	float2 pc[2][T];                     /* shared-phase case: (cos p, -sin p) per sample column, mark / space */
	float carry[2][G];
	float2 taps[SONDE_FIR_TAPS];
	int   zflag[NX];
	int   uniform;                       /* all channels of the CTA share the NCO phases          */
	unsigned long long negzero2;
	unsigned long long xfull[NX], sfull[NS2], sfree[NS2], vfull[NS2], vfree[NS2], yfull[NS2], yfree[NS2];
	unsigned long long pnfull[2], pnfree[2], dfull, bfull;
};

/* cabsf as glibc computes it.  Tried and measured slower (tools/stalls.py 5: PW 45.8 -> 51.3 busy cycles/sample): doing the
 * float <-> double conversions with integer operations (re-biased exponent, shifted mantissa, round-to-nearest-even on the
 * dropped bits) to take the three F2F per magnitude off the special-function pipe — the parallel warps of this kernel are
 * bound by issue slots, not by that pipe, and the ~16 extra integer instructions cost more than the conversions. */
__device__ __forceinline__ float cabs_exact(float re, float im)
{
	return (float)sqrt(__dadd_rn(__dmul_rn((double)re, (double)re), __dmul_rn((double)im, (double)im)));
}

__device__ __forceinline__ float nco_step(float p, float f)
{
	/* p' = (float)fmod((double)(p + f), 2 pi)   (afsk.c:127: float add, double fmod, float store).
	 * The smallest float >= 2 pi (double) is 0x40C90FDB, so for 0 <= x < that float fmod returns x itself and no
	 * double arithmetic is needed; a wrap (once per tone period) takes the exact double subtraction:
	 * fmod(x, 2 pi) for 2 pi <= x < 4 pi is x - 2 pi, exact in double (Sterbenz). */
	const double two_pi = 2.0 * 3.14159265358979323846;
	const float x = fadd(p, f);
	if (x >= 0.0f && x < __uint_as_float(0x40C90FDBu)) return x;
	const double xd = (double)x;
	return (xd >= two_pi && xd < 2.0 * two_pi) ? (float)__dsub_rn(xd, two_pi) : (float)fmod(xd, two_pi);
}

template <bool IQ, bool SOFT, int LAYOUT>
__global__ void __launch_bounds__(aroles<LAYOUT>::NWARPS * 32, 1)
demod_pipe_afsk_kernel(const demod_params p, const int group_base, const int n_here)
{
	if ((int)blockIdx.x >= n_here) return;            /* the padding CTA of an odd group count (launch_tpc_pairs) */
	extern __shared__ __align__(16) unsigned char smem_raw[];
	asmem_t &sm = *reinterpret_cast<asmem_t *>(smem_raw);

	using RL = aroles<LAYOUT>;
	constexpr int NTHREADS = RL::NWARPS * 32;
	constexpr int W_A1 = RL::W_A1, W_A2 = RL::W_A2, W_TM = RL::W_TM, W_NC = RL::W_NC, W_BX = RL::W_BX;
	constexpr int N = 24;                     /* NCO slots per timing round (20 or 40 slots per symbol) */
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int grp = group_base + blockIdx.x;
	const sonde_modem &md = c_modem[p.group_type[grp]];
	const int *chans = p.group_chan + (size_t)grp * G;
	const int L = p.len;
	const int ntiles = (L + T - 1) / T;
	const int blen = md.boxcar_len;
	const bool prof_on = p.prof != nullptr;
	long long wacc[2] = {0, 0};
	long long n_rounds = 0, n_slow = 0;
	const long long t_start = prof_on ? clock64() : 0;

	/* ---- prologue ------------------------------------------------------------------------- */
	if (tid == 0) {
		sm.negzero2 = 0x8000000080000000ull;
		for (int i = 0; i < NX; i++) { mbar_init(&sm.xfull[i], NPW); sm.zflag[i] = 0; }
		for (int i = 0; i < NS2; i++) {
			mbar_init(&sm.sfull[i], 1); mbar_init(&sm.sfree[i], 1 + NPW);
			mbar_init(&sm.vfull[i], 1); mbar_init(&sm.vfree[i], NPW);
			mbar_init(&sm.yfull[i], NPW); mbar_init(&sm.yfree[i], 1);
			mbar_init(&sm.pnfull[i], 1); mbar_init(&sm.pnfree[i], NPW);
		}
		mbar_init(&sm.dfull, NPW); mbar_init(&sm.bfull, 1);
		/* do all active channels share the NCO phases (bit for bit)? */
		int first = -1, uni = 1;
		for (int g = 0; g < G; g++) {
			const int ch = chans[g];
			if (ch < 0) continue;
			if (first < 0) { first = ch; continue; }
			uni &= __float_as_uint(p.ast[ch].p_mark) == __float_as_uint(p.ast[first].p_mark) &&
			       __float_as_uint(p.ast[ch].p_space) == __float_as_uint(p.ast[first].p_space);
		}
		sm.uniform = uni && chans[0] >= 0;       /* row 0 is the representative */
	}
	for (int i = tid; i < SONDE_FIR_TAPS; i += NTHREADS) sm.taps[i] = make_float2(md.taps[i], md.taps[i]);
	for (int i = tid; i < G * SONDE_FIR_HIST; i += NTHREADS) {
		const int g = i / SONDE_FIR_HIST, k = i % SONDE_FIR_HIST;
		const int ch = chans[g];
		sm.a[1][g][T + k] = (ch >= 0) ? p.st[ch].hist[k] : 0.0f;       /* tile 0 reads its head from "slot 1's tail" */
	}
	for (int i = tid; i < G * blen; i += NTHREADS) {
		const int g = i / blen, k = i % blen;
		const int ch = chans[g];
		sm.o[g][k] = (ch >= 0) ? make_float4(p.ast[ch].mark_hist[2 * k], p.ast[ch].mark_hist[2 * k + 1],
		                                      p.ast[ch].space_hist[2 * k], p.ast[ch].space_hist[2 * k + 1])
		                       : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	if (tid < G) sm.carry[0][tid] = (chans[tid] >= 0) ? p.st[chans[tid]].disc_prev : 0.0f;
	__syncthreads();

	if (warp != W_A1 && warp != W_A2 && warp != W_TM && warp != W_NC && warp != W_BX) {
		/* =============================== PW ================================================= */
		const int pw = RL::pw_index(warp);
		const int pt = pw * 32 + lane;
		const int t = pt % T;                    /* sample column owned in the per-sample stages  */
		const int g0 = (pt / T) * CPT;           /* first of the CPT channels owned there         */
		const int fir_g = pw % G;
		const int fir_seg = (pw / G) * 32 + lane;
		const bool uniform = sm.uniform != 0;
		const float flen = (float)blen;
		int ch_of[CPT], row_of[CPT];
#pragma unroll
		for (int c = 0; c < CPT; c++) {
			ch_of[c] = chans[g0 + c];
			row_of[c] = ch_of[c] >= 0 ? p.in_row[ch_of[c]] : 0;
		}
		const unsigned long long negzero2 = *reinterpret_cast<volatile unsigned long long *>(&sm.negzero2);

		float2 q[CPT];                           /* register prefetch, one tile ahead            */
		auto prefetch = [&](int tile) {
			if (tile >= ntiles) return;
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
		/* S1 of `tile` (same as demod_pipe.cu) */
		auto stage1 = [&](int tile) {
			const int slot = tile % NX;
			const int n = min(T, L - tile * T);
			float cur[CPT];
			if (pt == 0) sm.zflag[slot] = 0;
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
			prefetch(tile + 1);
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

		for (int k = 0; k <= ntiles; k++) {
			/* ---- S1(k+1) ---- */
			if (k + 1 < ntiles) stage1(k + 1);

			const int n = (k < ntiles) ? min(T, L - k * T) : 0;
			const int xs = k % NX, ss = k % NS2;
			const uint32_t par = (k / NS2) & 1;

			if (k < ntiles) {
				/* ---- S3m(k): gain, / len * 2, the two mixers (afsk.c:104-118) ---- */
				mbar_wait_t(&sm.vfull[ss], par, wacc[0], prof_on);            /* implies sfull[ss] */
				mbar_wait_t(&sm.pnfull[ss], par, wacc[0], prof_on);
				const bool zslow = sm.zflag[xs] != 0;
				float csm = 0.0f, snm = 0.0f, css = 0.0f, sns = 0.0f;
				if (uniform) {
					/* one sincos per thread: thread group 0 evaluates the mark tone, group 1 the space tone */
					const int tone = pt / T;
					double sn, cs;
					sincos((double)sm.pn[ss][tone][0][t], &sn, &cs);
					sm.pc[tone][t] = make_float2((float)cs, -(float)sn);
					pw_barrier();
					const float2 pm = sm.pc[0][t], psp = sm.pc[1][t];
					csm = pm.x; snm = pm.y; css = psp.x; sns = psp.y;
				}
#pragma unroll
				for (int c = 0; c < CPT; c++) {
					const int g = g0 + c;
					float4 out = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
					if (t < n && ch_of[c] >= 0) {
						float ov = 0.0f;
						if (!(zslow && sm.x[xs][g][t] == 0.0f)) ov = fmul(sm.s[ss][g][t], fdiv(5.0f, sm.v[ss][g][t]));
						const float sv = fmul(fdiv(ov, flen), 2.0f);
						if (!uniform) {
							double sn, cs;
							sincos((double)sm.pn[ss][0][g][t], &sn, &cs); csm = (float)cs; snm = -(float)sn;
							sincos((double)sm.pn[ss][1][g][t], &sn, &cs); css = (float)cs; sns = -(float)sn;
						}
						out = make_float4(fmul(sv, csm), fmul(sv, snm), fmul(sv, css), fmul(sv, sns));
					}
					sm.o[g][blen + t] = out;
				}
				warp_arrive(&sm.sfree[ss], lane);
				if (lane == 0) { mbar_arrive(&sm.vfree[ss]); mbar_arrive(&sm.pnfree[ss]); }
				pw_barrier();
				/* boxcar differences; d is free: S3c(k-1) ran before this iteration's barriers */
#pragma unroll
				for (int c = 0; c < CPT; c++) {
					const int g = g0 + c;
					const float4 nw = sm.o[g][blen + t], od = sm.o[g][t];
					sm.d[g][t] = make_float4(fsub(nw.x, od.x), fsub(nw.y, od.y), fsub(nw.z, od.z), fsub(nw.w, od.w));
				}
				warp_arrive(&sm.dfull, lane);
				/* slide the mixer history: o[g][0..len) <- o[g][n..n+len) */
				float4 hv[(G * SONDE_AFSK_MAXLEN + NPWT - 1) / NPWT];
				pw_barrier();
#pragma unroll
				for (int j = 0; j < (G * SONDE_AFSK_MAXLEN + NPWT - 1) / NPWT; j++) {
					const int i = pt + j * NPWT;
					if (i < G * blen) hv[j] = sm.o[i / blen][n + i % blen];
				}
				pw_barrier();
#pragma unroll
				for (int j = 0; j < (G * SONDE_AFSK_MAXLEN + NPWT - 1) / NPWT; j++) {
					const int i = pt + j * NPWT;
					if (i < G * blen) sm.o[i / blen][i % blen] = hv[j];
				}
			}

			/* ---- S4(k-1): FIR of the previous tile (runs while BX integrates tile k) ---- */
			if (k >= 1) {
				const int fs = (k - 1) % NS2;
				mbar_wait_t(&sm.yfree[fs], (((k - 1) / NS2) & 1) ^ 1, wacc[1], prof_on);
				fir_segment<1>(sm.a[fs][fir_g], sm.y[fs], sm.taps, negzero2, fir_g, fir_seg);
				warp_arrive(&sm.yfull[fs], lane);
			}

			if (k < ntiles) {
				/* ---- S3c(k): filter input = |mark_sum| - |space_sum| (afsk.c:121) ---- */
				mbar_wait_t(&sm.bfull, k & 1, wacc[0], prof_on);
#pragma unroll
				for (int c = 0; c < CPT; c++) {
					const int g = g0 + c;
					float ov = 0.0f;
					if (t < n && ch_of[c] >= 0) {
						const float4 sv = sm.d[g][t];
						ov = fsub(cabs_exact(sv.x, sv.y), cabs_exact(sv.z, sv.w));
					}
					sm.a[ss][g][SONDE_FIR_HIST + t] = ov;
				}
				/* head = tail of the previous tile */
				for (int i = pt; i < G * SONDE_FIR_HIST; i += NPWT) {
					const int g = i / SONDE_FIR_HIST, j = i % SONDE_FIR_HIST;
					sm.a[ss][g][j] = sm.a[ss ^ 1][g][T + j];
				}
				pw_barrier();
			}
		}

		/* save the filter memory (last 48 inputs), the discriminator phase and the mixer history */
		{
			const int ls = (ntiles - 1) % NS2;
			const int nl = L - (ntiles - 1) * T;
			for (int i = pt; i < G * SONDE_FIR_HIST; i += NPWT) {
				const int g = i / SONDE_FIR_HIST, j = i % SONDE_FIR_HIST;
				if (chans[g] >= 0) p.st[chans[g]].hist[j] = sm.a[ls][g][nl + j];
			}
			if (IQ && pt < G && chans[pt] >= 0) p.st[chans[pt]].disc_prev = sm.carry[ntiles & 1][pt];
			for (int i = pt; i < G * blen; i += NPWT) {
				const int g = i / blen, j = i % blen;
				const int ch = chans[g];
				if (ch >= 0) {
					const float4 hv = sm.o[g][j];
					p.ast[ch].mark_hist[2 * j] = hv.x; p.ast[ch].mark_hist[2 * j + 1] = hv.y;
					p.ast[ch].space_hist[2 * j] = hv.z; p.ast[ch].space_hist[2 * j + 1] = hv.w;
				}
			}
		}
	} else if (warp == W_A1) {
		/* =============================== A1: bias recurrence (agc.c:24-25) ================== */
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
			if (lane < G) {
				if (sm.zflag[xs] != 0) agc_bias_tile(x, s, n, bias, true);
				else agc_bias_tile_fast(x, s, n, bias);              /* register-rotated blocks of 8 (pipe_common.cuh) */
			}
			warp_arrive(&sm.sfull[ss], lane);
		}
		if (own) p.st[chans[g]].agc_bias = bias;
	} else if (warp == W_A2) {
		/* =============================== A2: level recurrence (agc.c:27-28) ================= */
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
			if (lane < G) {
				if (sm.zflag[xs] != 0) agc_level_tile(s, x, v, n, avg, true);
				else agc_level_tile_fast(s, v, n, avg);
			}
			warp_arrive(&sm.vfull[ss], lane);
			if (lane == 0) mbar_arrive(&sm.sfree[ss]);
		}
		if (own) p.st[chans[g]].agc_avg = avg;
	} else if (warp == W_NC) {
		/* =============================== NC: the two NCO phases (afsk.c:127-128) =============
		 * lane = 2 * row + tone; data independent, so it simply runs ahead of the rest. */
		const int g = (lane >> 1) & (G - 1), m = lane & 1;
		const bool own = lane < 2 * G && chans[g] >= 0;
		const float f = m ? md.f_space : md.f_mark;
		float ph = own ? (m ? p.ast[chans[g]].p_space : p.ast[chans[g]].p_mark) : 0.0f;
		const bool in_range = ph >= 0.0f && ph < __uint_as_float(0x40C90FDBu) && f > 0.0f && f < 4.0f;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int ps = k & 1;
			mbar_wait_t(&sm.pnfree[ps], ((k >> 1) & 1) ^ 1, wacc[1], prof_on);
			if (lane < 2 * G) {
				float *__restrict__ dst = sm.pn[ps][m][g];
				if (in_range) {
					/* 0 <= p < 2 pi and 0 < f < 2 pi: p + f < 4 pi, the wrap is one exact double subtraction */
					const float lim = __uint_as_float(0x40C90FDBu);          /* smallest float >= 2 pi */
					/* Branch-free wrap.  For every float x in [2 pi, 2 pi + 4.28]
					 *     (float)((double)x - 2 pi)  ==  fl(fl(x - HI) + DL),   HI = 0x40C90FDB, DL = fl(HI - 2 pi)
					 * (x - HI is exact by Sterbenz; the identity is checked exhaustively over all 6.3 M floats of
					 * that interval in tests/test_oracle.py::test_nco_wrap_identity), so the reference's double
					 * fmod costs two float adds and a select and the serial chain is 4 dependent ops per sample. */
					const float HI = lim, DL = 1.74845553e-07f;
#pragma unroll 4
					for (int i = 0; i < n; i++) {
						dst[i] = ph;
						const float x = fadd(ph, f);
						const float w = fadd(fsub(x, HI), DL);
						ph = (x >= lim) ? w : x;
					}
				} else {
					for (int i = 0; i < n; i++) {
						dst[i] = ph;
						ph = nco_step(ph, f);
					}
				}
			}
			warp_arrive(&sm.pnfull[ps], lane);
		}
		if (own) { if (m) p.ast[chans[g]].p_space = ph; else p.ast[chans[g]].p_mark = ph; }
	} else if (warp == W_BX) {
		/* =============================== BX: running boxcar sums (afsk.c:111,116) ============
		 * lane = 4 * channel + component; sum += out[n] - out[n-len], the difference comes from PW. */
		const int g = lane >> 2, comp = lane & 3;
		const bool own = chans[g] >= 0;
		float sum = 0.0f;
		if (own) {
			const afsk_state &as = p.ast[chans[g]];
			sum = comp == 0 ? as.mark_re : comp == 1 ? as.mark_im : comp == 2 ? as.space_re : as.space_im;
		}
		float *__restrict__ dv = reinterpret_cast<float *>(&sm.d[g][0]) + comp;
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			mbar_wait_t(&sm.dfull, k & 1, wacc[0], prof_on);
			int i = 0;
			float d0 = dv[0], d1 = dv[4], d2 = dv[8], d3 = dv[12];       /* the row has one float4 of slack */
			for (; i + 4 <= n; i += 4) {
				const int j = min(i + 4, T - 3);                            /* next group, loaded before this one is stored */
				const float e0 = dv[4 * j], e1 = dv[4 * j + 4], e2 = dv[4 * j + 8], e3 = dv[4 * j + 12];
				sum = fadd(sum, d0); dv[4 * i] = sum;
				sum = fadd(sum, d1); dv[4 * i + 4] = sum;
				sum = fadd(sum, d2); dv[4 * i + 8] = sum;
				sum = fadd(sum, d3); dv[4 * i + 12] = sum;
				d0 = e0; d1 = e1; d2 = e2; d3 = e3;
			}
			for (; i < n; i++) {
				sum = fadd(sum, dv[4 * i]);
				dv[4 * i] = sum;
			}
			warp_arrive(&sm.bfull, lane);
		}
		if (own) {
			afsk_state &as = p.ast[chans[g]];
			(comp == 0 ? as.mark_re : comp == 1 ? as.mark_im : comp == 2 ? as.space_re : as.space_im) = sum;
		}
	} else {
		/* =============================== TM: timing + slicer ==================================
		 * The chain-compare rounds of demod_pipe.cu (timing_exact.cuh) on a tile at a time: rounds whose candidate slots lie
		 * inside the tile take the fast path, the one or two that straddle its end the literal slot loop (which carries
		 * phase / state / interm across the boundary exactly as the reference's per-sample loop does).  Windows: iMet-1/4
		 * (1200 baud, 40 NCO slots per symbol) mid-symbol hit on slot 18..21, symbol on 38..41; SRS-C50 (2380 baud, 20.2
		 * slots) 8..11 and 18..21.  A round outside its window is replayed literally, so the choice only affects speed. */
		const int g = lane & (G - 1);
		const bool own = lane < G && chans[g] >= 0;
		const int ch = own ? chans[g] : 0;
		tmx_regs tr = {};
		const tmx_consts tc = {md.freq0, md.alpha, md.beta, md.max_fdev};
		uint8_t *ring = p.ring + (size_t)ch * p.ring_bytes;
		float *soft = (own && p.soft) ? p.soft + (size_t)ch * p.soft_stride : nullptr;
		const uint32_t ring_mask = p.ring_bytes - 1;
		uint64_t nbits0 = 0;
		if (own) {
			const demod_state &st = p.st[ch];
			tr.prev = st.t_prev; tr.phase = st.t_phase; tr.freq = st.t_freq;
			tr.target = (float)st.t_state;
			tr.interm = 0.0f;                               /* afsk.c:95 */
			nbits0 = st.nbits; tr.nb = (uint32_t)nbits0; tr.nsoft = 0;
			/* rebuild the 32-bit word in progress: its complete bytes are in the ring, the last < 8 bits in the state */
			const uint32_t cnt = tr.nb & 31u, wordoff = ((tr.nb >> 5) << 2) & ring_mask;
			uint32_t acc = 0;
			for (uint32_t bb = 0; bb < (cnt >> 3); bb++) acc = (acc << 8) | ring[wordoff + bb];
			tr.acc = (acc << (cnt & 7u)) | (st.bit_acc & ((1u << (cnt & 7u)) - 1u));
		} else {
			tr.freq = tc.center; tr.target = 1.0f;
		}
		const bool wide = md.freq0 < 0.075f;               /* 40 slots per symbol */
		constexpr int NOWRAP = 1 << 30;                      /* tiles are not a ring: the wrap of tmx_run never triggers */
		for (int k = 0; k < ntiles; k++) {
			const int n = min(T, L - k * T);
			const int ss = k % NS2;
			mbar_wait_t(&sm.yfull[ss], (k / NS2) & 1, wacc[0], prof_on);
			if (own) {
				const float *yrow = &sm.y[ss][0][g][0];
				int sabs = 0, sring = 0;
				if (wide) tmx_run<18, 4, 38, 4, NOWRAP, SOFT>(tr, yrow, sabs, sring, n, tc, ring, ring_mask, soft, p.soft_stride, n_slow);
				else      tmx_run<8, 4, 18, 4, NOWRAP, SOFT>(tr, yrow, sabs, sring, n, tc, ring, ring_mask, soft, p.soft_stride, n_slow);
				while (sabs < n)
					sabs += tmx_literal_round<SOFT>(tr, yrow + sabs, min(41, n - sabs), tc, ring, ring_mask, soft, p.soft_stride);
			}
			__syncwarp();
			warp_arrive(&sm.yfree[ss], lane);
		}
		if (!SOFT) tr.nsoft = (int)(tr.nb - (uint32_t)nbits0);
		n_rounds = tr.nsoft;
		if (own) {
			demod_state &st = p.st[ch];
			const uint64_t nbits = nbits0 + (uint64_t)(tr.nb - (uint32_t)nbits0);
			const uint32_t cnt = tr.nb & 31u, wordoff = ((tr.nb >> 5) << 2) & ring_mask, rem = cnt & 7u;
			for (uint32_t bb = 0; bb < (cnt >> 3); bb++) ring[wordoff + bb] = (uint8_t)(tr.acc >> (cnt - 8u * (bb + 1u)));
			if (rem) ring[wordoff + (cnt >> 3)] = (uint8_t)(tr.acc << (8u - rem));
			st.t_prev = tr.prev; st.t_phase = tr.phase; st.t_freq = tr.freq; st.t_state = (int)tr.target;
			st.bit_acc = tr.acc & ((1u << rem) - 1u); st.bit_cnt = (int)rem; st.nbits = nbits; st.nsoft = tr.nsoft;
			p.nbits_out[ch] = nbits;
		}
	}
	if (prof_on && lane == 0) {
		/* per CTA: [role*4 + {wait0, wait1, total}] ; roles: 0 = PW (pw 0), 1 = A1, 2 = BX, 3 = TM */
		const int role = warp == W_A1 ? 1 : warp == W_BX ? 2 : warp == W_TM ? 3
		                 : (warp != W_A2 && warp != W_NC && RL::pw_index(warp) == 0 ? 0 : -1);
		if (role >= 0) {
			long long *o = p.prof + (size_t)(group_base + blockIdx.x) * 16 + role * 4;
			o[0] = wacc[0]; o[1] = wacc[1]; o[2] = clock64() - t_start;
			o[3] = (n_rounds << 32) | n_slow;
		}
	}
}

template <bool IQ, bool SOFT, int LAYOUT>
cudaError_t launch2(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	static std::atomic<unsigned long long> attr_done{0};
	const cudaError_t ea = sonde_ensure_dynamic_smem(demod_pipe_afsk_kernel<IQ, SOFT, LAYOUT>, (int)sizeof(asmem_t), attr_done);
	if (ea != cudaSuccess) return ea;
	return launch_tpc_pairs(demod_pipe_afsk_kernel<IQ, SOFT, LAYOUT>, n_groups, aroles<LAYOUT>::NWARPS * 32, sizeof(asmem_t), stream,
	                              p->tpc_pairs != 0, *p, group_base, n_groups);
}

template <int LAYOUT>
cudaError_t launch(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	if (p->is_iq)
		return p->soft ? launch2<true, true, LAYOUT>(p, group_base, n_groups, stream)
		               : launch2<true, false, LAYOUT>(p, group_base, n_groups, stream);
	return p->soft ? launch2<false, true, LAYOUT>(p, group_base, n_groups, stream)
	               : launch2<false, false, LAYOUT>(p, group_base, n_groups, stream);
}

}  // namespace

extern "C" size_t sonde_demod_pipe_afsk_smem_bytes(void) { return sizeof(asmem_t); }

/* AFSK sondes (iMet-1/4, SRS-C50): 1 polyphase branch at 48 kS/s.  `layout` 0 (default) = PW warps spread over the four
 * schedulers (aroles<1>), 1 = serial lanes on their own scheduler (aroles<0>); within 2 % of each other on B200. */
extern "C" cudaError_t sonde_launch_demod_pipe_afsk(const demod_params *p, int group_base, int n_groups, int layout,
                                                    cudaStream_t stream)
{
	if (n_groups <= 0) return cudaSuccess;
	return layout ? launch<0>(p, group_base, n_groups, stream) : launch<1>(p, group_base, n_groups, stream);
}
