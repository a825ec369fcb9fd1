/*
 * demod.cu — K1: IQ / FM samples -> hard bits (+ optional soft symbols), batched over channels.
 *
 * One CTA owns DEMOD_G channels of the same sonde type and walks the chunk in tiles of
 * DEMOD_T samples.  Per tile:
 *
 *   S1  load + discriminate   all threads, coalesced        (upstream dsp::demod::FM, src/main.cpp:57)
 *   S2  AGC recurrences       1 lane / channel (serial IIR)  (demod/dsp/agc.c:19-34)
 *   S3  AGC gain apply        all threads (5/avg is off the recurrence)
 *   S4  49-tap FIR at EVERY sample position and polyphase branch, all threads,
 *       accumulated in the reference's order                 (demod/dsp/filter.c:41-64)
 *   S5  Gardner loop + slicer 1 lane / channel; only *selects* precomputed FIR outputs
 *                                                            (demod/dsp/timing.c:28-76, demod/gfsk.c:75-125)
 *
 * The only truly serial parts of the reference chain are the two AGC recurrences and the
 * timing NCO; filter_get() is a pure function of the AGC output at a position, so it is
 * evaluated speculatively everywhere in parallel and the timing lane picks what it needs
 * (SURVEY.md §7 "Design note for K1").  All arithmetic goes through strict_math.cuh, so soft
 * symbols are bit-identical to the CPU reference.
 *
 * The demodulator free-runs over the chunk: `interm` is cleared at chunk
 * start (gfsk.c:73), bits are appended to the channel's bit ring at absolute positions; the
 * framer is a separate kernel (frame.cu).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_state.h"
#include "strict_math.cuh"
#include "timing_round.cuh"

static __constant__ sonde_modem c_modem[SONDE_NTYPES_];

extern "C" cudaError_t sonde_upload_modems(const sonde_modem *m)
{
	return cudaMemcpyToSymbol(c_modem, m, sizeof(sonde_modem) * SONDE_NTYPES_);
}

namespace {

constexpr int G  = DEMOD_G;
constexpr int T  = DEMOD_T;
constexpr int NT = DEMOD_THREADS;
constexpr int RS = T + 4;                       /* row stride of the per-tile arrays (bank spread, 16 B aligned) */
constexpr int AS = SONDE_FIR_HIST + T + 4;      /* row stride of the FIR input array (48 history + tile)         */
constexpr int R  = 8;                           /* FIR outputs per thread                                         */

static_assert(G * T / R == NT, "one FIR segment per thread");
static_assert(T % 32 == 0 && T / R == 32, "a warp covers one channel row in S4");

struct smem_t {
	float x[G][RS];                /* S1 out: discriminator output / FM input             */
	float s[G][RS];                /* S2 out: bias-removed sample                         */
	float v[G][RS];                /* S2 out: moving_avg BEFORE this sample's update      */
	float a[G][AS];                /* S3 out: AGC output, [0,48) = previous tile's tail   */
	float y[SONDE_MAX_PHASES][G][RS];   /* S4 out: FIR output per polyphase branch       */
	float ph[G][RS];               /* S1 scratch: phases, ph[g][0] = previous sample      */
	float taps[SONDE_MAX_PHASES * SONDE_FIR_TAPS];
};

/* extra shared memory of the AFSK variant */
constexpr int OS2 = SONDE_AFSK_MAXLEN + T + 2;   /* row stride (float2) of the mixer outputs: `len` history + tile */
struct afsk_smem_t {
	float  pn[2][G][RS];           /* NCO phase used for each sample, [0] mark, [1] space        */
	float2 om[G][OS2], os[G][OS2]; /* mixer outputs; [0, len) = the previous `len` outputs       */
	float2 msum[G][RS], ssum[G][RS];   /* running boxcar sums after each sample                  */
};

/* ---- S2: the two AGC recurrences for one channel over n samples ---------------------- */
template <bool CHECK_ZERO>
__device__ __forceinline__ void agc_chains(const float *x, float *s, float *v, int n, float &bias, float &avg)
{
	const float kb1 = fsub(1.0f, 0.01f), kb0 = 0.01f;     /* agc.c:8,25  */
	const float kg1 = fsub(1.0f, 0.001f), kg0 = 0.001f;   /* agc.c:9,28  */
#pragma unroll 4
	for (int i = 0; i < n; i++) {
		const float xi = x[i];
		if (CHECK_ZERO && xi == 0.0f) {               /* agc.c:23: exact zero bypasses the AGC */
			s[i] = 0.0f;
			v[i] = 5.0f;
			continue;
		}
		const float si = fsub(xi, bias);
		bias = fadd(fmul(bias, kb1), fmul(si, kb0));
		s[i] = si;
		v[i] = avg;
		avg = fadd(fmul(avg, kg1), fmul(fabsf(si), kg0));
	}
}

/* ---- AFSK front end (afsk.c:104-141) -------------------------------------------------------
 * s = agc(x) / len * 2 ; mix with the mark and space NCOs ; running boxcar sums over one symbol ;
 * filter input = |mark_sum| - |space_sum|.  Only three things are serial in time: the AGC recurrences
 * (shared with GFSK), the two NCO phases p = fmod(p + f, 2 pi) — data independent — and the running
 * sums sum += out[n] - out[n - len].  They run on a few lanes; the expensive parts (sincos, sqrt) are
 * evaluated for all samples of the tile in parallel.
 * The reference calls glibc cexpf/cabsf/fmod per sample:
 *   - fmod(p + f, 2 pi) in double (p + f < 4 pi, so it is one exact conditional subtraction) and
 *     cabsf == (float)sqrt((double)re*re + (double)im*im) are reproduced exactly;
 *   - sin/cos are evaluated in double and rounded to float, which equals glibc's sinf/cosf except
 *     where glibc's own result is not the correctly rounded one (< 1 ulp apart).  The boxcar sums
 *     integrate such differences, so AFSK soft symbols are NOT claimed bit-exact;
 *     frame bytes are what tests/test_gpu_parity.py checks for the two AFSK sondes. */
__device__ __forceinline__ float cabs_exact(float re, float im)
{
	return (float)sqrt(__dadd_rn(__dmul_rn((double)re, (double)re), __dmul_rn((double)im, (double)im)));
}

__device__ __forceinline__ float nco_step(float p, float f)
{
	/* fmod(x, 2 pi) for 0 <= x < 4 pi is x or x - 2 pi, and that subtraction is exact in double (Sterbenz) */
	const double two_pi = 2.0 * 3.14159265358979323846;
	const double x = (double)fadd(p, f);
	return (x >= 0.0 && x < 2.0 * two_pi) ? (float)(x >= two_pi ? __dsub_rn(x, two_pi) : x) : (float)fmod(x, two_pi);
}

/* ---- S4: FIR at R consecutive positions, reference summation order -------------------- */
template <int P>
__device__ __forceinline__ void fir_segment(smem_t &sm, int g, int seg)
{
	/* window of inputs a[n-48 .. n] for n = seg*R .. seg*R+R-1  ->  R+48 values */
	float w[R + SONDE_FIR_HIST];
	const float4 *src = reinterpret_cast<const float4 *>(&sm.a[g][seg * R]);
#pragma unroll
	for (int k = 0; k < (R + SONDE_FIR_HIST) / 4; k++) {
		const float4 q = src[k];
		w[4 * k + 0] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
	}
#pragma unroll
	for (int br = 0; br < P; br++) {
		float acc[R];
#pragma unroll
		for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll
		for (int i = 0; i < SONDE_FIR_TAPS; i++) {
			const float c = sm.taps[br * SONDE_FIR_TAPS + i];
#pragma unroll
			for (int r = 0; r < R; r++) acc[r] = fadd(acc[r], fmul(w[r + i], c));   /* filter.c:59-61 */
		}
		float4 *dst = reinterpret_cast<float4 *>(&sm.y[br][g][seg * R]);
		dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
		dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
	}
}

template <int P, bool AFSK>
__global__ void __launch_bounds__(NT, 1)
demod_gfsk_kernel(const demod_params p, const int group_base)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	smem_t &sm = *reinterpret_cast<smem_t *>(smem_raw);

	const int tid = threadIdx.x;
	const int grp = group_base + blockIdx.x;
	const int type = p.group_type[grp];
	const sonde_modem &md = c_modem[type];
	const int *chans = p.group_chan + (size_t)grp * G;

	for (int i = tid; i < P * SONDE_FIR_TAPS; i += NT) sm.taps[i] = md.taps[i];

	/* channel owned by this thread in the serial stages (warp 0, lane g) */
	const bool serial_lane = tid < G;
	const int my_ch = serial_lane ? chans[tid] : -1;
	float bias = 0, avg = 0;
	tm_regs tr = {};
	uint64_t nbits0 = 0;
	float rf = 1.0f;
	long long n_rounds = 0, n_slow = 0;
	constexpr int NTM = (P == 2) ? 12 : 24;
	const float center = md.freq0, alpha = md.alpha, beta = md.beta, max_fdev = md.max_fdev;
	uint8_t *my_ring = nullptr;
	float *my_soft = nullptr;
	if (my_ch >= 0) {
		const demod_state &st = p.st[my_ch];
		bias = st.agc_bias; avg = st.agc_avg;
		tr.prev = st.t_prev; tr.phase = st.t_phase; tr.freq = st.t_freq; tr.target = (float)st.t_state;
		tr.interm = 0.0f;                                 /* gfsk.c:73 */
		tr.acc = st.bit_acc; nbits0 = st.nbits; tr.nb = (uint32_t)nbits0; tr.nsoft = 0;
		rf = rcp_approx(tr.freq);
		my_ring = p.ring + (size_t)my_ch * p.ring_bytes;
		if (p.soft) my_soft = p.soft + (size_t)my_ch * p.soft_stride;
		sm.ph[tid][0] = st.disc_prev;
	}
	/* AFSK: dynamic shared memory continues after smem_t */
	afsk_smem_t &am = *reinterpret_cast<afsk_smem_t *>(smem_raw + ((sizeof(smem_t) + 15) & ~(size_t)15));
	const int blen = AFSK ? md.boxcar_len : 1;
	float nco_p = 0.0f, box_sum = 0.0f;        /* lane state: NCO phase (warp 1, lane = 2g+m) / running sum (warp 2, lane = 4g+comp) */
	const int warp_id = tid >> 5, lane_id = tid & 31;
	if (AFSK) {
		if (warp_id == 1 && lane_id < 2 * G && chans[lane_id >> 1] >= 0) {
			const afsk_state &as = p.ast[chans[lane_id >> 1]];
			nco_p = (lane_id & 1) ? as.p_space : as.p_mark;
		}
		if (warp_id == 2 && chans[lane_id >> 2] >= 0) {
			const afsk_state &as = p.ast[chans[lane_id >> 2]];
			const int comp = lane_id & 3;
			box_sum = comp == 0 ? as.mark_re : comp == 1 ? as.mark_im : comp == 2 ? as.space_re : as.space_im;
		}
		/* the last `len` mixer outputs, oldest first */
		for (int i = tid; i < G * blen; i += NT) {
			const int g = i / blen, k = i % blen;
			const int ch = chans[g];
			am.om[g][k] = (ch >= 0) ? make_float2(p.ast[ch].mark_hist[2 * k], p.ast[ch].mark_hist[2 * k + 1]) : make_float2(0, 0);
			am.os[g][k] = (ch >= 0) ? make_float2(p.ast[ch].space_hist[2 * k], p.ast[ch].space_hist[2 * k + 1]) : make_float2(0, 0);
		}
	}
	/* FIR history */
	for (int i = tid; i < G * SONDE_FIR_HIST; i += NT) {
		const int g = i / SONDE_FIR_HIST, k = i % SONDE_FIR_HIST;
		const int ch = chans[g];
		sm.a[g][k] = (ch >= 0) ? p.st[ch].hist[k] : 0.0f;
	}
	__syncthreads();

	const uint32_t ring_mask = p.ring_bytes - 1;

	for (int base = 0; base < p.len; base += T) {
		const int n = min(T, p.len - base);

		/* ---- S1: load (+ discriminate) ------------------------------------------------ */
		int any_zero;
		if (p.is_iq) {
			const float2 *in = static_cast<const float2 *>(p.in);
#pragma unroll
			for (int k = 0; k < G * T / NT; k++) {
				const int idx = tid + k * NT;
				const int g = idx / T, i = idx % T;
				const int ch = chans[g];
				float phv = 0.0f;
				if (ch >= 0 && i < n) {
					const float2 q = __ldg(&in[(size_t)p.in_row[ch] * p.row_stride + base + i]);
					phv = det_phase(q.x, q.y);
				}
				sm.ph[g][i + 1] = phv;
			}
			__syncthreads();
			int zero = 0;
#pragma unroll
			for (int k = 0; k < G * T / NT; k++) {
				const int idx = tid + k * NT;
				const int g = idx / T, i = idx % T;
				const float xv = disc_step(sm.ph[g][i + 1], sm.ph[g][i], p.fm_gain);
				sm.x[g][i] = xv;
				zero |= (i < n && chans[g] >= 0 && xv == 0.0f);
			}
			any_zero = __syncthreads_or(zero);
			if (tid < G) sm.ph[tid][0] = sm.ph[tid][n];      /* carry the last phase */
		} else {
			const float *in = static_cast<const float *>(p.in);
			int zero = 0;
#pragma unroll
			for (int k = 0; k < G * T / NT; k++) {
				const int idx = tid + k * NT;
				const int g = idx / T, i = idx % T;
				const int ch = chans[g];
				float xv = 0.0f;
				if (ch >= 0 && i < n) xv = __ldg(&in[(size_t)p.in_row[ch] * p.row_stride + base + i]);
				sm.x[g][i] = xv;
				zero |= (i < n && ch >= 0 && xv == 0.0f);
			}
			any_zero = __syncthreads_or(zero);
		}

		if (AFSK) {
			/* ---- serial: AGC recurrences (warp 0) and the two NCOs per channel (warp 1) -------- */
			if (my_ch >= 0) {
				if (any_zero) agc_chains<true>(sm.x[tid], sm.s[tid], sm.v[tid], n, bias, avg);
				else          agc_chains<false>(sm.x[tid], sm.s[tid], sm.v[tid], n, bias, avg);
			}
			if (warp_id == 1 && lane_id < 2 * G) {
				const int g = lane_id >> 1, m = lane_id & 1;
				const float f = m ? md.f_space : md.f_mark;
				for (int i = 0; i < n; i++) {
					am.pn[m][g][i] = nco_p;
					nco_p = nco_step(nco_p, f);
				}
			}
			__syncthreads();
			/* ---- parallel: gain, /len*2, mixers --------------------------------------------- */
			const float flen = (float)blen;
#pragma unroll
			for (int k = 0; k < G * T / NT; k++) {
				const int idx = tid + k * NT;
				const int g = idx / T, i = idx % T;
				float2 m2 = make_float2(0.0f, 0.0f), s2 = m2;
				if (i < n && chans[g] >= 0) {
					float o = 0.0f;
					if (!(any_zero && sm.x[g][i] == 0.0f)) o = fmul(sm.s[g][i], fdiv(5.0f, sm.v[g][i]));
					const float sv = fmul(fdiv(o, flen), 2.0f);
					double sn, cs;
					sincos((double)am.pn[0][g][i], &sn, &cs);
					m2 = make_float2(fmul(sv, (float)cs), fmul(sv, -(float)sn));
					sincos((double)am.pn[1][g][i], &sn, &cs);
					s2 = make_float2(fmul(sv, (float)cs), fmul(sv, -(float)sn));
				}
				am.om[g][blen + i] = m2;
				am.os[g][blen + i] = s2;
			}
			__syncthreads();
			/* ---- serial: running sums, one lane per (channel, component) (warp 2) -------------- */
			if (warp_id == 2) {
				const int g = lane_id >> 2, comp = lane_id & 3;
				const float *o = reinterpret_cast<const float *>(comp < 2 ? &am.om[g][0] : &am.os[g][0]) + (comp & 1);
				float *dst = reinterpret_cast<float *>(comp < 2 ? &am.msum[g][0] : &am.ssum[g][0]) + (comp & 1);
				for (int i = 0; i < n; i++) {
					box_sum = fadd(box_sum, fsub(o[2 * (blen + i)], o[2 * i]));
					dst[2 * i] = box_sum;
				}
			}
			__syncthreads();
			/* ---- parallel: |mark| - |space| -> filter input; slide the mixer history ----------- */
#pragma unroll
			for (int k = 0; k < G * T / NT; k++) {
				const int idx = tid + k * NT;
				const int g = idx / T, i = idx % T;
				float o = 0.0f;
				if (i < n && chans[g] >= 0)
					o = fsub(cabs_exact(am.msum[g][i].x, am.msum[g][i].y), cabs_exact(am.ssum[g][i].x, am.ssum[g][i].y));
				sm.a[g][SONDE_FIR_HIST + i] = o;
			}
			{
				float2 hm[2], hs[2];
#pragma unroll
				for (int k = 0; k < 2; k++) {
					const int idx = tid + k * NT;
					if (idx < G * blen) { hm[k] = am.om[idx / blen][n + idx % blen]; hs[k] = am.os[idx / blen][n + idx % blen]; }
				}
				__syncthreads();
#pragma unroll
				for (int k = 0; k < 2; k++) {
					const int idx = tid + k * NT;
					if (idx < G * blen) { am.om[idx / blen][idx % blen] = hm[k]; am.os[idx / blen][idx % blen] = hs[k]; }
				}
			}
			__syncthreads();
		} else {
		/* ---- S2: AGC recurrences ------------------------------------------------------ */
		if (my_ch >= 0) {
			if (any_zero) agc_chains<true>(sm.x[tid], sm.s[tid], sm.v[tid], n, bias, avg);
			else          agc_chains<false>(sm.x[tid], sm.s[tid], sm.v[tid], n, bias, avg);
		}
		__syncthreads();

		/* ---- S3: gain apply: out = s * (5 / avg_before)  (agc.c:27,31) ----------------- */
#pragma unroll
		for (int k = 0; k < G * T / NT; k++) {
			const int idx = tid + k * NT;
			const int g = idx / T, i = idx % T;
			float o = 0.0f;
			if (i < n && chans[g] >= 0) o = fmul(sm.s[g][i], fdiv(5.0f, sm.v[g][i]));
			sm.a[g][SONDE_FIR_HIST + i] = o;
		}
		__syncthreads();
		}

		/* ---- S4: FIR everywhere -------------------------------------------------------- */
		fir_segment<P>(sm, tid / 32, tid % 32);
		__syncthreads();

		/* ---- S5: timing + slicer ------------------------------------------------------- */
		if (my_ch >= 0) {
			const float DELTA = 8.0f * ((float)(NTM + 16) * 1.2e-7f) / center;       /* see demod_pipe.cu */
			tm_tile<P, NTM, true, G, RS>(tr, rf, sm.y, tid, n * P, center, alpha, beta, max_fdev, DELTA, my_ring, ring_mask,
			                             my_soft, p.soft_stride, n_rounds, n_slow, false);
		}

		/* ---- slide the FIR history: a[g][0..48) <- a[g][n..n+48) ----------------------- */
		float hv[2];
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const int idx = tid + k * NT;
			hv[k] = (idx < G * SONDE_FIR_HIST) ? sm.a[idx / SONDE_FIR_HIST][n + idx % SONDE_FIR_HIST] : 0.0f;
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const int idx = tid + k * NT;
			if (idx < G * SONDE_FIR_HIST) sm.a[idx / SONDE_FIR_HIST][idx % SONDE_FIR_HIST] = hv[k];
		}
		__syncthreads();
	}

	/* ---- save state ---------------------------------------------------------------------- */
	if (my_ch >= 0) {
		demod_state &st = p.st[my_ch];
		st.agc_bias = bias; st.agc_avg = avg;
		const uint64_t nbits = nbits0 + (uint64_t)(tr.nb - (uint32_t)nbits0);
		const int cnt = (int)(tr.nb & 7u);
		st.t_prev = tr.prev; st.t_phase = tr.phase; st.t_freq = tr.freq; st.t_state = (int)tr.target;
		st.bit_acc = tr.acc & ((1u << cnt) - 1u); st.bit_cnt = cnt; st.nbits = nbits; st.nsoft = tr.nsoft;
		p.nbits_out[my_ch] = nbits;
		if (p.is_iq) st.disc_prev = sm.ph[tid][0];
		if (cnt)                                          /* left-aligned partial byte, gfsk.c:78 */
			my_ring[(uint32_t)(nbits >> 3) & ring_mask] = (uint8_t)(tr.acc << (8 - cnt));
	}
	for (int i = tid; i < G * SONDE_FIR_HIST; i += NT) {
		const int g = i / SONDE_FIR_HIST, k = i % SONDE_FIR_HIST;
		const int ch = chans[g];
		if (ch >= 0) p.st[ch].hist[k] = sm.a[g][k];
	}
	if (AFSK) {
		if (warp_id == 1 && lane_id < 2 * G && chans[lane_id >> 1] >= 0) {
			afsk_state &as = p.ast[chans[lane_id >> 1]];
			if (lane_id & 1) as.p_space = nco_p; else as.p_mark = nco_p;
		}
		if (warp_id == 2 && chans[lane_id >> 2] >= 0) {
			afsk_state &as = p.ast[chans[lane_id >> 2]];
			const int comp = lane_id & 3;
			(comp == 0 ? as.mark_re : comp == 1 ? as.mark_im : comp == 2 ? as.space_re : as.space_im) = box_sum;
		}
		for (int i = tid; i < G * blen; i += NT) {
			const int g = i / blen, k = i % blen;
			const int ch = chans[g];
			if (ch >= 0) {
				p.ast[ch].mark_hist[2 * k] = am.om[g][k].x; p.ast[ch].mark_hist[2 * k + 1] = am.om[g][k].y;
				p.ast[ch].space_hist[2 * k] = am.os[g][k].x; p.ast[ch].space_hist[2 * k + 1] = am.os[g][k].y;
			}
		}
	}
}

}  // namespace

extern "C" size_t sonde_demod_smem_bytes(void) { return sizeof(smem_t); }

/* Launches the phase-by-phase demodulator over groups [group_base, group_base + n_groups) that
 * share the polyphase count `phases`. */
template <bool AFSK>
static constexpr size_t legacy_smem_bytes()
{
	return ((sizeof(smem_t) + 15) & ~(size_t)15) + (AFSK ? sizeof(afsk_smem_t) : 0);
}

template <int P, bool AFSK>
static cudaError_t launch_legacy(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	static std::atomic<unsigned long long> attr_done{0};
	const cudaError_t ea = sonde_ensure_dynamic_smem(demod_gfsk_kernel<P, AFSK>, (int)legacy_smem_bytes<AFSK>(), attr_done);
	if (ea != cudaSuccess) return ea;
	demod_gfsk_kernel<P, AFSK><<<n_groups, NT, legacy_smem_bytes<AFSK>(), stream>>>(*p, group_base);
	return cudaGetLastError();
}

extern "C" cudaError_t sonde_launch_demod_gfsk(const demod_params *p, int group_base, int n_groups,
                                               int phases, cudaStream_t stream)
{
	if (n_groups <= 0) return cudaSuccess;
	return phases == 1 ? launch_legacy<1, false>(p, group_base, n_groups, stream)
	                   : launch_legacy<2, false>(p, group_base, n_groups, stream);
}

/* AFSK sondes (iMet-1/4, SRS-C50): 1 polyphase branch at 48 kS/s. */
extern "C" cudaError_t sonde_launch_demod_afsk(const demod_params *p, int group_base, int n_groups, cudaStream_t stream)
{
	if (n_groups <= 0) return cudaSuccess;
	return launch_legacy<1, true>(p, group_base, n_groups, stream);
}
