/*
 * timing_round.cuh — the Gardner timing lane shared by both demodulator kernels:
 * predict-and-verify rounds over the precomputed FIR outputs of one tile (see the comment at tm_tile).
 * Reference: SD/demod/dsp/timing.c:28-76, SD/demod/gfsk.c:75-125.
 */
#ifndef SONDE_TIMING_ROUND_CUH
#define SONDE_TIMING_ROUND_CUH

#include <stdint.h>
#include "strict_math.cuh"

/* ---- TM helpers ----------------------------------------------------------------------------- */
struct tm_regs {
	float prev, phase, freq, interm, target;     /* target = (float)state : 1 = mid-symbol, 2 = symbol */
	uint32_t acc;                                /* bits of the byte being assembled                   */
	uint32_t nb;                                 /* bits demodulated so far, low 32 bits of the stream position */
	int nsoft;
};

__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

/* retime() at a symbol hit (timing.c:45-76) */
__device__ __forceinline__ void retime(tm_regs &t, const float yv, const float center, const float alpha,
                                       const float beta, const float max_fdev)
{
	const float err = (fmul(yv, t.prev) < 0.0f) ? fmul(fsub(yv, t.prev), t.interm) : 0.0f;
	t.prev = yv;
	float fd = fsub(t.freq, center);
	const float ea = fmul(err, alpha);
	const float lo = (2.0f < ea) ? 2.0f : ea;
	const float cl = (-2.0f > lo) ? -2.0f : lo;
	t.phase = fsub(t.phase, fsub(2.0f, cl));
	fd = fadd(fd, fmul(err, beta));
	const float fl = (max_fdev < fd) ? max_fdev : fd;
	fd = (-max_fdev > fl) ? -max_fdev : fl;
	t.freq = fadd(center, fd);
	t.target = 1.0f;
}

/* slicer + bit packing (gfsk.c:107-115).  `ring` points at the channel's ring; byte index =
 * (stream bit position >> 3) & mask, the position's low 32 bits are t.nb. */
template <bool SOFT>
__device__ __forceinline__ void emit_symbol(tm_regs &t, const float yv, uint8_t *ring, const uint32_t ring_mask,
                                            float *soft, const int soft_cap)
{
	t.acc = (t.acc << 1) | (yv > 0.0f ? 1u : 0u);
	if (SOFT) {
		if (soft && t.nsoft < soft_cap) soft[t.nsoft] = yv;
	}
	t.nsoft++;
	if ((t.nb & 7u) == 7u) ring[(t.nb >> 3) & ring_mask] = (uint8_t)t.acc;
	t.nb++;
}

/* filter_get(phase) of slot index sl: sample sl / P, polyphase branch P-1-(sl % P) (filter.c:54) */
template <int P, int GG, int RRS>
__device__ __forceinline__ float y_at(const float (*y)[GG][RRS], const int g, const int sl)
{
	return (P == 1) ? y[0][g][sl] : y[P - 1 - (sl % P)][g][sl / P];
}


/* One tile of the timing lane for one channel (lane): ns NCO slots over the FIR outputs y[branch][g][sample].
 *
 * Per "round" the lane advances its NCO by up to N slots, i.e. to its next symbol instant (timing.c:28-43),
 * then retimes and slices (timing.c:45-76, gfsk.c:99-115).  The reference finds the hit slots by comparing
 * after every add.  Here the slot numbers are PREDICTED arithmetically (c = ceil((threshold - phase) / freq)),
 * the reference's chain of adds is then run for exactly that many slots as predicated FADDs (same
 * operations, same order, so the phase value is the reference's), and the symbol prediction is VERIFIED on
 * the chain values (p[c-1] < 2 <= p[c]; the chain is monotone because freq > 0), the mid-symbol prediction
 * by an error-bound margin (DELTA).  If a check fails the round is replayed with the literal slot-by-slot
 * loop, so the result is exact in every case. */
template <int P, int N, bool SOFT, int GG, int RRS, int KMIN = 0>
__device__ __forceinline__ void tm_tile(tm_regs &tr, float &rf, const float (*y)[GG][RRS], const int g, const int ns,
                                        const float center, const float alpha, const float beta, const float max_fdev,
                                        const float DELTA, uint8_t *ring, const uint32_t ring_mask, float *soft,
                                        const int soft_cap, long long &n_rounds, long long &n_slow, const bool prof_on)
{
	int s = 0;
	/* Lanes run their rounds independently (no warp votes on the critical path); the few lanes that
	 * need the replay or finish the tile earlier simply diverge and reconverge. */
	while (s < ns) {
		/* ---- steady-state round (KMIN > 0 variants only) -----------------------------------------------------
		 * The common round — it starts waiting for a mid-symbol hit, lies entirely inside the tile and ends with a
		 * symbol after KMIN..N slots — needs none of the general round's bookkeeping (partial rounds, "no hit in this
		 * round", mid-symbol already pending) and its first KMIN-1 adds need no predicate.  Same arithmetic, same
		 * verification; whatever does not fit falls through, state untouched, to the general round below.  Measured:
		 * M10/M20 (51 rounds per tile) 82 -> 66 cycles/sample; RS41 and the 2400-baud sondes gain nothing (fewer rounds
		 * per tile, so the lanes of the warp are more often in different kinds of round), see profiles/README.md. */
		if (KMIN > 0 && s + N <= ns && tr.target == 1.0f) {
			const float p0 = tr.phase, f = tr.freq;
			const float x1 = fmul(fsub(1.0f, p0), rf), x2 = fmul(fsub(2.0f, p0), rf);
			const float x1c = ceilf(x1);
			const float m1 = x1c - x1;
			const int c1 = (int)x1c;
			const int c2 = __float2int_ru(x2);
			float pa = p0;
#pragma unroll
			for (int i = 1; i < KMIN; i++) pa = fadd(pa, f);
#pragma unroll
			for (int i = KMIN; i < N; i++)
				asm("{\n.reg .pred q;\nsetp.lt.s32 q, %2, %3;\n@q add.rn.f32 %0, %0, %1;\n}"
				    : "+f"(pa) : "f"(f), "r"(i), "r"(c2));
			const float pl = fadd(pa, f);
			/* indices are clamped only to keep the speculative loads inside the tile; a clamped index fails `ok` */
			const int i1 = min(max(c1, 1), N), i2 = min(max(c2, 1), N);
			const float y_mid = y_at<P, GG, RRS>(y, g, s + i1 - 1);
			const float y_sym = y_at<P, GG, RRS>(y, g, s + i2 - 1);
			const float err = (fmul(y_sym, tr.prev) < 0.0f) ? fmul(fsub(y_sym, tr.prev), y_mid) : 0.0f;
			const float ea = fmul(err, alpha);
			const float lo = (2.0f < ea) ? 2.0f : ea;
			const float cl = (-2.0f > lo) ? -2.0f : lo;
			float fd = fadd(fsub(f, center), fmul(err, beta));
			const float fl = (max_fdev < fd) ? max_fdev : fd;
			fd = (-max_fdev > fl) ? -max_fdev : fl;
			const float f_new = fadd(center, fd);
			const float ph_new = fsub(pl, fsub(2.0f, cl));
			/* mid-symbol slot by the error-bound margin, symbol slot on the chain values (see below);
			 * c2 - 1 > c1 keeps the slot before the symbol clear of the mid-symbol hit */
			const bool ok = x1 > 0.0f && m1 > DELTA && m1 < 1.0f - DELTA && c2 >= KMIN && c2 <= N && c2 - 1 > c1 &&
			                pa < 2.0f && pl >= 2.0f;
			if (ok) {
				tr.interm = y_mid;
				tr.phase = ph_new; tr.freq = f_new; tr.prev = y_sym;
				rf = rcp_approx(f_new);
				emit_symbol<SOFT>(tr, y_sym, ring, ring_mask, soft, soft_cap);
				s += c2;
				if (prof_on) n_rounds++;
				continue;
			}
		}
		const int lim = min(N, ns - s);                 /* slots this round may consume */
		const float p0 = tr.phase, f = tr.freq;
		const bool want_mid = tr.target == 1.0f;
		/* predicted slots (1-based) of the mid-symbol and symbol hits */
		const float x1 = fmul(fsub(1.0f, p0), rf), x2 = fmul(fsub(2.0f, p0), rf);
		const float x1c = ceilf(x1);
		const float m1 = x1c - x1;
		const bool mid_ok = !want_mid || x1 <= 0.0f || (m1 > DELTA && m1 < 1.0f - DELTA);
		const int c1 = want_mid ? max(1, (int)x1c) : 0;
		const int c2 = max(c1 + 1, __float2int_ru(x2));
		const bool mid_in = want_mid && c1 <= lim;
		const bool sym_in = c2 <= lim;
		const int K = sym_in ? c2 : lim;                 /* slots consumed */
		const float y_mid = mid_in ? y_at<P, GG, RRS>(y, g, s + c1 - 1) : tr.interm;
		const float y_sym = sym_in ? y_at<P, GG, RRS>(y, g, s + c2 - 1) : 0.0f;
		/* the reference's adds for slots 1 .. K-1, then slot K */
		float pa = p0;
#pragma unroll
		for (int i = 1; i < N; i++)
			asm("{\n.reg .pred q;\nsetp.lt.s32 q, %2, %3;\n@q add.rn.f32 %0, %0, %1;\n}"
			    : "+f"(pa) : "f"(f), "r"(i), "r"(K));
		const float pl = fadd(pa, f);                    /* lim >= 1 here, so K >= 1 */
		/* speculative retime on the predicted symbol (timing.c:45-76); independent of the add chain
		 * except for the final phase correction */
		const float err = (fmul(y_sym, tr.prev) < 0.0f) ? fmul(fsub(y_sym, tr.prev), y_mid) : 0.0f;
		const float ea = fmul(err, alpha);
		const float lo = (2.0f < ea) ? 2.0f : ea;
		const float cl = (-2.0f > lo) ? -2.0f : lo;
		float fd = fadd(fsub(f, center), fmul(err, beta));
		const float fl = (max_fdev < fd) ? max_fdev : fd;
		fd = (-max_fdev > fl) ? -max_fdev : fl;
		const float f_new = fadd(center, fd);
		const float ph_new = fsub(pl, fsub(2.0f, cl));
		/* verify the symbol prediction on the chain values (monotone chain: freq > 0) */
		const bool hit_ok = (c2 - 1 == c1 || pa < 2.0f) && pl >= 2.0f;
		const bool none_ok = (want_mid && !mid_in) ? (pl < 1.0f)              /* not even the mid-symbol hit */
		                                           : (lim <= c1 || pl < 2.0f);
		const bool ok = mid_ok && (sym_in ? hit_ok : none_ok);
		if (ok) {
			if (mid_in) { tr.interm = y_mid; tr.target = 2.0f; }
			tr.phase = pl;
			if (sym_in) {
				tr.phase = ph_new; tr.freq = f_new; tr.prev = y_sym; tr.target = 1.0f;
				rf = rcp_approx(f_new);
				emit_symbol<SOFT>(tr, y_sym, ring, ring_mask, soft, soft_cap);
			}
			s += K;
		} else {
			/* exact slot-by-slot replay of this round (timing.c:28-43) */
			if (prof_on) n_slow++;
			float ph = p0;
			int used = lim;
			bool sym = false;
			for (int i = 1; i <= lim; i++) {
				ph = fadd(ph, f);
				if (ph >= tr.target) {
					if (tr.target == 1.0f) {
						tr.interm = y_at<P, GG, RRS>(y, g, s + i - 1);
						tr.target = 2.0f;
					} else {
						used = i;
						sym = true;
						break;
					}
				}
			}
			tr.phase = ph;
			if (sym) {
				const float yv = y_at<P, GG, RRS>(y, g, s + used - 1);
				retime(tr, yv, center, alpha, beta, max_fdev);
				rf = rcp_approx(tr.freq);
				emit_symbol<SOFT>(tr, yv, ring, ring_mask, soft, soft_cap);
			}
			s += used;
		}
		if (prof_on) n_rounds++;
	}
}

#endif
