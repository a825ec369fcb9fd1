/*
 * timing_exact.cuh — the Gardner timing lane of K1 (demod_pipe.cu): one lane = one channel, the lane walks the
 * FIR outputs of its channel (a ring in shared memory, indexed by NCO slot) from symbol to symbol.
 * Reference: SD/demod/dsp/timing.c:28-76 (advance_timeslot, retime, gardner_err, update_estimate),
 * SD/demod/gfsk.c:87-115 (the slot loop and the slicer).
 *
 * Fast round ("chain compare").  The reference adds `freq` to `phase` once per slot and compares after every add.
 * With freq within +-1/256 of its centre (timing.c:18,73-75) and the phase left after a symbol in [-0.15, 0.35],
 * the mid-symbol hit of the next symbol can only fall on NM known slots (KM0 .. KM0+NM-1) and the symbol hit on NS
 * known slots (KS0 .. KS0+NS-1).  The round therefore runs the reference's chain of adds unconditionally for
 * KS0+NS-1 slots (same operations, same order: p[i] are the reference's phase values), takes the hit slots from the
 * reference's own comparisons on those values (p[k] >= 1, p[k] >= 2) and only has to check that the hits are inside
 * the windows (p[KM0-1] < 1 <= p[KM0+NM-1], p[KS0-1] < 2 <= p[KS0+NS-1]).  Nothing is predicted: when the check
 * holds the round IS the reference's computation; when it does not (acquisition, a noise burst, the first and last
 * slots of a buffer) the lane runs the literal slot-by-slot loop for that round.  The FIR outputs of the candidate
 * slots are loaded at fixed offsets from the round's first slot, before the chain resolves.
 */
#ifndef SONDE_TIMING_EXACT_CUH
#define SONDE_TIMING_EXACT_CUH

#include <stdint.h>
#include "strict_math.cuh"

struct tmx_regs {
	float prev, phase, freq, interm;
	float target;            /* (float)state : 1 = waiting for the mid-symbol hit, 2 = for the symbol hit */
	uint32_t acc;            /* bits of the 32-bit word being assembled (low bits = newest)               */
	uint32_t nb;             /* bits demodulated so far, low 32 bits of the stream position               */
	int nsoft;
};

struct tmx_consts {
	float center, alpha, beta, max_fdev;
};

/* MIN()/MAX() of SD/utils.h:46-47 are `(a) < (b) ? (a) : (b)` / `(a) > (b) ? (a) : (b)` with the constant first:
 * a NaN second operand comes out as NaN.  min.NaN / max.NaN give exactly that in one instruction. */
__device__ __forceinline__ float min_nan(float a, float b)
{
	float r;
	asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ float max_nan(float a, float b)
{
	float r;
	asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
	return r;
}

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
	return v;
}

/* retime() + update_estimate() at a symbol hit whose phase is `ph` (timing.c:45-76) */
__device__ __forceinline__ void tmx_retime(tmx_regs &t, const float ph, const float yv, const tmx_consts &c)
{
	const float err = (fmul(yv, t.prev) < 0.0f) ? fmul(fsub(yv, t.prev), t.interm) : 0.0f;
	t.prev = yv;
	float fd = fsub(t.freq, c.center);
	t.phase = fsub(ph, fsub(2.0f, max_nan(-2.0f, min_nan(2.0f, fmul(err, c.alpha)))));
	fd = fadd(fd, fmul(err, c.beta));
	fd = max_nan(-c.max_fdev, min_nan(c.max_fdev, fd));
	t.freq = fadd(c.center, fd);
	t.target = 1.0f;
}

/* slicer + bit packing (gfsk.c:107-115): bits are collected MSB-first into 32-bit words and stored big-endian, so
 * the ring holds the same byte stream the reference writes.  The ring is a power of two of at least 4 bytes. */
template <bool SOFT>
__device__ __forceinline__ void tmx_emit(tmx_regs &t, const float yv, uint8_t *ring, const uint32_t ring_mask,
                                         float *soft, const int soft_cap)
{
	t.acc = (t.acc << 1) | (yv > 0.0f ? 1u : 0u);
	if (SOFT) {
		if (soft && t.nsoft < soft_cap) soft[t.nsoft] = yv;
	}
	t.nsoft++;
	t.nb++;
	/* rare (one round in 32): keep it off the fall-through path, a taken branch costs the lane ~10 cycles */
	if (__builtin_expect((t.nb & 31u) == 0u, 0))
		*reinterpret_cast<uint32_t *>(ring + (((t.nb - 32u) >> 3) & ring_mask)) = __byte_perm(t.acc, 0, 0x0123);
}

/* The literal loop of gfsk.c:87-115 for at most `maxslots` slots, stopping after a symbol.  y[i] is the FIR output
 * of slot i counted from the lane's current slot.  Returns the slots consumed. */
template <bool SOFT>
__device__ __forceinline__ int tmx_literal_round(tmx_regs &t, const float *y, const int maxslots, const tmx_consts &c,
                                                 uint8_t *ring, const uint32_t ring_mask, float *soft, const int soft_cap)
{
	float ph = t.phase;
	for (int i = 0; i < maxslots; i++) {
		ph = fadd(ph, t.freq);
		if (ph >= t.target) {
			if (t.target == 1.0f) {
				t.interm = y[i];
				t.target = 2.0f;
			} else {
				const float yv = y[i];
				tmx_retime(t, ph, yv, c);
				tmx_emit<SOFT>(t, yv, ring, ring_mask, soft, soft_cap);
				return i + 1;
			}
		}
	}
	t.phase = ph;
	return maxslots;
}

/* retime() + update_estimate() of a fast round: as tmx_retime(), with the mid-symbol value passed in and without the
 * state writes a completed fast round makes redundant (target stays 1; interm is dead after a symbol, gfsk.c:73,93,99) */
__device__ __forceinline__ void tmx_retime_fast(tmx_regs &t, const float ph, const float yv, const float interm, const tmx_consts &c)
{
	const float err = (fmul(yv, t.prev) < 0.0f) ? fmul(fsub(yv, t.prev), interm) : 0.0f;
	t.prev = yv;
	float fd = fsub(t.freq, c.center);
	t.phase = fsub(ph, fsub(2.0f, max_nan(-2.0f, min_nan(2.0f, fmul(err, c.alpha)))));
	fd = fadd(fd, fmul(err, c.beta));
	fd = max_nan(-c.max_fdev, min_nan(c.max_fdev, fd));
	t.freq = fadd(c.center, fd);
}

/* All rounds of one lane that start at slot `sabs` (ring position `sring`) and whose candidate slots end before slot
 * `end`: W = KS0 + NS - 1 slots must be readable from the start of each round (the ring has a mirror of its first
 * slots behind its last one).
 *
 * Shape of the loop (ncu on a first version that handled the two rare events inside the round showed a quarter of
 * every round in BSSY/BSYNC reconvergence, branch resolution and predicate latency): the inner loop is straight-line
 * code with one backward branch; a lane that meets a rare event (the round does not fit the windows -> literal round;
 * a 32-bit word of bits is complete -> store it) leaves the inner loop, handles it and re-enters.
 *
 * Round 2, second pass (ncu: 71 SASS instructions per round, the half after the add chain issue-bound; and lanes =
 * channels whose 32-bit words of bits complete at different rounds left the loop one by one, after which the warp ran
 * the loop body once per straggler).  Now: the inner loop holds TWO rounds with the candidate registers alternating
 * between two sets (the six loop-carried MOVs of the one-round form disappear); the bookkeeping a completed fast round
 * implies is left out of it (target is 1 on entry and stays 1, the bit and soft-symbol counts advance by the number of
 * rounds); new bits are funnel-shifted into a 64-bit accumulator and a stretch ends after at most 32 rounds — a count
 * that is the same in every lane — so that the completed word is stored AFTER the loop by all lanes together. */
template <int KM0, int NM, int KS0, int NS, int RING, bool SOFT>
__device__ __forceinline__ void tmx_run(tmx_regs &t, const float *yrow, int &sabs, int &sring, const int end,
                                        const tmx_consts &c, uint8_t *ring, const uint32_t ring_mask, float *soft,
                                        const int soft_cap, long long &n_slow)
{
	constexpr int W = KS0 + NS - 1;
	static_assert(KM0 >= 1 && KM0 + NM - 1 < KS0, "the mid-symbol window must end before the symbol window starts");
	int left = end - W - sabs;                   /* rounds may start while left >= 0 */
	/* 32-bit shared-memory address of the row, taken once: left to itself the compiler rebuilt the shared window base
	 * (S2R SR_CgaCtaId ...) in every round */
	const uint32_t ybase = (uint32_t)__cvta_generic_to_shared(yrow);
	/* The candidate FIR outputs of a round are fetched at the END of the round before it (as soon as its first slot is
	 * known) and carried in registers: the loads are in flight during the loop branch, and ptxas cannot turn them into
	 * loads predicated on this round's comparisons, which would put the shared-memory latency behind the add chain. */
	float ymA[NM], ysA[NS], ymB[NM], ysB[NS];
	auto fetch = [&](float (&ym)[NM], float (&ys)[NS], const int at) {
		const uint32_t ya = ybase + 4u * (uint32_t)at;
#pragma unroll
		for (int j = 0; j < NM; j++) ym[j] = lds_f32(ya + 4u * (KM0 - 1 + j));   /* slot k (1-based) reads y[k - 1] */
#pragma unroll
		for (int j = 0; j < NS; j++) ys[j] = lds_f32(ya + 4u * (KS0 - 1 + j));
	};
	int room = 0;                                    /* fast rounds this stretch may still run (32 at its start) */
	uint32_t acc_hi = 0;                             /* bits shifted out of t.acc during the stretch */
	/* one fast round on the candidates (ymc, ysc); prefetches the next round's into (ymn, ysn).
	 * 1: the round does not fit the windows, nothing consumed; 2: consumed, but no further round may start here; 0: go on */
	auto round = [&](const float (&ymc)[NM], const float (&ysc)[NS], float (&ymn)[NM], float (&ysn)[NS]) -> int {
		/* the reference's chain of adds */
		float p[W + 1];
		p[0] = t.phase;
#pragma unroll
		for (int i = 1; i <= W; i++) p[i] = fadd(p[i - 1], t.freq);
		const bool ok = p[KM0 - 1] < 1.0f && p[KM0 + NM - 1] >= 1.0f && p[KS0 - 1] < 2.0f && p[W] >= 2.0f;
		/* earliest slot at or above the threshold wins (timing.c:35) */
		float ym = ymc[NM - 1];
#pragma unroll
		for (int j = NM - 2; j >= 0; j--) ym = (p[KM0 + j] >= 1.0f) ? ymc[j] : ym;
		float ys = ysc[NS - 1], pl = p[W];
		int ks = W;
#pragma unroll
		for (int j = NS - 2; j >= 0; j--) {
			const bool hit = p[KS0 + j] >= 2.0f;
			ys = hit ? ysc[j] : ys;
			pl = hit ? p[KS0 + j] : pl;
			ks = hit ? KS0 + j : ks;
		}
		if (__builtin_expect(!ok, 0)) return 1;
		left -= ks;
		sring += ks;
		sring = (int)min((unsigned)sring, (unsigned)(sring - RING));      /* wrap without a predicate */
		room--;
		const bool again = (left | (room - 1)) >= 0;                      /* left >= 0 and room >= 1 */
		fetch(ymn, ysn, sring);                          /* next round's candidates (harmless if there is no next round) */
		tmx_retime_fast(t, pl, ys, ym, c);
		acc_hi = __funnelshift_l(t.acc, acc_hi, 1);                       /* (hi:lo) = (hi:lo) << 1 | bit */
		t.acc = __funnelshift_l(ys > 0.0f ? 0x80000000u : 0u, t.acc, 1);
		if (SOFT) {
			if (soft && t.nsoft < soft_cap) soft[t.nsoft] = ys;
			t.nsoft++;
		}
		return again ? 0 : 2;
	};
	if (left >= 0) {
		for (;;) {
			int ev = 1;                              /* 1: literal round needed; else: stretch over */
			if (t.target == 1.0f) {                  /* a fast round starts at a symbol boundary */
				room = 32;
				fetch(ymA, ysA, sring);
				for (;;) {
					int r = round(ymA, ysA, ymB, ysB);
					if (__builtin_expect(r != 0, 0)) { ev = r; break; }
					r = round(ymB, ysB, ymA, ysA);
					if (__builtin_expect(r != 0, 0)) { ev = r; break; }
				}
				/* one bit per completed round; with at most 31 bits pending and 32 new ones exactly one word can be complete */
				const uint32_t pend = (t.nb & 31u) + (uint32_t)(32 - room);
				t.nb += (uint32_t)(32 - room);
				if (pend >= 32u) {
					const uint32_t word = __funnelshift_r(t.acc, acc_hi, pend - 32u);
					*reinterpret_cast<uint32_t *>(ring + ((((t.nb - pend) >> 3)) & ring_mask)) = __byte_perm(word, 0, 0x0123);
				}
			}
			if (ev == 1) {
				n_slow++;
				const int used = tmx_literal_round<SOFT>(t, yrow + sring, W, c, ring, ring_mask, soft, soft_cap);
				left -= used;
				sring += used;
				sring = (sring >= RING) ? sring - RING : sring;
			}
			if (left < 0) break;
		}
	}
	sabs = end - W - left;
}

#endif
