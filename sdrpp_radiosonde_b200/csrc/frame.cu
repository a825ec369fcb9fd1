/*
 * frame.cu — K2: demodulated bit stream -> aligned frames -> descramble -> FEC / checksum.
 *
 * One warp per channel.  The warp replays the reference's framer over the channel's bit
 * ring as a window scan: for each window it
 *   1. stages the F+S bits the framer would hold after its READ step  (decode/framer.c:70-77)
 *   2. searches the sync word: XOR + popcount at every bit offset, lanes over offsets,
 *      warp-shuffle arg-min with earliest-offset tie-break, both polarities
 *                                                             (decode/correlator/correlator.c:19-89)
 *   3. waits until the REALIGN step's F+sync_offset bits exist (framer.c:84-90), extracts
 *      the frame and undoes the inversion                      (framer.c:92-106)
 *   4. runs the per-sonde post-framer pipeline up to (not including) telemetry parsing:
 *        RS41    rs41/frame.c:21-78   + decode/ecc/rs.c:99-212 (RS(255,231) x2)
 *        DFM     manchester.c:6-26, dfm09/frame.c:10-109
 *        M10     manchester, m10/frame.c:7-54
 *        iMS-100 manchester, ims100/frame.c:10-113 + rs.c (BCH(63,51) x12)
 *        MRZ-N1  manchester, mrz-n1/frame.c:6-21, decode/ecc/crc.c:26
 *        iMet-4  imet4/frame.c:5-24, subframe.c:8-30, imet4.c:91-117 (incl. framer_adjust)
 *        C50     c50/frame.c:9-39
 *   5. writes one sonde_frame_rec and advances the framer state (framer.c:57-68,114-137).
 * Frames per channel and chunk are few (RS41: 1.16/s), so the per-channel sequential walk
 * costs microseconds; all the data-parallel work inside a window is spread over the lanes.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "device_state.h"

/* frame-side copy of the modem table (frame_bits, sync_len, syncword, data_len are used here) */
static __constant__ sonde_modem c_modem[SONDE_NTYPES_];

extern "C" cudaError_t sonde_upload_modems_frame(const sonde_modem *m)
{
	return cudaMemcpyToSymbol(c_modem, m, sizeof(sonde_modem) * SONDE_NTYPES_);
}

namespace {

constexpr int WARPS_PER_CTA = 2;    /* 64 threads x 128 registers: small enough to sit beside a K1 CTA (768 x 72) on the same SM */
constexpr int MAX_SKIP_WARPS = 0;   /* idle warps in front of the working ones (SMSP placement experiments: compile with 3) */
constexpr int WIN_WORDS = 264;            /* >= 2 * 4144 / 32 + 4 : the framer buffer holds up to 2F bits */
constexpr int WORK_BYTES = 1024;
constexpr unsigned FULL = 0xffffffffu;

struct gf_tables {
	uint8_t exp256[256], log256[256];   /* GF(256)/0x11D : alpha[], logtable[] of rs_init_internal (rs.c:66-89) */
	uint8_t exp64[64], log64[64];       /* GF(64)/0x61                                                         */
};

struct warp_smem {
	uint32_t win[WIN_WORDS];            /* framer buffer bits, MSB first inside each 32-bit word */
	uint32_t carry[FRAMER_CARRY_WORDS];
	uint8_t  raw[SONDE_REC_BYTES + 8];  /* aligned, de-inverted frame                            */
	uint8_t  work[WORK_BYTES];          /* descrambled / decoded frame (zero beyond what is written) */
	uint8_t  blk[256];                  /* RS block                                              */
	uint8_t  syn[32], lam[16], omg[32], lpr[16];
	uint8_t  root[16], pos[16];
};

struct cta_smem {
	gf_tables gf;
	warp_smem w[WARPS_PER_CTA];
};

__constant__ uint8_t c_rs41_prn[64] = {       /* rs41/frame.c:9-18 */
	0x96, 0x83, 0x3e, 0x51, 0xb1, 0x49, 0x08, 0x98, 0x32, 0x05, 0x59, 0x0e, 0xf9, 0x44, 0xc6, 0x26,
	0x21, 0x60, 0xc2, 0xea, 0x79, 0x5d, 0x6d, 0xa1, 0x54, 0x69, 0x47, 0x0c, 0xdc, 0xe8, 0x5c, 0xf1,
	0xf7, 0x76, 0x82, 0x7f, 0x07, 0x99, 0xa2, 0x2c, 0x93, 0x7c, 0x30, 0x63, 0xf5, 0x10, 0x2e, 0x61,
	0xd0, 0xbc, 0xb4, 0xb6, 0x06, 0xaa, 0xf4, 0x23, 0x78, 0x6e, 0x3b, 0xae, 0xbf, 0x7b, 0x4c, 0xc1,
};

/* ---- bit helpers ---------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t ring_bit(const uint8_t *ring, uint32_t mask, uint64_t pos)
{
	return (ring[(uint32_t)(pos >> 3) & mask] >> (7 - (uint32_t)(pos & 7))) & 1u;
}

/* 32 consecutive stream bits starting at `pos` (MSB = first bit) */
__device__ __forceinline__ uint32_t ring_word(const uint8_t *ring, uint32_t mask, uint64_t pos)
{
	const uint32_t b = (uint32_t)(pos >> 3), sh = (uint32_t)(pos & 7);
	const uint64_t v = ((uint64_t)ring[b & mask] << 32) | ((uint64_t)ring[(b + 1) & mask] << 24) |
	                   ((uint64_t)ring[(b + 2) & mask] << 16) | ((uint64_t)ring[(b + 3) & mask] << 8) |
	                   (uint64_t)ring[(b + 4) & mask];
	return (uint32_t)(v >> (8 - sh));
}

__device__ __forceinline__ uint32_t win_bit(const uint32_t *win, int i)
{
	return (win[i >> 5] >> (31 - (i & 31))) & 1u;
}

/* up to 64 bits starting at bit i of the window (left aligned in the result) */
__device__ __forceinline__ uint64_t win_bits64(const uint32_t *win, int i)
{
	const int w = i >> 5, sh = i & 31;
	const uint64_t hi = ((uint64_t)win[w] << 32) | win[w + 1];
	const uint64_t lo = (uint64_t)win[w + 2] << 32;
	return sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
}

__device__ __forceinline__ uint32_t buf_bit(const uint8_t *b, int i) { return (b[i >> 3] >> (7 - (i & 7))) & 1u; }

__device__ __forceinline__ void buf_set(uint8_t *b, int i, uint32_t v)
{
	const uint8_t m = (uint8_t)(0x80u >> (i & 7));
	b[i >> 3] = v ? (b[i >> 3] | m) : (b[i >> 3] & ~m);
}

/* 8 bits starting at bit i of a byte buffer */
__device__ __forceinline__ uint32_t buf_byte_at(const uint8_t *b, int i)
{
	const int k = i >> 3, sh = i & 7;
	return (((uint32_t)b[k] << 8 | b[k + 1]) >> (8 - sh)) & 0xffu;
}

__device__ __forceinline__ uint32_t bitrev8(uint32_t x) { return __brev(x) >> 24; }

/* ---- GF arithmetic, exactly the reference's log/antilog formulation (rs.c:258-282) ----- */

template <int N>
struct gf {
	const uint8_t *ex, *lg;
	__device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t y) const
	{
		return (x == 0 || y == 0) ? 0u : ex[(lg[x] + lg[y]) % N];
	}
	__device__ __forceinline__ uint32_t div(uint32_t x, uint32_t y) const
	{
		return (x == 0 || y == 0) ? 0u : ex[(lg[x] - lg[y] + N) % N];
	}
	__device__ __forceinline__ uint32_t pw(uint32_t x, int e) const
	{
		return x == 0 ? 0u : ex[(lg[x] * e) % N];
	}
};

/* The log/antilog tables are built once on the host with the reference's recurrence (rs.c:78-87; note that
 * logtable[1] ends up = n, not 0, because alpha[n] = 1 overwrites it — kept on purpose) and copied to
 * shared memory by every CTA. */
__device__ gf_tables g_gf;

__device__ __forceinline__ void load_gf_tables(gf_tables &t, int tid, int nthreads)
{
	const uint32_t *src = reinterpret_cast<const uint32_t *>(&g_gf);
	uint32_t *dst = reinterpret_cast<uint32_t *>(&t);
	for (int i = tid; i < (int)(sizeof(gf_tables) / 4); i += nthreads) dst[i] = src[i];
}

/* ---- Reed-Solomon RS(255,231), warp-cooperative (rs.c:99-212, first_root 0, root_skip 1) -
 * blk[255] in shared memory is corrected in place; returns the reference's return value. */
__device__ int rs255_fix_warp(warp_smem &ws, const gf_tables &gt, int lane)
{
	constexpr int N = 255, TT = 24, T2 = 12;
	const gf<N> f = {gt.exp256, gt.log256};
	uint8_t *data = ws.blk;

	/* syndromes S_j = sum_i data[i] * alpha^(i*j), j = 0..23 (== the reference's Horner evaluation at
	 * zeroes[j] = alpha^j, rs.c:120-123,215-224; GF arithmetic is exact, so the summation order is free).
	 * Lane l takes symbols 8l .. 8l+7 and all 24 syndromes, then the lanes are XOR-reduced. */
	uint32_t part[TT];
#pragma unroll
	for (int j = 0; j < TT; j++) part[j] = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		const int i = 8 * lane + k;
		const uint32_t d = (i < N) ? data[i] : 0u;
		const uint32_t ld = gt.log256[d];
		uint32_t e = (ld >= N) ? ld - N : ld;    /* (log d + i*j) mod 255, stepped by i per syndrome; log 1 == 255 */
		const uint32_t step = (i >= N) ? i - N : i;
#pragma unroll
		for (int j = 0; j < TT; j++) {
			part[j] ^= d ? gt.exp256[e] : 0u;
			e += step;
			if (e >= N) e -= N;
		}
	}
	uint32_t syn = 0;
#pragma unroll
	for (int j = 0; j < TT; j++) {
		uint32_t v = part[j];
#pragma unroll
		for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(FULL, v, o);
		if (lane == j) syn = v;
	}
	if (lane < TT) ws.syn[lane] = (uint8_t)syn;
	if (!__any_sync(FULL, syn != 0)) return 0;
	__syncwarp();

	/* Berlekamp-Massey, lane i holds lambda[i] / prev_lambda[i] for i <= T2 (rs.c:130-173) */
	uint32_t lam = (lane == 0), plam = (lane == 0);
	int deg = 0, m = 1;
	uint32_t prev_delta = 1;
	for (int n = 0; n < TT; n++) {
		uint32_t d = (lane >= 1 && lane <= deg && lane <= T2) ? f.mul(ws.syn[n - lane], lam) : 0u;
#pragma unroll
		for (int o = 16; o; o >>= 1) d ^= __shfl_xor_sync(FULL, d, o);
		d ^= ws.syn[n];
		if (d == 0) {
			m++;
			continue;
		}
		const uint32_t scale = f.div(d, prev_delta);
		const uint32_t shifted = __shfl_sync(FULL, plam, (lane - m) & 31);
		const uint32_t upd = (lane >= m && lane <= T2) ? f.mul(scale, shifted) : 0u;
		if (2 * deg <= n) {
			plam = lam;
			lam ^= upd;
			prev_delta = d;
			deg = n + 1 - deg;
			m = 1;
			/* with more than t/2 errors the locator degree overshoots the stored polynomial:
			 * the reference then always ends in "error_count != lambda_deg" (rs.c:185) */
			if (deg > T2) return -1;
		} else {
			lam ^= upd;
			m++;
		}
	}
	if (lane <= T2) ws.lam[lane] = (uint8_t)lam;
	__syncwarp();

	/* roots: x = 1..255 in increasing order (rs.c:175-183) */
	/* lambda(x) = sum_k lam[k] x^k as thirteen INDEPENDENT table reads per point (exp[(log lam_k + k log x) mod 255]) — the
	 * same field element as the reference's Horner evaluation (rs.c:177-181), without its chain of 12 dependent
	 * multiplications (24 dependent shared-memory reads) per point */
	uint32_t llam[T2 + 1];
#pragma unroll
	for (int k = 0; k <= T2; k++) llam[k] = ws.lam[k] ? gt.log256[ws.lam[k]] : 0xffffu;
	int count = 0;
	for (int base = 1; base <= N; base += 32) {
		const int x = base + lane;
		bool is_root = false;
		if (x <= N) {
			const uint32_t lx = gt.log256[x];                 /* log 1 == 255 (the table quirk): harmless modulo 255 */
			uint32_t r = 0;
#pragma unroll
			for (int k = 0; k <= T2; k++) {
				uint32_t e = llam[k] + (uint32_t)k * lx;      /* <= 255 + 12 * 255; 256 == 1 (mod 255) */
				e = (e & 255u) + (e >> 8);
				e = (e >= (uint32_t)N) ? e - N : e;
				r ^= (llam[k] != 0xffffu) ? gt.exp256[e & 255u] : 0u;
			}
			is_root = (r == 0);
		}
		const unsigned bal = __ballot_sync(FULL, is_root);
		if (is_root) {
			const int idx = count + __popc(bal & ((1u << lane) - 1));
			if (idx < 16) {
				ws.root[idx] = (uint8_t)x;
				ws.pos[idx] = gt.log256[f.div(1, x)];           /* gaproots is the identity for root_skip 1 */
			}
		}
		count += __popc(bal);
	}
	if (count != deg) return -1;

	/* omega = syndrome * lambda mod x^24 ; lambda' (rs.c:189-190,226-256) */
	if (lane < TT) {
		uint32_t o = 0;
		for (int j = 0; j <= T2 && j <= lane; j++) o ^= f.mul(ws.syn[lane - j], ws.lam[j]);
		ws.omg[lane] = (uint8_t)o;
	}
	if (lane < T2) ws.lpr[lane] = ((lane + 1) & 1) ? ws.lam[lane + 1] : 0;
	__syncwarp();

	/* Forney (rs.c:193-204) */
	if (lane < count) {
		const uint32_t x = ws.root[lane];
		const uint32_t fcr = f.pw(x, (0 - 1 + N) % N);
		/* omega(x) and lambda'(x) the same way: independent terms instead of Horner chains (rs.c:226-256) */
		const uint32_t lx = gt.log256[x];
		uint32_t num = 0, den = 0;
#pragma unroll 8
		for (int k = 0; k < TT; k++) {
			const uint32_t c = ws.omg[k];
			uint32_t e = gt.log256[c] + (uint32_t)k * lx;         /* <= 255 + 23 * 255 */
			e = (e & 255u) + (e >> 8);
			e = (e >= (uint32_t)N) ? e - N : e;
			num ^= c ? gt.exp256[e] : 0u;
		}
#pragma unroll
		for (int k = 0; k < T2; k++) {
			const uint32_t c = ws.lpr[k];
			uint32_t e = gt.log256[c] + (uint32_t)k * lx;
			e = (e & 255u) + (e >> 8);
			e = (e >= (uint32_t)N) ? e - N : e;
			den ^= c ? gt.exp256[e] : 0u;
		}
		const int p = ws.pos[lane];
		/* p == 255 (locator of symbol 0, because logtable[1] == n) is one past the block in the
		 * reference (rs41/frame.c:43): that write never reaches the frame. */
		if (p < N) data[p] ^= (uint8_t)f.div(f.mul(num, fcr), den);
	}
	__syncwarp();
	return count;
}

/* ---- BCH(63,51) over GF(64), one lane per message (rs.c:99-212 with first_root < 0) ------
 * msg: symbol i in bit i (i = 0..63).  Returns the reference's return value. */
__device__ int bch63_fix_lane(uint64_t &msg, const gf_tables &gt)
{
	constexpr int N = 63, TT = 4, T2 = 2;
	const gf<N> f = {gt.exp64, gt.log64};
	/* syndromes at the roots {2,4,8,16} = alpha^1..alpha^4 (ims100/protocol.h:62).  The reference evaluates the message
	 * polynomial by Horner's rule (rs.c:215-224: 63 dependent table multiplications per root); over GF(2) coefficients that
	 * is the XOR of alpha^((j+1) k) over the set bits k — the same field element, from independent table reads. */
	uint32_t syn[TT] = {0, 0, 0, 0};
	for (uint64_t m = msg & ~(1ull << 63); m; m &= m - 1) {
		const uint32_t k = (uint32_t)__ffsll((long long)m) - 1u;
#pragma unroll
		for (int j = 0; j < TT; j++) {
			uint32_t e = (uint32_t)(j + 1) * k;              /* < 252; 64 == 1 (mod 63) */
			e = (e & 63u) + (e >> 6);
			e = (e >= (uint32_t)N) ? e - N : e;
			syn[j] ^= gt.exp64[e];
		}
	}
	if (!(syn[0] | syn[1] | syn[2] | syn[3])) return 0;

	uint32_t lam[T2 + 1] = {1, 0, 0}, plam[T2 + 1] = {1, 0, 0}, tmp[T2 + 1];
	int deg = 0, m = 1;
	uint32_t prev_delta = 1;
	for (int n = 0; n < TT; n++) {
		uint32_t d = syn[n];
		for (int i = 1; i <= deg && i <= T2; i++) d ^= f.mul(syn[n - i], lam[i]);
		if (d == 0) {
			m++;
		} else if (2 * deg <= n) {
			for (int i = 0; i <= T2; i++) tmp[i] = lam[i];
			for (int i = m; i <= T2; i++) lam[i] ^= f.mul(f.div(d, prev_delta), plam[i - m]);
			for (int i = 0; i <= T2; i++) plam[i] = tmp[i];
			prev_delta = d;
			deg = n + 1 - deg;
			m = 1;
			if (deg > T2) return -1;
		} else {
			for (int i = m; i <= T2; i++) lam[i] ^= f.mul(f.div(d, prev_delta), plam[i - m]);
			m++;
		}
	}
	/* roots of lambda among x = 1..63 (rs.c:175-183 walks them in this order and stops once it has `deg` of them; a
	 * polynomial of degree <= 2 has no more, so scanning all of them finds the same set).  lambda(x) = lam2 x^2 + lam1 x +
	 * lam0 from independent table reads instead of 63 dependent Horner evaluations. */
	const uint32_t l1 = gt.log64[lam[1]], l2 = gt.log64[lam[2]];
	uint64_t roots = 0;
#pragma unroll 7
	for (int x = 1; x <= N; x++) {
		const uint32_t lx = gt.log64[x];                        /* log 1 == 63 (the table quirk): harmless modulo 63 */
		uint32_t e1 = l1 + lx, e2 = l2 + 2u * lx;               /* <= 126, <= 189 */
		e1 = (e1 >= (uint32_t)N) ? e1 - N : e1;
		e1 = (e1 >= (uint32_t)N) ? e1 - N : e1;
		e2 = (e2 & 63u) + (e2 >> 6);
		e2 = (e2 >= (uint32_t)N) ? e2 - N : e2;
		const uint32_t r = (lam[2] ? gt.exp64[e2] : 0u) ^ (lam[1] ? gt.exp64[e1] : 0u) ^ lam[0];
		roots |= (uint64_t)(r == 0u) << x;
	}
	const int count = __popcll(roots);
	if (count != deg) return -1;
	for (uint64_t m = roots; m; m &= m - 1) {
		const uint32_t x = (uint32_t)__ffsll((long long)m) - 1u;
		msg ^= 1ull << gt.log64[f.div(1, x)];                   /* position may be 63: message[64] in ims100/frame.c:28-31 */
	}
	return count;
}

/* ---- checksums (decode/ecc/crc.c) -------------------------------------------------------- */
__device__ uint32_t crc16_msb(uint32_t crc, const uint8_t *p, int n)
{
	for (int i = 0; i < n; i++) {
		crc ^= (uint32_t)p[i] << 8;
		for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xffff : (crc << 1) & 0xffff;
	}
	return crc;
}

__device__ uint32_t crc16_modbus(const uint8_t *p, int n)
{
	uint32_t crc = 0xffff;
	for (int i = 0; i < n; i++) {
		crc ^= p[i];
		for (int b = 0; b < 8; b++) crc = (crc & 1) ? (crc >> 1) ^ 0xA001 : crc >> 1;
	}
	return crc;
}

__device__ uint32_t m10_crc_step(uint32_t c, uint32_t b)              /* m10/frame.c:38-54 */
{
	const uint32_t c1 = c & 0xff;
	b = ((b >> 1) | ((b & 1) << 7)) & 0xff;
	b ^= (b >> 2) & 0xff;
	const uint32_t t6 = (c & 1) ^ ((c >> 2) & 1) ^ ((c >> 4) & 1);
	const uint32_t t7 = ((c >> 1) & 1) ^ ((c >> 3) & 1) ^ ((c >> 5) & 1);
	const uint32_t t = (c & 0x3f) | (t6 << 6) | (t7 << 7);
	uint32_t s = (c >> 7) & 0xff;
	s ^= (s >> 2) & 0xff;
	return ((c1 << 8) | (b ^ t ^ s)) & 0xffff;
}

/* Manchester: keep the 2nd bit of each pair (manchester.c:6-26); writes nbits/16 bytes and
 * the trailing byte the reference always emits. */
__device__ void manchester_warp(uint8_t *dst, const uint8_t *raw, int nbits, int lane)
{
	const int nbytes = nbits / 16;
	for (int i = lane; i <= nbytes; i += 32) {
		uint32_t o = 0;
		if (i < nbytes) {
			const uint32_t v = ((uint32_t)raw[2 * i] << 8) | raw[2 * i + 1];
#pragma unroll
			for (int b = 0; b < 8; b++) o |= ((v >> (2 * b)) & 1u) << b;
		} else {
			const int rem = (nbits / 2) & 7;           /* partial output byte, right aligned like the reference */
			for (int b = 0; b < rem; b++) o = (o << 1) | buf_bit(raw, 2 * (8 * nbytes + b) + 1);
		}
		dst[i] = (uint8_t)o;
	}
}

/* ---- per-sonde post-framer pipelines ------------------------------------------------------
 * Each fills rec->status / ok / aux / data_len / data from ws.raw.  ws.work is zero on entry
 * beyond whatever the pipeline writes.  Returns the iMet-4 framer_adjust bit count (0 = none). */

__device__ void deframe_rs41(warp_smem &ws, const gf_tables &gt, int lane, int &status)
{
	uint8_t *fr = ws.work;
	for (int i = lane; i < 518; i += 32) fr[i] = (uint8_t)(bitrev8(ws.raw[i]) ^ c_rs41_prn[i & 63]);
	__syncwarp();

	const bool ext = fr[56] == 0xF0;
	const int chunk = ext ? 231 : 132;                       /* rs41/frame.c:45-50 */
	for (int i = lane; i < 256; i += 32) ws.blk[i] = 0;
	__syncwarp();
	int errors = 0;
	for (int b = 0; b < 2; b++) {
		for (int i = lane; i < chunk; i += 32) ws.blk[i] = fr[57 + 2 * i + b - 1];
		if (lane < 24) ws.blk[231 + lane] = fr[8 + lane + 24 * b];
		__syncwarp();
		const int ne = rs255_fix_warp(ws, gt, lane);
		__syncwarp();
		errors = (ne < 0 || errors < 0) ? -1 : errors + ne;
		for (int i = lane; i < chunk; i += 32) fr[57 + 2 * i + b - 1] = ws.blk[i];
		if (lane < 24) fr[8 + lane + 24 * b] = ws.blk[231 + lane];
		__syncwarp();
	}
	status = errors;
}

__device__ void deframe_dfm(warp_smem &ws, int lane, int &status, int &ok, int &aux)
{
	uint8_t *man = ws.work + 640;       /* manchester output, 35 (+1) bytes */
	uint8_t *ecc = ws.work;             /* de-interleaved ECC frame         */
	manchester_warp(man, ws.raw, 560, lane);
	__syncwarp();
	/* de-interleave: sync 2 B | ptu 7 codewords depth 7 | gps 2 x 13 codewords depth 13 */
	for (int k = lane; k < 35; k += 32) {
		uint32_t v;
		if (k < 2) {
			v = man[k];
		} else {
			int base, depth, cw;
			if (k < 9)       { base = 16;            depth = 7;  cw = k - 2; }
			else if (k < 22) { base = 16 + 56;       depth = 13; cw = k - 9; }
			else             { base = 16 + 56 + 104; depth = 13; cw = k - 22; }
			v = 0;
#pragma unroll
			for (int pbit = 0; pbit < 8; pbit++) v = (v << 1) | buf_bit(man, base + pbit * depth + cw);
		}
		ecc[k] = (uint8_t)v;
	}
	__syncwarp();
	/* Hamming(8,4): a codeword with errpos > 7 aborts its block (ptu / gps) at that byte */
	int errs[2];
	for (int blk = 0; blk < 2; blk++) {
		const int off = blk ? 9 : 2, len = blk ? 26 : 7;
		int errpos = 0;
		if (lane < len) {
			const uint32_t d = ecc[off + lane];
			errpos = (__popc(d & 0xaa) & 1) + 2 * (__popc(d & 0x66) & 1) + 4 * (__popc(d & 0x1e) & 1) +
			         8 * (__popc(d & 0xff) & 1);
		}
		const unsigned fail = __ballot_sync(FULL, errpos > 7);
		const int first_fail = fail ? __ffs(fail) - 1 : 32;
		const bool fix = lane < first_fail && errpos > 0;
		if (fix) ecc[off + lane] ^= (uint8_t)(1u << (8 - errpos));
		const int nfix = __popc(__ballot_sync(FULL, fix));
		errs[blk] = fail ? -1 : nfix;
	}
	__syncwarp();
	status = (errs[0] < 0 || errs[1] < 0) ? -1 : errs[0] + errs[1];
	ok = 0;
	aux = 0;
	if (!(status < 0 || status > 8)) {
		uint8_t *un = ws.work + 64;
		if (lane < 18) {
			uint32_t v;
			if (lane == 0)       v = ecc[2] >> 4;
			else if (lane < 4)   v = (ecc[2 + 1 + 2 * (lane - 1)] & 0xF0) | (ecc[2 + 2 + 2 * (lane - 1)] >> 4);
			else if (lane == 4)  v = ecc[9 + 12] >> 4;
			else if (lane < 11)  v = (ecc[9 + 2 * (lane - 5)] & 0xF0) | (ecc[9 + 2 * (lane - 5) + 1] >> 4);
			else if (lane == 11) v = ecc[9 + 25] >> 4;
			else                 v = (ecc[9 + 13 + 2 * (lane - 12)] & 0xF0) | (ecc[9 + 13 + 2 * (lane - 12) + 1] >> 4);
			un[lane] = (uint8_t)v;
		}
		__syncwarp();
		const unsigned nz = __ballot_sync(FULL, lane < 18 && un[lane] != 0);
		aux = nz ? 1 : 0;
		ok = aux;
	}
	/* the manchester scratch is not part of the record */
	for (int i = lane; i < 40; i += 32) man[i] = 0;
	__syncwarp();
}

__device__ void deframe_m10(warp_smem &ws, int lane, int &status)
{
	uint8_t *man = ws.work + 640;
	uint8_t *fr = ws.work;
	manchester_warp(man, ws.raw, 1664, lane);
	__syncwarp();
	/* out[i] = ~(in[i] ^ (in_stream >> 1)) with the top bit carried from the previous byte */
	for (int i = lane; i < 104; i += 32) {
		const uint32_t cur = man[i], prev = i ? man[i - 1] : 0;
		fr[i] = (uint8_t)(cur ^ 0xFF ^ (((prev << 7) & 0x80) | (cur >> 1)));
	}
	/* byte 104 = the trailing Manchester byte (0), bytes beyond read as 0 */
	for (int i = lane; i < 112; i += 32) man[i] = 0;
	__syncwarp();
	int st = 0;
	if (lane == 0) {
		const int len = fr[3];
		const uint8_t *base = fr + 3;
		const uint8_t *crc_ptr = base + len - 1;
		const uint32_t expected = ((uint32_t)crc_ptr[0] << 8) | crc_ptr[1];
		uint32_t crc = 0;
		for (const uint8_t *q = base; q < crc_ptr; q++) crc = m10_crc_step(crc, *q);
		st = (crc == expected) ? 0 : -1;
	}
	status = __shfl_sync(FULL, st, 0);
}

__device__ void deframe_ims100(warp_smem &ws, const gf_tables &gt, int lane, int &status, int &ok, int &aux)
{
	uint8_t *man = ws.work + 640;
	uint8_t *fr = ws.work;
	manchester_warp(man, ws.raw, 1200, lane);                 /* 75 bytes + trailing 0 */
	__syncwarp();
	for (int i = lane; i < 75; i += 32) {
		const uint32_t cur = man[i], nxt = man[i + 1];
		fr[i] = (uint8_t)(cur ^ ((cur << 1) | (nxt >> 7)));
	}
	__syncwarp();
	for (int i = lane; i < 80; i += 32) man[i] = 0;
	__syncwarp();

	/* 12 messages of 46 bits at bit offsets sf*300 + 24 + 46*k; symbols 17..62 of a zero-padded block */
	const int moff = (lane / 6) * 300 + 24 + 46 * (lane % 6);
	uint64_t msg = 0;
	int ret = 0;
	if (lane < 12) {
		for (int k = 0; k < 46; k++) msg |= (uint64_t)buf_bit(fr, moff + k) << (17 + k);
		ret = bch63_fix_lane(msg, gt);
	}
	/* The reference keeps the 17 padding symbols (and symbol 63) across the 12 messages of a
	 * frame (ims100/frame.c:33): if a mis-correction touched them, later messages see it.
	 * Re-run sequentially in that (rare) case so the result is exact. */
	const uint64_t padmask = 0x1FFFFull | (1ull << 63);
	const bool touched = lane < 12 && ret > 0 && (msg & padmask);
	if (__any_sync(FULL, touched)) {
		uint64_t pad = 0;
		for (int mi = 0; mi < 12; mi++) {
			uint64_t m2 = 0;
			int r2 = 0;
			if (lane == mi) {
				for (int k = 0; k < 46; k++) m2 |= (uint64_t)buf_bit(fr, moff + k) << (17 + k);
				m2 |= pad;
				r2 = bch63_fix_lane(m2, gt);
				msg = m2;
				ret = r2;
			}
			pad = __shfl_sync(FULL, m2, mi) & padmask;
		}
	}
	/* write back in message order (bitpack / bitclear of ims100/frame.c:57-64) */
	int errcount = 0;
	for (int mi = 0; mi < 12; mi++) {
		const int r = __shfl_sync(FULL, ret, mi);
		const uint64_t mm = __shfl_sync(FULL, msg, mi);
		const int off = (mi / 6) * 300 + 24 + 46 * (mi % 6);
		errcount = (r < 0 || errcount < 0) ? -1 : errcount + r;
		if (lane == 0) {
			if (r < 0)      for (int k = 0; k < 34; k++) buf_set(fr, off + k, 0);
			else if (r > 0) for (int k = 0; k < 46; k++) buf_set(fr, off + k, (uint32_t)((mm >> (17 + k)) & 1));
		}
	}
	__syncwarp();
	status = errcount;
	ok = errcount >= 0;
	aux = 0;
	if (errcount >= 0) {
		/* unpack 24 x (16 data bits + odd parity) -> 48 bytes + 24-bit valid mask (ims100/frame.c:71-113) */
		uint8_t *un = ws.work + 80;
		uint32_t vbit = 0;
		if (lane < 24) {
			const int off = (lane / 12) * 300 + 24 + 46 * ((lane % 12) / 2) + 17 * (lane & 1);
			const uint32_t b0 = buf_byte_at(fr, off), b1 = buf_byte_at(fr, off + 8), par = buf_bit(fr, off + 16);
			un[2 * lane] = (uint8_t)b0;
			un[2 * lane + 1] = (uint8_t)b1;
			vbit = ((__popc(b0) + __popc(b1)) & 1) != par ? 1u : 0u;
		}
		const unsigned bal = __ballot_sync(FULL, vbit);
		const uint32_t valid = __brev(bal) >> 8;              /* first value ends up in bit 23 */
		if (lane == 0) {
			un[48] = valid & 0xff; un[49] = (valid >> 8) & 0xff; un[50] = (valid >> 16) & 0xff; un[51] = 0;
		}
		aux = (int)valid;
	}
	__syncwarp();
}

__device__ void deframe_mrzn1(warp_smem &ws, int lane, int &status)
{
	manchester_warp(ws.work, ws.raw, 816, lane);               /* 51 bytes + trailing 0 at [51] */
	__syncwarp();
	int st = 0;
	if (lane == 0) {
		const uint32_t expected = ws.work[49] | ((uint32_t)ws.work[50] << 8);
		st = (crc16_modbus(ws.work + 4, 45) == expected) ? 0 : -1;
	}
	status = __shfl_sync(FULL, st, 0);
}

__device__ int imet4_subframe_len(const uint8_t *sf)           /* imet4/subframe.c:8-30 */
{
	if (sf[0] != 0x01) return 0;
	switch (sf[1]) {
	case 0x01: return 12 + 2;
	case 0x02: return 16 + 2;
	case 0x03: return 2 + 1 + sf[2] + 2;
	case 0x04: return 18 + 2;
	case 0x05: return 28 + 2;
	default:   return 0;
	}
}

__device__ int deframe_imet4(warp_smem &ws, int lane, int &status, int &ok, int &aux)
{
	uint8_t *fr = ws.work;
	/* 8N1: skip the 8-bit sync byte, then start bit + 8 data bits LSB first + stop bit */
	for (int i = lane; i < 60; i += 32) fr[i] = (uint8_t)bitrev8(buf_byte_at(ws.raw, 8 + 10 * i + 1));
	__syncwarp();
	int n = 0, i = 0;
	if (lane == 0) {
		int sflen = 0;
		for (i = 0; i < 72; i += sflen) {
			const uint8_t *sf = fr + i;
			sflen = imet4_subframe_len(sf);
			if (!sflen) break;
			if (crc16_msb(0x1D0F, sf, sflen) == 0) n++;
			else if (sf[1] == 0x03) break;
		}
	}
	n = __shfl_sync(FULL, n, 0);
	i = __shfl_sync(FULL, i, 0);
	status = n;
	ok = n > 0;
	aux = i;
	return i > 0 ? 10 * (1 + i) : 0;
}

__device__ void deframe_c50(warp_smem &ws, int lane, int &status)
{
	uint8_t *fr = ws.work;
	if (lane < 9) fr[lane] = (uint8_t)bitrev8(buf_byte_at(ws.raw, 10 * lane + 1));
	__syncwarp();
	int st = 0;
	if (lane == 0) {
		uint32_t s0 = 0, s1 = 0;
		for (int k = 2; k < 7; k++) { s0 = (s0 + fr[k]) & 0xff; s1 = (s1 + s0) & 0xff; }
		const uint32_t expected = ((uint32_t)fr[7] << 8) | (fr[8] ^ 0xFFu);
		st = (((s0 << 8) | s1) == expected) ? 0 : -1;
	}
	status = __shfl_sync(FULL, st, 0);
}

/* ---- the framer walk ------------------------------------------------------------------------ */

/* the framer walk of one channel by one warp */
__device__ __forceinline__ void frame_channel(const frame_params &p, const int ch, warp_smem &ws, cta_smem &sm, const int lane)
{
	if (p.active && !p.active[ch]) {
		if (lane == 0 && p.counts) { p.counts[2 * ch] = 0; p.counts[2 * ch + 1] = 0; }
		return;
	}

	const int type = p.types[ch];
	const sonde_modem &md = c_modem[type];
	const int F = md.frame_bits, S = md.sync_len;
	const uint64_t syncword = md.syncword;
	const uint64_t syncmask = (S < 64) ? ((1ull << S) - 1) : ~0ull;
	const int nq = 8 * (F / 8 + S / 8) - S;                   /* candidate offsets: correlator.c:36-63 */
	const uint8_t *ring = p.ring + (size_t)ch * p.ring_bytes;
	const uint32_t rmask = p.ring_bytes - 1;
	const uint64_t B = p.nbits[ch];

	struct { uint64_t d_pos; int n_carry, zero_prefix, frames_total, ok_total, frames_last, ok_last; } fs;
	{
		const framer_state *in = &p.fst[ch];
		fs.d_pos = in->d_pos; fs.n_carry = in->n_carry; fs.zero_prefix = in->zero_prefix;
		fs.frames_total = in->frames_total; fs.ok_total = in->ok_total;
	}
	sonde_frame_rec *recs = p.recs + (size_t)ch * p.max_frames;
	int nrec = 0, nok = 0;

	for (int i = lane; i < WORK_BYTES; i += 32) ws.work[i] = 0;
	if (lane < FRAMER_CARRY_WORDS) ws.carry[lane] = p.fst[ch].carry[lane];
	__syncwarp();

	for (;;) {
		/* READ: the framer holds n_carry carried bits followed by stream bits from d_pos */
		const int64_t avail = (int64_t)fs.n_carry + (int64_t)(B - fs.d_pos);
		if (avail < F + S) break;

		/* stage the buffer: up to 2F bits can be needed (F + sync_offset); bits past `avail` are
		 * never used */
		const int nw = (2 * F + 31) / 32 + 3;
		if (fs.n_carry == 0 && fs.zero_prefix == 0) {
			/* the common case, straight from the ring: all of a lane's aligned 32-bit loads are issued before the first
			 * use (one round trip to HBM/L2 per window instead of one per 32 words), then each window word is funnel-
			 * shifted out of two neighbouring ring words (bytes are stream order, MSB first) */
			constexpr int NI = (WIN_WORDS + 31) / 32;
			const uint32_t *ring32 = reinterpret_cast<const uint32_t *>(ring);
			const uint32_t wmask = (p.ring_bytes >> 2) - 1u;
			const uint32_t w0 = (uint32_t)(fs.d_pos >> 5) & wmask, sh = (uint32_t)(fs.d_pos & 31);
			uint32_t a[NI], b[NI];
#pragma unroll
			for (int i = 0; i < NI; i++) {
				const int w = lane + 32 * i;
				a[i] = b[i] = 0;
				if (w < nw) {
					a[i] = ring32[(w0 + w) & wmask];
					b[i] = ring32[(w0 + w + 1) & wmask];
				}
			}
#pragma unroll
			for (int i = 0; i < NI; i++) {
				const int w = lane + 32 * i;
				if (w < nw) ws.win[w] = __funnelshift_l(__byte_perm(b[i], 0, 0x0123), __byte_perm(a[i], 0, 0x0123), sh);
			}
		} else
		for (int w = lane; w < nw; w += 32) {
			const int b0 = w * 32;
			uint32_t v;
			{
				v = 0;
				for (int k = 0; k < 32; k++) {
					const int bi = b0 + k;
					uint32_t bit;
					if (bi < fs.zero_prefix)    bit = 0;
					else if (bi < fs.n_carry)   bit = win_bit(ws.carry, bi);
					else                        bit = ring_bit(ring, rmask, fs.d_pos + (bi - fs.n_carry));
					v = (v << 1) | bit;
				}
			}
			ws.win[w] = v;
		}
		__syncwarp();

		/* sync search */
		/* a lane takes 32 consecutive offsets at a time: three window words in registers, every offset a funnel shift away
		 * (the first version read three shared-memory words per offset: 38 % of the kernel's stall samples) */
		uint32_t best = 0xffffffffu;
		if (nq < 512) {
			/* short frames (SRS-C50, iMet-4: many per call, few offsets each): one offset per lane and step */
			for (int q = lane; q < nq; q += 32) {
				const uint64_t wbits = win_bits64(ws.win, q) >> (64 - S);
				const int d = __popcll((wbits ^ syncword) & syncmask);
				const int di = S - d;
				const uint32_t key = (di < d) ? (((uint32_t)di << 18) | ((uint32_t)q << 1) | 1u)
				                              : (((uint32_t)d << 18) | ((uint32_t)q << 1));
				best = min(best, key);
			}
		} else
		for (int w = lane; 32 * w < nq; w += 32) {
			const uint32_t w0 = ws.win[w], w1 = ws.win[w + 1], w2 = ws.win[w + 2];
#pragma unroll 8
			for (int b = 0; b < 32; b++) {
				const int q = 32 * w + b;
				const uint64_t win64 = ((uint64_t)__funnelshift_l(w1, w0, b) << 32) | __funnelshift_l(w2, w1, b);
				const uint64_t wbits = win64 >> (64 - S);
				const int d = __popcll((wbits ^ syncword) & syncmask);
				const int di = S - d;
				const uint32_t key = (di < d) ? (((uint32_t)di << 18) | ((uint32_t)q << 1) | 1u)
				                              : (((uint32_t)d << 18) | ((uint32_t)q << 1));
				best = (q < nq) ? min(best, key) : best;
			}
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
		const int s_off = (int)((best >> 1) & 0x1ffff);
		const int inverted = (int)(best & 1u);

		/* REALIGN needs F + sync_offset bits in the buffer */
		if (avail < F + s_off) break;
		const int offset_bits = max(F + S, F + s_off);

		/* extract + de-invert: frame bit j = buffer bit s_off + j */
		const int nbytes = (F + 7) / 8;
		for (int k = lane; k < nbytes; k += 32) {
			uint32_t v = (uint32_t)(win_bits64(ws.win, s_off + 8 * k) >> 56);
			if (inverted) v ^= 0xff;
			if (8 * k + 8 > F) {
				/* F % 8 != 0 (C50): the byte holding the frame's last bits (framer.c:94-104, bitops.c:22-29) */
				const int keep = F - 8 * k;
				const uint32_t topmask = (0xff00u >> keep) & 0xff;
				if (s_off) {
					v &= topmask;
				} else {
					/* no realign copy happened: the trailing bits are the raw buffer bits */
					const uint32_t rawv = inverted ? (v ^ 0xff) : v;
					v = (v & topmask) | (rawv & ~topmask & 0xff);
				}
			}
			ws.raw[k] = (uint8_t)v;
		}
		__syncwarp();

		/* iMet-4's 60th byte spans frame bits 599..606: the 7 bits past the frame are whatever the framer
		 * buffer holds there, i.e. the UNSHIFTED, un-inverted buffer byte 75 (imet4/frame.c:14, framer.c:94-104) */
		if (type == SONDE_IMET4 && lane == 0) ws.raw[75] = (uint8_t)(win_bits64(ws.win, 600) >> 56);
		__syncwarp();

		/* post-framer pipeline */
		int status = 0, ok = 0, aux = 0, adjust = 0;
		switch (type) {
		case SONDE_RS41:   deframe_rs41(ws, sm.gf, lane, status); ok = status >= 0; break;
		case SONDE_DFM09:  deframe_dfm(ws, lane, status, ok, aux); break;
		case SONDE_M10:    deframe_m10(ws, lane, status); ok = status >= 0; break;
		case SONDE_IMS100: deframe_ims100(ws, sm.gf, lane, status, ok, aux); break;
		case SONDE_MRZN1:  deframe_mrzn1(ws, lane, status); ok = status >= 0; break;
		case SONDE_IMET4:  adjust = deframe_imet4(ws, lane, status, ok, aux); break;
		default:           deframe_c50(ws, lane, status); ok = status >= 0; break;
		}
		__syncwarp();

		/* record */
		if (nrec < p.max_frames) {
			sonde_frame_rec *r = recs + nrec;
			if (lane == 0) {
				r->type = type;
				r->chunk = p.chunk_index;
				r->sync_offset = s_off;
				r->inverted = inverted;
				r->status = status;
				r->ok = ok;
				r->aux = aux;
				r->data_len = md.data_len;
				r->bit_pos = fs.d_pos;
			}
			uint32_t *dr = reinterpret_cast<uint32_t *>(r->raw);
			uint32_t *dd = reinterpret_cast<uint32_t *>(r->data);
			const uint32_t *sr = reinterpret_cast<const uint32_t *>(ws.raw);
			const uint32_t *sd = reinterpret_cast<const uint32_t *>(ws.work);
			for (int i = lane; i < SONDE_REC_BYTES / 4; i += 32) {
				dr[i] = (4 * i < nbytes) ? sr[i] : 0u;
				dd[i] = sd[i];
			}
		}
		nrec++;
		nok += ok;
		__syncwarp();
		for (int i = lane; i < WORK_BYTES; i += 32) ws.work[i] = 0;
		for (int i = lane; i < SONDE_REC_BYTES + 8; i += 32) ws.raw[i] = 0;
		__syncwarp();

		/* advance the framer */
		if (adjust) {
			/* framer_adjust() (framer.c:114-137): keep frame bits [adjust, F) in raw polarity, then go on
			 * with the stream bits the demodulator has not produced yet (READ_PRE is skipped). */
			const int keep = F - adjust;
			for (int w = lane; w < FRAMER_CARRY_WORDS; w += 32) {
				uint32_t v = 0;
				for (int k = 0; k < 32; k++) {
					const int j = 32 * w + k;
					v = (v << 1) | (j < keep ? win_bit(ws.win, s_off + adjust + j) : 0u);
				}
				ws.carry[w] = v;
			}
			fs.d_pos += (uint64_t)(offset_bits - fs.n_carry);
			fs.n_carry = keep;
		} else if (fs.n_carry > 0) {
			/* READ_PRE (framer.c:57-68) after a carried buffer: from buffer bit F on everything is stream */
			fs.d_pos += (uint64_t)(F - fs.n_carry);
			fs.n_carry = 0;
		} else {
			fs.d_pos += (uint64_t)F;
		}
		__syncwarp();
		fs.zero_prefix = (type == SONDE_C50 && s_off != 0) ? 6 : 0;
	}

	if (lane < FRAMER_CARRY_WORDS) p.fst[ch].carry[lane] = ws.carry[lane];
	if (lane == 0) {
		fs.frames_last = nrec;
		fs.ok_last = nok;
		fs.frames_total += nrec;
		fs.ok_total += nok;
		framer_state *o = &p.fst[ch];
		o->d_pos = fs.d_pos; o->n_carry = fs.n_carry; o->zero_prefix = fs.zero_prefix;
		o->frames_total = fs.frames_total; o->ok_total = fs.ok_total;
		o->frames_last = fs.frames_last; o->ok_last = fs.ok_last;
		if (p.counts) { p.counts[2 * ch] = nrec; p.counts[2 * ch + 1] = nok; }
	}
}

/* One warp per channel.  Experiment hooks (profiles/r2_step_experiments.md, exp52): `skip_warps` idle warps in front of the
 * `work_warps` working ones choose the SMSPs the working warps sit on (warp id % 4); with `persist` the grid is smaller than
 * the batch and every working warp walks over its channels one after the other — a framer that runs beside the next call's
 * demodulator as one quiet warp per SM.  Every such placement made the step slower than the framer simply running in
 * order behind the demodulator (0.70-0.87 ms against 0.654), so that is what the library does. */
__global__ void __launch_bounds__((WARPS_PER_CTA + MAX_SKIP_WARPS) * 32)
frame_kernel(const frame_params p)
{
	__shared__ cta_smem sm;
	const int lane = threadIdx.x & 31, wid = (int)(threadIdx.x >> 5) - p.skip_warps;
	load_gf_tables(sm.gf, threadIdx.x, blockDim.x);
	__syncthreads();
	if (wid < 0) return;
	const int work = p.work_warps > 0 ? p.work_warps : WARPS_PER_CTA;
	const int stride = p.persist ? (int)gridDim.x * work : p.n_channels;
	for (int ch = blockIdx.x * work + wid; ch < p.n_channels; ch += stride) {
		frame_channel(p, ch, sm.w[wid], sm, lane);
		__syncwarp();
	}
}

}  // namespace

extern "C" cudaError_t sonde_upload_gf_tables(void)
{
	gf_tables t;
	memset(&t, 0, sizeof(t));
	unsigned a = 1;
	t.exp256[0] = 1;
	for (int i = 1; i < 256; i++) {
		a <<= 1;
		if (a >= 256) a ^= 0x11D;
		t.exp256[i] = (uint8_t)a;
		t.log256[a] = (uint8_t)i;
	}
	a = 1;
	t.exp64[0] = 1;
	for (int i = 1; i < 64; i++) {
		a <<= 1;
		if (a >= 64) a ^= 0x61;
		t.exp64[i] = (uint8_t)a;
		t.log64[a] = (uint8_t)i;
	}
	return cudaMemcpyToSymbol(g_gf, &t, sizeof(t));
}

extern "C" cudaError_t sonde_launch_frames(const frame_params *p, cudaStream_t stream)
{
	if (p->skip_warps < 0 || p->skip_warps > MAX_SKIP_WARPS || p->work_warps < 0 || p->work_warps > WARPS_PER_CTA) return cudaErrorInvalidValue;
	const int work = p->work_warps > 0 ? p->work_warps : WARPS_PER_CTA;
	int ctas = (p->n_channels + work - 1) / work;
	if (p->persist > 0 && ctas > p->persist) ctas = p->persist;
	frame_kernel<<<ctas, (work + p->skip_warps) * 32, 0, stream>>>(*p);
	return cudaGetLastError();
}
