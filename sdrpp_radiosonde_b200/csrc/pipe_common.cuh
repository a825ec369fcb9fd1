/*
 * pipe_common.cuh — pieces shared by the warp-specialised pipeline kernels (demod_pipe.cu, demod_pipe_afsk.cu):
 * CTA-scope mbarrier helpers, the bulk-copy (TMA) primitives, the packed fp32x2 FIR segment and the
 * tile geometry.
 */
#ifndef SONDE_PIPE_COMMON_CUH
#define SONDE_PIPE_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "device_state.h"
#include "strict_math.cuh"

/* Launch a pipeline kernel as thread-block clusters of two, so that both SMs of a TPC run the SAME kernel variant.
 * Measured (tools/mixprobe.py): when a batch mixes kernel variants, CTAs whose TPC sibling runs a different variant are
 * 30-50 % slower than beside a sibling of their own kind (the two SMs of a TPC share instruction-fetch resources and the
 * parallel-work warps stream through a long unrolled body).  An odd group count is padded with one CTA that exits at once
 * (`n_here` = real groups). */
template <class Kernel, class... Args>
static inline cudaError_t launch_tpc_pairs(Kernel kern, int n_groups, int threads, size_t smem, cudaStream_t stream, bool pairs,
                                           Args... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(pairs ? (unsigned)((n_groups + 1) & ~1) : (unsigned)n_groups, 1, 1);
	cfg.blockDim = dim3((unsigned)threads, 1, 1);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = 2;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pairs ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kern, args...);
}

namespace pipe {

constexpr int G = 8;                     /* channels per CTA                       */
constexpr int T = 256;                   /* samples per tile                       */
constexpr int NPW = 16;                  /* parallel-work warps                    */
constexpr int NPWT = NPW * 32;           /* PW threads                             */
constexpr int CPT = G * T / NPWT;        /* channels per PW thread in the per-sample stages (one sample column each) */
constexpr int RS = T + 4;                /* row stride: 16 B aligned, lanes (= rows) hit distinct banks */
constexpr int AS = SONDE_FIR_HIST + T + 4;
constexpr int R = G * T / NPWT;           /* FIR outputs per thread                 */
constexpr int SEGS = T / R;              /* FIR segments per channel row           */
constexpr int NX = 3, NS2 = 2;           /* ring depths                            */
constexpr unsigned FULL = 0xffffffffu;

static_assert(NPWT % T == 0 && CPT * (NPWT / T) == G && R % 4 == 0 && SEGS % 32 == 0 && (SEGS / 32) * G == NPW,
              "thread <-> work mappings of the pipeline kernels rely on this");

/* ---- mbarrier helpers (CTA scope) --------------------------------------------------------- */
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *b, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"W_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra D_%=;\n"
		"bra W_%=;\n"
		"D_%=:\n"
		"}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}
/* wait that charges the stalled cycles to a diagnostics counter */
__device__ __forceinline__ void mbar_wait_t(unsigned long long *b, uint32_t parity, long long &acc, bool on)
{
	if (!on) { mbar_wait(b, parity); return; }
	const long long t0 = clock64();
	mbar_wait(b, parity);
	acc += clock64() - t0;
}
/* TMA: one elected thread starts a bulk global->shared copy whose completion is signalled on an mbarrier
 * (cp.async.bulk, SASS UBLKCP).  Addresses and size must be multiples of 16 bytes. */
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *b)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(s32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(s32(b)) : "memory");
}
/* TMA, tensor form: one elected thread copies a 2-D box (columns x rows of a row-major array described by a CUtensorMap,
 * out-of-bounds elements zero-filled) into shared memory (cp.async.bulk.tensor.2d, SASS UTMALDG) */
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const void *tmap, int col, int row, unsigned long long *b)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(s32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(s32(b)), "r"(col), "r"(row) : "memory");
}
/* the same for a 3-D tensor (column, row residue, row quotient): rows a fixed step apart are consecutive in the third dimension */
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const void *tmap, int c0, int c1, int c2, unsigned long long *b)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
	             ::"r"(s32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(s32(b)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
/* all lanes of a warp finished their writes -> one arrival */
__device__ __forceinline__ void warp_arrive(unsigned long long *b, int lane)
{
	__syncwarp();
	if (lane == 0) mbar_arrive(b);
}
__device__ __forceinline__ void pw_barrier()
{
	asm volatile("bar.sync 1, %0;" ::"n"(NPWT) : "memory");
}

/* ---- S4: FIR at R consecutive positions, reference summation order (filter.c:59-61) ----------
 * Two neighbouring outputs share one packed fp32x2 multiply and one packed add (sm_100 FMUL2/FADD2,
 * round-to-nearest per lane, never fused), which halves the issue slots of the dominant loop. */
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
	unsigned long long r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
	unsigned long long r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

template <int P>
__device__ __forceinline__ void fir_segment(const float *arow, float (*y)[G][RS], const float2 *taps2,
                                            const unsigned long long negzero2, int g, int seg)
{
	float w[R + SONDE_FIR_HIST];
	const float4 *src = reinterpret_cast<const float4 *>(arow + seg * R);
#pragma unroll
	for (int k = 0; k < (R + SONDE_FIR_HIST) / 4; k++) {
		const float4 q = src[k];
		w[4 * k + 0] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
	}
#pragma unroll
	for (int br = 0; br < P; br++) {
		unsigned long long acc[R / 2];
#pragma unroll
		for (int r = 0; r < R / 2; r++) acc[r] = 0ull;                 /* (+0, +0) */
#pragma unroll
		for (int i = 0; i < SONDE_FIR_TAPS; i++) {
			const unsigned long long c = reinterpret_cast<const unsigned long long *>(taps2)[br * SONDE_FIR_TAPS + i];
#pragma unroll
			for (int r = 0; r < R / 2; r++) {
				/* a*c + (-0) == fl(a*c): the product rounded once, then the reference's add */
				const unsigned long long prod = fma2(pack2(w[2 * r + i], w[2 * r + i + 1]), c, negzero2);
				acc[r] = add2(acc[r], prod);
			}
		}
		unsigned long long *dst = reinterpret_cast<unsigned long long *>(&y[br][g][seg * R]);
#pragma unroll
		for (int r = 0; r < R / 2; r++) dst[r] = acc[r];
	}
}

/* ---- one first-order recurrence over a tile, one channel per lane: out[i] = f(in[i], state) -----------------------------
 * Blocks of 8 samples rotate through three register sets and a block is stored only after the NEXT one has been computed
 * (the instruction after an STS.128 that overwrites one of its source registers waits until the store has read them).
 * Rows have slack behind them: the float4 prefetch may run up to 8 floats past n. */
struct rec_block {
	float4 o0, o1;
};
template <class F>
__device__ __forceinline__ void rec_tile(const float *__restrict__ in, float *__restrict__ out, const int n, float &st, F f)
{
	const int nb = n >> 3;                       /* whole blocks of 8 */
	int i = 0;
	if (nb > 0) {
		const float4 *ip = reinterpret_cast<const float4 *>(in);
		float4 xa = ip[0], xb = ip[1];
		rec_block A, B, C;
		auto blk = [&](rec_block &o) {
			o.o0.x = f(xa.x, st); o.o0.y = f(xa.y, st); o.o0.z = f(xa.z, st); o.o0.w = f(xa.w, st);
			o.o1.x = f(xb.x, st); o.o1.y = f(xb.y, st); o.o1.z = f(xb.z, st); o.o1.w = f(xb.w, st);
		};
		auto put = [&](const int at, const rec_block &o) {
			*reinterpret_cast<float4 *>(out + at) = o.o0;
			*reinterpret_cast<float4 *>(out + at + 4) = o.o1;
		};
		int b = 0;
		{
			const float4 na = ip[2], nbv = ip[3];      /* rows have slack behind them: the prefetch may run past n */
			blk(A);
			xa = na; xb = nbv;
		}
		for (b = 1; b + 2 < nb; b += 3) {
			{ const float4 na = ip[2 * b + 2], nbv = ip[2 * b + 3]; blk(B); put(8 * (b - 1), A); xa = na; xb = nbv; }
			{ const float4 na = ip[2 * b + 4], nbv = ip[2 * b + 5]; blk(C); put(8 * b, B); xa = na; xb = nbv; }
			{ const float4 na = ip[2 * b + 6], nbv = ip[2 * b + 7]; blk(A); put(8 * (b + 1), C); xa = na; xb = nbv; }
		}
		if (b < nb) {
			const float4 na = ip[2 * b + 2], nbv = ip[2 * b + 3];
			blk(B); put(8 * (b - 1), A);
			xa = na; xb = nbv;
			if (b + 1 < nb) { blk(C); put(8 * b, B); put(8 * (b + 1), C); }
			else put(8 * b, B);
		} else {
			put(8 * (b - 1), A);
		}
		i = nb << 3;
	}
	for (; i < n; i++) out[i] = f(in[i], st);
}

/* the two AGC recurrences in that form, for tiles without exact-zero samples (agc.c:24-28) */
__device__ __forceinline__ void agc_bias_tile_fast(const float *__restrict__ x, float *__restrict__ s, const int n, float &bias)
{
	const float b1 = fsub(1.0f, 0.01f), b0 = 0.01f;
	rec_tile(x, s, n, bias, [=](const float xi, float &b) { const float o = fsub(xi, b); b = fadd(fmul(b, b1), fmul(o, b0)); return o; });
}
__device__ __forceinline__ void agc_level_tile_fast(const float *__restrict__ s, float *__restrict__ v, const int n, float &avg)
{
	const float g1 = fsub(1.0f, 0.001f), g0 = 0.001f;
	rec_tile(s, v, n, avg, [=](const float si, float &a) { const float o = a; a = fadd(fmul(a, g1), fmul(fabsf(si), g0)); return o; });
}

/* ---- A1 / A2: the two AGC recurrences of one channel over one tile (SD/demod/dsp/agc.c:23-28) ----------------
 * bias:  s = x - bias ; bias = bias * (1 - 0.01) + s * 0.01          avg (level):  v = avg ; avg = avg * (1 - 0.001) + |s| * 0.001
 * `check_zero`: the tile contains exact-zero samples, which bypass the AGC and do not update its state (agc.c:23).
 * Rows have 4 floats of slack, so the float4 prefetch past the tile end stays inside the row. */
__device__ __forceinline__ void agc_bias_tile(const float *__restrict__ x, float *__restrict__ s, const int n, float &bias,
                                              const bool check_zero)
{
	const float k1 = fsub(1.0f, 0.01f), k0 = 0.01f;
	if (!check_zero) {
		int i = 0;
		float4 xv = *reinterpret_cast<const float4 *>(x);
		for (; i + 4 <= n; i += 4) {
			const float4 nx = *reinterpret_cast<const float4 *>(x + i + 4);
			float4 o;
			o.x = fsub(xv.x, bias); bias = fadd(fmul(bias, k1), fmul(o.x, k0));
			o.y = fsub(xv.y, bias); bias = fadd(fmul(bias, k1), fmul(o.y, k0));
			o.z = fsub(xv.z, bias); bias = fadd(fmul(bias, k1), fmul(o.z, k0));
			o.w = fsub(xv.w, bias); bias = fadd(fmul(bias, k1), fmul(o.w, k0));
			*reinterpret_cast<float4 *>(s + i) = o;
			xv = nx;
		}
		for (; i < n; i++) {
			const float o = fsub(x[i], bias);
			bias = fadd(fmul(bias, k1), fmul(o, k0));
			s[i] = o;
		}
	} else {
		for (int i = 0; i < n; i++) {
			const float xi = x[i];
			if (xi == 0.0f) { s[i] = 0.0f; continue; }
			const float o = fsub(xi, bias);
			bias = fadd(fmul(bias, k1), fmul(o, k0));
			s[i] = o;
		}
	}
}

__device__ __forceinline__ void agc_level_tile(const float *__restrict__ s, const float *__restrict__ x, float *__restrict__ v,
                                               const int n, float &avg, const bool check_zero)
{
	const float k1 = fsub(1.0f, 0.001f), k0 = 0.001f;
	if (!check_zero) {
		int i = 0;
		float4 sv = *reinterpret_cast<const float4 *>(s);
		for (; i + 4 <= n; i += 4) {
			const float4 nx = *reinterpret_cast<const float4 *>(s + i + 4);
			float4 o;
			o.x = avg; avg = fadd(fmul(avg, k1), fmul(fabsf(sv.x), k0));
			o.y = avg; avg = fadd(fmul(avg, k1), fmul(fabsf(sv.y), k0));
			o.z = avg; avg = fadd(fmul(avg, k1), fmul(fabsf(sv.z), k0));
			o.w = avg; avg = fadd(fmul(avg, k1), fmul(fabsf(sv.w), k0));
			*reinterpret_cast<float4 *>(v + i) = o;
			sv = nx;
		}
		for (; i < n; i++) {
			v[i] = avg;
			avg = fadd(fmul(avg, k1), fmul(fabsf(s[i]), k0));
		}
	} else {
		for (int i = 0; i < n; i++) {
			v[i] = avg;
			if (x[i] == 0.0f) continue;
			avg = fadd(fmul(avg, k1), fmul(fabsf(s[i]), k0));
		}
	}
}

}  // namespace pipe

#endif
