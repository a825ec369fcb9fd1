/*
 * channelizer.cu — K3: wideband IQ -> C narrowband channels at 48 kS/s (include/sonde_b200_channelizer.h,
 * SURVEY.md §8 f-2; replaces the SDR++ VFO + resampler in front of the plugin, src/main.cpp:55-60).
 *
 *     y_c[m] = sum_k h[k] x[n0 - k] e^{-j w_c (n0 - k)},  n0 = mD + D - 1
 *            = e^{-j w_c n0} * sum_k' W_c[k'] x[s_m + k'],   s_m = n0 - Kp + 1,  W_c[k'] = h[Kp-1-k'] e^{+j w_c (Kp-1-k')}
 *
 * The sum is a dense contraction, shared-input across channels: one real GEMM on the tensor cores
 *
 *     Y[m][n] = sum_j A[m][j] * B[n][j]      m: output sample (M), n = 2c + {re, im} (N = 2C), j = 2k' + {re, im} (2Kp)
 *
 *     A[m][j] = the raw interleaved bf16 input stream at element 2 (mD) + j — overlapping windows; never
 *               materialised: a 2-D TMA tensor map whose row pitch (2D elements) is smaller than its row
 *               length (2Kp) reads them straight out of the sample buffer (pitch must be a multiple of 16 bytes:
 *               D % 4 == 0; other decimations store the samples in every 2nd / 4th slot, see sonde_chan_create)
 *     B[2c][2k'] = Re W, B[2c][2k'+1] = -Im W, B[2c+1][2k'] = Im W, B[2c+1][2k'+1] = Re W      (bf16, built at create)
 *
 * Kernel: persistent, one CTA per SM, warp specialised:
 *     warp 0      TMA producer: A tile 128 x 64 and B tile 256 x 64 (128-byte swizzle) into a 4-stage smem ring
 *     warp 1      one thread issues tcgen05.mma (M 128, N 256, K 16, kind::f16 bf16 -> fp32) into TMEM;
 *                 tcgen05.commit releases smem stages and publishes finished accumulators
 *     warps 4-19  epilogue: tcgen05.ld the 128 x 256 fp32 accumulator (two accumulators = all 512 TMEM columns,
 *                 so the epilogue of tile i overlaps the MMAs of tile i+1), rotate by e^{-j w_c n0} (32-bit
 *                 phase accumulator, so the phase is exact for any stream length), store complex64 [C][M]
 *                 (lanes = consecutive m: 256-byte coalesced rows)
 *
 * Roofline: 2 * M * 2C * 2Kp flops per chunk on the tensor cores; 8 B per output sample written to HBM.
 *
 * Options (sonde_chan_create_ex):
 *   precision 1  "split bf16": every operand is the sum of two bf16 numbers (hi + lo, 16 significant bits); the GEMM
 *                runs three passes into the same TMEM accumulator (x_hi w_hi + x_hi w_lo + x_lo w_hi, the dropped
 *                x_lo w_lo term is 2^-18 relative) — fp32-grade results for 3x the tensor work.
 *   interp L     rational resampling fs_out = fs_in L / M.  Output m = L q + r is branch r of a polyphase filter: the
 *                same overlapping-window GEMM with the window shifted by (e_r - e_0) samples (a TMA coordinate) and its
 *                own weight rows, y[Lq + r] = sum_k g[phi_r + k L] x[qM + e_r - k] e^{-j w (qM + e_r - k)},
 *                e_r = (rM + M - 1) div L, phi_r = (rM + M - 1) mod L.
 *   cutoff_hz[c] per-channel prototype (the per-type VFO bandwidths of src/main.hpp:45-51) — only the weight rows differ.
 */
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sonde_b200.h"
#include "../../include/sonde_b200_channelizer.h"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UK = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int N_EPI_WARPS = 16, FIRST_EPI_WARP = 4;
constexpr int CHUNKS_PER_WARP = (BN / 32) / (N_EPI_WARPS / 4);     /* 32-column chunks of the accumulator per epilogue warp */
constexpr int NTHREADS = (FIRST_EPI_WARP + N_EPI_WARPS) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

constexpr int MAX_INTERP = 16;

struct gemm_params {
	float2 *out;               /* [C][out_stride] */
	size_t out_stride;
	const uint32_t *step;      /* [C] */
	int n_channels;
	int m_rows;                /* output samples per polyphase branch this call (n_in / M) */
	int n_mt, n_nt, n_kb;      /* n_kb: K blocks of one pass */
	int npass;                 /* 1: bf16 operands; 3: split bf16 (hi hi + hi lo + lo hi) */
	int np;                    /* rows of one weight matrix (padded 2C) */
	int decim, interp;
	uint32_t n0_base;          /* low 32 bits of the absolute index of the chunk's first input sample */
	int e_abs[MAX_INTERP];     /* branch r: newest input sample of output L q + r is q M + e_abs[r]          */
	int e_rel[MAX_INTERP];     /* branch r: window shift in elements inside its sample-stream copy (a multiple of 8:
	                              TMA box rows must start on 16-byte boundaries)                              */
	int copy_of[MAX_INTERP];   /* branch r: which shifted copy of the sample stream it reads                   */
};

/* The window of polyphase branch r starts S (e_r - e_0) slots into a row; TMA needs that start 16-byte aligned
 * (4 slots), so the sample stream is kept in up to four copies shifted by 0..3 slots and a branch reads the copy that
 * makes its shift a multiple of four.  L = 1 needs one copy. */
constexpr int MAX_COPIES = 4;
struct chan_maps {
	CUtensorMap a[MAX_COPIES];      /* samples (hi part)          */
	CUtensorMap alo[MAX_COPIES];    /* split precision: residuals */
};
struct copy_ptrs {
	__nv_bfloat162 *hi[MAX_COPIES];
	__nv_bfloat162 *lo[MAX_COPIES];    /* nullptr without split precision */
	int shift[MAX_COPIES];
	int n;
};

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t b)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t b, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity)
{
	asm volatile(
		"{\n.reg .pred p;\nW_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
/* K-major operand tile, 128-byte swizzle (rows of 64 bf16 = 128 B, 8-row atoms of 1024 B): UMMA shared-memory
 * matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading byte offset [16,30)
 * (unused for swizzled K-major), stride byte offset >> 4 in [32,46) = 1024 B between 8-row groups,
 * version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64). */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr)
{
	return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
/* instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A/B bf16 (1 << 7, 1 << 10), both K-major,
 * N >> 3 in [17,23), M >> 4 in [24,29) */
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
	asm volatile(
		"{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
		::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* 32 consecutive fp32 columns of this thread's TMEM lane */
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
		  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
		  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
		  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
		: "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
chan_gemm_kernel(const __grid_constant__ chan_maps maps, const __grid_constant__ CUtensorMap tmB, const gemm_params p)
{
	extern __shared__ unsigned char smem_raw[];
	const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;            /* swizzle atoms need 1024-byte alignment */
	const uint32_t bars = base + STAGES * STAGE_BYTES;
	/* barrier layout (8 B each): full[STAGES], empty[STAGES], tfull[2], tempty[2], then the TMEM base address */
	auto full = [&](int s) { return bars + 8u * s; };
	auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
	auto tfull = [&](int a) { return bars + 8u * (2 * STAGES + a); };
	auto tempty = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
	const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int tiles_per_branch = p.n_mt * p.n_nt;
	const int n_tiles = tiles_per_branch * p.interp;

	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
		for (int a = 0; a < 2; a++) { mbar_init(tfull(a), 1); mbar_init(tempty(a), N_EPI_WARPS); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

	if (warp == 0) {
		/* ===== TMA producer ===== */
		if (lane == 0) {
			int stage = 0;
			uint32_t phase = 0;
			for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
				const int r = tile / tiles_per_branch, rem = tile - r * tiles_per_branch;
				const int mt = rem / p.n_nt, nt = rem % p.n_nt;
				for (int pass = 0; pass < p.npass; pass++) {
					/* pass 0: x_hi w_hi, 1: x_hi w_lo, 2: x_lo w_hi; weight rows: [hi | lo][branch][np] */
					const CUtensorMap *ma = (pass == 2) ? &maps.alo[p.copy_of[r]] : &maps.a[p.copy_of[r]];
					const int brow = ((pass == 1 ? p.interp : 0) + r) * p.np + nt * BN;
					for (int kb = 0; kb < p.n_kb; kb++) {
						mbar_wait(empty(stage), phase ^ 1);
						mbar_expect_tx(full(stage), STAGE_BYTES);
						const uint32_t sa = base + stage * STAGE_BYTES;
						tma_load_2d(sa, ma, kb * BK + p.e_rel[r], mt * BM, full(stage));
						tma_load_2d(sa + A_BYTES, &tmB, kb * BK, brow, full(stage));
						if (++stage == STAGES) { stage = 0; phase ^= 1; }
					}
				}
			}
		}
	} else if (warp == 1) {
		/* ===== MMA issuer ===== */
		if (lane == 0) {
			int stage = 0, as = 0;
			uint32_t phase = 0, aphase = 0;
			for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
				mbar_wait(tempty(as), aphase ^ 1);              /* epilogue drained this accumulator */
				tc_fence_after();
				const uint32_t d = tmem_base + (uint32_t)as * BN;
				const int n_kb_all = p.n_kb * p.npass;
				for (int kb = 0; kb < n_kb_all; kb++) {
					mbar_wait(full(stage), phase);
					tc_fence_after();
					const uint32_t sa = base + stage * STAGE_BYTES;
					const uint64_t da = umma_desc(sa), db = umma_desc(sa + A_BYTES);
#pragma unroll
					for (int k = 0; k < BK / UK; k++)        /* 16 bf16 = 32 B along K inside the swizzle atom: +2 in the address field */
						umma_bf16(d, da + 2u * k, db + 2u * k, (kb | k) != 0);
					umma_commit(empty(stage));                /* frees the smem stage when these MMAs retire */
					if (++stage == STAGES) { stage = 0; phase ^= 1; }
				}
				umma_commit(tfull(as));                       /* accumulator complete */
				if (++as == 2) { as = 0; aphase ^= 1; }
			}
		}
	} else if (warp >= FIRST_EPI_WARP) {
		/* ===== epilogue ===== */
		const int q = warp & 3;                               /* TMEM lane quadrant this warp may access */
		const int part = (warp - FIRST_EPI_WARP) >> 2;        /* which slice of the 256 accumulator columns */
		int as = 0;
		uint32_t aphase = 0;
		const float k_ang = 3.14159265358979323846f / 2147483648.0f;
		for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
			const int r = tile / tiles_per_branch, rem = tile - r * tiles_per_branch;
			const int mt = rem / p.n_nt, nt = rem % p.n_nt;
			mbar_wait(tfull(as), aphase);
			tc_fence_after();
			const int row = mt * BM + q * 32 + lane;                /* output sample L row + r of the call */
			const int m = row * p.interp + r;
			const uint32_t n0 = p.n0_base + (uint32_t)row * (uint32_t)p.decim + (uint32_t)p.e_abs[r];
			const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * BN;
#pragma unroll 1
			for (int ck = part * CHUNKS_PER_WARP; ck < (part + 1) * CHUNKS_PER_WARP; ck++) {
				uint32_t v[32];
				tmem_ld32(trow + ck * 32, v);
				const int c0 = nt * (BN / 2) + ck * 16;
#pragma unroll
				for (int i = 0; i < 16; i++) {
					const int c = c0 + i;
					if (c < p.n_channels && row < p.m_rows) {
						const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
						const uint32_t ph = __ldg(p.step + c) * n0;             /* phase mod 2^32: exact */
						float sn, cs;
						__sincosf((float)(int32_t)ph * k_ang, &sn, &cs);
						p.out[(size_t)c * p.out_stride + m] = make_float2(re * cs + im * sn, im * cs - re * sn);
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(tempty(as));
			if (++as == 2) { as = 0; aphase ^= 1; }
		}
	}

	tc_fence_before();
	__syncthreads();
	if (warp == 2) {
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
	}
}

/* ---- input conversion into the persistent bf16 sample buffer -------------------------------- */
/* `stride`: complex samples are stored every `stride`-th slot of the sample buffer (the slots in between stay zero),
 * which makes the window pitch a multiple of 16 bytes for decimations that are not multiples of 4 (see create). */
/* x -> bf16(x) and, for the split-precision mode (lo != nullptr), the residual bf16(x - hi), into every shifted copy */
__device__ __forceinline__ void put_sample(const copy_ptrs &cp, size_t slot, float2 v)
{
	const __nv_bfloat162 h = __float22bfloat162_rn(v);
	const float2 hf = __bfloat1622float2(h);
	const __nv_bfloat162 l = __float22bfloat162_rn(make_float2(v.x - hf.x, v.y - hf.y));
	for (int c = 0; c < cp.n; c++) {
		cp.hi[c][slot - cp.shift[c]] = h;
		if (cp.lo[c]) cp.lo[c][slot - cp.shift[c]] = l;
	}
}
__global__ void __launch_bounds__(256) to_bf16_c64_kernel(const float2 *__restrict__ in, const copy_ptrs cp, size_t n, int stride)
{
	const size_t step = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) put_sample(cp, i * stride, __ldg(in + i));
}
__global__ void __launch_bounds__(256) to_bf16_s16_kernel(const short2 *__restrict__ in, const copy_ptrs cp, size_t n, int stride, float scale)
{
	const size_t step = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
		const short2 a = __ldg(in + i);
		put_sample(cp, i * stride, make_float2((float)a.x * scale, (float)a.y * scale));
	}
}
/* offset-binary 8-bit IQ (RTL-SDR): sample = (u8 - 127.5) / 128 (exact in bf16: the residual is zero) */
__global__ void __launch_bounds__(256) to_bf16_u8_kernel(const uchar2 *__restrict__ in, const copy_ptrs cp, size_t n, int stride)
{
	const size_t step = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
		const uchar2 a = __ldg(in + i);
		put_sample(cp, i * stride, make_float2(((float)a.x - 127.5f) * 0.0078125f, ((float)a.y - 127.5f) * 0.0078125f));
	}
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct sonde_chan {
	sonde_chan_config cfg;
	int C = 0, D = 0, L = 1, S = 1, K = 0, Kp = 0, H = 0, Np = 0;   /* D: decimation M; L: interpolation; S: slot stride of
	                                                                 the sample buffer; K: taps per branch; Kp, H in slots */
	int npass = 1;
	int max_out = 0;
	int e_abs[MAX_INTERP] = {0}, e_rel[MAX_INTERP] = {0}, copy_of[MAX_INTERP] = {0};
	int n_copies = 1, copy_shift[MAX_COPIES] = {0};
	size_t nx = 0;                      /* slots per copy (without the 4 slots of front padding) */
	std::vector<float> cutoffs;                 /* [C] */
	std::vector<std::vector<double>> protos;    /* distinct prototypes g (L K taps at rate L fs_in) */
	std::vector<int> proto_of;                  /* [C] index into protos */
	std::vector<uint32_t> steps;
	/* sample stream, complex bf16, [H history | chunk]; copy c is the stream shifted by copy_shift[c] slots (slot s of
	 * the stream is d_x[c][4 + s - shift]); d_xlo: the residuals of the split-precision mode, same layout */
	__nv_bfloat162 *d_x[MAX_COPIES] = {nullptr, nullptr, nullptr, nullptr};
	__nv_bfloat162 *d_xlo[MAX_COPIES] = {nullptr, nullptr, nullptr, nullptr};
	__nv_bfloat162 *d_tail = nullptr;   /* [H] */
	__nv_bfloat16 *d_w = nullptr;       /* [hi | lo][L][Np][2 Kp] */
	uint32_t *d_step = nullptr;
	float2 *d_out[2] = {nullptr, nullptr};
	void *d_in = nullptr;               /* staging of the host entry points */
	size_t out_stride = 0;
	chan_maps maps;
	CUtensorMap tmB;
	uint64_t n_consumed = 0;            /* wideband samples consumed so far */
	unsigned long long peers_enabled = 0;   /* source devices already given peer access (process_c64_peer) */
	long n_calls = 0;
	cudaEvent_t ev[2] = {nullptr, nullptr};
	bool have_timing = false;
	int n_sms = 148;
	std::string err;
};

static int cfail(sonde_chan *h, int code, const char *msg)
{
	if (h) h->err = msg;
	return code;
}
#define CCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { if (h) h->err = std::string(#x) + ": " + cudaGetErrorString(e_); return SONDE_ERR_CUDA; } } while (0)

static uint16_t bf16_bits(float v)
{
	const __nv_bfloat16 b = __float2bfloat16_rn(v);
	uint16_t u;
	memcpy(&u, &b, 2);
	return u;
}
static float bf16_value(uint16_t u)
{
	const uint32_t w = (uint32_t)u << 16;
	float f;
	memcpy(&f, &w, 4);
	return f;
}

static int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

/* prototype low-pass at the up-sampled rate L fs_in: Hamming-windowed sinc, -6 dB at fc, gain L (unit DC gain per
 * polyphase branch) */
static std::vector<double> make_prototype(int n, double fc, double fs_up, int L)
{
	std::vector<double> g(n);
	double sum = 0;
	for (int k = 0; k < n; k++) {
		const double t = k - 0.5 * (n - 1);
		const double x = 2.0 * fc / fs_up * t;
		const double sinc = fabs(x) < 1e-12 ? 1.0 : sin(M_PI * x) / (M_PI * x);
		const double w = 0.54 - 0.46 * cos(2.0 * M_PI * k / (n - 1));
		g[k] = sinc * w;
		sum += g[k];
	}
	for (double &v : g) v *= (double)L / sum;
	return g;
}

extern "C" {

int sonde_chan_create_ex(sonde_chan **out, const sonde_chan_config *cfg, const sonde_chan_options *opt)
{
	if (!out || !cfg || !cfg->freq_hz) return SONDE_ERR_ARG;
	*out = nullptr;
	if (cfg->n_channels <= 0 || cfg->decim < 2 || cfg->fs_out <= 0 || cfg->max_in_len <= 0 || cfg->max_in_len % cfg->decim)
		return SONDE_ERR_ARG;
	const int L = (opt && opt->interp > 0) ? opt->interp : 1;
	if (L > MAX_INTERP || L >= cfg->decim || gcd_int(L, cfg->decim) != 1) return SONDE_ERR_ARG;
	if (opt && (opt->precision < 0 || opt->precision > 1)) return SONDE_ERR_ARG;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return SONDE_ERR_NODEVICE;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) return SONDE_ERR_NODEVICE;
	if (cudaSetDevice(cfg->device) != cudaSuccess) return SONDE_ERR_CUDA;

	sonde_chan *h = new sonde_chan;
	h->cfg = *cfg;
	h->cfg.freq_hz = nullptr;
	h->C = cfg->n_channels;
	h->D = cfg->decim;
	h->L = L;
	h->npass = (opt && opt->precision == 1) ? 3 : 1;
	const int tpp = cfg->taps_per_phase > 0 ? cfg->taps_per_phase : 8;
	h->K = (tpp * h->D + L - 1) / L;                 /* taps per branch: tpp output samples of input */
	/* The TMA window view needs a row pitch of 4 D S bytes that is a multiple of 16.  D % 4 == 0: S = 1.  Otherwise the
	 * samples are stored every 2nd (even D) or 4th slot with zero slots in between and the weight rows carry zeros at
	 * those positions: the same GEMM over a 2x / 4x longer K dimension (2x / 4x the tensor work, still exact). */
	h->S = (h->D % 4 == 0) ? 1 : (h->D % 2 == 0) ? 2 : 4;
	h->Kp = (h->K * h->S + 31) / 32 * 32;            /* window length in slots; 2 Kp a multiple of the 64-element K block */
	for (int r = 0; r < L; r++) h->e_abs[r] = (r * h->D + h->D - 1) / L;
	for (int r = 0; r < L; r++) {
		const int delta = h->S * (h->e_abs[r] - h->e_abs[0]), j = delta % 4;       /* window shift of the branch in slots */
		int ci = -1;
		for (int c = 0; c < h->n_copies && ci < 0; c++)
			if (h->copy_shift[c] == j) ci = c;
		if (ci < 0) { ci = h->n_copies++; h->copy_shift[ci] = j; }                 /* copy 0 has shift 0 (branch 0) */
		h->copy_of[r] = ci;
		h->e_rel[r] = 2 * (delta - j);
	}
	h->H = h->Kp - h->S * (1 + h->e_abs[0]);         /* history slots in front of the chunk (L = 1: Kp - D S) */
	h->Np = (2 * h->C + BN - 1) / BN * BN;
	h->max_out = cfg->max_in_len / h->D * L;
	h->out_stride = ((size_t)h->max_out + 3) & ~(size_t)3;
	h->n_sms = prop.multiProcessorCount;
	const double fs_in = (double)cfg->fs_out * h->D / L;

	/* prototypes: one per distinct cut-off (the per-type channel bandwidths) */
	h->cutoffs.resize(h->C);
	h->proto_of.resize(h->C);
	for (int c = 0; c < h->C; c++) {
		float fc = (opt && opt->cutoff_hz && opt->cutoff_hz[c] > 0) ? opt->cutoff_hz[c] : cfg->cutoff_hz;
		if (!(fc > 0)) fc = 0.42f * cfg->fs_out;
		h->cutoffs[c] = fc;
		int found = -1;
		for (int d = 0; d < c && found < 0; d++)
			if (h->cutoffs[d] == fc) found = h->proto_of[d];
		if (found < 0) {
			found = (int)h->protos.size();
			h->protos.push_back(make_prototype(h->K * L, fc, fs_in * L, L));
		}
		h->proto_of[c] = found;
	}

	/* oscillator steps and the weight matrices B[pass][r][n][j] (see the header comment) */
	h->steps.resize(h->C);
	const size_t wmat = (size_t)h->Np * 2 * h->Kp;
	std::vector<uint16_t> w((size_t)(h->npass > 1 ? 2 : 1) * L * wmat, 0);
	for (int c = 0; c < h->C; c++) {
		const double f = cfg->freq_hz[c] / fs_in;             /* cycles per input sample */
		if (!(fabs(f) < 0.5)) { delete h; return SONDE_ERR_ARG; }
		const long long st = llround(f * 4294967296.0);
		h->steps[c] = (uint32_t)st;
		const std::vector<double> &g = h->protos[h->proto_of[c]];
		for (int r = 0; r < L; r++) {
			const int phi = (r * h->D + h->D - 1) % L;
			uint16_t *row_re = &w[(size_t)r * wmat + (size_t)(2 * c) * 2 * h->Kp], *row_im = row_re + 2 * h->Kp;
			uint16_t *lo_re = h->npass > 1 ? row_re + (size_t)L * wmat : nullptr, *lo_im = lo_re ? lo_re + 2 * h->Kp : nullptr;
			for (int kk = 0; kk < h->Kp; kk++) {
				if ((h->Kp - 1 - kk) % h->S) continue;            /* a zero slot between samples     */
				const int k = (h->Kp - 1 - kk) / h->S;            /* tap index of window position kk */
				if (k >= h->K) continue;                          /* zero padding (oldest samples)   */
				/* phase reduced before the trig call: w_c k mod 2 pi via the integer step */
				const uint32_t ph = h->steps[c] * (uint32_t)k;
				const double a = 2.0 * M_PI * (double)(int32_t)ph / 4294967296.0;
				const double tap = g[phi + (size_t)k * L];
				const float v[4] = {(float)(tap * cos(a)), (float)(-tap * sin(a)), (float)(tap * sin(a)), (float)(tap * cos(a))};
				uint16_t hi[4];
				for (int i = 0; i < 4; i++) hi[i] = bf16_bits(v[i]);
				row_re[2 * kk] = hi[0];  row_re[2 * kk + 1] = hi[1];
				row_im[2 * kk] = hi[2];  row_im[2 * kk + 1] = hi[3];
				if (lo_re) {
					lo_re[2 * kk] = bf16_bits(v[0] - bf16_value(hi[0]));  lo_re[2 * kk + 1] = bf16_bits(v[1] - bf16_value(hi[1]));
					lo_im[2 * kk] = bf16_bits(v[2] - bf16_value(hi[2]));  lo_im[2 * kk + 1] = bf16_bits(v[3] - bf16_value(hi[3]));
				}
			}
		}
	}

	auto bail = [&](int code, const char *msg) { h->err = msg; sonde_chan_destroy(h); return code; };
	/* the tensor map has at least one full tile of rows, so the buffer covers BM windows even for tiny max_in_len */
	const size_t rows = (size_t)(cfg->max_in_len / h->D);
	const size_t rows_dim = rows > BM ? rows : (size_t)BM;
	const size_t row_len = (size_t)h->Kp + (size_t)h->S * h->D;             /* window + the largest branch shift, in slots */
	const size_t nx = rows_dim * h->D * h->S + row_len + 2 * BK;            /* + slack */
	h->nx = nx;
	for (int c = 0; c < h->n_copies; c++) {
		if (cudaMalloc(&h->d_x[c], (nx + 4) * sizeof(__nv_bfloat162)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
		if (cudaMemset(h->d_x[c], 0, (nx + 4) * sizeof(__nv_bfloat162)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMemset");
		if (h->npass > 1) {
			if (cudaMalloc(&h->d_xlo[c], (nx + 4) * sizeof(__nv_bfloat162)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
			if (cudaMemset(h->d_xlo[c], 0, (nx + 4) * sizeof(__nv_bfloat162)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMemset");
		}
	}
	if (cudaMalloc(&h->d_tail, ((size_t)h->H + 1) * sizeof(__nv_bfloat162)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
	if (cudaMalloc(&h->d_w, w.size() * 2) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
	if (cudaMemcpy(h->d_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMemcpy");
	if (cudaMalloc(&h->d_step, (size_t)h->C * 4) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
	if (cudaMemcpy(h->d_step, h->steps.data(), (size_t)h->C * 4, cudaMemcpyHostToDevice) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMemcpy");
	for (int b = 0; b < 2; b++)
		if (cudaMalloc(&h->d_out[b], (size_t)h->C * h->out_stride * sizeof(float2)) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaMalloc");
	for (int b = 0; b < 2; b++)
		if (cudaEventCreate(&h->ev[b]) != cudaSuccess) return bail(SONDE_ERR_CUDA, "cudaEventCreate");

	/* tensor maps (driver entry point through the runtime: libcuda is not linked) */
	encode_tiled_fn encode = nullptr;
	cudaDriverEntryPointQueryResult qres;
	if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode)
		return bail(SONDE_ERR_CUDA, "cuTensorMapEncodeTiled not available");
	{
		/* A: rows = output samples of one branch (pitch 2 D S elements), columns = the interleaved re/im of the window plus
		 * the branch shifts */
		const cuuint64_t gdim[2] = {(cuuint64_t)(2 * row_len), (cuuint64_t)rows_dim};      /* rows past it: zero fill, no access */
		const cuuint64_t gstr[1] = {(cuuint64_t)(2 * h->D * h->S) * 2};
		const cuuint32_t box[2] = {BK, BM}, estr[2] = {1, 1};
		memset(&h->maps, 0, sizeof(h->maps));
		for (int c = 0; c < h->n_copies; c++) {
			if (encode(&h->maps.a[c], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->d_x[c] + 4, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
				return bail(SONDE_ERR_CUDA, "tensor map A");
			h->maps.alo[c] = h->maps.a[c];
			if (h->d_xlo[c] &&
			    encode(&h->maps.alo[c], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->d_xlo[c] + 4, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
				return bail(SONDE_ERR_CUDA, "tensor map A (residuals)");
		}
		for (int c = h->n_copies; c < MAX_COPIES; c++) { h->maps.a[c] = h->maps.a[0]; h->maps.alo[c] = h->maps.alo[0]; }
		const cuuint64_t gdimb[2] = {(cuuint64_t)(2 * h->Kp), (cuuint64_t)((h->npass > 1 ? 2 : 1) * L * h->Np)};
		const cuuint64_t gstrb[1] = {(cuuint64_t)(2 * h->Kp) * 2};
		const cuuint32_t boxb[2] = {BK, BN};
		if (encode(&h->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->d_w, gdimb, gstrb, boxb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
			return bail(SONDE_ERR_CUDA, "tensor map B");
	}
	if (cudaFuncSetAttribute(chan_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
		return bail(SONDE_ERR_CUDA, "smem attribute");
	*out = h;
	return SONDE_OK;
}

int sonde_chan_create(sonde_chan **out, const sonde_chan_config *cfg) { return sonde_chan_create_ex(out, cfg, nullptr); }

void sonde_chan_destroy(sonde_chan *h)
{
	if (!h) return;
	cudaSetDevice(h->cfg.device);
	cudaDeviceSynchronize();
	for (int c = 0; c < MAX_COPIES; c++) { cudaFree(h->d_x[c]); cudaFree(h->d_xlo[c]); } cudaFree(h->d_tail); cudaFree(h->d_w); cudaFree(h->d_step); cudaFree(h->d_in);
	for (int b = 0; b < 2; b++) { cudaFree(h->d_out[b]); if (h->ev[b]) cudaEventDestroy(h->ev[b]); }
	delete h;
}

/* kind: 0 = float2 on the device, 1 = float2 on the host, 2 = short2 on the host, 3 = uchar2 on the host,
 * 4 = float2 on another device (`src_device`), pulled with the copy engine */
static int chan_run(sonde_chan *h, const void *src, size_t n_in, int kind, float scale, void *stream_, void **d_out, size_t *out_stride,
                    int src_device = -1)
{
	if (!h) return SONDE_ERR_ARG;
	if (!src || !d_out || !out_stride || n_in == 0) return cfail(h, SONDE_ERR_ARG, "null argument or zero length");
	if (n_in % h->D) return cfail(h, SONDE_ERR_ARG, "n_in is not a multiple of the decimation");
	if (n_in > (size_t)h->cfg.max_in_len) return cfail(h, SONDE_ERR_TOOLONG, "n_in > max_in_len");
	CCK(cudaSetDevice(h->cfg.device));
	cudaStream_t st = (cudaStream_t)stream_;
	const int blocks = h->n_sms * 4;
	copy_ptrs cp;                                               /* sample i lives in slot H + S i + (S - 1) of the stream */
	memset(&cp, 0, sizeof(cp));
	cp.n = h->n_copies;
	for (int c = 0; c < h->n_copies; c++) {
		cp.hi[c] = h->d_x[c] + 4 + h->H + (h->S - 1);
		cp.lo[c] = h->d_xlo[c] ? h->d_xlo[c] + 4 + h->H + (h->S - 1) : nullptr;
		cp.shift[c] = h->copy_shift[c];
	}
	if (kind == 0) {
		to_bf16_c64_kernel<<<blocks, 256, 0, st>>>(static_cast<const float2 *>(src), cp, n_in, h->S);
	} else {
		const size_t esz = (kind == 1 || kind == 4) ? sizeof(float2) : kind == 2 ? sizeof(short2) : sizeof(uchar2);
		if (!h->d_in) CCK(cudaMalloc(&h->d_in, (size_t)h->cfg.max_in_len * sizeof(float2)));
		if (kind == 4) {
			if (src_device != h->cfg.device && !(src_device >= 0 && src_device < 64 && ((h->peers_enabled >> src_device) & 1ull))) {
				int can = 0;
				CCK(cudaDeviceCanAccessPeer(&can, h->cfg.device, src_device));
				if (can) {
					const cudaError_t e = cudaDeviceEnablePeerAccess(src_device, 0);
					if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cfail(h, SONDE_ERR_CUDA, "cudaDeviceEnablePeerAccess");
					(void)cudaGetLastError();
				}
				if (src_device >= 0 && src_device < 64) h->peers_enabled |= 1ull << src_device;   /* without P2P the driver stages the copy */
			}
			CCK(cudaMemcpyPeerAsync(h->d_in, h->cfg.device, src, src_device, n_in * esz, st));
		} else {
			CCK(cudaMemcpyAsync(h->d_in, src, n_in * esz, cudaMemcpyHostToDevice, st));
		}
		if (kind == 1 || kind == 4) to_bf16_c64_kernel<<<blocks, 256, 0, st>>>(static_cast<const float2 *>(h->d_in), cp, n_in, h->S);
		else if (kind == 2) to_bf16_s16_kernel<<<blocks, 256, 0, st>>>(static_cast<const short2 *>(h->d_in), cp, n_in, h->S, scale);
		else                to_bf16_u8_kernel<<<blocks, 256, 0, st>>>(static_cast<const uchar2 *>(h->d_in), cp, n_in, h->S);
	}
	CCK(cudaGetLastError());

	gemm_params gp;
	const int par = (int)(h->n_calls & 1);
	gp.out = h->d_out[par];
	gp.out_stride = h->out_stride;
	gp.step = h->d_step;
	gp.n_channels = h->C;
	gp.m_rows = (int)(n_in / h->D);
	gp.n_mt = (gp.m_rows + BM - 1) / BM;
	gp.n_nt = h->Np / BN;
	gp.n_kb = 2 * h->Kp / BK;
	gp.npass = h->npass;
	gp.np = h->Np;
	gp.decim = h->D;
	gp.interp = h->L;
	gp.n0_base = (uint32_t)h->n_consumed;
	for (int r = 0; r < MAX_INTERP; r++) { gp.e_abs[r] = h->e_abs[r]; gp.e_rel[r] = h->e_rel[r]; gp.copy_of[r] = h->copy_of[r]; }
	const int n_tiles = gp.n_mt * gp.n_nt * h->L;
	const int grid = n_tiles < h->n_sms ? n_tiles : h->n_sms;
	CCK(cudaEventRecord(h->ev[0], st));
	chan_gemm_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(h->maps, h->tmB, gp);
	CCK(cudaGetLastError());
	CCK(cudaEventRecord(h->ev[1], st));
	h->have_timing = true;
	/* carry the last H samples to the front for the next call (through a side buffer: the ranges may overlap) */
	/* every copy is the same stream at its own shift: slots [n_in S, n_in S + H) of the stream move to [0, H); the few
	 * slots in front of a shifted copy's origin belong to samples older than any window and are never read */
	for (int c = 0; c < h->n_copies && h->H > 0; c++) {
		for (__nv_bfloat162 *buf : {h->d_x[c], h->d_xlo[c]}) {
			if (!buf) continue;
			CCK(cudaMemcpyAsync(h->d_tail, buf + 4 + n_in * h->S, (size_t)h->H * sizeof(__nv_bfloat162), cudaMemcpyDeviceToDevice, st));
			CCK(cudaMemcpyAsync(buf + 4, h->d_tail, (size_t)h->H * sizeof(__nv_bfloat162), cudaMemcpyDeviceToDevice, st));
		}
	}
	h->n_consumed += n_in;
	h->n_calls++;
	*d_out = gp.out;
	*out_stride = h->out_stride;
	return SONDE_OK;
}

int sonde_chan_process_c64(sonde_chan *h, const float *wide_iq, size_t n_in, void *stream, void **d_out, size_t *out_stride)
{
	return chan_run(h, wide_iq, n_in, 1, 1.0f, stream, d_out, out_stride);
}
int sonde_chan_process_c64_device(sonde_chan *h, const void *d_wide_iq, size_t n_in, void *stream, void **d_out, size_t *out_stride)
{
	return chan_run(h, d_wide_iq, n_in, 0, 1.0f, stream, d_out, out_stride);
}
int sonde_chan_process_c64_peer(sonde_chan *h, int src_device, const void *d_wide_iq, size_t n_in, void *stream, void **d_out, size_t *out_stride)
{
	return chan_run(h, d_wide_iq, n_in, 4, 1.0f, stream, d_out, out_stride, src_device);
}
int sonde_chan_process_s16(sonde_chan *h, const int16_t *wide_iq, size_t n_in, float scale, void *stream, void **d_out, size_t *out_stride)
{
	return chan_run(h, wide_iq, n_in, 2, scale, stream, d_out, out_stride);
}

int sonde_chan_process_u8(sonde_chan *h, const uint8_t *wide_iq, size_t n_in, void *stream, void **d_out, size_t *out_stride)
{
	return chan_run(h, wide_iq, n_in, 3, 1.0f, stream, d_out, out_stride);
}

int sonde_chan_num_taps(const sonde_chan *h) { return h ? h->K * h->L : SONDE_ERR_ARG; }
int sonde_chan_taps_of(const sonde_chan *h, int channel, double *taps, int cap)
{
	if (!h || !taps || channel < 0 || channel >= h->C) return SONDE_ERR_ARG;
	const std::vector<double> &g = h->protos[h->proto_of[channel]];
	const int n = cap < (int)g.size() ? cap : (int)g.size();
	memcpy(taps, g.data(), (size_t)n * sizeof(double));
	return n;
}
int sonde_chan_taps(const sonde_chan *h, double *taps, int cap) { return sonde_chan_taps_of(h, 0, taps, cap); }
int sonde_chan_steps(const sonde_chan *h, uint32_t *steps, int cap)
{
	if (!h || !steps) return SONDE_ERR_ARG;
	const int n = cap < h->C ? cap : h->C;
	memcpy(steps, h->steps.data(), (size_t)n * 4);
	return n;
}
int sonde_chan_last_kernel_ms(sonde_chan *h, float *gemm_ms)
{
	if (!h || !gemm_ms) return SONDE_ERR_ARG;
	*gemm_ms = -1.0f;
	if (!h->have_timing) return SONDE_OK;
	CCK(cudaSetDevice(h->cfg.device));
	CCK(cudaEventSynchronize(h->ev[1]));
	CCK(cudaEventElapsedTime(gemm_ms, h->ev[0], h->ev[1]));
	return SONDE_OK;
}
const char *sonde_chan_last_error(const sonde_chan *h) { return h ? h->err.c_str() : "null handle"; }

}  /* extern "C" */
