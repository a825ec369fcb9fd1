// The bias recurrence loop of K1 (rec_tile / bias_tile of demod_pipe.cu) alone on an SM: cycles per sample.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
constexpr uint32_t ZSENT = 0x7fc5a5a5u;
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

__device__ __forceinline__ void bias_tile(const float *__restrict__ x, float *__restrict__ s, const int n, float &bias, const bool check_zero)
{
	const float b1 = fsub(1.0f, 0.01f), b0 = 0.01f;
	if (!check_zero) {
		rec_tile(x, s, n, bias, [=](const float xi, float &b) { const float o = fsub(xi, b); b = fadd(fmul(b, b1), fmul(o, b0)); return o; });
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

__global__ void k(float *out, long long *cyc, int tiles, int lanes)
{
	__shared__ __align__(16) float x[8][268], s[8][260];
	for (int i = threadIdx.x; i < 8 * 268; i += 32) (&x[0][0])[i] = 0.001f * i;
	__syncwarp();
	const int g = threadIdx.x & 7;
	float bias = 0.0f;
	long long t0 = clock64();
	for (int k = 0; k < tiles; k++) {
		if ((int)threadIdx.x < lanes) bias_tile(x[g], s[g], 256, bias, false);
		__syncwarp();
	}
	long long t1 = clock64();
	out[threadIdx.x] = bias + s[g][5];
	if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	float *d; long long *c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
	for (int lanes : {1, 7, 8}) {
		for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(d, c, 188, lanes); cudaDeviceSynchronize(); }
		long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		printf("bias_tile, %d active lanes: %.2f cycles per sample\n", lanes, (double)h / (188.0 * 256.0));
	}
	return 0;
}
