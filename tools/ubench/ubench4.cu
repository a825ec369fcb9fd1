// Dependent-chain latency of fp32 add / sub in their register-register and FFMA-with-immediate forms (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false ubench4.cu -o ubench4
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float add_fma(float p, float q) { float r; asm volatile("fma.rn.f32 %0, %1, 0f3F800000, %2;" : "=f"(r) : "f"(p), "f"(q)); return r; }
__device__ __forceinline__ float sub_fma(float x, float b) { float r; asm volatile("fma.rn.f32 %0, %1, 0fBF800000, %2;" : "=f"(r) : "f"(b), "f"(x)); return r; }
__device__ __forceinline__ float add_rr(float p, float q) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(p), "f"(q)); return r; }
__device__ __forceinline__ float sub_rr(float p, float q) { float r; asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(p), "f"(q)); return r; }
template <int MODE>
__global__ void k(float *out, long long *cyc, float f, int iters)
{
	float p = out[threadIdx.x], b = out[32 + threadIdx.x];
	const float x = out[64 + threadIdx.x];
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 16; u++) {
			if (MODE == 0) p = add_rr(p, f);                       // FADD R,R,R chain
			if (MODE == 1) p = add_fma(p, f);                      // FFMA R,R,1,R chain
			if (MODE == 2) { const float o = sub_rr(x, b); b = add_rr(__fmul_rn(b, 0.99f), __fmul_rn(o, 0.01f)); p += o; }   // bias chain as coded
			if (MODE == 3) { const float o = sub_fma(x, b); b = add_fma(__fmul_rn(b, 0.99f), __fmul_rn(o, 0.01f)); p += o; } // bias chain, FFMA forms
			if (MODE == 4) p = __fmul_rn(p, 0.999f);                // FMUL imm chain
		}
	}
	long long t1 = clock64();
	out[threadIdx.x] = p + b;
	if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
	float *d; long long *c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8); cudaMemset(d, 0, 4096);
	const int iters = 4096;
	const char *names[5] = {"FADD R,R,R chain", "FFMA R,R,1.0,R chain", "bias chain (FADD/FADD)", "bias chain (FFMA imm forms)", "FMUL R,R,imm chain"};
	const int per[5] = {1, 1, 1, 1, 1};
	for (int m = 0; m < 5; m++) {
		for (int rep = 0; rep < 2; rep++) {
			if (m == 0) k<0><<<1, 32>>>(d, c, 0.2f, iters);
			if (m == 1) k<1><<<1, 32>>>(d, c, 0.2f, iters);
			if (m == 2) k<2><<<1, 32>>>(d, c, 0.2f, iters);
			if (m == 3) k<3><<<1, 32>>>(d, c, 0.2f, iters);
			if (m == 4) k<4><<<1, 32>>>(d, c, 0.2f, iters);
			cudaDeviceSynchronize();
		}
		long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		printf("%-30s %.2f cycles per %s\n", names[m], (double)h / (iters * 16.0 * per[m]), (m == 2 || m == 3) ? "sample (3 dependent ops)" : "op");
	}
	return 0;
}
