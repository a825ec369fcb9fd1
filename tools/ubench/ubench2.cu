// dependent-chain latencies of the fp32 ops of the AGC / timing recurrences, single warp on an idle SM
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void chain(long long *out, int iters, float b, float c, float x)
{
	float a = threadIdx.x * 1e-3f + 1.0f, o = 0.f;
	long long t0 = clock64();
	for (int k = 0; k < iters; k++) {
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
			if (MODE == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c));
			if (MODE == 2) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(c)); }
			if (MODE == 3) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); asm volatile("mul.rn.f32 %0, %0, 0f3F7D70A4;" : "+f"(a)); }
			if (MODE == 4) {   // AGC bias step, constants as immediates (what nvcc emits)
				asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(o) : "f"(x), "f"(a));
				float t1, t2;
				asm volatile("mul.rn.f32 %0, %1, 0f3F7D70A4;" : "=f"(t1) : "f"(a));
				asm volatile("mul.rn.f32 %0, %1, 0f3C23D70A;" : "=f"(t2) : "f"(o));
				asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a) : "f"(t1), "f"(t2));
			}
			if (MODE == 5) {   // same, constants in registers
				asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(o) : "f"(x), "f"(a));
				float t1, t2;
				asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(t1) : "f"(a), "f"(c));
				asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(t2) : "f"(o), "f"(b));
				asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(a) : "f"(t1), "f"(t2));
			}
			if (MODE == 6) asm volatile("add.ftz.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
			if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(c), "f"(b));
			if (MODE == 8) { asm volatile("{.reg .pred p; setp.ge.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(a) : "f"(b)); }
			if (MODE == 9) asm volatile("min.NaN.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
		}
	}
	long long t1 = clock64();
	if (a == 12345.f || o == 54321.f) out[1] = 1;
	if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
int run(long long *d, const char *name, int ops)
{
	const int iters = 4096;
	long long h;
	for (int nthr : {32, 256}) {
		chain<MODE><<<1, nthr>>>(d, iters, 1e-9f, 0.99f, 0.3f);
		CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
		printf("%-40s threads %3d : %.2f cyc per dependent op\n", name, nthr, (double)h / (iters * 8.0 * ops));
	}
	return 0;
}

int main()
{
	long long *d; CK(cudaMalloc(&d, 64)); CK(cudaMemset(d, 0, 64));
	run<0>(d, "FADD chain", 1);
	run<1>(d, "FMUL chain", 1);
	run<2>(d, "FADD -> FMUL (regs)", 2);
	run<3>(d, "FADD -> FMUL imm", 2);
	run<4>(d, "AGC bias step, imm constants (3 ops)", 3);
	run<5>(d, "AGC bias step, reg constants (3 ops)", 3);
	run<6>(d, "FADD.FTZ chain", 1);
	run<7>(d, "FFMA chain", 1);
	run<8>(d, "FSETP -> FSEL", 2);
	run<9>(d, "FMNMX.NAN chain", 1);
	return 0;
}
