// Micro-benchmarks that drive the K1 design (run on the B200 through gpurun):
//  1. pipe throughput per SMSP of the fp32 ops the chain uses (scalar and packed f32x2)
//  2. arbiter priority: a latency-bound dependent chain next to throughput warps, as the highest / lowest warp id
//  3. single-warp issue rate with independent instructions
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum { OP_FADD, OP_FMUL, OP_FFMA, OP_FFMA2, OP_FADD2, OP_FMUL2, OP_FMNMX, OP_IADD, OP_LOP, OP_MIX_FADD_FMUL, OP_FSETP_SEL, OP_F2I, OP_NOPS };
static const char *opname[] = {"FADD", "FMUL", "FFMA", "FFMA2", "FADD2", "FMUL2", "FMNMX", "IADD3", "LOP3", "FADD+FMUL", "FSETP+FSEL", "F2I.CEIL"};

template <int OP>
__device__ __forceinline__ void body(float (&a)[8], unsigned long long (&q)[8], int (&n)[8], float b, float c, unsigned long long qb)
{
#pragma unroll
	for (int i = 0; i < 8; i++) {
		if (OP == OP_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
		if (OP == OP_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
		if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(b));
		if (OP == OP_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(qb));
		if (OP == OP_FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(qb));
		if (OP == OP_FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(qb));
		if (OP == OP_FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
		if (OP == OP_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(n[(i + 1) & 7] | 1));
		if (OP == OP_LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[i]) : "r"(n[(i + 1) & 7] | 1));
		if (OP == OP_MIX_FADD_FMUL) {
			if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
			else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
		}
		if (OP == OP_FSETP_SEL) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(a[i]) : "f"(b));
		if (OP == OP_F2I) asm volatile("cvt.rpi.s32.f32 %0, %1;" : "=r"(n[i]) : "f"(a[i]));
	}
}

template <int OP>
__global__ void tput(long long *out, int iters, float b, float c)
{
	float a[8]; unsigned long long q[8]; int n[8];
	for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; q[i] = 0x3f8000003f800000ull + i; n[i] = threadIdx.x + i; }
	unsigned long long qb = 0x3f8000013f800001ull;
	__syncthreads();
	long long t0 = clock64();
	for (int k = 0; k < iters; k++) body<OP>(a, q, n, b, c, qb);
	long long t1 = clock64();
	float s = 0; unsigned long long sq = 0; int sn = 0;
	for (int i = 0; i < 8; i++) { s += a[i]; sq += q[i]; sn += n[i]; }
	if (s == 12345.f) out[1] = 1;
	if (sq == 77) out[2] = 1;
	if (sn == 99) out[3] = 1;
	if (threadIdx.x == 0) out[0] = t1 - t0;
}

// priority test: warp `serial_wid` runs a dependent FADD chain; the other warps on the same SMSP (wid % 4 == serial_wid % 4)
// run FFMA2 throughput code until the serial warp is done.
__global__ void prio(long long *out, int iters, int serial_wid, int mode, float b)
{
	__shared__ volatile int done;
	const int warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) done = 0;
	__syncthreads();
	if (warp == serial_wid) {
		float a = threadIdx.x;
		long long t0 = clock64();
		for (int k = 0; k < iters; k++) {
#pragma unroll
			for (int i = 0; i < 8; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
		}
		long long t1 = clock64();
		if (a == 12345.f) out[1] = 1;
		if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; done = 1; }
	} else if ((warp & 3) == (serial_wid & 3)) {
		float a[8]; unsigned long long q[8]; int n[8];
		for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; q[i] = 0x3f8000003f800000ull + i; n[i] = i; }
		unsigned long long qb = 0x3f8000013f800001ull;
		while (!done) {
			for (int k = 0; k < 16; k++) {
				if (mode == 0) body<OP_FFMA2>(a, q, n, b, b, qb);
				else if (mode == 1) body<OP_FADD>(a, q, n, b, b, qb);
				else body<OP_IADD>(a, q, n, b, b, qb);
			}
		}
		float s = 0; unsigned long long sq = 0;
		for (int i = 0; i < 8; i++) { s += a[i]; sq += q[i] + n[i]; }
		if (s == 12345.f && sq == 77) out[2] = 1;
	}
}

// single warp, independent instruction mix (FADD chain x4 interleaved with LDS + integer ops): cycles per instruction
__global__ void single_ipc(long long *out, int iters, float b)
{
	float a[4] = {1, 2, 3, 4}; int n[4] = {1, 2, 3, 4};
	long long t0 = clock64();
	for (int k = 0; k < iters; k++) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
				asm volatile("add.s32 %0, %0, %1;" : "+r"(n[i]) : "r"(n[(i + 1) & 3] | 1));
			}
		}
	}
	long long t1 = clock64();
	if (a[0] + a[1] + a[2] + a[3] == 12345.f && n[0] + n[1] + n[2] + n[3] == 5) out[1] = 1;
	if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
int run_tput(long long *d, int wps)
{
	const int iters = 2048;
	long long h = 0;
	tput<OP><<<1, 128 * wps>>>(d, iters, 1e-9f, 1.0000001f);
	CK(cudaDeviceSynchronize());
	tput<OP><<<1, 128 * wps>>>(d, iters, 1e-9f, 1.0000001f);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
	const int per = (OP == OP_FSETP_SEL) ? 16 : 8;
	printf("%-11s warps/SMSP %d : %.2f cyc per warp-instr per SMSP   (%.2f cyc per instr per warp)\n", opname[OP], wps,
	       (double)h / (iters * per * wps), (double)h / (iters * per));
	return 0;
}

int main()
{
	long long *d; CK(cudaMalloc(&d, 64)); CK(cudaMemset(d, 0, 64));
	for (int wps : {1, 2, 4, 6}) {
		run_tput<OP_FADD>(d, wps); run_tput<OP_FMUL>(d, wps); run_tput<OP_FFMA>(d, wps); run_tput<OP_FFMA2>(d, wps);
		run_tput<OP_FADD2>(d, wps); run_tput<OP_FMUL2>(d, wps); run_tput<OP_FMNMX>(d, wps); run_tput<OP_IADD>(d, wps);
		run_tput<OP_LOP>(d, wps); run_tput<OP_MIX_FADD_FMUL>(d, wps); run_tput<OP_FSETP_SEL>(d, wps); run_tput<OP_F2I>(d, wps);
	}
	long long h;
	const int iters = 4096;
	for (int mode = 0; mode < 3; mode++)
		for (int nw : {8, 20}) {
			for (int sw : {0, nw - 4}) {
				prio<<<1, nw * 32>>>(d, iters, sw, mode, 1e-9f);
				CK(cudaDeviceSynchronize());
				CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
				printf("prio: %d warps (%d on the SMSP), hog=%s, serial warp id %d : %.2f cyc per dependent FADD\n", nw, nw / 4,
				       mode == 0 ? "FFMA2" : mode == 1 ? "FADD" : "IADD", sw, (double)h / (iters * 8));
			}
		}
	single_ipc<<<1, 32>>>(d, iters, 1e-9f);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
	printf("single warp, 4 FADD chains + 4 IADD chains interleaved: %.2f cyc per instr\n", (double)h / (iters * 32));
	return 0;
}
