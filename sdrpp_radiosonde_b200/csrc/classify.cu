/*
 * classify.cu — AUTO pre-classifier: which decoder families can this channel possibly be?
 *
 * The reference's autodetect (SD/decode.c:174-224) runs all seven decoders on a channel until one yields data — 7x the
 * demodulation cost while a channel is unlocked.  The seven modems differ strongly in the run lengths between zero
 * crossings of the (DC-removed, 3-tap smoothed) FM discriminator output at 48 kS/s:
 *
 *     RS41  4800 Bd NRZ          runs of 10, 20, 30, 40 ... samples
 *     M10   9600 Bd Manchester   runs of 5 and 10
 *     DFM / iMS-100 / MRZ-N1     2400-2500 Bd Manchester: runs of ~20 and ~40
 *     iMet  AFSK 1200 / 2200 Hz  half periods: ~10.9 and 20, nothing longer
 *     C50   AFSK 2900 / 4700 Hz  half periods: ~5.1 and ~8.3
 *
 * One warp per channel histograms the runs of the first <= 8192 samples of the buffer into a few bands and returns a
 * bit mask of plausible decoder types.  The rule only ever narrows the set when the histogram is unambiguous; anything
 * else (noise, weak signal, a transmission gap) returns all seven, i.e. the reference's behaviour.  On the synthetic
 * signals (SNR 4..25 dB) it never excluded the true type and fell back to "all" below ~10 dB.  The host still
 * re-enables every decoder for a channel that has not locked a few seconds after a narrowing decision
 * (sonde_b200.cu), so a wrong guess costs time, never the decode.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sonde_b200.h"

namespace {

constexpr int NCLS = 8192;                 /* samples examined */
constexpr uint32_t ALL = (1u << SONDE_NTYPES) - 1u;

__global__ void __launch_bounds__(32) auto_classify_kernel(const void *in, size_t row_stride, int len, int is_iq,
                                                           const int32_t *rows, int n, uint32_t *mask_out)
{
	__shared__ float d[NCLS + 2];
	__shared__ uint32_t words[NCLS / 32];
	const int ch = blockIdx.x, lane = threadIdx.x;
	if (ch >= n) return;
	const int m = len < NCLS ? len : NCLS;
	if (m < 2048) {                        /* too little signal to judge */
		if (lane == 0) mask_out[ch] = ALL;
		return;
	}
	const size_t base = (size_t)rows[ch] * row_stride;
	float sum = 0.0f;
	for (int i = lane; i < m; i += 32) {
		float v;
		if (is_iq) {
			const float2 *x = static_cast<const float2 *>(in) + base;
			const float2 c = __ldg(x + i), p = __ldg(x + (i ? i - 1 : 0));
			v = atan2f(c.y * p.x - c.x * p.y, c.x * p.x + c.y * p.y);      /* arg(x[i] conj(x[i-1])) */
		} else {
			v = __ldg(static_cast<const float *>(in) + base + i);
		}
		d[i + 1] = v;
		sum += v;
	}
	if (lane == 0) { d[0] = 0.0f; d[m + 1] = 0.0f; }
	for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
	const float mean = sum / (float)m;
	__syncwarp();
	for (int w = 0; w < m / 32; w++) {
		const int i = 32 * w + lane;
		const float v = (d[i] + d[i + 1] + d[i + 2]) * (1.0f / 3.0f) - mean;
		const uint32_t bits = __ballot_sync(0xffffffffu, v > 0.0f);
		if (lane == 0) words[w] = bits;
	}
	__syncwarp();
	if (lane != 0) return;
	int f5 = 0, f8 = 0, f10 = 0, f20 = 0, f30 = 0, total = 0;
	int prev = -1;
	uint32_t carry = words[0] & 1u;        /* no crossing at sample 0 */
	for (int w = 0; w < m / 32; w++) {
		const uint32_t b = words[w];
		uint32_t x = b ^ ((b << 1) | carry);                       /* bit j: sample 32w+j differs from its predecessor */
		carry = b >> 31;
		while (x) {
			const int pos = 32 * w + __ffs(x) - 1;
			x &= x - 1;
			if (prev >= 0) {
				const int run = pos - prev;
				total++;
				f5 += run >= 4 && run <= 6;
				f8 += run >= 7 && run <= 9;
				f10 += run >= 10 && run <= 12;
				f20 += run >= 18 && run <= 22;
				f30 += run >= 27 && run <= 33;
			}
			prev = pos;
		}
	}
	uint32_t mask = ALL;
	if (total >= 64) {
		const float t = (float)total;
		const float r5 = f5 / t, r8 = f8 / t, r10 = f10 / t, r20 = f20 / t, r30 = f30 / t;
		if (r5 >= 0.3f) {
			if (r8 >= 0.2f) mask = 1u << SONDE_C50;
			else if (r10 >= 0.15f && r8 < 0.1f) mask = 1u << SONDE_M10;
		} else if (r20 >= 0.45f && r10 < 0.1f) {
			mask = (1u << SONDE_DFM09) | (1u << SONDE_IMS100) | (1u << SONDE_MRZN1);
		} else if (r10 >= 0.25f) {
			if (r30 >= 0.06f) mask = 1u << SONDE_RS41;
			else if (r20 >= 0.1f && r30 < 0.02f) mask = 1u << SONDE_IMET4;
			else mask = (1u << SONDE_RS41) | (1u << SONDE_IMET4);
		}
	}
	mask_out[ch] = mask;
}

}  // namespace

/* rows[n]: input row of each channel to classify; mask_out[n] (device): bit t set = decoder type t plausible */
extern "C" cudaError_t sonde_launch_auto_classify(const void *d_in, size_t row_stride, int len, int is_iq,
                                                  const int32_t *d_rows, int n, uint32_t *d_mask, cudaStream_t stream)
{
	if (n <= 0) return cudaSuccess;
	auto_classify_kernel<<<n, 32, 0, stream>>>(d_in, row_stride, len, is_iq, d_rows, n, d_mask);
	return cudaGetLastError();
}
