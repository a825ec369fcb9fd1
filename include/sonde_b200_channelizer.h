/*
 * sonde_b200_channelizer.h — C ABI of the wideband front end.
 *
 * Replaces, for C channels at once, what the plugin gets from SDR++ in front of its decoders
 * (paths relative to /root/reference):
 *
 *     vfo = sigpath::vfoManager.createVFO(name, REF_CENTER, 0, bw, bw, bw, bw, true)   src/main.cpp:55
 *           -> frequency translation to the channel centre + decimating low-pass, one VFO per sonde
 *     resampler.init(&fmDemod.out, bw, OUT_SAMPLE_RATE)                               src/main.cpp:60
 *           -> rate change to the decoders' 48 kS/s
 *
 * i.e. wideband complex IQ at fs_in = D * fs_out  ->  C narrowband channels of complex64 IQ at fs_out
 * (48 kS/s), laid out [C][out_stride] exactly as sonde_b200_process_iq_device() takes them:
 *
 *     y_c[m] = sum_{k=0}^{K-1} h[k] * x[mD + D-1 - k] * exp(-j w_c (mD + D-1 - k)),   w_c = 2 pi step_c / 2^32
 *
 * SDR++'s VFO/resampler sources are not part of the reference repository (and are unpinned upstream,
 * SURVEY.md §8c), so this stage has no reference implementation to be bit-compared with: "parity unpinned".
 * Its oracle is the formula above evaluated in double precision (oracle/channelizer_oracle.py).
 *
 * The contraction over k is dense, so this is the one stage of the chain that runs on the tensor cores
 * (tcgen05, bf16 operands, fp32 accumulation in TMEM; csrc/channelizer.cu).
 */
#ifndef SONDE_B200_CHANNELIZER_H
#define SONDE_B200_CHANNELIZER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SONDE_API __attribute__((visibility("default")))
#else
#define SONDE_API
#endif

typedef struct sonde_chan sonde_chan;   /* opaque */

typedef struct {
	int32_t n_channels;        /* C                                                                     */
	int32_t decim;             /* D = fs_in / fs_out >= 2 (M of L/M with sonde_chan_options.interp).  Multiples of 4 run at full speed; even D costs 2x, odd D 4x
	                              the tensor work (zero slots keep the TMA window pitch 16-byte aligned)             */
	int32_t fs_out;            /* output rate per channel, 48000                                        */
	int32_t taps_per_phase;    /* prototype low-pass length K = taps_per_phase * D; 0 -> 8              */
	float   cutoff_hz;         /* -6 dB point of the channel filter; 0 -> 0.42 * fs_out                 */
	int32_t max_in_len;        /* largest n_in (wideband samples, multiple of D) per process call       */
	int32_t device;            /* CUDA device ordinal                                                   */
	const double *freq_hz;     /* [C] channel centre offsets from the wideband centre, |f| < fs_in / 2 */
} sonde_chan_config;

/* Optional settings of sonde_chan_create_ex (all zero = sonde_chan_create):
 *
 *  precision   SONDE_CHAN_BF16: samples and weights are rounded to bfloat16 (8 significant bits, a noise floor ~57 dB
 *              below the signal), one tensor pass.  SONDE_CHAN_SPLIT_BF16: every operand is carried as hi + lo bfloat16
 *              (16 significant bits) and the GEMM runs three passes into the same fp32 accumulator
 *              (hi*hi + hi*lo + lo*hi): within ~1e-5 rms of the exact formula, i.e. the fp32 grade of the SDR++ chain
 *              it replaces, for 3x the tensor work.
 *  interp      L of a rational resampler, fs_out = fs_in * L / decim (1 <= L <= 16, gcd(L, decim) = 1) — sources whose
 *              rate is not an integer multiple of 48 kS/s (src/main.cpp:60 uses a RationalResampler for the same
 *              reason): 2.048 MS/s -> L/M = 3/128, 2.5 MS/s -> 12/625, 10 MS/s -> 3/625.  The prototype g then has
 *              L * K taps at the rate L * fs_in and
 *                  y_c[m] = sum_n g[m*decim + decim-1 - n*L] * x[n] * exp(-j w_c n)
 *              (L = 1 is the formula at the top).  A call with n_in input samples yields n_in / decim * L outputs.
 *  cutoff_hz   [C] per-channel -6 dB point of the channel filter, e.g. half the VFO bandwidth the plugin uses for the
 *              channel's sonde type (src/main.hpp:45-51: 10 / 15 / 20 / 50 / 20 / 20 / 20 kHz); NULL or a non-positive
 *              entry -> sonde_chan_config.cutoff_hz. */
#define SONDE_CHAN_BF16        0
#define SONDE_CHAN_SPLIT_BF16  1
typedef struct {
	int32_t precision;
	int32_t interp;
	const float *cutoff_hz;
	int32_t reserved[4];
} sonde_chan_options;

/* Errors: the SONDE_ERR_* codes of sonde_b200.h. */
SONDE_API int  sonde_chan_create(sonde_chan **out, const sonde_chan_config *cfg);
SONDE_API int  sonde_chan_create_ex(sonde_chan **out, const sonde_chan_config *cfg, const sonde_chan_options *opt);
SONDE_API void sonde_chan_destroy(sonde_chan *h);

/* One chunk of wideband IQ: n_in complex samples (a multiple of D; at least K - D except for the very first
 * chunks) -> n_in / D output samples per channel, continuing the stream of the previous calls (filter history
 * and oscillator phases carry over).  Work is enqueued on `stream` (a cudaStream_t, e.g. sonde_b200_stream() of
 * the decoder that consumes the result, so that process_iq_device() may be called right after on the same
 * stream without a host sync).  *d_out receives the device address of the result, complex64 [C][*out_stride];
 * it stays valid until the second next process call (results are double buffered).  A host input buffer is copied
 * asynchronously when it is pinned (sonde_b200_host_alloc) and must then stay untouched until the stream has passed
 * the call. */
SONDE_API int  sonde_chan_process_c64(sonde_chan *h, const float *wide_iq /* host [n_in][2] */, size_t n_in,
                                      void *stream, void **d_out, size_t *out_stride);
SONDE_API int  sonde_chan_process_c64_device(sonde_chan *h, const void *d_wide_iq /* device [n_in][2] float */,
                                             size_t n_in, void *stream, void **d_out, size_t *out_stride);
/* The wideband chunk lives on ANOTHER GPU of the box (one front end feeding all GPUs): it is pulled over
 * NVLink by the copy engine (cudaMemcpyPeerAsync on `stream`: no SMs), then channelised here.  The source buffer must
 * stay untouched until the stream has passed the call; across processes it is mapped with CUDA IPC. */
SONDE_API int  sonde_chan_process_c64_peer(sonde_chan *h, int src_device, const void *d_wide_iq /* on src_device */,
                                           size_t n_in, void *stream, void **d_out, size_t *out_stride);
/* int16 interleaved I,Q as SDR hardware delivers it; sample = i16 * scale */
SONDE_API int  sonde_chan_process_s16(sonde_chan *h, const int16_t *wide_iq /* host [n_in][2] */, size_t n_in,
                                      float scale, void *stream, void **d_out, size_t *out_stride);

/* offset-binary 8-bit IQ as RTL-SDR dongles deliver it; sample = (u8 - 127.5) / 128 */
SONDE_API int  sonde_chan_process_u8(sonde_chan *h, const uint8_t *wide_iq /* host [n_in][2] */, size_t n_in,
                                     void *stream, void **d_out, size_t *out_stride);

/* Parameters the oracle needs (they are inputs of the algorithm, not results): the K prototype taps, the
 * quantised oscillator steps (w_c = 2 pi step_c / 2^32), K itself. */
SONDE_API int  sonde_chan_num_taps(const sonde_chan *h);                       /* L * K */
SONDE_API int  sonde_chan_taps(const sonde_chan *h, double *taps, int cap);   /* channel 0's prototype */
SONDE_API int  sonde_chan_taps_of(const sonde_chan *h, int channel, double *taps, int cap);
SONDE_API int  sonde_chan_steps(const sonde_chan *h, uint32_t *steps, int cap);
/* duration of the last GEMM kernel in ms (CUDA events on `stream`), < 0 if none */
SONDE_API int  sonde_chan_last_kernel_ms(sonde_chan *h, float *gemm_ms);
SONDE_API const char *sonde_chan_last_error(const sonde_chan *h);

#ifdef __cplusplus
}
#endif
#endif
