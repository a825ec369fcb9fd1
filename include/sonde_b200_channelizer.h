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
	int32_t decim;             /* D = fs_in / fs_out >= 2.  Multiples of 4 run at full speed; even D costs 2x, odd D 4x
	                              the tensor work (zero slots keep the TMA window pitch 16-byte aligned)             */
	int32_t fs_out;            /* output rate per channel, 48000                                        */
	int32_t taps_per_phase;    /* prototype low-pass length K = taps_per_phase * D; 0 -> 8              */
	float   cutoff_hz;         /* -6 dB point of the channel filter; 0 -> 0.42 * fs_out                 */
	int32_t max_in_len;        /* largest n_in (wideband samples, multiple of D) per process call       */
	int32_t device;            /* CUDA device ordinal                                                   */
	const double *freq_hz;     /* [C] channel centre offsets from the wideband centre, |f| < fs_in / 2 */
} sonde_chan_config;

/* Errors: the SONDE_ERR_* codes of sonde_b200.h. */
SONDE_API int  sonde_chan_create(sonde_chan **out, const sonde_chan_config *cfg);
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
/* int16 interleaved I,Q as SDR hardware delivers it; sample = i16 * scale */
SONDE_API int  sonde_chan_process_s16(sonde_chan *h, const int16_t *wide_iq /* host [n_in][2] */, size_t n_in,
                                      float scale, void *stream, void **d_out, size_t *out_stride);

/* offset-binary 8-bit IQ as RTL-SDR dongles deliver it; sample = (u8 - 127.5) / 128 */
SONDE_API int  sonde_chan_process_u8(sonde_chan *h, const uint8_t *wide_iq /* host [n_in][2] */, size_t n_in,
                                     void *stream, void **d_out, size_t *out_stride);

/* Parameters the oracle needs (they are inputs of the algorithm, not results): the K prototype taps, the
 * quantised oscillator steps (w_c = 2 pi step_c / 2^32), K itself. */
SONDE_API int  sonde_chan_num_taps(const sonde_chan *h);
SONDE_API int  sonde_chan_taps(const sonde_chan *h, double *taps, int cap);
SONDE_API int  sonde_chan_steps(const sonde_chan *h, uint32_t *steps, int cap);
/* duration of the last GEMM kernel in ms (CUDA events on `stream`), < 0 if none */
SONDE_API int  sonde_chan_last_kernel_ms(sonde_chan *h, float *gemm_ms);
SONDE_API const char *sonde_chan_last_error(const sonde_chan *h);

#ifdef __cplusplus
}
#endif
#endif
