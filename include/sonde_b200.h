/*
 * sonde_b200.h — C ABI of the B200 batched radiosonde demod + sync + FEC path.
 *
 * This is the drop-in boundary for the hot path of dbdexter-dev/sdrpp_radiosonde
 * (reference paths below are relative to /root/reference, "SD/" =
 * src/decode/sondedump/).  Plain pointers and sizes only; no C++/torch types.
 *
 * What each entry point replaces in the reference:
 *
 *   sonde_b200_create / _destroy
 *       N x  xxx_decoder_init(samplerate) / xxx_decoder_deinit()
 *       (SD/include/rs41.h:14,21  dfm09.h  m10.h  ims100.h  mrzn1.h  imet4.h  c50.h;
 *        called from src/decode/decoder.hpp:39,28)
 *
 *   sonde_b200_process_fm[_device]
 *       the "while (xxx_decode(dec, &frag, buf, len) != PROCEED)" loop of
 *       src/decode/decoder.hpp:61 / SD/main.c:333 for every channel of the batch:
 *       one call == one reference buffer of `len` float FM samples per channel.
 *
 *   sonde_b200_process_iq[_device]
 *       the same, preceded by dsp::demod::FM<float> (src/main.hpp:33, src/main.cpp:57)
 *       i.e. the quadrature discriminator over complex IQ.
 *
 *   sonde_b200_fetch
 *       the per-PARSED results: one sonde_frame_rec per framer window
 *       (raw frame after framer_read SD/decode/framer.c:51, frame bytes after
 *       descramble + FEC, FEC status) in the order the reference would return them.
 *
 * Threading: a handle is not thread safe; distinct handles are independent
 * (same contract as the reference decoders).
 */
#ifndef SONDE_B200_H
#define SONDE_B200_H

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

/* Sonde (decoder) types: the seven decoders of the reference (the headers under SD/include/). */
enum sonde_type {
	SONDE_RS41   = 0,   /* Vaisala RS41-SG          SD/sonde/rs41/   */
	SONDE_DFM09  = 1,   /* GRAW DFM06/09/17         SD/sonde/dfm09/  */
	SONDE_M10    = 2,   /* Meteomodem M10 / M20     SD/sonde/m10/    */
	SONDE_IMS100 = 3,   /* Meisei iMS-100 / RS-11G  SD/sonde/ims100/ */
	SONDE_MRZN1  = 4,   /* MRZ-N1                   SD/sonde/mrz-n1/ */
	SONDE_IMET4  = 5,   /* InterMet iMet-1/4        SD/sonde/imet4/  */
	SONDE_C50    = 6,   /* Meteolabor SRS-C50       SD/sonde/c50/    */
	SONDE_NTYPES = 7,
	/* AUTO: run all seven decoders on the channel until one yields a decodable frame, then lock to it
	 * (sondedump's autodetect, SD/decode.c:174-224; order rs41, m10, ims100, dfm09, imet4, c50, mrzn1) */
	SONDE_AUTO   = -1
};

/* Error codes (all entry points return 0 on success, <0 on failure). */
enum sonde_err {
	SONDE_OK            =  0,
	SONDE_ERR_ARG       = -1,   /* bad argument                          */
	SONDE_ERR_CUDA      = -2,   /* CUDA runtime failure (see last_error) */
	SONDE_ERR_NODEVICE  = -3,   /* no sm_100 device: there is NO CPU fallback */
	SONDE_ERR_TOOLONG   = -4,   /* len > max_chunk_len given at create   */
	SONDE_ERR_STATE     = -5    /* call sequence violated                */
};

#define SONDE_REC_BYTES 520     /* >= largest frame (RS41: 518 B, SD/sonde/rs41/protocol.h:19) */

/*
 * One record per framer window (== one PARSED return of xxx_decode()).
 *
 *  status  RS41   : rs41_frame_correct() result (sum of RS errors, -1 on failure)  SD/sonde/rs41/frame.c:39
 *          DFM    : dfm09_frame_correct() result                                   SD/sonde/dfm09/frame.c:41
 *          M10    : m10_frame_correct()   0 / -1                                   SD/sonde/m10/frame.c:22
 *          iMS-100: ims100_frame_error_correct() result                            SD/sonde/ims100/frame.c:21
 *          MRZ-N1 : mrzn1_frame_correct() 0 / -1                                   SD/sonde/mrz-n1/frame.c:6
 *          iMet-4 : number of subframes whose CRC16 matched                        SD/sonde/imet4/imet4.c:97-106
 *          C50    : c50_frame_correct()   0 / -1                                   SD/sonde/c50/frame.c:27
 *  ok      1 when the window passes the reference's "decodable" gate
 *  aux     DFM: 1 if the unpacked frame is not all-zero; iMS: 24-bit valid mask;
 *          iMet: byte index `i` reached by the subframe walk (framer_adjust when >0)
 *  raw     first frame_len bits of the aligned, de-inverted frame (framer output)
 *  data    frame bytes after descramble + FEC:
 *          RS41 518 B | DFM 35 B ECC frame, 18 B unpacked at +64 | M10 104 B |
 *          iMS 75 B ECC frame, 52 B unpacked at +80 | MRZ 51 B | iMet 60 B | C50 9 B
 */
typedef struct {
	int32_t  type;
	int32_t  chunk;        /* index of the process_*() call in which the frame completed */
	int32_t  sync_offset;  /* correlate() result, SD/decode/correlator/correlator.c:19 */
	int32_t  inverted;
	int32_t  status;
	int32_t  ok;
	int32_t  aux;
	int32_t  data_len;
	uint64_t bit_pos;      /* index, in the channel's demodulated bit stream, of the window start */
	uint8_t  raw[SONDE_REC_BYTES];
	uint8_t  data[SONDE_REC_BYTES];
} sonde_frame_rec;

typedef struct sonde_b200 sonde_b200;   /* opaque batch decoder */

typedef struct {
	int32_t n_channels;       /* C                                                   */
	int32_t samplerate;       /* per-channel input rate, e.g. 48000                  */
	int32_t max_chunk_len;    /* largest `len` that will be passed to process_*()    */
	int32_t device;           /* CUDA device ordinal                                 */
	const int32_t *types;     /* [C] enum sonde_type per channel                     */
	float   fm_gain;          /* discriminator gain (1/deviation); 0 -> 2/pi, i.e.
	                             dsp::demod::FM(samplerate=bw, bandwidth=bw/2)  src/main.cpp:57 */
	int32_t keep_soft;        /* !=0: keep soft symbols for sonde_b200_fetch_soft()  */
	int32_t reserved;         /* bit 0: use the phase-by-phase demod kernel (cross-check only); bit 1: never stage with TMA;
	                             bit 2: alternative warp placement of the AFSK pipeline kernel (diagnostics);
	                             bit 3: AUTO pre-classifier: before an unlocked AUTO channel is demodulated for the
	                             first time a cheap kernel looks at the run lengths between zero crossings of its discriminator
	                             output; if they identify one modem family unambiguously only that family's decoders are tried
	                             (instead of all seven), and all seven are restored if the channel has not locked 3 s later.
	                             Off by default: the default is the reference's try-all (SD/decode.c:174-224). */
} sonde_b200_config;

SONDE_API int  sonde_b200_create(sonde_b200 **out, const sonde_b200_config *cfg);
SONDE_API void sonde_b200_destroy(sonde_b200 *h);

/* Host-buffer entry points: H2D copy + kernels are enqueued (asynchronously when the buffer is pinned, see
 * sonde_b200_host_alloc; the buffer must then stay untouched until the call has been fetched or synced). */
SONDE_API int  sonde_b200_process_iq(sonde_b200 *h, const float *iq /*[C][len][2]*/, size_t len);
SONDE_API int  sonde_b200_process_fm(sonde_b200 *h, const float *fm /*[C][len]*/,    size_t len);

/* 16-bit IQ as SDR hardware delivers it (interleaved I,Q int16): sample = (float)i16 * scale, which is exact for a
 * power-of-two scale, so the result equals sonde_b200_process_iq() on the converted floats bit for bit.  Halves
 * the bytes crossing PCIe — the bound of the host entry points (the complex64 path replaces what SDR++ hands the
 * plugin, src/main.cpp:55-57; this one is for callers that own the front end).  The conversion runs on the GPU. */
SONDE_API int  sonde_b200_process_iq_s16(sonde_b200 *h, const int16_t *iq /*[C][len][2]*/, size_t len, float scale);

/* Device-buffer entry points (inputs already resident in HBM). row_stride in samples. */
SONDE_API int  sonde_b200_process_iq_device(sonde_b200 *h, const void *d_iq, size_t len, size_t row_stride);
SONDE_API int  sonde_b200_process_fm_device(sonde_b200 *h, const void *d_fm, size_t len, size_t row_stride);

/* Input resident on ANOTHER GPU of the box (a front-end GPU feeding its peers; SURVEY.md §8e, the north_star's
 * "scatter the channel batch"): this handle's [C][len] complex64 block is pulled from device `src_device` over NVLink
 * by the copy engine (cudaMemcpyPeerAsync, no SMs involved) and decoded.  Double buffered like the host entry points:
 * process_iq_peer(i+1) may be issued before fetch(i). */
SONDE_API int  sonde_b200_process_iq_peer(sonde_b200 *h, int src_device, const void *d_iq_src, size_t len, size_t row_stride);

/* Capacity (records per channel per process call) of the fetch arrays. */
SONDE_API int  sonde_b200_max_frames(const sonde_b200 *h);

/* Copy the records of one process call: recs[C][max_frames], counts[C].  Blocks until that call is done.
 * Results are double buffered: fetch() serves the OLDEST call not fetched yet among the last two issued, so
 *     process(0); process(1); fetch() -> call 0; process(2); fetch() -> call 1; ...
 * overlaps the host->device copy of call i+1 and the fetch of call i with the kernels in between, while the
 * plain process(); fetch(); process(); fetch(); sequence keeps returning the call just made.  Results of calls
 * older than the last two are dropped if they were never fetched. */
SONDE_API int  sonde_b200_fetch(sonde_b200 *h, sonde_frame_rec *recs, int32_t *counts);

/* Only the per-channel counters (cheap D2H): frames[C], ok[C]; same call selection as fetch(). */
SONDE_API int  sonde_b200_fetch_counts(sonde_b200 *h, int32_t *frames, int32_t *ok);

/* AUTO channels: types[C] receives the decoder each channel is locked to (SONDE_AUTO while undetermined; fixed
 * channels report their own type).  A channel locks at the fetch() of the first call in which one of the seven
 * decoders, tried in the reference's order, produced a frame passing its gate; until then fetch() reports no
 * records for it, from then on the records of the locked decoder; the six others stop running. */
SONDE_API int  sonde_b200_detected_types(sonde_b200 *h, int32_t *types);

/* Running totals since create, per channel: framer windows, windows passing the gate, demodulated bits. */
/* masks[C]: bit t set = decoder type t is currently run for the channel (one bit for fixed and locked channels,
 * seven for an unlocked AUTO channel unless the pre-classifier narrowed it) */
SONDE_API int  sonde_b200_auto_plausible(sonde_b200 *h, uint32_t *masks);

SONDE_API int  sonde_b200_fetch_totals(sonde_b200 *h, int64_t *frames, int64_t *ok, int64_t *bits);

/* Parity taps for the last process call.
 *  bits : demodulated hard bits, MSB first, bits[C][bits_stride_bytes]; nbits[C]
 *  soft : soft symbols at symbol instants (needs keep_soft), soft[C][soft_stride]; nsoft[C] */
SONDE_API int  sonde_b200_bits_stride(const sonde_b200 *h);
SONDE_API int  sonde_b200_fetch_bits(sonde_b200 *h, uint8_t *bits, int32_t *nbits);
SONDE_API int  sonde_b200_soft_stride(const sonde_b200 *h);
SONDE_API int  sonde_b200_fetch_soft(sonde_b200 *h, float *soft, int32_t *nsoft);

/* Loop state after the last call, state[C][8] = agc.bias, agc.moving_avg, timing.phase,
 * timing.freq, timing.prev, timing.state, discriminator phase, 0  (parity tap). */
SONDE_API int  sonde_b200_fetch_state(sonde_b200 *h, float *state);

/* Host-only (no GPU needed): the demodulator constants the kernels use for (type, samplerate),
 * i.e. what gfsk_init()/afsk_init() derive (SD/demod/gfsk.c:17-35, afsk.c:16-48).
 * taps[phase*49+i] as SD/demod/dsp/filter.c:21-25 lays them out; consts[8] = timing.freq, alpha,
 * beta, max_fdev, num_phases, afsk boxcar len, f_mark, f_space.  Returns the tap count or <0. */
SONDE_API int  sonde_b200_modem_info(int type, int samplerate, float *taps, int taps_cap, float *consts);

/* Pinned host memory for the host-buffer entry points (plain malloc'd memory also works, slower). */
SONDE_API void *sonde_b200_host_alloc(size_t bytes);
/* The same as write-combined memory, for staging buffers the host only WRITES (reading them back on the CPU is very slow):
 * the GPU's reads then cause no cache snooping on the host.  Whether that helps depends on the host: with two B200s of
 * one box copying at once a bare copy ran at 55.5 GB/s per GPU from write-combined and 48 GB/s from ordinary pinned
 * memory, with all eight at once at 13-16 GB/s against 23-35 GB/s (bench.py e2e.h2d_link_gbs_all_ranks_at_once reports
 * both for the box it runs on).  The library itself stages through ordinary pinned memory. */
SONDE_API void *sonde_b200_host_alloc_wc(size_t bytes);
SONDE_API void  sonde_b200_host_free(void *p);

/* Diagnostics, host only (no GPU needed): how a batch of `types` would be grouped into CTAs — per kernel variant v = 0..3
 * out17[4 v + {0,1,2,3}] = groups, channels per group, rows of the TMA box (0: one bulk copy per row), input-row step of a
 * group's channels; out17[16] = CTAs of one call including cluster padding. */
SONDE_API int  sonde_b200_debug_plan(const int32_t *types, int n_channels, int samplerate, int n_sms, int32_t *out17);

/* Diagnostics: the first call switches on the pipeline kernel's per-CTA stall counters; later calls copy them
 * out: out[groups][16] cycles = [role*4 + {wait for input, wait for output slot, total}], roles 0..3 = parallel
 * warps, AGC bias lane, AGC level lane, timing lane.  Returns the number of CTA groups. */
SONDE_API int  sonde_b200_debug_stalls(sonde_b200 *h, long long *out, int cap_groups);

/* Diagnostics: copies the raw per-channel demodulator state (AGC, timing loop, FIR memory; 256 bytes per channel,
 * layout of csrc/device_state.h) into out.  Returns the record size, or a negative error. */
SONDE_API int  sonde_b200_debug_demod_state(sonde_b200 *h, void *out, size_t cap_bytes);

/* The CUDA stream (cudaStream_t) all work of this handle is enqueued on. */
SONDE_API void *sonde_b200_stream(sonde_b200 *h);

/* Device-side join: the stream returned by sonde_b200_stream() waits for the framer kernels of every call made
 * so far.  The framer runs in order on that stream by default, so this is a no-op then; with the experiment switch
 * SONDE_FRAME_OVERLAP=1 it runs on an internal stream beside the next call's demodulator.  Does not block the host. */
SONDE_API int  sonde_b200_join(sonde_b200 *h);

/* Block until everything enqueued so far has finished. */
SONDE_API int  sonde_b200_sync(sonde_b200 *h);

/* CUDA-event timing of the kernels of the last process call, in ms (demod, frame). */
SONDE_API int  sonde_b200_last_kernel_ms(sonde_b200 *h, float *demod_ms, float *frame_ms);

/* Number of kernels launched by this handle so far. */
SONDE_API long sonde_b200_launch_count(const sonde_b200 *h);

SONDE_API const char *sonde_b200_last_error(const sonde_b200 *h);
SONDE_API const char *sonde_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SONDE_B200_H */
