/*
 * device_state.h — per-channel streaming state kept in HBM between process() calls,
 * and the launch parameter blocks of the two kernels.
 *
 * This is the state the reference keeps inside each decoder object:
 *   Agc {bias, moving_avg}                      demod/dsp/agc.h:5-8
 *   Filter {mem[2*49], idx}                     demod/dsp/filter.h:6-13  (kept here as the last 48 inputs)
 *   Timing {prev, phase, freq, state}           demod/dsp/timing.h:6-12
 *   AFSKDemod {p_mark, p_space, sums, history}  demod/afsk.h:16-34
 *   Framer {offset, sync_offset, state, ...}    decode/framer.h:14-27   (kept as stream positions)
 */
#ifndef SONDE_DEVICE_STATE_H
#define SONDE_DEVICE_STATE_H

#include <stdint.h>
#include "sonde_params.h"
#include "../../include/sonde_b200.h"

/* demodulator state, one per channel (256 B, 16 B aligned) */
struct __align__(16) demod_state {
	float    disc_prev;       /* previous sample's phase (discriminator)          */
	float    agc_bias;
	float    agc_avg;
	float    t_prev;
	float    t_phase;
	float    t_freq;
	int32_t  t_state;
	uint32_t bit_acc;         /* pending (<8) bits of the byte being assembled    */
	int32_t  bit_cnt;
	int32_t  nsoft;           /* soft symbols written by the last call            */
	uint64_t nbits;           /* bits demodulated since create (absolute position)*/
	uint32_t pad[4];
	float    hist[SONDE_FIR_HIST];   /* last 48 filter inputs, oldest first        */
};

/* AFSK extra state (only allocated when the batch has AFSK channels) */
struct __align__(16) afsk_state {
	float    p_mark, p_space;
	float    mark_re, mark_im, space_re, space_im;
	int32_t  idx;
	int32_t  pad;
	float    mark_hist[2 * SONDE_AFSK_MAXLEN];    /* interleaved re,im */
	float    space_hist[2 * SONDE_AFSK_MAXLEN];
};

/* framer state, one per channel: the reference's bit buffer (decode/framer.h:14-27) expressed
 * against the channel's demodulated bit stream:
 *   buffer bit i  ==  carry bit i                         for i <  n_carry
 *                     stream[d_pos + (i - n_carry)]       for i >= n_carry
 * n_carry is non-zero only after an iMet-4 framer_adjust() (framer.c:114-137), which keeps the
 * tail of the last frame and then skips the stream bits the demodulator had already consumed. */
#define FRAMER_CARRY_WORDS 20
struct __align__(16) framer_state {
	uint64_t d_pos;
	int32_t  n_carry;
	int32_t  zero_prefix;     /* C50: leading buffer bits forced to 0 (bitops.c:22-29 masking) */
	int32_t  frames_total;    /* windows emitted since create                                  */
	int32_t  ok_total;
	int32_t  frames_last;     /* windows emitted by the last call                              */
	int32_t  ok_last;
	uint32_t carry[FRAMER_CARRY_WORDS];   /* MSB-first bits, raw polarity                     */
};

#define DEMOD_G      8        /* channels per CTA                                   */
#define DEMOD_T      256      /* samples per tile                                   */
#define DEMOD_THREADS 256

struct demod_params {
	const void   *in;             /* [C][row_stride] float2 IQ or float FM           */
	size_t        row_stride;     /* in samples                                       */
	int32_t       len;            /* samples per channel this call                    */
	int32_t       is_iq;
	int32_t       use_tma;        /* rows 16-byte aligned: stage tiles with cp.async.bulk */
	int32_t       tma_box_rows;   /* > 0: the rows of every group are consecutive input rows and a tile is staged with ONE 2-D
	                                 tensor-map copy of this many rows (K1: k1_maps kernel argument); 0: one 1-D bulk copy per row */
	int32_t       tma_row_step;   /* the channels of a group are this many input rows apart (1 = consecutive; 2 = two sonde types
	                                 on alternating channels, ...)                                              */
	int32_t       n_rows;         /* rows of the input array                                                   */
	float         fm_gain;
	int32_t       n_groups;
	const int32_t *group_chan;    /* [n_groups][DEMOD_G] channel ids, -1 = empty      */
	const int32_t *group_type;    /* [n_groups]                                       */
	demod_state  *st;             /* [C]                                              */
	afsk_state   *ast;            /* [C] or NULL                                      */
	uint8_t      *ring;           /* [C][ring_bytes] demodulated bits, MSB first      */
	uint32_t      ring_bytes;     /* power of two                                     */
	float        *soft;           /* [C][soft_stride] or NULL                         */
	int32_t       soft_stride;
	long long    *prof;           /* optional [n_groups][16] cycle counters (diagnostics) */
	const int32_t *in_row;        /* [C] input row of each (virtual) channel          */
	uint32_t      pw_mask;        /* K1: bit w set = warp w of the CTA is a parallel-work warp (SMSP = w % 4)  */
	int32_t       tpc_pairs;      /* launch as clusters of two CTAs: TPC siblings run the same kernel variant  */
	uint64_t     *nbits_out;      /* [C] stream length after this call (snapshot for the framer, which
	                                 may run concurrently with the next call's demodulator)             */
};

struct frame_params {
	int32_t        n_channels;
	const int32_t *types;         /* [C]                                              */
	const uint64_t *nbits;        /* [C] demodulated stream length at the end of this call */
	framer_state  *fst;           /* [C]                                              */
	const uint8_t *ring;
	uint32_t       ring_bytes;
	sonde_frame_rec *recs;        /* [C][max_frames]                                  */
	int32_t        max_frames;
	int32_t        chunk_index;
	int32_t       *counts;        /* [C][2] frames / ok of this call                 */
	const int32_t *active;        /* [C] 0 = channel switched off (AUTO loser)        */
	int32_t        skip_warps;    /* idle warps in front of the working ones of each CTA (SMSP placement, frame.cu) */
	int32_t        work_warps;    /* working warps per CTA (0 = default)                                              */
	int32_t        persist;       /* > 0: at most this many CTAs, every working warp loops over its channels          */
};


#ifdef __CUDACC__
#include <atomic>
/* Opt a kernel into its dynamic shared-memory size once per device.  The attribute is per device, and handles may
 * live on different devices of one process (sonde_b200_config.device) and be created from different threads, so the
 * "done" state is a per-kernel bit mask indexed by the current device (devices >= 64 simply set it every time). */
template <class Kernel>
static inline cudaError_t sonde_ensure_dynamic_smem(Kernel kernel, int bytes, std::atomic<unsigned long long> &done)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	/* keep the SM's L1 / shared split at its shared-memory maximum so that small CTAs of other kernels (the framer of
	 * the previous call) can be resident beside a demodulator CTA */
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	if (e == cudaSuccess && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
	return e;
}
#endif

#endif
