/*
 * sonde_oracle.h — CPU restatement ("port") of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from this file.
 * The product (sdrpp_radiosonde_b200/) never includes, links or calls it.
 *
 * Parity status: PINNED for everything from float FM samples to frame bytes
 * (checked against the unmodified reference compiled into oracle/_ref and against
 * the committed fixtures in tests/golden/).  The FM discriminator (orc_discriminate)
 * is "parity unpinned": its upstream (SDR++ core dsp::demod::FM) is not vendored
 * in the reference tree, see DESIGN.md.
 */
#ifndef SONDE_ORACLE_H
#define SONDE_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "sonde_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* same signatures as the ref_* entry points of oracle/ref_harness.c */
int  orc_frames_run(int type, int samplerate, const float *fm, size_t n, size_t chunk,
                    sonde_frame_rec *recs, int max_recs);
int  orc_frames_run_ragged(int type, int samplerate, const float *fm, size_t n, const size_t *chunks, int n_chunks,
                           sonde_frame_rec *recs, int max_recs);
long orc_demod_bits(int type, int samplerate, const float *fm, size_t n, size_t chunk,
                    uint8_t *bits, size_t bits_cap_bytes);
long orc_gfsk_soft(int samplerate, int baud, const float *fm, size_t n, size_t chunk,
                   float *soft, size_t soft_cap, float *state_out);
int  orc_gfsk_taps(int samplerate, int baud, float *taps, int cap);
void orc_gfsk_timing(int samplerate, int baud, float *out);
int  orc_rs41_correct(uint8_t *frame518);
int  orc_bch_fix(uint8_t *message64);
int  orc_rs255_fix(uint8_t *block255);
void orc_rs41_descramble(uint8_t *dst518, const uint8_t *src518);
int  orc_correlate(uint64_t syncword, int sync_len, const uint8_t *bits, int len_bytes, int *inverted);
unsigned orc_crc16_ccitt_false(const uint8_t *p, size_t n);
unsigned orc_crc16_aug_ccitt(const uint8_t *p, size_t n);
unsigned orc_crc16_modbus(const uint8_t *p, size_t n);
unsigned orc_fcs16(const uint8_t *p, size_t n);

/* FM discriminator (restated upstream semantics, deterministic fp32) */
void orc_discriminate(const float *iq /*[n][2]*/, size_t n, float gain, float *prev_phase, float *out);
/* glibc-atan2f variant, only used to REPORT how far an upstream-like discriminator drifts */
void orc_discriminate_libm(const float *iq, size_t n, float gain, float *prev_phase, float *out);

/* IQ front end + frames */
int  orc_frames_run_iq(int type, int samplerate, const float *iq, size_t n, size_t chunk, float gain,
                       sonde_frame_rec *recs, int max_recs);

/*
 * One channel as a stream: the loops of orc_frames_run / orc_frames_run_iq, one buffer per call (any length).  Records
 * of the frames that complete in a call are returned with chunk = call_index; the return value is their number (which
 * may exceed max_recs: the surplus is decoded and dropped).  gain 0 -> 2/pi as everywhere.
 */
typedef struct orc_chan orc_chan;
orc_chan *orc_chan_open(int type, int samplerate, float gain);
int  orc_chan_push_fm(orc_chan *s, const float *fm, size_t len, int call_index, sonde_frame_rec *recs, int max_recs);
int  orc_chan_push_iq(orc_chan *s, const float *iq /*[len][2]*/, size_t len, int call_index, sonde_frame_rec *recs, int max_recs);
void orc_chan_close(orc_chan *s);

/*
 * Multi-threaded batch runner for the CPU baseline: channels statically partitioned
 * over nthreads.  in = [C][n] float FM (is_iq=0) or [C][n][2] float IQ (is_iq=1).
 * frames_out/ok_out (may be NULL) receive per-channel counts.  Returns 0.
 */
int  orc_batch_run(const int32_t *types, int n_channels, int samplerate, const float *in, int is_iq,
                   size_t n, size_t chunk, float gain, int nthreads,
                   int32_t *frames_out, int32_t *ok_out);

#ifdef __cplusplus
}
#endif
#endif
