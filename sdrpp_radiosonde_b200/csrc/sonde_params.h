/*
 * sonde_params.h — per-sonde modem / frame constants and the derived demodulator
 * constants, shared by the host side and the kernels of libsonde_b200.
 *
 * Constants are the ones of the reference's protocol headers (paths relative to
 * /root/reference/src/decode/sondedump):
 *   sonde/rs41/protocol.h:8-19      sonde/dfm09/protocol.h:9-15   sonde/m10/protocol.h:11-14
 *   sonde/ims100/protocol.h:11-23   sonde/mrz-n1/protocol.h:8-11  sonde/imet4/protocol.h:8-13
 *   sonde/c50/protocol.h:8-13
 * Derived values follow demod/gfsk.c:17-35, demod/afsk.c:16-48, demod/dsp/timing.c:14-25,79-87
 * and demod/dsp/filter.c:10-32,67-93 (see modem_tables.cpp).
 */
#ifndef SONDE_PARAMS_H
#define SONDE_PARAMS_H

#include <stdint.h>

#define SONDE_NTYPES_       7
#define SONDE_FIR_TAPS      49      /* demod/gfsk.h:11 : order 24 -> 2*24+1 taps            */
#define SONDE_FIR_HIST      48      /* samples of history a 49-tap output needs              */
#define SONDE_MAX_PHASES    2       /* M10 is the only 2-phase modem at 48 kS/s             */
#define SONDE_AFSK_MAXLEN   64      /* boxcar length bound (39 for iMet-4, 20 for C50)      */

/* Everything the kernels need to know about one decoder type at one sample rate. */
typedef struct {
	int32_t  type;
	int32_t  afsk;            /* 0 = GFSK chain, 1 = AFSK chain                              */
	int32_t  baud;
	int32_t  frame_bits;      /* F  : framer frame length in raw bits                        */
	int32_t  sync_len;        /* S  : sync word length in bits                               */
	int32_t  num_phases;      /* P  : polyphase branches (gfsk.c:20)                         */
	int32_t  boxcar_len;      /* AFSK boxcar length (afsk.c:41), 0 for GFSK                  */
	int32_t  data_len;        /* bytes of the post-FEC frame stored in sonde_frame_rec.data  */
	uint64_t syncword;
	/* timing loop (timing.c:14-25,79-87) */
	float    freq0;           /* initial / centre NCO increment                              */
	float    alpha, beta, max_fdev;
	/* AFSK mixers (afsk.c:22-23) */
	float    f_mark, f_space; /* radians / sample                                            */
	/* FIR taps, coeffs[phase*49 + i] exactly as filter_init_lpf lays them out            */
	float    taps[SONDE_MAX_PHASES * SONDE_FIR_TAPS];
} sonde_modem;

#ifdef __cplusplus
extern "C" {
#endif
/* Fills `m` for (type, samplerate).  Returns 0, or -1 for an unsupported combination. */
int sonde_modem_init(sonde_modem *m, int type, int samplerate);
#ifdef __cplusplus
}
#endif

#endif
