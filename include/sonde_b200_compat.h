/*
 * sonde_b200_compat.h — the seven single-channel entry points of the reference's C ABI
 * (SD/include/{rs41,dfm09,m10,ims100,mrzn1,imet4,c50}.h), served by the GPU path.
 *
 * Same names, same signatures, same call protocol (SD/include/rs41.h:14-34):
 *   X_decoder_init(samplerate) -> opaque decoder (NULL on failure — there is no CPU fallback)
 *   X_decode(dec, dst, src, len): call repeatedly with the SAME (src, len) until it returns PROCEED;
 *       every PARSED return is one framer window and *dst is valid (fields == 0: nothing decodable)
 *   X_decoder_deinit(dec)
 * `src` is float, already FM-demodulated, like the reference.  These exist so that code written against
 * libradiosonde links unchanged (libsonde_b200_compat.so); batch users should use sonde_b200.h.
 * X_last_frame(dec) additionally exposes the frame record behind the last PARSED return.
 */
#ifndef SONDE_B200_COMPAT_H
#define SONDE_B200_COMPAT_H
#include <stddef.h>
#include "sonde_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef sondedump_data_h      /* skipped when the reference's own SD/include/data.h was included first */
#define sondedump_data_h
#include <time.h>
typedef enum { PROCEED, PARSED } ParserStatus;
typedef enum {
	DATA_SEQ = 1 << 0, DATA_SERIAL = 1 << 1, DATA_POS = 1 << 2, DATA_SPEED = 1 << 3,
	DATA_TIME = 1 << 4, DATA_PTU = 1 << 5, DATA_OZONE = 1 << 6, DATA_SHUTDOWN = 1 << 7
} DataBitmask;
typedef struct {
	int fields;          /* DataBitmask bits (an int here so that C++ callers can OR into it) */
	int seq;
	char serial[32];
	float lat, lon, alt, speed, climb, heading;
	time_t time;
	float calib_percent, temp, rh, pressure, o3_mpa;
	int shutdown;
} SondeData;
#endif

#define SONDE_COMPAT_DECL(X, T)                                                                   \
	typedef struct sonde_compat_decoder T;                                                        \
	SONDE_API T *X##_decoder_init(int samplerate);                                                \
	SONDE_API void X##_decoder_deinit(T *d);                                                      \
	SONDE_API ParserStatus X##_decode(T *d, SondeData *dst, const float *src, size_t len);        \
	SONDE_API const sonde_frame_rec *X##_last_frame(const T *d);

SONDE_COMPAT_DECL(rs41, RS41Decoder)        /* SD/include/rs41.h:14,21,34   */
SONDE_COMPAT_DECL(dfm09, DFM09Decoder)      /* SD/include/dfm09.h           */
SONDE_COMPAT_DECL(m10, M10Decoder)          /* SD/include/m10.h             */
SONDE_COMPAT_DECL(ims100, IMS100Decoder)    /* SD/include/ims100.h          */
SONDE_COMPAT_DECL(mrzn1, MRZN1Decoder)      /* SD/include/mrzn1.h           */
SONDE_COMPAT_DECL(imet4, IMET4Decoder)      /* SD/include/imet4.h           */
SONDE_COMPAT_DECL(c50, C50Decoder)          /* SD/include/c50.h             */

/* Telemetry parsers alone (host only, no GPU): frame record -> SondeData with the per-channel state the
 * reference keeps in its decoder structs (SD/sonde/<type>/<type>.c, parser.c). */
SONDE_API void *sonde_telemetry_create(int type);
SONDE_API void  sonde_telemetry_destroy(void *t);
SONDE_API int   sonde_telemetry_parse(void *t, const sonde_frame_rec *rec, SondeData *out);

#ifdef __cplusplus
}
#endif
#endif
