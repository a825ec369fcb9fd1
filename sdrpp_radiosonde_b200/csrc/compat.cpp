/*
 * compat.cpp — the reference's seven xxx_decoder_init / xxx_decode / xxx_decoder_deinit entry points
 * (SD/include/*.h shape) as a thin adapter over the batch C ABI with C = 1.
 *
 * Call protocol (SD/include/rs41.h:23-33, src/decode/decoder.hpp:61): the caller re-invokes xxx_decode
 * with the same buffer until PROCEED.  The first call for a buffer runs the GPU path over the buffer — over its first
 * kMaxChunk samples if it is longer — and queues the frame records; each further call pops one (PARSED); when the
 * queue is empty the next piece of the buffer is decoded, and the call after the last record of the last piece answers
 * PROCEED and re-arms.  Buffers of any length are decoded completely, as by the reference.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/sonde_b200_compat.h"
#include "../host/gpu_decoder.hpp"
#include "../host/telemetry.hpp"

struct sonde_compat_decoder {
	sonde_b200 *h = nullptr;
	int type = 0, max_chunk = 0, max_frames = 0;
	std::vector<sonde_frame_rec> recs;
	int n_recs = 0, next = 0;
	bool armed = false;              /* the caller is working through a buffer: `done` of its samples have been decoded */
	size_t done = 0;
	const sonde_frame_rec *last = nullptr;
	radiosonde::Telemetry tele;
};

namespace {

/* Samples per GPU call.  A caller's buffer may be any length (the plugin's dsp::stream hands over up to 1,000,000
 * samples, the tool 1024): longer buffers are decoded in pieces of this size while the caller repeats its
 * xxx_decode(d, dst, src, len) call, which is how the reference works through a buffer too (SD/sonde/rs41/rs41.c:107-125
 * returns PARSED from the middle of `src` and continues there on the next call). */
constexpr int kMaxChunk = 1 << 16;

sonde_compat_decoder *make(int type, int samplerate)
{
	sonde_compat_decoder *d = new (std::nothrow) sonde_compat_decoder();
	if (!d) return nullptr;
	const int32_t t = type;
	sonde_b200_config cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.n_channels = 1;
	cfg.samplerate = samplerate;
	cfg.max_chunk_len = kMaxChunk;
	cfg.types = &t;
	const int rc = sonde_b200_create(&d->h, &cfg);
	if (rc != SONDE_OK) {
		/* the reference's init has no error channel besides NULL: say why, loudly */
		fprintf(stderr, "sonde_b200_compat: decoder init failed (%d): an sm_100 CUDA device is required, there is no CPU fallback\n", rc);
		delete d;
		return nullptr;
	}
	d->type = type;
	d->tele.reset(type);
	d->max_chunk = cfg.max_chunk_len;
	d->max_frames = sonde_b200_max_frames(d->h);
	d->recs.resize(d->max_frames);
	return d;
}

void destroy(sonde_compat_decoder *d)
{
	if (!d) return;
	sonde_b200_destroy(d->h);
	delete d;
}

ParserStatus step(sonde_compat_decoder *d, SondeData *dst, const float *src, size_t len)
{
	if (!d) {
		fprintf(stderr, "sonde_b200_compat: xxx_decode() called with a NULL decoder (init failed: no CUDA device, no CPU fallback)\n");
		abort();
	}
	if (!src) return PROCEED;
	if (!d->armed) {
		d->n_recs = d->next = 0;
		d->done = 0;
		d->armed = true;
	}
	for (;;) {
		if (d->next < d->n_recs) {
			d->last = &d->recs[d->next++];
			if (dst) {
				d->tele.parse(*d->last, dst);
			}
			return PARSED;
		}
		if (d->done >= len) break;
		/* the records of the last piece are used up: decode the next piece of this buffer */
		const size_t piece = (len - d->done < (size_t)d->max_chunk) ? len - d->done : (size_t)d->max_chunk;
		int32_t count = 0;
		d->n_recs = d->next = 0;
		const int rc = sonde_b200_process_fm(d->h, src + d->done, piece);
		d->done += piece;
		if (rc != SONDE_OK || sonde_b200_fetch(d->h, d->recs.data(), &count) != SONDE_OK) break;
		d->n_recs = count;
	}
	d->armed = false;
	return PROCEED;
}

}  // namespace

#define SONDE_COMPAT_IMPL(X, T, TYPE)                                                                   \
	extern "C" T *X##_decoder_init(int samplerate) { return make(TYPE, samplerate); }                   \
	extern "C" void X##_decoder_deinit(T *d) { destroy(d); }                                            \
	extern "C" ParserStatus X##_decode(T *d, SondeData *dst, const float *src, size_t len)              \
	{                                                                                                   \
		return step(d, dst, src, len);                                                                  \
	}                                                                                                   \
	extern "C" const sonde_frame_rec *X##_last_frame(const T *d) { return d ? d->last : nullptr; }

SONDE_COMPAT_IMPL(rs41, RS41Decoder, SONDE_RS41)
SONDE_COMPAT_IMPL(dfm09, DFM09Decoder, SONDE_DFM09)
SONDE_COMPAT_IMPL(m10, M10Decoder, SONDE_M10)
SONDE_COMPAT_IMPL(ims100, IMS100Decoder, SONDE_IMS100)
SONDE_COMPAT_IMPL(mrzn1, MRZN1Decoder, SONDE_MRZN1)
SONDE_COMPAT_IMPL(imet4, IMET4Decoder, SONDE_IMET4)
SONDE_COMPAT_IMPL(c50, C50Decoder, SONDE_C50)

/* Host-only telemetry entry points (no GPU needed): one parser state per channel, fed with frame records. */
extern "C" SONDE_API void *sonde_telemetry_create(int type) { return new (std::nothrow) radiosonde::Telemetry(type); }
extern "C" SONDE_API void sonde_telemetry_destroy(void *t) { delete static_cast<radiosonde::Telemetry *>(t); }
extern "C" SONDE_API int sonde_telemetry_parse(void *t, const sonde_frame_rec *rec, SondeData *out)
{
	if (!t || !rec || !out) return SONDE_ERR_ARG;
	static_cast<radiosonde::Telemetry *>(t)->parse(*rec, out);
	return SONDE_OK;
}
