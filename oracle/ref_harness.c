/*
 * ref_harness.c — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Thin driver around the UNMODIFIED reference library (sondedump, compiled from
 * /root/reference/src/decode/sondedump by oracle/Makefile into oracle/_ref/).
 * Nothing here re-implements the algorithm: every stage below is a call into the
 * reference's own non-static functions.  The glue mirrors the pre-parser half of
 * each xxx_decode():
 *     rs41.c:126-143   dfm09.c:69-108   m10.c:52-73   ims100.c:81-106
 *     mrzn1.c:62-75    imet4.c:65-117   c50.c:57-81
 * because the decoder structs are private to those .c files, and the frame bytes
 * (the parity criterion) are not reachable through the public API.  ref_decode_run()
 * additionally drives the real public xxx_decode() API so tests can cross-check
 * the glue (same number of PARSED returns per buffer).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load the resulting library.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sonde_b200.h"                 /* record layout shared with the GPU path */

/* reference headers (resolved through -I<reference>/src/decode/sondedump) */
#include "include/data.h"
#include "include/rs41.h"
#include "include/dfm09.h"
#include "include/m10.h"
#include "include/ims100.h"
#include "include/mrzn1.h"
#include "include/imet4.h"
#include "include/c50.h"
#include "bitops.h"
#include "decode/framer.h"
#include "decode/manchester.h"
#include "decode/ecc/rs.h"
#include "decode/ecc/crc.h"
#include "demod/gfsk.h"
#include "demod/afsk.h"
#include "sonde/rs41/frame.h"
#include "sonde/dfm09/frame.h"
#include "sonde/m10/frame.h"
#include "sonde/ims100/frame.h"
#include "sonde/mrz-n1/frame.h"
#include "sonde/imet4/frame.h"
#include "sonde/imet4/subframe.h"
#include "sonde/c50/frame.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct {
	int type;
	Framer f;
	RSDecoder rs;
	int has_rs;
	size_t framelen;
	/* generously sized scratch; zeroed at init so that reads past the frame are defined */
	uint8_t raw[4096];
	uint8_t work[4096];
	uint8_t work2[4096];
} RefChan;

static int
chan_init(RefChan *c, int type, int samplerate)
{
	memset(c, 0, sizeof(*c));
	c->type = type;
	switch (type) {
	case SONDE_RS41:
		framer_init_gfsk(&c->f, samplerate, RS41_BAUDRATE, RS41_FRAME_LEN, RS41_SYNCWORD, RS41_SYNC_LEN);
		rs_init(&c->rs, RS41_REEDSOLOMON_N, RS41_REEDSOLOMON_K, RS41_REEDSOLOMON_POLY,
		        RS41_REEDSOLOMON_FIRST_ROOT, RS41_REEDSOLOMON_ROOT_SKIP);
		c->has_rs = 1;
		c->framelen = RS41_FRAME_LEN;
		break;
	case SONDE_DFM09:
		framer_init_gfsk(&c->f, samplerate, DFM09_BAUDRATE, DFM09_FRAME_LEN, DFM09_SYNCWORD, DFM09_SYNC_LEN);
		c->framelen = DFM09_FRAME_LEN;
		break;
	case SONDE_M10:
		framer_init_gfsk(&c->f, samplerate, M10_BAUDRATE, M10_FRAME_LEN, M10_SYNCWORD, M10_SYNC_LEN);
		c->framelen = M10_FRAME_LEN;
		break;
	case SONDE_IMS100:
		framer_init_gfsk(&c->f, samplerate, IMS100_BAUDRATE, IMS100_FRAME_LEN, IMS100_SYNCWORD, IMS100_SYNC_LEN);
		bch_init(&c->rs, IMS100_REEDSOLOMON_N, IMS100_REEDSOLOMON_K, IMS100_REEDSOLOMON_POLY,
		         ims100_bch_roots, IMS100_REEDSOLOMON_T);
		c->has_rs = 1;
		c->framelen = IMS100_FRAME_LEN;
		break;
	case SONDE_MRZN1:
		framer_init_gfsk(&c->f, samplerate, MRZN1_BAUDRATE, MRZN1_FRAME_LEN, MRZN1_SYNCWORD, MRZN1_SYNC_LEN);
		c->framelen = MRZN1_FRAME_LEN;
		break;
	case SONDE_IMET4:
		framer_init_afsk(&c->f, samplerate, IMET4_BAUDRATE, IMET4_FRAME_LEN, IMET4_MARK_FREQ, IMET4_SPACE_FREQ,
		                 IMET4_SYNCWORD, IMET4_SYNC_LEN);
		c->framelen = IMET4_FRAME_LEN;
		break;
	case SONDE_C50:
		framer_init_afsk(&c->f, samplerate, C50_BAUDRATE, C50_FRAME_LEN, C50_MARK_FREQ, C50_SPACE_FREQ,
		                 C50_SYNCWORD, C50_SYNC_LEN);
		c->framelen = C50_FRAME_LEN;
		break;
	default:
		return -1;
	}
	return 0;
}

static void
chan_deinit(RefChan *c)
{
	framer_deinit(&c->f);
	if (c->has_rs) rs_deinit(&c->rs);
}

/* Post-framer half of xxx_decode(), up to (not including) telemetry parsing. */
static void
chan_deframe(RefChan *c, sonde_frame_rec *r)
{
	size_t i;
	int n, errcount, sflen, nonzero;

	memset(r->data, 0, sizeof(r->data));
	memcpy(r->raw, c->raw, (c->framelen + 7) / 8);
	r->aux = 0;

	switch (c->type) {
	case SONDE_RS41: {
		RS41Frame *fr = (RS41Frame*)c->work;
		rs41_frame_descramble(fr, (RS41Frame*)c->raw);
		r->status = rs41_frame_correct(fr, &c->rs);
		r->ok = r->status >= 0;
		r->data_len = sizeof(RS41Frame);
		memcpy(r->data, fr, sizeof(RS41Frame));
		break;
	}
	case SONDE_DFM09: {
		DFM09ECCFrame *fr = (DFM09ECCFrame*)c->work;
		DFM09Frame *un = (DFM09Frame*)c->work2;
		manchester_decode(c->work, c->raw, DFM09_FRAME_LEN);
		dfm09_frame_deinterleave(fr);
		errcount = dfm09_frame_correct(fr);
		r->status = errcount;
		r->data_len = sizeof(DFM09ECCFrame);
		memcpy(r->data, fr, sizeof(DFM09ECCFrame));
		r->ok = 0;
		if (!(errcount < 0 || errcount > 8)) {
			dfm09_frame_unpack(un, fr);
			nonzero = 0;
			for (i=0; i<sizeof(DFM09Frame); i++) nonzero |= ((uint8_t*)un)[i];
			memcpy(r->data + 64, un, sizeof(DFM09Frame));
			r->aux = nonzero ? 1 : 0;
			r->ok = r->aux;
		}
		break;
	}
	case SONDE_M10: {
		M10Frame *fr = (M10Frame*)c->work;
		memset(c->work, 0, 512);      /* m10_frame_correct may read up to 258 B past &len */
		manchester_decode(c->work, c->raw, M10_FRAME_LEN);
		m10_frame_descramble(fr);
		r->status = m10_frame_correct(fr);
		r->ok = r->status >= 0;
		r->data_len = sizeof(M10Frame);
		memcpy(r->data, fr, sizeof(M10Frame));
		break;
	}
	case SONDE_IMS100: {
		IMS100ECCFrame *fr = (IMS100ECCFrame*)c->work;
		IMS100Frame *un = (IMS100Frame*)c->work2;
		manchester_decode(c->work, c->raw, IMS100_FRAME_LEN);
		ims100_frame_descramble(fr);
		errcount = ims100_frame_error_correct(fr, &c->rs);
		r->status = errcount;
		r->ok = errcount >= 0;
		r->data_len = sizeof(IMS100ECCFrame);
		memcpy(r->data, fr, sizeof(IMS100ECCFrame));
		if (errcount >= 0) {
			ims100_frame_unpack(un, fr);
			memcpy(r->data + 80, un, sizeof(IMS100Frame));
			r->aux = (int)un->valid;
		}
		break;
	}
	case SONDE_MRZN1: {
		MRZN1Frame *fr = (MRZN1Frame*)c->work;
		manchester_decode(c->work, c->raw, MRZN1_FRAME_LEN);
		r->status = mrzn1_frame_correct(fr);
		r->ok = r->status >= 0;
		r->data_len = sizeof(MRZN1Frame);
		memcpy(r->data, fr, sizeof(MRZN1Frame));
		break;
	}
	case SONDE_IMET4: {
		IMET4Frame *fr = (IMET4Frame*)c->work;
		IMET4Subframe *sf;
		imet4_frame_descramble(fr, (IMET4Frame*)c->raw);
		n = 0;
		sflen = 0;
		for (i = 0; i < sizeof(fr->data); i += sflen) {
			sf = (IMET4Subframe*)&fr->data[i];
			sflen = (int)imet4_subframe_len(sf);
			if (!sflen) break;
			if (!crc16_aug_ccitt((uint8_t*)sf, sflen)) {
				n++;
			} else if (sf->type == IMET4_SFTYPE_XDATA) {
				break;
			}
		}
		if (i > 0) {
			framer_adjust(&c->f, c->raw, 10 * (offsetof(IMET4Frame, data) + i));
		}
		r->status = n;
		r->ok = n > 0;
		r->aux = (int)i;
		r->data_len = IMET4_FRAME_LEN / 10;
		memcpy(r->data, fr->data, IMET4_FRAME_LEN / 10);
		break;
	}
	case SONDE_C50: {
		C50Frame *fr = (C50Frame*)c->work;
		c50_frame_descramble(fr, (C50RawFrame*)c->raw);
		r->status = c50_frame_correct(fr);
		r->ok = r->status >= 0;
		r->data_len = sizeof(C50Frame);
		memcpy(r->data, fr, sizeof(C50Frame));
		break;
	}
	}
}

/*
 * Run one channel of `type` over fm[0..n) in buffers of `chunk` samples, exactly as
 * src/decode/decoder.hpp:59-61 / SD/main.c:330-333 do.  Returns the number of framer
 * windows (PARSED returns); at most max_recs records are stored.
 */
EXPORT int
ref_frames_run(int type, int samplerate, const float *fm, size_t n, size_t chunk,
               sonde_frame_rec *recs, int max_recs)
{
	RefChan *c = malloc(sizeof(*c));
	size_t pos, len;
	int count = 0, chunk_idx = 0;
	sonde_frame_rec tmp;

	if (!c || chan_init(c, type, samplerate)) { free(c); return -1; }

	for (pos = 0; pos < n; pos += chunk, chunk_idx++) {
		len = (n - pos < chunk) ? n - pos : chunk;
		while (framer_read(&c->f, c->raw, fm + pos, len) != PROCEED) {
			sonde_frame_rec *r = (count < max_recs) ? &recs[count] : &tmp;
			r->type = type;
			r->chunk = chunk_idx;
			r->sync_offset = c->f.sync_offset;
			r->inverted = c->f.inverted;
			r->bit_pos = 0;
			chan_deframe(c, r);
			count++;
		}
	}

	chan_deinit(c);
	free(c);
	return count;
}

/* The public API of the reference, untouched: xxx_decoder_init / xxx_decode / deinit. */
EXPORT int
ref_decode_run(int type, int samplerate, const float *fm, size_t n, size_t chunk,
               SondeData *out, int32_t *out_chunk, int max_out)
{
	void *d;
	size_t pos, len;
	int count = 0, chunk_idx = 0;
	SondeData tmp;
	ParserStatus st;

	switch (type) {
	case SONDE_RS41:   d = rs41_decoder_init(samplerate); break;
	case SONDE_DFM09:  d = dfm09_decoder_init(samplerate); break;
	case SONDE_M10:    d = m10_decoder_init(samplerate); break;
	case SONDE_IMS100: d = ims100_decoder_init(samplerate); break;
	case SONDE_MRZN1:  d = mrzn1_decoder_init(samplerate); break;
	case SONDE_IMET4:  d = imet4_decoder_init(samplerate); break;
	case SONDE_C50:    d = c50_decoder_init(samplerate); break;
	default: return -1;
	}
	if (!d) return -1;

	for (pos = 0; pos < n; pos += chunk, chunk_idx++) {
		len = (n - pos < chunk) ? n - pos : chunk;
		for (;;) {
			SondeData *dst = (count < max_out) ? &out[count] : &tmp;
			memset(dst, 0, sizeof(*dst));
			switch (type) {
			case SONDE_RS41:   st = rs41_decode(d, dst, fm + pos, len); break;
			case SONDE_DFM09:  st = dfm09_decode(d, dst, fm + pos, len); break;
			case SONDE_M10:    st = m10_decode(d, dst, fm + pos, len); break;
			case SONDE_IMS100: st = ims100_decode(d, dst, fm + pos, len); break;
			case SONDE_MRZN1:  st = mrzn1_decode(d, dst, fm + pos, len); break;
			case SONDE_IMET4:  st = imet4_decode(d, dst, fm + pos, len); break;
			default:           st = c50_decode(d, dst, fm + pos, len); break;
			}
			if (st == PROCEED) break;
			if (count < max_out && out_chunk) out_chunk[count] = chunk_idx;
			count++;
		}
	}

	switch (type) {
	case SONDE_RS41:   rs41_decoder_deinit(d); break;
	case SONDE_DFM09:  dfm09_decoder_deinit(d); break;
	case SONDE_M10:    m10_decoder_deinit(d); break;
	case SONDE_IMS100: ims100_decoder_deinit(d); break;
	case SONDE_MRZN1:  mrzn1_decoder_deinit(d); break;
	case SONDE_IMET4:  imet4_decoder_deinit(d); break;
	default:           c50_decoder_deinit(d); break;
	}
	return count;
}

/*
 * Throughput leg (cpu_baseline "reference"): just run the public decode loop and
 * count PARSED returns and frames with fields != 0; nothing is stored.
 */
EXPORT int
ref_decode_count(int type, int samplerate, const float *fm, size_t n, size_t chunk, int *n_fields)
{
	int nf = 0, total;
	SondeData *out;
	int i, cap = (int)(n / 64) + 16;

	out = malloc(sizeof(*out) * cap);
	if (!out) return -1;
	total = ref_decode_run(type, samplerate, fm, n, chunk, out, NULL, cap);
	for (i = 0; i < total && i < cap; i++) nf += out[i].fields != 0;
	free(out);
	if (n_fields) *n_fields = nf;
	return total;
}

/*
 * Free-running demodulator: the reference's gfsk_demod()/afsk_demod() asked for
 * "as many bits as the buffer gives" per chunk.  bits are
 * MSB-first in `bits`; returns the number of bits produced.
 */
EXPORT long
ref_demod_bits(int type, int samplerate, const float *fm, size_t n, size_t chunk,
               uint8_t *bits, size_t bits_cap_bytes)
{
	RefChan *c = malloc(sizeof(*c));
	size_t pos, len, bit_offset = 0;
	const size_t want = bits_cap_bytes * 8;

	if (!c || chan_init(c, type, samplerate)) { free(c); return -1; }
	memset(bits, 0, bits_cap_bytes);

	for (pos = 0; pos < n; pos += chunk) {
		len = (n - pos < chunk) ? n - pos : chunk;
		if (c->f.type == GFSK)
			gfsk_demod(&c->f.demod.gfsk, bits, &bit_offset, want, fm + pos, len);
		else
			afsk_demod(&c->f.demod.afsk, bits, &bit_offset, want, fm + pos, len);
	}
	chan_deinit(c);
	free(c);
	return (long)bit_offset;
}

/*
 * Soft symbols at the symbol instants of a GFSK chain, obtained by driving the
 * reference's own public primitives in the call order of gfsk_demod()
 * (SD/demod/gfsk.c:75-125); interm is cleared at every chunk start like gfsk.c:73.
 * Also returns the final loop state for state-parity checks.
 * state_out[0..5] = agc.bias, agc.moving_avg, timing.phase, timing.freq, timing.prev, timing.state
 */
EXPORT long
ref_gfsk_soft(int samplerate, int baud, const float *fm, size_t n, size_t chunk,
              float *soft, size_t soft_cap, float *state_out)
{
	GFSKDemod g;
	size_t pos, i, len, count = 0;
	int phase;
	float s, interm, sym;

	if (gfsk_init(&g, samplerate, baud)) return -1;
	for (pos = 0; pos < n; pos += chunk) {
		len = (n - pos < chunk) ? n - pos : chunk;
		interm = 0;
		for (i = 0; i < len; i++) {
			s = agc_apply(&g.agc, fm[pos + i]);
			filter_fwd_sample(&g.lpf, s);
			for (phase = 0; phase < g.lpf.num_phases; phase++) {
				switch (advance_timeslot(&g.timing)) {
				case 1:
					interm = filter_get(&g.lpf, phase);
					break;
				case 2:
					sym = filter_get(&g.lpf, phase);
					retime(&g.timing, interm, sym);
					if (count < soft_cap) soft[count] = sym;
					count++;
					break;
				default:
					break;
				}
			}
		}
	}
	if (state_out) {
		state_out[0] = g.agc.bias;
		state_out[1] = g.agc.moving_avg;
		state_out[2] = g.timing.phase;
		state_out[3] = g.timing.freq;
		state_out[4] = g.timing.prev;
		state_out[5] = (float)g.timing.state;
	}
	gfsk_deinit(&g);
	return (long)count;
}

/*
 * Soft symbols of the reference's AFSK chain (SD/demod/afsk.c:104-151): its own AFSKDemod state and its own
 * agc_apply / filter_fwd_sample / filter_get / advance_timeslot / retime, with the per-sample mixer + boxcar statements
 * between them written out exactly as afsk.c has them (same expression types, this container's libm cexpf / cabsf /
 * fmod), because afsk_demod() does not export the symbol values.  The mirror is validated in tests/test_oracle.py: its
 * hard decisions equal afsk_demod()'s bits.  One entry = one afsk_demod() call sequence over a buffer of `chunk`
 * samples (interm = 0 at every buffer start, afsk.c:94).
 * type: SONDE_IMET4 or SONDE_C50 (mark / space / baud of the protocol headers).
 */
EXPORT long
ref_afsk_soft(int type, int samplerate, const float *fm, size_t n, size_t chunk, float *soft, size_t soft_cap,
              float *state_out)
{
	AFSKDemod d;
	size_t pos, i, len, count = 0;
	int phase, rc;
	float symbol, interm, sym;
	float p_mark, p_space;
	float complex out, mark_sum, space_sum;

	if (type == 5) rc = afsk_init(&d, samplerate, IMET4_BAUDRATE, IMET4_MARK_FREQ, IMET4_SPACE_FREQ);
	else if (type == 6) rc = afsk_init(&d, samplerate, C50_BAUDRATE, C50_MARK_FREQ, C50_SPACE_FREQ);
	else return -1;
	if (rc) return -1;
	p_mark = d.p_mark; p_space = d.p_space;
	mark_sum = d.mark_sum; space_sum = d.space_sum;
	for (pos = 0; pos < n; pos += chunk) {
		len = (n - pos < chunk) ? n - pos : chunk;
		interm = 0;
		for (i = 0; i < len; i++) {
			symbol = fm[pos + i];
			symbol = agc_apply(&d.agc, symbol) / d.len * 2;

			out = symbol * cexpf(-I * p_mark);
			mark_sum += out - d.mark_history[d.idx];
			d.mark_history[d.idx] = out;

			out = symbol * cexpf(-I * p_space);
			space_sum += out - d.space_history[d.idx];
			d.space_history[d.idx] = out;

			symbol = cabsf(mark_sum) - cabsf(space_sum);
			d.idx = (d.idx + 1) % d.len;
			p_mark = fmod(p_mark + d.f_mark, 2*M_PI);
			p_space = fmod(p_space + d.f_space, 2*M_PI);
			filter_fwd_sample(&d.lpf, symbol);

			for (phase = 0; phase < d.lpf.num_phases; phase++) {
				switch (advance_timeslot(&d.timing)) {
				case 1:
					interm = filter_get(&d.lpf, phase);
					break;
				case 2:
					sym = filter_get(&d.lpf, phase);
					retime(&d.timing, interm, sym);
					if (count < soft_cap) soft[count] = sym;
					count++;
					break;
				default:
					break;
				}
			}
		}
	}
	if (state_out) {
		state_out[0] = d.agc.bias;
		state_out[1] = d.agc.moving_avg;
		state_out[2] = d.timing.phase;
		state_out[3] = d.timing.freq;
		state_out[4] = d.timing.prev;
		state_out[5] = (float)d.timing.state;
	}
	afsk_deinit(&d);
	return (long)count;
}

/* FIR taps exactly as the reference computes them (SD/demod/dsp/filter.c:10-32). */
EXPORT int
ref_gfsk_taps(int samplerate, int baud, float *taps, int cap)
{
	GFSKDemod g;
	int n, i;
	if (gfsk_init(&g, samplerate, baud)) return -1;
	n = g.lpf.size * g.lpf.num_phases;
	for (i = 0; i < n && i < cap; i++) taps[i] = g.lpf.coeffs[i];
	gfsk_deinit(&g);
	return n;
}

/* Timing-loop constants as the reference computes them (SD/demod/dsp/timing.c:14-25,79-87). */
EXPORT void
ref_gfsk_timing(int samplerate, int baud, float *out /*[5]: freq, alpha, beta, max_fdev, num_phases*/)
{
	GFSKDemod g;
	gfsk_init(&g, samplerate, baud);
	out[0] = g.timing.freq;
	out[1] = g.timing.alpha;
	out[2] = g.timing.beta;
	out[3] = g.timing.max_fdev;
	out[4] = (float)g.lpf.num_phases;
	gfsk_deinit(&g);
}

/* FEC known-answer entry points. */
EXPORT int
ref_rs41_correct(uint8_t *frame518)
{
	RSDecoder rs;
	int ret;
	rs_init(&rs, RS41_REEDSOLOMON_N, RS41_REEDSOLOMON_K, RS41_REEDSOLOMON_POLY,
	        RS41_REEDSOLOMON_FIRST_ROOT, RS41_REEDSOLOMON_ROOT_SKIP);
	ret = rs41_frame_correct((RS41Frame*)frame518, &rs);
	rs_deinit(&rs);
	return ret;
}

EXPORT int
ref_bch_fix(uint8_t *message64)
{
	RSDecoder rs;
	int ret;
	bch_init(&rs, IMS100_REEDSOLOMON_N, IMS100_REEDSOLOMON_K, IMS100_REEDSOLOMON_POLY,
	         ims100_bch_roots, IMS100_REEDSOLOMON_T);
	ret = rs_fix_block(&rs, message64);
	rs_deinit(&rs);
	return ret;
}

EXPORT int
ref_rs255_fix(uint8_t *block255)
{
	RSDecoder rs;
	int ret;
	rs_init(&rs, RS41_REEDSOLOMON_N, RS41_REEDSOLOMON_K, RS41_REEDSOLOMON_POLY,
	        RS41_REEDSOLOMON_FIRST_ROOT, RS41_REEDSOLOMON_ROOT_SKIP);
	ret = rs_fix_block(&rs, block255);
	rs_deinit(&rs);
	return ret;
}

EXPORT void
ref_rs41_descramble(uint8_t *dst518, const uint8_t *src518)
{
	rs41_frame_descramble((RS41Frame*)dst518, (RS41Frame*)src518);
}

EXPORT int
ref_correlate(uint64_t syncword, int sync_len, const uint8_t *bits, int len_bytes, int *inverted)
{
	Correlator c;
	correlator_init(&c, syncword, sync_len);
	return correlate(&c, inverted, bits, len_bytes);
}

EXPORT unsigned ref_crc16_ccitt_false(const uint8_t *p, size_t n) { return crc16_ccitt_false(p, n); }
EXPORT unsigned ref_crc16_aug_ccitt(const uint8_t *p, size_t n)   { return crc16_aug_ccitt(p, n); }
EXPORT unsigned ref_crc16_modbus(const uint8_t *p, size_t n)      { return crc16_modbus(p, n); }
EXPORT unsigned ref_fcs16(const uint8_t *p, size_t n)             { return fcs16(p, n); }
EXPORT int      ref_sizeof_sondedata(void)                        { return (int)sizeof(SondeData); }
