/*
 * stub_sonde_b200.c — TEST INFRASTRUCTURE ONLY: the handful of batch-ABI entry points the host layer calls
 * (sonde_batch.cpp, host/gpu_decoder.hpp), served by the CPU oracle (oracle/_build/libsonde_oracle.so, its streaming
 * form orc_chan_*), so that the HOST logic — reading recordings, padding the last buffer, the two-deep submit / fetch
 * pipeline, backlogs of streams that deliver unequal lengths, fragment aggregation, the CSV / GPX / KML writers — can be
 * checked against the reference where there is no GPU (tests/test_batch_host_logic.py).
 *
 * Built by that test into build/stub/libbatch_abi_stub.so and linked ONLY into test copies of the host programs under
 * build/stub/.  The product library and the product binaries never see this file; without an sm_100 device they fail
 * with SONDE_ERR_NODEVICE / exit code 3 (tests/test_cli_dropin.py, tests/test_host_cpp.py: *_fails_loudly_without_gpu).
 *
 * Semantics kept from include/sonde_b200.h: buffers of any length up to max_chunk_len, at most two calls in flight,
 * fetch() returns the records of the oldest call not fetched yet, [C][max_frames] records + [C] counts.  SONDE_AUTO
 * channels follow the library's rule (csrc/sonde_b200.cu auto_update / report_slot): seven decoders in the order
 * RS41, M10, iMS-100, DFM, iMet, C50, MRZ-N1 run side by side and report nothing until, in some call, one of them
 * delivers a frame that passes its FEC / checksum gate; the first such decoder in that order is locked from that call
 * on (its records of that call included) and the others stop.
 */
#include <stdlib.h>
#include <string.h>

#include "sonde_b200.h"
#include "../../oracle/sonde_oracle.h"

#define STUB_MAX_FRAMES 64

static const int32_t kAutoOrder[SONDE_NTYPES] = {SONDE_RS41, SONDE_M10, SONDE_IMS100, SONDE_DFM09, SONDE_IMET4, SONDE_C50, SONDE_MRZN1};

struct sonde_b200 {
	int C;
	size_t max_len;
	int32_t *types;                 /* as configured */
	int32_t *locked;                /* SONDE_AUTO until an AUTO channel locks */
	orc_chan **chan;                /* [C][SONDE_NTYPES]: slot 0 for a typed channel, kAutoOrder for an AUTO one */
	sonde_frame_rec *scratch;       /* [STUB_MAX_FRAMES] */
	sonde_frame_rec *recs[2];       /* [C][STUB_MAX_FRAMES] per call in flight */
	int32_t *counts[2];
	long n_calls, n_fetched;
	char err[128];
};

int sonde_b200_create(sonde_b200 **out, const sonde_b200_config *cfg)
{
	if (!out || !cfg || cfg->n_channels <= 0 || cfg->max_chunk_len <= 0) return SONDE_ERR_ARG;
	for (int c = 0; c < cfg->n_channels; c++)
		if (cfg->types[c] != SONDE_AUTO && (cfg->types[c] < 0 || cfg->types[c] >= SONDE_NTYPES)) return SONDE_ERR_ARG;
	sonde_b200 *h = calloc(1, sizeof(*h));
	h->C = cfg->n_channels;
	h->max_len = (size_t)cfg->max_chunk_len;
	h->types = malloc(sizeof(int32_t) * h->C);
	memcpy(h->types, cfg->types, sizeof(int32_t) * h->C);
	h->locked = malloc(sizeof(int32_t) * h->C);
	memcpy(h->locked, cfg->types, sizeof(int32_t) * h->C);
	h->chan = calloc((size_t)h->C * SONDE_NTYPES, sizeof(*h->chan));
	h->scratch = malloc(sizeof(sonde_frame_rec) * STUB_MAX_FRAMES);
	for (int c = 0; c < h->C; c++) {
		if (cfg->types[c] != SONDE_AUTO) { h->chan[c * SONDE_NTYPES] = orc_chan_open(cfg->types[c], cfg->samplerate, cfg->fm_gain); continue; }
		for (int k = 0; k < SONDE_NTYPES; k++) h->chan[c * SONDE_NTYPES + k] = orc_chan_open(kAutoOrder[k], cfg->samplerate, cfg->fm_gain);
	}
	for (int k = 0; k < 2; k++) {
		h->recs[k] = malloc(sizeof(sonde_frame_rec) * STUB_MAX_FRAMES * h->C);
		h->counts[k] = calloc(h->C, sizeof(int32_t));
	}
	*out = h;
	return SONDE_OK;
}

void sonde_b200_destroy(sonde_b200 *h)
{
	if (!h) return;
	for (int c = 0; c < h->C * SONDE_NTYPES; c++) orc_chan_close(h->chan[c]);
	for (int k = 0; k < 2; k++) { free(h->recs[k]); free(h->counts[k]); }
	free(h->chan);
	free(h->scratch);
	free(h->locked);
	free(h->types);
	free(h);
}

static int process(sonde_b200 *h, const float *in, size_t len, int is_iq, size_t row_stride)
{
	if (len > h->max_len) { strcpy(h->err, "stand-in: len > max_chunk_len"); return SONDE_ERR_TOOLONG; }
	if (h->n_calls - h->n_fetched >= 2) { strcpy(h->err, "stand-in: more than two calls in flight"); return SONDE_ERR_STATE; }
	const int slot = (int)(h->n_calls & 1);
	for (int c = 0; c < h->C; c++) {
		sonde_frame_rec *dst = h->recs[slot] + (size_t)c * STUB_MAX_FRAMES;
		const float *row = in + (is_iq ? 2 : 1) * (size_t)c * row_stride;
		h->counts[slot][c] = 0;
		for (int k = 0; k < SONDE_NTYPES; k++) {
			orc_chan *ch = h->chan[c * SONDE_NTYPES + k];
			if (!ch) continue;
			const int unlocked = h->types[c] == SONDE_AUTO && h->locked[c] == SONDE_AUTO;
			if (h->types[c] == SONDE_AUTO && !unlocked && kAutoOrder[k] != h->locked[c]) continue;       /* a loser: stopped */
			sonde_frame_rec *out = unlocked ? h->scratch : dst;
			int n = is_iq ? orc_chan_push_iq(ch, row, len, (int)h->n_calls, out, STUB_MAX_FRAMES)
			              : orc_chan_push_fm(ch, row, len, (int)h->n_calls, out, STUB_MAX_FRAMES);
			if (n < 0) { strcpy(h->err, "stand-in: oracle failure"); return SONDE_ERR_CUDA; }
			if (n > STUB_MAX_FRAMES) n = STUB_MAX_FRAMES;
			if (!unlocked) { h->counts[slot][c] = n; continue; }
			/* still undetermined: every decoder sees the buffer; the first in order with a good frame takes the channel */
			int ok = 0;
			for (int i = 0; i < n; i++) ok += out[i].ok;
			if (ok > 0 && h->counts[slot][c] == 0 && h->locked[c] == SONDE_AUTO) {
				h->locked[c] = kAutoOrder[k];
				memcpy(dst, out, sizeof(sonde_frame_rec) * n);
				h->counts[slot][c] = n;
			}
		}
	}
	h->n_calls++;
	return SONDE_OK;
}

int sonde_b200_process_fm(sonde_b200 *h, const float *fm, size_t len) { return process(h, fm, len, 0, len); }
int sonde_b200_process_iq(sonde_b200 *h, const float *iq, size_t len) { return process(h, iq, len, 1, len); }
/* "device" memory is host memory here (stub_sonde_chan.c hands out host pointers) */
int sonde_b200_process_iq_device(sonde_b200 *h, const void *d_iq, size_t len, size_t row_stride) { return process(h, d_iq, len, 1, row_stride); }
void *sonde_b200_stream(sonde_b200 *h) { (void)h; return NULL; }

int sonde_b200_max_frames(const sonde_b200 *h) { (void)h; return STUB_MAX_FRAMES; }

int sonde_b200_fetch(sonde_b200 *h, sonde_frame_rec *recs, int32_t *counts)
{
	if (h->n_fetched >= h->n_calls) { strcpy(h->err, "stand-in: nothing to fetch"); return SONDE_ERR_STATE; }
	const int slot = (int)(h->n_fetched++ & 1);
	memcpy(counts, h->counts[slot], sizeof(int32_t) * h->C);
	for (int c = 0; c < h->C; c++)
		memcpy(recs + (size_t)c * STUB_MAX_FRAMES, h->recs[slot] + (size_t)c * STUB_MAX_FRAMES, sizeof(sonde_frame_rec) * counts[c]);
	return SONDE_OK;
}

int sonde_b200_detected_types(sonde_b200 *h, int32_t *types)
{
	memcpy(types, h->locked, sizeof(int32_t) * h->C);
	return SONDE_OK;
}

void *sonde_b200_host_alloc(size_t bytes) { return malloc(bytes); }
void sonde_b200_host_free(void *p) { free(p); }
const char *sonde_b200_last_error(const sonde_b200 *h) { return h ? h->err : ""; }
