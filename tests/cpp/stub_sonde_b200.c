/*
 * stub_sonde_b200.c — TEST INFRASTRUCTURE ONLY: the handful of batch-ABI entry points sonde_batch.cpp calls, served by
 * the CPU oracle (oracle/_build/libsonde_oracle.so), so that the HOST logic of the batch runner — reading recordings,
 * padding the last buffer, the two-deep submit / fetch pipeline, fragment aggregation, the CSV / GPX / KML writers — can
 * be checked against the reference's command-line tool where there is no GPU (tests/test_batch_host_logic.py).
 *
 * Built by that test into build/stub/libbatch_abi_stub.so and linked ONLY into a test copy of the runner
 * (build/stub/sonde_b200_batch_stub).  The product library and the product runner never see this file; the product
 * runner exits with code 3 without an sm_100 device (tests/test_cli_dropin.py::test_batch_runner_fails_loudly_without_gpu).
 *
 * Streaming on top of the oracle's whole-recording call: every process_fm() appends its buffer to the channel's
 * recording; fetch() for call k re-decodes the first (k+1) buffers with the same chunking and hands out the records
 * whose `chunk` is k.  Quadratic, which is fine for recordings of a few seconds.
 */
#include <stdlib.h>
#include <string.h>

#include "sonde_b200.h"
#include "../../oracle/sonde_oracle.h"

#define STUB_MAX_FRAMES 64

struct sonde_b200 {
	int C, samplerate;
	size_t len;               /* buffer length: every call must use the same */
	int32_t *types;
	float **rec;              /* [C] growing recordings */
	size_t n_calls, n_fetched, cap_calls;
	char err[128];
};

int sonde_b200_create(sonde_b200 **out, const sonde_b200_config *cfg)
{
	if (!out || !cfg || cfg->n_channels <= 0) return SONDE_ERR_ARG;
	for (int c = 0; c < cfg->n_channels; c++)
		if (cfg->types[c] < 0 || cfg->types[c] >= SONDE_NTYPES) return SONDE_ERR_ARG;      /* no AUTO in the stub */
	sonde_b200 *h = calloc(1, sizeof(*h));
	h->C = cfg->n_channels;
	h->samplerate = cfg->samplerate;
	h->len = (size_t)cfg->max_chunk_len;
	h->types = malloc(sizeof(int32_t) * h->C);
	memcpy(h->types, cfg->types, sizeof(int32_t) * h->C);
	h->rec = calloc(h->C, sizeof(float *));
	*out = h;
	return SONDE_OK;
}

void sonde_b200_destroy(sonde_b200 *h)
{
	if (!h) return;
	for (int c = 0; c < h->C; c++) free(h->rec[c]);
	free(h->rec);
	free(h->types);
	free(h);
}

int sonde_b200_process_fm(sonde_b200 *h, const float *fm, size_t len)
{
	if (len != h->len) { strcpy(h->err, "stub: every call must pass max_chunk_len samples"); return SONDE_ERR_ARG; }
	if (h->n_calls - h->n_fetched >= 2) { strcpy(h->err, "stub: more than two calls in flight"); return SONDE_ERR_STATE; }
	if (h->n_calls == h->cap_calls) {
		h->cap_calls = h->cap_calls ? 2 * h->cap_calls : 64;
		for (int c = 0; c < h->C; c++) h->rec[c] = realloc(h->rec[c], h->cap_calls * len * sizeof(float));
	}
	for (int c = 0; c < h->C; c++) memcpy(h->rec[c] + h->n_calls * len, fm + (size_t)c * len, len * sizeof(float));
	h->n_calls++;
	return SONDE_OK;
}

int sonde_b200_process_iq(sonde_b200 *h, const float *iq, size_t len)
{
	(void)iq; (void)len;
	strcpy(h->err, "stub: FM input only");
	return SONDE_ERR_ARG;
}

int sonde_b200_max_frames(const sonde_b200 *h) { (void)h; return STUB_MAX_FRAMES; }

int sonde_b200_fetch(sonde_b200 *h, sonde_frame_rec *recs, int32_t *counts)
{
	if (h->n_fetched >= h->n_calls) { strcpy(h->err, "stub: nothing to fetch"); return SONDE_ERR_STATE; }
	const size_t k = h->n_fetched++;
	const int cap = 4096;
	sonde_frame_rec *all = malloc(sizeof(sonde_frame_rec) * cap);
	for (int c = 0; c < h->C; c++) {
		int n = orc_frames_run(h->types[c], h->samplerate, h->rec[c], (k + 1) * h->len, h->len, all, cap);
		if (n > cap) n = cap;
		counts[c] = 0;
		for (int i = 0; i < n; i++)
			if ((size_t)all[i].chunk == k && counts[c] < STUB_MAX_FRAMES) recs[(size_t)c * STUB_MAX_FRAMES + counts[c]++] = all[i];
	}
	free(all);
	return SONDE_OK;
}

int sonde_b200_detected_types(sonde_b200 *h, int32_t *types)
{
	memcpy(types, h->types, sizeof(int32_t) * h->C);
	return SONDE_OK;
}

void *sonde_b200_host_alloc(size_t bytes) { return malloc(bytes); }
void sonde_b200_host_free(void *p) { free(p); }
const char *sonde_b200_last_error(const sonde_b200 *h) { return h ? h->err : ""; }
