/*
 * stub_sonde_chan.c — TEST INFRASTRUCTURE ONLY: the channelizer entry points radiosonde::GpuWidebandBank calls
 * (include/sonde_b200_channelizer.h) plus the two device-side entry points of the decoder it pairs them with, as plain
 * C on host memory, so that the bank's host logic — staging, the n mod D carry between buffers of odd lengths, the
 * hand-over of [C][stride] channel rows to the decoder — can be checked where there is no GPU
 * (tests/test_batch_host_logic.py).  Linked with stub_sonde_b200.c into build/stub/libbatch_abi_stub.so; the product
 * never sees it.
 *
 * The formula is the header's (L = 1):  y_c[m] = sum_n g[m*D + D-1 - n] * x[n] * exp(-j w_c n), filter history and
 * oscillator phase carried across calls; g is a Hamming-windowed sinc of taps_per_phase * D taps with the configured
 * -6 dB point.  It is a stand-in, not a model of the tensor-core kernel's arithmetic (tests/test_channelizer.py and
 * oracle/channelizer_oracle.py are that, on the GPU): "device" pointers are host pointers here.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sonde_b200.h"
#include "sonde_b200_channelizer.h"

struct sonde_chan {
	int C, D, K;               /* K = taps in total */
	double fs_in;
	double *g;                 /* [C][K] */
	double *w;                 /* [C] rad / sample */
	float *hist;               /* [K - 1][2] last input samples */
	unsigned long long n_seen;
	float *out[2];             /* double-buffered [C][stride][2] */
	size_t stride;
	int flip;
	char err[96];
};

int sonde_chan_create_ex(sonde_chan **out, const sonde_chan_config *cfg, const sonde_chan_options *opt)
{
	if (!out || !cfg || cfg->n_channels <= 0 || cfg->decim < 2 || cfg->max_in_len <= 0) return SONDE_ERR_ARG;
	if (opt && opt->interp > 1) return SONDE_ERR_ARG;                       /* integer rate changes only in the stand-in */
	sonde_chan *h = calloc(1, sizeof(*h));
	h->C = cfg->n_channels;
	h->D = cfg->decim;
	h->K = (cfg->taps_per_phase ? cfg->taps_per_phase : 8) * cfg->decim;
	h->fs_in = (double)cfg->fs_out * cfg->decim;
	h->g = malloc(sizeof(double) * h->C * h->K);
	h->w = malloc(sizeof(double) * h->C);
	h->hist = calloc((size_t)(h->K - 1) * 2, sizeof(float));
	h->stride = (size_t)cfg->max_in_len / cfg->decim + 1;
	for (int k = 0; k < 2; k++) h->out[k] = malloc(sizeof(float) * 2 * h->C * h->stride);
	for (int c = 0; c < h->C; c++) {
		double fc = cfg->cutoff_hz > 0 ? cfg->cutoff_hz : 0.42 * cfg->fs_out, sum = 0;
		if (opt && opt->cutoff_hz && opt->cutoff_hz[c] > 0) fc = opt->cutoff_hz[c];
		for (int k = 0; k < h->K; k++) {
			const double t = k - 0.5 * (h->K - 1), x = 2 * fc / h->fs_in * t;
			const double s = fabs(x) < 1e-12 ? 1.0 : sin(M_PI * x) / (M_PI * x);
			h->g[c * h->K + k] = s * (0.54 - 0.46 * cos(2 * M_PI * k / (h->K - 1)));
			sum += h->g[c * h->K + k];
		}
		for (int k = 0; k < h->K; k++) h->g[c * h->K + k] /= sum;
		h->w[c] = 2 * M_PI * cfg->freq_hz[c] / h->fs_in;
	}
	*out = h;
	return SONDE_OK;
}

int sonde_chan_create(sonde_chan **out, const sonde_chan_config *cfg) { return sonde_chan_create_ex(out, cfg, NULL); }

void sonde_chan_destroy(sonde_chan *h)
{
	if (!h) return;
	free(h->g); free(h->w); free(h->hist); free(h->out[0]); free(h->out[1]);
	free(h);
}

const char *sonde_chan_last_error(const sonde_chan *h) { return h ? h->err : ""; }

int sonde_chan_process_c64(sonde_chan *h, const float *x, size_t n_in, void *stream, void **d_out, size_t *out_stride)
{
	(void)stream;
	if (n_in % (size_t)h->D || n_in / h->D > h->stride) { strcpy(h->err, "stand-in: n_in must be a multiple of D within max_in_len"); return SONDE_ERR_ARG; }
	const int K = h->K, D = h->D;
	const size_t n_out = n_in / D;
	float *dst = h->out[h->flip ^= 1];
	/* per channel: mix the K-1 history samples and this call's samples down, then filter and decimate */
	double *mix = malloc(sizeof(double) * 2 * (n_in + K - 1));
	for (int c = 0; c < h->C; c++) {
		const double *g = h->g + (size_t)c * K;
		for (long n = -(long)(K - 1); n < (long)n_in; n++) {
			const double ph = -h->w[c] * (double)((long long)h->n_seen + n);
			const double cr = cos(ph), ci = sin(ph);
			const double xr = n < 0 ? h->hist[2 * (n + K - 1)] : x[2 * (size_t)n], xi = n < 0 ? h->hist[2 * (n + K - 1) + 1] : x[2 * (size_t)n + 1];
			mix[2 * (n + K - 1)] = xr * cr - xi * ci;
			mix[2 * (n + K - 1) + 1] = xr * ci + xi * cr;
		}
		for (size_t m = 0; m < n_out; m++) {
			double re = 0, im = 0;
			const double *top = mix + 2 * (m * D + D - 1 + K - 1);          /* newest sample of the window */
			for (int k = 0; k < K; k++) { re += g[k] * top[-2 * k]; im += g[k] * top[-2 * k + 1]; }
			dst[2 * ((size_t)c * h->stride + m)] = (float)re;
			dst[2 * ((size_t)c * h->stride + m) + 1] = (float)im;
		}
	}
	free(mix);
	/* keep the last K - 1 samples */
	if (n_in >= (size_t)(K - 1)) {
		memcpy(h->hist, x + 2 * (n_in - (K - 1)), sizeof(float) * 2 * (K - 1));
	} else {
		memmove(h->hist, h->hist + 2 * n_in, sizeof(float) * 2 * (K - 1 - n_in));
		memcpy(h->hist + 2 * (K - 1 - n_in), x, sizeof(float) * 2 * n_in);
	}
	h->n_seen += n_in;
	*d_out = dst;
	*out_stride = h->stride;
	return SONDE_OK;
}
