/*
 * host_block_test.cpp — drives radiosonde::GpuDecoder (the dsp::block drop-in) and the reference-signature
 * compat API with an RS41 IQ/FM recording read from files written by the Python test.
 *
 *   host_block_test <iq.c64> <fm.f32> <n_samples> <chunk>
 * prints:  BLOCK callbacks=<n> frames=<n> ok=<n> seq=<last seq> serial=<last serial>
 *          COMPAT parsed=<n> with_fields=<n> seq=<...> serial=<...>
 * exit code 3 when the CUDA path is unavailable (no CPU fallback).
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../sdrpp_radiosonde_b200/host/gpu_decoder.hpp"

static int n_cb = 0, n_frames = 0, n_ok = 0, last_seq = -1;
static std::string last_serial;

static void on_data(SondeFullData *d, void *) { n_cb++; last_seq = d->seq; last_serial = d->serial; }
static void on_frame(int, const sonde_frame_rec *r, void *) { n_frames++; n_ok += r->ok; }

template <class T>
static std::vector<T> slurp(const char *path, size_t n)
{
	std::vector<T> v(n);
	FILE *f = fopen(path, "rb");
	if (!f || fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
	fclose(f);
	return v;
}

int main(int argc, char **argv)
{
	if (argc < 5) return 2;
	const size_t n = strtoul(argv[3], nullptr, 10), chunk = strtoul(argv[4], nullptr, 10);
	auto iq = slurp<dsp::complex_t>(argv[1], n);
	auto fm = slurp<float>(argv[2], n);

	/* --- the block: producer thread = this thread, consumer = the block's worker --- */
	dsp::stream<dsp::complex_t> vfo_out;
	radiosonde::GpuDecoder dec;
	try {
		dec.init(&vfo_out, 48000.0, SONDE_RS41, on_data, nullptr);
	} catch (const std::exception &e) {
		printf("NOGPU %s\n", e.what());
		return 3;
	}
	dec.setFrameCallback(on_frame, nullptr);
	dec.start();
	for (size_t pos = 0; pos < n; pos += chunk) {
		const size_t len = n - pos < chunk ? n - pos : chunk;
		memcpy(vfo_out.writeBuf, iq.data() + pos, len * sizeof(dsp::complex_t));
		if (!vfo_out.swap((int)len)) break;
	}
	/* wait until the last buffer has been consumed, then stop */
	memset(vfo_out.writeBuf, 0, sizeof(dsp::complex_t));
	vfo_out.swap(0);
	dec.stop();
	printf("BLOCK callbacks=%d frames=%d ok=%d seq=%d serial=%s\n", n_cb, n_frames, n_ok, last_seq, last_serial.c_str());
	dec.deinit();

	/* --- reference call protocol through the compat API --- */
	RS41Decoder *d = rs41_decoder_init(48000);
	if (!d) { printf("NOGPU compat\n"); return 3; }
	int parsed = 0, with_fields = 0, seq = -1;
	std::string serial;
	SondeData out, last_pos;
	memset(&last_pos, 0, sizeof(last_pos));
	for (size_t pos = 0; pos < n; pos += chunk) {
		const size_t len = n - pos < chunk ? n - pos : chunk;
		while (rs41_decode(d, &out, fm.data() + pos, len) != PROCEED) {
			parsed++;
			if (out.fields) { with_fields++; seq = out.seq; serial = out.serial; }
			if ((out.fields & (DATA_POS | DATA_TIME)) == (DATA_POS | DATA_TIME)) last_pos = out;
		}
	}
	printf("COMPAT parsed=%d with_fields=%d seq=%d serial=%s\n", parsed, with_fields, seq, serial.c_str());
	printf("TELEM fields=%d lat=%.6f lon=%.6f alt=%.3f speed=%.4f heading=%.4f climb=%.4f time=%lld\n", last_pos.fields,
	       last_pos.lat, last_pos.lon, last_pos.alt, last_pos.speed, last_pos.heading, last_pos.climb, (long long)last_pos.time);
	rs41_decoder_deinit(d);
	return 0;
}
