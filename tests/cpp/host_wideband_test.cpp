/*
 * host_wideband_test.cpp — drives radiosonde::GpuWidebandBank (wideband dsp::stream in, C sondes out).
 *
 *   host_wideband_test <wide.c64> <n_samples> <decim> <buffer_len> <freq0> <type0> [<freq1> <type1> ...]
 * prints one line per channel:  CH <c> frames=<n> ok=<n> callbacks=<n> seq=<last seq> serial=<last serial>
 * exit code 3 when the CUDA path is unavailable (no CPU fallback).
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../sdrpp_radiosonde_b200/host/gpu_wideband.hpp"

struct chan_stat { int frames = 0, ok = 0; };
static std::vector<chan_stat> g_stat;
static int g_cb = 0;
static std::vector<std::string> g_serials;

static void on_data(SondeFullData *d, void *) { g_cb++; if (!d->serial.empty()) g_serials.push_back(d->serial + ":" + std::to_string(d->seq)); }
static void on_frame(int c, const sonde_frame_rec *r, void *) { g_stat[c].frames++; g_stat[c].ok += r->ok; }

int main(int argc, char **argv)
{
	if (argc < 7 || (argc - 5) % 2) return 2;
	const size_t n = strtoul(argv[2], nullptr, 10);
	const int D = atoi(argv[3]);
	const size_t buflen = strtoul(argv[4], nullptr, 10);
	std::vector<double> freqs;
	std::vector<int> types;
	for (int i = 5; i + 1 < argc; i += 2) { freqs.push_back(atof(argv[i])); types.push_back(atoi(argv[i + 1])); }
	g_stat.resize(freqs.size());
	std::vector<dsp::complex_t> wide(n);
	FILE *f = fopen(argv[1], "rb");
	if (!f || fread(wide.data(), sizeof(dsp::complex_t), n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
	fclose(f);

	dsp::stream<dsp::complex_t> src;
	radiosonde::GpuWidebandBank bank;
	try {
		bank.init(&src, 48000.0 * D, freqs, types, on_data, nullptr, 48000 * D);
	} catch (const std::exception &e) {
		printf("NOGPU %s\n", e.what());
		return 3;
	}
	bank.setFrameCallback(on_frame, nullptr);
	bank.start();
	for (size_t pos = 0; pos < n; pos += buflen) {
		const size_t len = n - pos < buflen ? n - pos : buflen;
		memcpy(src.writeBuf, wide.data() + pos, len * sizeof(dsp::complex_t));
		if (!src.swap((int)len)) break;
	}
	src.swap(0);
	bank.stop();
	for (size_t c = 0; c < freqs.size(); c++) printf("CH %zu frames=%d ok=%d\n", c, g_stat[c].frames, g_stat[c].ok);
	printf("CALLBACKS %d\n", g_cb);
	for (auto &s : g_serials) printf("SERIAL %s\n", s.c_str());
	bank.deinit();
	return 0;
}
