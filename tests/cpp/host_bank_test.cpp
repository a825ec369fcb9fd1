/*
 * host_bank_test.cpp — radiosonde::GpuChannelBank with streams that do NOT deliver in lock step: per pass every
 * channel's producer hands over a buffer of a different length, some longer than the bank's max_chunk.  Nothing may
 * be dropped (VERDICT r1 / ADVICE r1: run() used to decode min(count) and flush the rest away).
 *
 *   host_bank_test <n_channels> <n_samples> <max_chunk> <type0> <iq0.c64> <type1> <iq1.c64> ...
 * With BANK_TEST_SEED=<n> in the environment the buffer lengths are random (1 .. 40000, a third of them below 64)
 * instead of the fixed patterns.
 * prints per channel:  CH <c> frames=<n> ok=<n> callbacks=<n> pressure_ok=<0|1> digest=<FNV-1a over status and frame
 *                      bytes of every record but the first, in order: equal digests = same frames in the same order.
 *                      The first record is left out: the reference's demodulator forgets its mid-symbol sample at
 *                      the start of every call (SD/demod/gfsk.c:73), so where a buffer ends matters to the timing loop
 *                      while it is still acquiring (a first buffer of 55-57 samples flips the third bit of an RS41
 *                      stream, in the compiled reference and in the oracle alike), and the bank chooses the call lengths>
 * and                  BACKLOG <samples left in the backlogs after the last pass>
 * exit code 3 when the CUDA path is unavailable (no CPU fallback).
 */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../sdrpp_radiosonde_b200/host/gpu_decoder.hpp"

struct Tally { int frames = 0, ok = 0, cb = 0; bool pressure_ok = true; unsigned long long digest = 1469598103934665603ull; };
static void fnv(unsigned long long &h, const void *p, size_t n)
{
	for (size_t i = 0; i < n; i++) { h ^= ((const unsigned char *)p)[i]; h *= 1099511628211ull; }
}
static std::vector<Tally> tally;
static std::vector<SondeFullData *> slots;

static void on_frame(int c, const sonde_frame_rec *r, void *)
{
	tally[c].frames++;
	tally[c].ok += r->ok;
	if (tally[c].frames == 1) return;      /* the first window holds the demodulator's start-up bits, see below */
	fnv(tally[c].digest, &r->status, sizeof(r->status));
	fnv(tally[c].digest, r->data, (size_t)r->data_len);
}
static void on_data(SondeFullData *d, void *)
{
	/* the bank hands out one persistent SondeFullData per channel: identify the channel by address */
	for (size_t c = 0; c < slots.size(); c++) {
		if (slots[c] == nullptr) { slots[c] = d; }
		if (slots[c] == d) {
			tally[c].cb++;
			if (!(d->pressure > 0)) tally[c].pressure_ok = false;      /* decoder.hpp:108-110 fallback */
			return;
		}
	}
}

int main(int argc, char **argv)
{
	if (argc < 4) return 2;
	const size_t C = strtoul(argv[1], nullptr, 10), n = strtoul(argv[2], nullptr, 10);
	const int max_chunk = atoi(argv[3]);
	if ((size_t)argc < 4 + 2 * C) return 2;
	std::vector<int> types(C);
	std::vector<std::vector<dsp::complex_t>> iq(C, std::vector<dsp::complex_t>(n));
	for (size_t c = 0; c < C; c++) {
		types[c] = atoi(argv[4 + 2 * c]);
		FILE *f = fopen(argv[5 + 2 * c], "rb");
		if (!f || fread(iq[c].data(), sizeof(dsp::complex_t), n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[5 + 2 * c]); return 2; }
		fclose(f);
	}
	tally.assign(C, Tally());
	slots.assign(C, nullptr);

	std::vector<dsp::stream<dsp::complex_t>> streams(C);
	std::vector<dsp::stream<dsp::complex_t> *> in;
	for (auto &s : streams) in.push_back(&s);
	radiosonde::GpuChannelBank bank;
	try {
		bank.init(in, 48000, types, on_data, nullptr, max_chunk);
	} catch (const std::exception &e) {
		printf("NOGPU %s\n", e.what());
		return 3;
	}
	bank.setFrameCallback(on_frame, nullptr);
	bank.start();
	/* channel c delivers buffers of len_c(pass) samples; the patterns differ per channel and include buffers longer
	 * than max_chunk.  Every channel delivers n samples in total. */
	static const int pattern[3][5] = {{4096, 1000, 12000, 7, 3001}, {1024, 9000, 333, 5000, 2048}, {2500, 2500, 16000, 1, 640}};
	std::vector<size_t> pos(C, 0);
	const char *seed_env = getenv("BANK_TEST_SEED");
	unsigned long long lcg = seed_env ? strtoull(seed_env, nullptr, 10) * 2862933555777941757ull + 3037000493ull : 0;
	auto rnd = [&lcg]() { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(lcg >> 33); };
	for (int pass = 0;; pass++) {
		bool any = false;
		for (size_t c = 0; c < C; c++) {
			size_t len = (size_t)pattern[c % 3][pass % 5];
			if (seed_env) len = (rnd() % 3 == 0) ? 1 + rnd() % 63 : 1 + rnd() % 40000;
			if (len > n - pos[c]) len = n - pos[c];
			any |= len > 0;
			memcpy(streams[c].writeBuf, iq[c].data() + pos[c], len * sizeof(dsp::complex_t));
			pos[c] += len;
			if (!streams[c].swap((int)len)) return 1;
		}
		if (!any) break;
	}
	/* one more empty pass so that the worker has consumed everything before it is stopped */
	for (size_t c = 0; c < C; c++) streams[c].swap(0);
	bank.stop();
	size_t left = 0;
	for (size_t c = 0; c < C; c++) left += bank.backlog(c);
	for (size_t c = 0; c < C; c++)
		printf("CH %zu frames=%d ok=%d callbacks=%d pressure_ok=%d digest=%016llx\n", c, tally[c].frames, tally[c].ok, tally[c].cb,
		       (int)tally[c].pressure_ok, tally[c].digest);
	printf("BACKLOG %zu\n", left);
	bank.deinit();
	return 0;
}
