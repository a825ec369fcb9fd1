/*
 * host_multibank_test.cpp — radiosonde::GpuMultiBank: the channel batch sharded over the devices of the box (or two
 * shards on one device when there is only one), fed (a) from a host buffer and (b) from a buffer resident on device 0
 * that the other shards pull with the copy engine; the gathered records must equal a single-handle run.
 *
 *   host_multibank_test <n_channels> <n_samples> <chunk> <n_shards> <type0> <iq0.c64> <type1> <iq1.c64> ...
 * prints  MULTI shards=<n> devices=<n> host_equal=<0|1> peer_equal=<0|1> frames=<n> ok=<n>
 * exit code 3 when the CUDA path is unavailable (no CPU fallback).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../sdrpp_radiosonde_b200/host/gpu_multibank.hpp"

/* framer frame length in bytes per decoder type (4144, 560, 1664, 1200, 816, 600, 90 bits) */
static int raw_bytes(int type)
{
	static const int bits[7] = {4144, 560, 1664, 1200, 816, 600, 90};
	return type >= 0 && type < 7 ? (bits[type] + 7) / 8 : 0;
}

/* the fields a record defines (bytes past data_len / the raw frame are scratch), as tests/gpu_util.py:rec_key */
static bool same(const sonde_frame_rec &a, const sonde_frame_rec &b)
{
	const int n = a.data_len > 132 ? a.data_len : 132;
	return a.type == b.type && a.chunk == b.chunk && a.sync_offset == b.sync_offset && a.inverted == b.inverted &&
	       a.status == b.status && a.ok == b.ok && a.aux == b.aux && a.data_len == b.data_len && a.bit_pos == b.bit_pos &&
	       memcmp(a.data, b.data, (size_t)n) == 0 && memcmp(a.raw, b.raw, (size_t)raw_bytes(a.type)) == 0;
}

int main(int argc, char **argv)
{
	if (argc < 5) return 2;
	const size_t C = strtoul(argv[1], nullptr, 10), n = strtoul(argv[2], nullptr, 10), chunk = strtoul(argv[3], nullptr, 10);
	const int n_shards = atoi(argv[4]);
	if ((size_t)argc < 5 + 2 * C) return 2;
	std::vector<int32_t> types(C);
	std::vector<std::vector<float>> iq(C, std::vector<float>(2 * n));
	for (size_t c = 0; c < C; c++) {
		types[c] = atoi(argv[5 + 2 * c]);
		FILE *f = fopen(argv[6 + 2 * c], "rb");
		if (!f || fread(iq[c].data(), 8, n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[6 + 2 * c]); return 2; }
		fclose(f);
	}
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { printf("NOGPU no CUDA device: there is no CPU fallback\n"); return 3; }
	std::vector<int> devices;
	for (int g = 0; g < n_shards; g++) devices.push_back(g % ndev);

	try {
		/* reference run: one handle on device 0 */
		std::vector<std::vector<sonde_frame_rec>> want(C);
		{
			radiosonde::GpuMultiBank one(types, 48000, (int)chunk, {0});
			std::vector<sonde_frame_rec> recs(C * one.max_frames());
			std::vector<int32_t> cnt(C);
			float *buf = (float *)sonde_b200_host_alloc(C * chunk * 8);
			for (size_t pos = 0; pos < n; pos += chunk) {
				const size_t len = n - pos < chunk ? n - pos : chunk;
				for (size_t c = 0; c < C; c++) memcpy(buf + c * len * 2, iq[c].data() + 2 * pos, len * 8);
				one.process_host(buf, len);
				one.fetch(recs.data(), cnt.data());
				for (size_t c = 0; c < C; c++)
					for (int k = 0; k < cnt[c]; k++) want[c].push_back(recs[c * one.max_frames() + k]);
			}
			sonde_b200_host_free(buf);
		}
		bool equal[2] = {true, true};
		long frames = 0, ok = 0;
		for (int mode = 0; mode < 2; mode++) {
			radiosonde::GpuMultiBank bank(types, 48000, (int)chunk, devices);
			std::vector<std::vector<sonde_frame_rec>> got(C);
			std::vector<sonde_frame_rec> recs(C * bank.max_frames());
			std::vector<int32_t> cnt(C);
			float *buf[2] = {(float *)sonde_b200_host_alloc(C * chunk * 8), (float *)sonde_b200_host_alloc(C * chunk * 8)};
			void *dsrc[2] = {nullptr, nullptr};
			if (mode == 1) {
				cudaSetDevice(0);
				cudaMalloc(&dsrc[0], C * chunk * 8);
				cudaMalloc(&dsrc[1], C * chunk * 8);
			}
			/* two-deep pipeline: issue buffer k+1 before fetching buffer k */
			std::vector<size_t> starts;
			for (size_t pos = 0; pos < n; pos += chunk) starts.push_back(pos);
			auto issue = [&](size_t k) {
				const size_t pos = starts[k], len = n - pos < chunk ? n - pos : chunk;
				float *b = buf[k & 1];
				for (size_t c = 0; c < C; c++) memcpy(b + c * len * 2, iq[c].data() + 2 * pos, len * 8);
				if (mode == 0) {
					bank.process_host(b, len);
				} else {
					cudaSetDevice(0);
					cudaMemcpy(dsrc[k & 1], b, C * len * 8, cudaMemcpyHostToDevice);      /* the front-end GPU's buffer */
					bank.process_peer(0, dsrc[k & 1], len);
				}
			};
			issue(0);
			for (size_t k = 0; k < starts.size(); k++) {
				if (k + 1 < starts.size()) issue(k + 1);
				bank.fetch(recs.data(), cnt.data());
				for (size_t c = 0; c < C; c++)
					for (int j = 0; j < cnt[c]; j++) got[c].push_back(recs[c * bank.max_frames() + j]);
			}
			bank.sync();
			for (size_t c = 0; c < C; c++) {
				if (got[c].size() != want[c].size()) { equal[mode] = false; continue; }
				for (size_t j = 0; j < got[c].size(); j++) equal[mode] = equal[mode] && same(got[c][j], want[c][j]);
				if (mode == 0) { frames += (long)got[c].size(); for (auto &r : got[c]) ok += r.ok; }
			}
			for (auto b : buf) sonde_b200_host_free(b);
			if (mode == 1) { cudaSetDevice(0); cudaFree(dsrc[0]); cudaFree(dsrc[1]); }
		}
		printf("MULTI shards=%d devices=%d host_equal=%d peer_equal=%d frames=%ld ok=%ld\n", n_shards, ndev, (int)equal[0],
		       (int)equal[1], frames, ok);
	} catch (const std::exception &e) {
		printf("NOGPU %s\n", e.what());
		return 3;
	}
	return 0;
}
