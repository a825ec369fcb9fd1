/*
 * gpu_multibank.hpp — C channels sharded over the GPUs of one box, in one host process.
 *
 * Channels share nothing (one decoder object per channel in the reference: src/main.hpp:36-42, SD/decode.c:24-30),
 * so the partition is the whole multi-GPU design: device g of G owns the contiguous block
 * [g*C/G, (g+1)*C/G) — the same rule as sdrpp_radiosonde_b200/shard.py:shard_range — with one sonde_b200 handle per
 * device.  There is no collective on the data path.  What moves is
 *   in:   host buffer [C][len]    -> every device copies ITS rows over its own PCIe link      (process_host)
 *         buffer on one GPU       -> every other device PULLS its rows over NVLink with the copy engine
 *                                    (sonde_b200_process_iq_peer = cudaMemcpyPeerAsync; no SMs, so the scatter of
 *                                    buffer i+1 runs beside the decode of buffer i)                (process_peer)
 *   out:  the frame records of all devices gathered into one [C][max_frames] array, global channel order  (fetch)
 * Calls are asynchronous per device and pipelined two deep exactly like the single-device C ABI:
 *     process(0); process(1); fetch() -> buffer 0; process(2); fetch() -> buffer 1; ...
 *
 * There is no CPU fallback: the constructor throws when a device cannot run the CUDA path.
 */
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sonde_b200.h"

namespace radiosonde {

class GpuMultiBank {
public:
	/* types[C]: decoder type per channel (SONDE_AUTO allowed); devices: CUDA device ordinals, one shard each (a device
	 * may be listed twice, which gives it two shards) */
	GpuMultiBank(const std::vector<int32_t> &types, int samplerate, int max_chunk_len, const std::vector<int> &devices)
	    : m_types(types), m_devices(devices)
	{
		const size_t C = types.size(), G = devices.size();
		if (!C || !G) throw std::invalid_argument("GpuMultiBank: no channels or no devices");
		m_lo.resize(G + 1);
		for (size_t g = 0; g <= G; g++) m_lo[g] = C * g / G;
		m_h.assign(G, nullptr);
		for (size_t g = 0; g < G; g++) {
			if (m_lo[g + 1] == m_lo[g]) continue;                        /* more devices than channels */
			sonde_b200_config cfg = {};
			cfg.n_channels = (int32_t)(m_lo[g + 1] - m_lo[g]);
			cfg.samplerate = samplerate;
			cfg.max_chunk_len = max_chunk_len;
			cfg.device = devices[g];
			cfg.types = m_types.data() + m_lo[g];
			const int rc = sonde_b200_create(&m_h[g], &cfg);
			if (rc != SONDE_OK) {
				close();
				throw std::runtime_error("GpuMultiBank: sonde_b200_create failed on device " + std::to_string(devices[g]) + " (" +
				                         std::to_string(rc) + "): the CUDA path is required, there is no CPU fallback");
			}
			m_max_frames = std::max(m_max_frames, sonde_b200_max_frames(m_h[g]));
		}
		m_stage.resize(G);
		m_cnt.resize(G);
		for (size_t g = 0; g < G; g++)
			if (m_h[g]) {
				m_stage[g].resize((m_lo[g + 1] - m_lo[g]) * (size_t)sonde_b200_max_frames(m_h[g]));
				m_cnt[g].resize(m_lo[g + 1] - m_lo[g]);
			}
	}
	~GpuMultiBank() { close(); }
	GpuMultiBank(const GpuMultiBank &) = delete;
	GpuMultiBank &operator=(const GpuMultiBank &) = delete;

	size_t channels() const { return m_types.size(); }
	size_t shards() const { return m_devices.size(); }
	int max_frames() const { return m_max_frames; }
	size_t shard_begin(size_t g) const { return m_lo[g]; }
	sonde_b200 *handle(size_t g) { return m_h[g]; }

	/* iq[C][len] interleaved float pairs in (ideally pinned) host memory */
	void process_host(const float *iq, size_t len)
	{
		for (size_t g = 0; g < m_h.size(); g++)
			if (m_h[g]) check(g, sonde_b200_process_iq(m_h[g], iq + m_lo[g] * len * 2, len));
	}

	/* d_iq[C][len] complex64 resident on GPU `src_device` */
	void process_peer(int src_device, const void *d_iq, size_t len)
	{
		for (size_t g = 0; g < m_h.size(); g++) {
			if (!m_h[g]) continue;
			const char *rows = static_cast<const char *>(d_iq) + m_lo[g] * len * 2 * sizeof(float);
			if (m_devices[g] == src_device) check(g, sonde_b200_process_iq_device(m_h[g], rows, len, len));
			else check(g, sonde_b200_process_iq_peer(m_h[g], src_device, rows, len, len));
		}
	}

	/* records of the oldest unfetched buffer of every shard: recs[C][max_frames()], counts[C], global channel order */
	void fetch(sonde_frame_rec *recs, int32_t *counts)
	{
		for (size_t g = 0; g < m_h.size(); g++) {
			if (!m_h[g]) continue;
			check(g, sonde_b200_fetch(m_h[g], m_stage[g].data(), m_cnt[g].data()));
			const int mf = sonde_b200_max_frames(m_h[g]);
			for (size_t c = m_lo[g]; c < m_lo[g + 1]; c++) {
				const int n = m_cnt[g][c - m_lo[g]];
				counts[c] = n;
				for (int k = 0; k < n; k++) recs[c * m_max_frames + k] = m_stage[g][(c - m_lo[g]) * mf + k];
			}
		}
	}

	void sync()
	{
		for (size_t g = 0; g < m_h.size(); g++)
			if (m_h[g]) check(g, sonde_b200_sync(m_h[g]));
	}

private:
	void check(size_t g, int rc)
	{
		if (rc != SONDE_OK)
			throw std::runtime_error("GpuMultiBank shard " + std::to_string(g) + ": " + sonde_b200_last_error(m_h[g]));
	}
	void close()
	{
		for (auto &h : m_h) {
			if (h) sonde_b200_destroy(h);
			h = nullptr;
		}
	}

	std::vector<int32_t> m_types;
	std::vector<int> m_devices;
	std::vector<size_t> m_lo;
	std::vector<sonde_b200 *> m_h;
	std::vector<std::vector<sonde_frame_rec>> m_stage;
	std::vector<std::vector<int32_t>> m_cnt;
	int m_max_frames = 0;
};

}  // namespace radiosonde
