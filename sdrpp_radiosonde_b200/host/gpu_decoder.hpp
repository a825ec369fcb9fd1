/*
 * gpu_decoder.hpp — the SDR++ decoder-block surface on top of the C ABI (include/sonde_b200.h).
 *
 *   radiosonde::GpuDecoder       : dsp::block     one channel, complex IQ in -> SondeFullData callback
 *   radiosonde::GpuChannelBank   : dsp::block     C channels batched into one GPU call per buffer
 *
 * GpuDecoder replaces, behind src/main.cpp:57-68, the chain
 *     dsp::demod::FM<float> -> dsp::multirate::RationalResampler<float> -> radiosonde::Decoder<...>
 * (src/main.hpp:33-42, src/decode/decoder.hpp:22-129): it is fed by vfo->output
 * (dsp::stream<dsp::complex_t> at the 48 kS/s channel rate) and fires the same
 * void(*)(SondeFullData*, void*) callback on the block's worker thread.
 *
 * run() mirrors decoder.hpp:53-119: read() -> for every framer window of the buffer
 * (== every PARSED return of the reference's xxx_decode loop) merge the fragment into the persistent
 * SondeFullData and call back iff fragment.fields != 0 -> flush().
 *
 * Telemetry: frame -> SondeData conversion is radiosonde::Telemetry (telemetry.hpp), one
 * stateful parser per channel; see its header for the coverage.  Every record is additionally handed to an
 * optional frame callback so a host parser can run on the exact bytes the reference's parser would see.
 *
 * There is no CPU fallback: init() throws std::runtime_error when the CUDA path is unavailable.
 */
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sonde_b200.h"
#include "dsp_standin.hpp"
#include "sonde_data.hpp"
#include "telemetry.hpp"

namespace radiosonde {

typedef void (*SondeCallback)(SondeFullData *data, void *ctx);
typedef void (*FrameCallback)(int channel, const sonde_frame_rec *rec, void *ctx);

inline uint16_t crc16_ccitt_false(const uint8_t *p, size_t n)
{
	uint16_t crc = 0xFFFF;
	for (; n; n--) {
		crc ^= (uint16_t)(*p++ << 8);
		for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
	}
	return crc;
}

inline float dewpoint(float temp, float rh)          /* Magnus formula, as src/decode/decoder.hpp:132-140 */
{
	const float t = (logf(rh / 100.0f) + (17.27f * temp / (237.3f + temp))) / 17.27f;
	return 237.3f * t / (1 - t);
}

/* merge of decoder.hpp:64-110 */
inline void merge_fragment(SondeFullData &d, const SondeData &f)
{
	if (f.fields & DATA_SEQ) d.seq = f.seq;
	if (f.fields & DATA_POS) { d.lat = f.lat; d.lon = f.lon; d.alt = f.alt; }
	if (f.fields & DATA_SPEED) { d.spd = f.speed; d.hdg = f.heading; d.climb = f.climb; }
	if (f.fields & DATA_TIME) d.time = f.time;
	if (f.fields & DATA_PTU) {
		d.calib_percent = f.calib_percent;
		d.calibrated = d.calib_percent >= 100.0f;
		d.temp = f.temp; d.rh = f.rh; d.pressure = f.pressure;
		d.dewpt = dewpoint(d.temp, d.rh);
	}
	if (f.fields & DATA_SERIAL) d.serial = f.serial;
	if (f.fields & DATA_SHUTDOWN) d.burstkill = f.shutdown;
	if (f.fields & DATA_OZONE) {
		char buf[48];
		snprintf(buf, sizeof(buf), "O3=%.2fmPa", f.o3_mpa);
		d.auxData = buf;
	}
	/* decoder.hpp:108-110: after every fragment, a sonde (or a frame) without a pressure sensor reports the
	 * barometric estimate of its altitude */
	if (d.pressure <= 0) d.pressure = tl::altitude_to_pressure(d.alt);
}

/* C channels, one input stream each, one GPU call per buffer. */
class GpuChannelBank : public dsp::block {
public:
	GpuChannelBank() {}
	~GpuChannelBank() { deinit(); }

	void init(const std::vector<dsp::stream<dsp::complex_t> *> &in, int samplerate, const std::vector<int> &types,
	          SondeCallback cb, void *ctx, int max_chunk = dsp::STREAM_BUFFER_SIZE / 8, int device = 0)
	{
		if (in.empty() || in.size() != types.size()) throw std::invalid_argument("GpuChannelBank: bad channel list");
		m_in = in;
		m_cb = cb;
		m_ctx = ctx;
		m_types.assign(types.begin(), types.end());
		sonde_b200_config cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.n_channels = (int32_t)in.size();
		cfg.samplerate = samplerate;
		cfg.max_chunk_len = max_chunk;
		cfg.device = device;
		cfg.types = m_types.data();
		const int rc = sonde_b200_create(&m_h, &cfg);
		if (rc != SONDE_OK)
			throw std::runtime_error(std::string("sonde_b200_create failed (") + std::to_string(rc) +
			                         "): the CUDA path is required, there is no CPU fallback");
		m_max_frames = sonde_b200_max_frames(m_h);
		m_max_chunk = max_chunk;
		m_recs.resize((size_t)in.size() * m_max_frames);
		m_counts.resize(in.size());
		m_data.resize(in.size());
		m_tele.clear();
		for (size_t c = 0; c < in.size(); c++) m_tele.emplace_back(types[c]);
		m_backlog.assign(in.size(), {});
		for (auto &st : m_stage) {
			st = (float *)sonde_b200_host_alloc((size_t)in.size() * max_chunk * 2 * sizeof(float));
			if (!st) throw std::runtime_error("pinned staging allocation failed");
		}
		for (auto *s : m_in) dsp::block::registerInput(s);
		dsp::block::_block_init = true;
	}

	void setFrameCallback(FrameCallback cb, void *ctx) { m_fcb = cb; m_fctx = ctx; }

	void deinit()
	{
		if (!dsp::block::_block_init) return;
		dsp::block::stop();
		for (auto *s : m_in) dsp::block::unregisterInput(s);
		dsp::block::_block_init = false;
		for (auto &st : m_stage) {
			if (st) sonde_b200_host_free(st);
			st = nullptr;
		}
		sonde_b200_destroy(m_h);
		m_h = nullptr;
	}

	/* One pass: take what every input stream delivers, decode as much as ALL channels have, keep the rest.
	 *
	 * The reference block consumes its whole buffer (decoder.hpp:59-117).  A bank has C streams that need not deliver
	 * the same count, and a dsp::stream buffer may carry more than max_chunk samples, so nothing may be cut off:
	 * every stream's samples are appended to a per-channel backlog, the common prefix is decoded in sub-chunks of at
	 * most max_chunk, and what one channel has beyond the others waits for the next pass.  Sub-chunks are pipelined
	 * two deep through the C ABI (process(k+1) is issued before fetch(k), sonde_b200.h), from two pinned staging
	 * buffers that the channels are copied into by a few threads. */
	int run() override
	{
		const size_t C = m_in.size();
		for (size_t c = 0; c < C; c++) {
			const int n = m_in[c]->read();
			if (n < 0) return -1;
			m_backlog[c].insert(m_backlog[c].end(), m_in[c]->readBuf, m_in[c]->readBuf + n);
			m_in[c]->flush();
		}
		size_t avail = m_backlog[0].size();
		for (size_t c = 1; c < C; c++) avail = std::min(avail, m_backlog[c].size());

		size_t done = 0;
		int in_flight = 0;                               /* sub-chunks processed but not fetched yet (0 or 1) */
		int slot = 0;
		while (done < avail) {
			const size_t count = std::min(avail - done, (size_t)m_max_chunk);
			stage(slot, done, count);
			if (sonde_b200_process_iq(m_h, m_stage[slot], count) != SONDE_OK)
				throw std::runtime_error(std::string("sonde_b200: ") + sonde_b200_last_error(m_h));
			if (in_flight) deliver();                    /* records of the sub-chunk before this one */
			in_flight = 1;
			done += count;
			slot ^= 1;
		}
		if (in_flight) deliver();
		if (done)
			for (size_t c = 0; c < C; c++) m_backlog[c].erase(m_backlog[c].begin(), m_backlog[c].begin() + done);
		return 0;
	}

	/* samples waiting because another channel's stream has delivered less so far */
	size_t backlog(size_t channel) const { return m_backlog[channel].size(); }

	sonde_b200 *handle() { return m_h; }

private:
	/* copy [first, first + count) of every channel's backlog into staging buffer `slot`, [C][count] row-major */
	void stage(int slot, size_t first, size_t count)
	{
		const size_t C = m_in.size();
		const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t)8, C, (size_t)std::thread::hardware_concurrency()}));
		auto work = [&](size_t t) {
			for (size_t c = t; c < C; c += nthreads)
				memcpy(m_stage[slot] + c * count * 2, m_backlog[c].data() + first, count * sizeof(dsp::complex_t));
		};
		if (nthreads == 1 || C * count < (size_t)1 << 16) {
			for (size_t t = 0; t < nthreads; t++) work(t);
			return;
		}
		std::vector<std::thread> pool;
		for (size_t t = 1; t < nthreads; t++) pool.emplace_back(work, t);
		work(0);
		for (auto &th : pool) th.join();
	}

	/* fetch the oldest unfetched call and fire the callbacks, in record order per channel */
	void deliver()
	{
		const size_t C = m_in.size();
		if (sonde_b200_fetch(m_h, m_recs.data(), m_counts.data()) != SONDE_OK)
			throw std::runtime_error(std::string("sonde_b200: ") + sonde_b200_last_error(m_h));
		for (size_t c = 0; c < C; c++) {
			for (int k = 0; k < m_counts[c]; k++) {
				const sonde_frame_rec &r = m_recs[c * m_max_frames + k];
				if (m_fcb) m_fcb((int)c, &r, m_fctx);
				SondeData fragment;
				m_tele[c].parse(r, &fragment);
				merge_fragment(m_data[c], fragment);
				if (fragment.fields && m_cb) m_cb(&m_data[c], m_ctx);
			}
		}
	}

	std::vector<dsp::stream<dsp::complex_t> *> m_in;
	std::vector<std::vector<dsp::complex_t>> m_backlog;
	std::vector<int32_t> m_types;
	SondeCallback m_cb = nullptr;
	FrameCallback m_fcb = nullptr;
	void *m_ctx = nullptr, *m_fctx = nullptr;
	sonde_b200 *m_h = nullptr;
	int m_max_frames = 0, m_max_chunk = 0;
	std::vector<sonde_frame_rec> m_recs;
	std::vector<int32_t> m_counts;
	std::vector<SondeFullData> m_data;
	std::vector<Telemetry> m_tele;
	float *m_stage[2] = {nullptr, nullptr};
};

/* One channel: the drop-in for one plugin instance (src/main.hpp:33-42). */
class GpuDecoder : public GpuChannelBank {
public:
	void init(dsp::stream<dsp::complex_t> *in, double samplerate, int type, SondeCallback cb, void *ctx)
	{
		GpuChannelBank::init({in}, (int)samplerate, {type}, cb, ctx);
	}
};

}  // namespace radiosonde
