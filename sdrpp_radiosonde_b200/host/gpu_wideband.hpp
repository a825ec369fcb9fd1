/*
 * gpu_wideband.hpp — radiosonde::GpuWidebandBank : dsp::block
 *
 * One wideband IQ stream in (the SDR source's dsp::stream<dsp::complex_t>), C sondes out: replaces, for C plugin
 * instances at once, the whole per-instance chain of the reference
 *
 *     VFO (sigpath::vfoManager.createVFO, src/main.cpp:55) -> dsp::demod::FM (:57) -> RationalResampler (:60)
 *     -> radiosonde::Decoder<...> (:62-68, src/decode/decoder.hpp:22-129)
 *
 * with the GPU channelizer (include/sonde_b200_channelizer.h) feeding the batched decoder (include/sonde_b200.h)
 * on the same CUDA stream — the narrowband channels never leave HBM.  Fires the same
 * void(*)(SondeFullData*, void*) callback as the reference's blocks, plus an optional per-frame callback that
 * carries the channel index.
 *
 * Buffers of arbitrary length are accepted: samples that do not fill a whole output sample (n mod D) are carried to
 * the next buffer.  No CPU fallback: init() throws when the CUDA path is unavailable.
 */
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sonde_b200_channelizer.h"
#include "gpu_decoder.hpp"

namespace radiosonde {

class GpuWidebandBank : public dsp::block {
public:
	GpuWidebandBank() {}
	~GpuWidebandBank() { deinit(); }

	/* fs_in * L / M = 48000 for integers M >= 2 and 1 <= L <= 16 (M = fs_in / 48000 when that is an integer; e.g.
	 * 2.048 MS/s: 3/128, 2.5 MS/s: 12/625 — the rational resampler of src/main.cpp:60; multiples of 4 for M run fastest);
	 * freq_hz[c] = centre of channel c relative to the centre of the wideband stream; types[c] = enum sonde_type or
	 * SONDE_AUTO.  Every channel's filter gets half the VFO bandwidth the plugin uses for its sonde type
	 * (src/main.hpp:45-51), AUTO channels the widest; `split_precision` selects SONDE_CHAN_SPLIT_BF16. */
	static float typeCutoffHz(int type)
	{
		/* VFO bandwidths: RS41 10 kHz, DFM 15, iMS-100 20, M10/M20 50, iMet-4 20, SRS-C50 20, MRZ-N1 20; a 48 kS/s
		 * channel carries at most +-24 kHz, the filter's transition band ends there */
		switch (type) {
		case SONDE_RS41:  return 5000.0f;
		case SONDE_DFM09: return 7500.0f;
		case SONDE_M10:   return 0.42f * 48000.0f;
		case SONDE_AUTO:  return 0.42f * 48000.0f;
		default:          return 10000.0f;
		}
	}

	void init(dsp::stream<dsp::complex_t> *in, double fs_in, const std::vector<double> &freq_hz, const std::vector<int> &types,
	          SondeCallback cb, void *ctx, int max_in = dsp::STREAM_BUFFER_SIZE, int device = 0, bool split_precision = false,
	          int taps_per_phase = 0)
	{
		if (!in || freq_hz.empty() || freq_hz.size() != types.size()) throw std::invalid_argument("GpuWidebandBank: bad channel list");
		int L = 0, D = 0;
		for (int l = 1; l <= 16 && !L; l++) {
			const double m = fs_in * l / 48000.0;
			const long long mi = (long long)(m + 0.5);
			if (mi >= 2 && mi > l && (double)mi * 48000.0 == fs_in * l) { L = l; D = (int)mi; }
		}
		if (!L) throw std::invalid_argument("GpuWidebandBank: fs_in * L / M must be 48000 with 1 <= L <= 16, M >= 2");
		m_in = in; m_cb = cb; m_ctx = ctx; m_D = D; m_L = L;
		m_types.assign(types.begin(), types.end());
		m_max_in = (max_in + D) / D * D;
		sonde_chan_config cc;
		memset(&cc, 0, sizeof(cc));
		cc.n_channels = (int32_t)freq_hz.size();
		cc.decim = D;
		cc.fs_out = 48000;
		cc.max_in_len = m_max_in;
		cc.device = device;
		cc.freq_hz = freq_hz.data();
		cc.taps_per_phase = taps_per_phase;
		std::vector<float> cut(types.size());
		for (size_t c = 0; c < types.size(); c++) cut[c] = typeCutoffHz(types[c]);
		sonde_chan_options co;
		memset(&co, 0, sizeof(co));
		co.precision = split_precision ? SONDE_CHAN_SPLIT_BF16 : SONDE_CHAN_BF16;
		co.interp = L;
		co.cutoff_hz = cut.data();
		int rc = sonde_chan_create_ex(&m_ch, &cc, &co);
		if (rc != SONDE_OK)
			throw std::runtime_error("sonde_chan_create failed (" + std::to_string(rc) + "): the CUDA path is required, there is no CPU fallback");
		sonde_b200_config cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.n_channels = (int32_t)types.size();
		cfg.samplerate = 48000;
		cfg.max_chunk_len = m_max_in / D * L;
		cfg.device = device;
		cfg.types = m_types.data();
		rc = sonde_b200_create(&m_h, &cfg);
		if (rc != SONDE_OK) {
			sonde_chan_destroy(m_ch); m_ch = nullptr;
			throw std::runtime_error("sonde_b200_create failed (" + std::to_string(rc) + "): the CUDA path is required, there is no CPU fallback");
		}
		m_max_frames = sonde_b200_max_frames(m_h);
		m_recs.resize(types.size() * (size_t)m_max_frames);
		m_counts.resize(types.size());
		m_data.resize(types.size());
		m_tele.clear();
		for (size_t c = 0; c < types.size(); c++) m_tele.emplace_back(types[c]);
		m_stage = (float *)sonde_b200_host_alloc((size_t)m_max_in * 2 * sizeof(float));
		if (!m_stage) throw std::runtime_error("pinned staging allocation failed");
		m_carry = 0;
		dsp::block::registerInput(m_in);
		dsp::block::_block_init = true;
	}

	void setFrameCallback(FrameCallback cb, void *ctx) { m_fcb = cb; m_fctx = ctx; }

	void deinit()
	{
		if (!dsp::block::_block_init) return;
		dsp::block::stop();
		dsp::block::unregisterInput(m_in);
		dsp::block::_block_init = false;
		sonde_b200_destroy(m_h); m_h = nullptr;           /* the decoder first: it reads the channelizer's buffers */
		sonde_chan_destroy(m_ch); m_ch = nullptr;
		if (m_stage) sonde_b200_host_free(m_stage);
		m_stage = nullptr;
	}

	int run() override
	{
		int n = m_in->read();
		if (n < 0) return -1;
		const dsp::complex_t *src = m_in->readBuf;
		while (n > 0) {
			/* staging = carried remainder + as much of this buffer as fits */
			int take = n;
			if (m_carry + take > m_max_in) take = m_max_in - m_carry;
			memcpy(m_stage + 2 * (size_t)m_carry, src, (size_t)take * sizeof(dsp::complex_t));
			const int have = m_carry + take;
			const int use = have / m_D * m_D;
			if (use > 0) process(use);
			m_carry = have - use;
			if (m_carry) memmove(m_stage, m_stage + 2 * (size_t)use, (size_t)m_carry * sizeof(dsp::complex_t));
			src += take;
			n -= take;
		}
		m_in->flush();
		return 0;
	}

	sonde_b200 *decoder() { return m_h; }
	sonde_chan *channelizer() { return m_ch; }

private:
	void process(int n_in)
	{
		void *d_out = nullptr;
		size_t stride = 0;
		if (sonde_chan_process_c64(m_ch, m_stage, (size_t)n_in, sonde_b200_stream(m_h), &d_out, &stride) != SONDE_OK)
			throw std::runtime_error(std::string("sonde_chan: ") + sonde_chan_last_error(m_ch));
		if (sonde_b200_process_iq_device(m_h, d_out, (size_t)(n_in / m_D * m_L), stride) != SONDE_OK ||
		    sonde_b200_fetch(m_h, m_recs.data(), m_counts.data()) != SONDE_OK)      /* fetch() also orders the reuse of m_stage */
			throw std::runtime_error(std::string("sonde_b200: ") + sonde_b200_last_error(m_h));
		for (size_t c = 0; c < m_types.size(); c++) {
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

	dsp::stream<dsp::complex_t> *m_in = nullptr;
	std::vector<int32_t> m_types;
	SondeCallback m_cb = nullptr;
	FrameCallback m_fcb = nullptr;
	void *m_ctx = nullptr, *m_fctx = nullptr;
	sonde_chan *m_ch = nullptr;
	sonde_b200 *m_h = nullptr;
	int m_D = 0, m_L = 1, m_max_in = 0, m_max_frames = 0, m_carry = 0;
	std::vector<sonde_frame_rec> m_recs;
	std::vector<int32_t> m_counts;
	std::vector<SondeFullData> m_data;
	std::vector<Telemetry> m_tele;
	float *m_stage = nullptr;
};

}  // namespace radiosonde
