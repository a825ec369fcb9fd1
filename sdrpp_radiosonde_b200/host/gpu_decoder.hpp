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
 * Telemetry: frame -> SondeData conversion is SURVEY.md §8 row f-1 ("next"); fragment_from_record()
 * below covers RS41 sequence number, serial, GPS position / velocity and GPS time (CRC-checked subframes as
 * rs41.c:165-175 does) and the M10 / MRZ-N1 sequence counters; PTU/XDATA and the other sondes' physical
 * values are not converted yet.  Every record is additionally handed to an optional frame callback so a
 * host parser can run on the exact bytes the reference's parser would see.
 *
 * There is no CPU fallback: init() throws std::runtime_error when the CUDA path is unavailable.
 */
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sonde_b200.h"
#include "dsp_standin.hpp"
#include "sonde_data.hpp"

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

/* ---- WGS-84 ECEF -> geodetic (Bowring's one-step method) and ENU velocity, as SD/gps/ecef.c:6-57.
 * Mixed float/double like the reference (double constants, float libm calls) so the results agree to
 * rounding. */
namespace wgs84 {
constexpr double A = 6378137.0, F = 1 / 298.257223563, B = A * (1 - F);
constexpr double E2 = (A * A - B * B) / (A * A), EP2 = (A * A - B * B) / (B * B);
constexpr double PI = 3.14159265358979323846;
}  // namespace wgs84

inline bool ecef_to_lla(float *lat, float *lon, float *alt, float x, float y, float z)
{
	const float lambda = atan2f(y, x);
	const float p = sqrtf(x * x + y * y);
	const float theta = atan2f((float)(z * wgs84::A), (float)(p * wgs84::B));
	const float st = sinf(theta), ct = cosf(theta);
	if (x == 0 || y == 0 || z == 0) {
		*lat = *lon = *alt = NAN;
		return false;
	}
	const float phi = atan2f((float)(z + wgs84::EP2 * wgs84::B * (st * st * st)),
	                         (float)(p - wgs84::E2 * wgs84::A * (ct * ct * ct)));
	const float sp = sinf(phi);
	const float n = (float)(wgs84::A / sqrtf((float)(1 - wgs84::E2 * sp * sp)));
	*lat = (float)(phi * 180 / wgs84::PI);
	*lon = (float)(lambda * 180 / wgs84::PI);
	*alt = p / cosf(phi) - n;
	return true;
}

inline void ecef_velocity(float *speed, float *heading, float *climb, float lat, float lon, float dx, float dy, float dz)
{
	lat = (float)(lat * (wgs84::PI / 180));
	lon = (float)(lon * (wgs84::PI / 180));
	if (dx == 0 && dy == 0 && dz == 0) {
		*speed = *heading = *climb = 0;
		return;
	}
	*climb = dx * cosf(lat) * cosf(lon) + dy * cosf(lat) * sinf(lon) + dz * sinf(lat);
	const float vn = -dx * sinf(lat) * cosf(lon) - dy * sinf(lat) * sinf(lon) + dz * cosf(lat);
	const float ve = -dx * sinf(lon) + dy * cosf(lon);
	*speed = sqrtf(vn * vn + ve * ve);
	*heading = (float)(atan2f(ve, vn) * 180 / wgs84::PI);
	if (*heading < 0) *heading += 360;
}

inline int32_t le32(const uint8_t *p) { return (int32_t)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
inline int16_t le16(const uint8_t *p) { return (int16_t)(p[0] | p[1] << 8); }

/* Frame record -> SondeData fragment.  RS41: the subframe walk of rs41.c:157-175 ({type, len, data[len],
 * crc16 LE}; CRC-valid subframes are parsed even if RS failed) with the status (rs41.c:219-236), GPS position
 * (rs41.c:249-266, parser.c:193-226, gps/ecef.c) and GPS time (rs41.c:267-273, gps/time.c:8-11) subframes.
 * PTU and XDATA need the 816-byte calibration state. */
inline void fragment_from_record(const sonde_frame_rec &r, SondeData *dst)
{
	memset(dst, 0, sizeof(*dst));
	switch (r.type) {
	case SONDE_RS41: {
		const uint8_t *data = r.data + 57;
		const int data_len = 263 + (r.data[56] == 0xF0 ? 198 : 0);
		int off = 0;
		const uint8_t *sf = data;
		off += sf[1] + 4;
		while (off < data_len && sf[1]) {
			const uint16_t want = (uint16_t)(sf[2 + sf[1]] | sf[3 + sf[1]] << 8);
			if (crc16_ccitt_false(sf + 2, sf[1]) == want) {
				const uint8_t *d = sf + 2;
				switch (sf[0]) {
				case 0x79:                                       /* RS41_SFTYPE_INFO */
					dst->seq = d[0] | d[1] << 8;
					memcpy(dst->serial, d + 2, 8);
					dst->serial[8] = 0;
					dst->fields |= DATA_SEQ | DATA_SERIAL;
					break;
				case 0x7B: {                                     /* RS41_SFTYPE_GPSPOS: ECEF cm, cm/s */
					const float x = (float)(le32(d) / 100.0), y = (float)(le32(d + 4) / 100.0), z = (float)(le32(d + 8) / 100.0);
					const float dx = (float)(le16(d + 12) / 100.0), dy = (float)(le16(d + 14) / 100.0), dz = (float)(le16(d + 16) / 100.0);
					dst->fields |= DATA_POS | DATA_SPEED;
					ecef_to_lla(&dst->lat, &dst->lon, &dst->alt, x, y, z);
					ecef_velocity(&dst->speed, &dst->heading, &dst->climb, dst->lat, dst->lon, dx, dy, dz);
					break;
				}
				case 0x7C: {                                     /* RS41_SFTYPE_GPSINFO: GPS week + ms of week */
					const uint16_t week = (uint16_t)(d[0] | d[1] << 8);
					const uint32_t ms = (uint32_t)le32(d + 2);
					dst->time = (time_t)(ms / 1000UL) + (86400UL * 7) * week + 315964800UL;
					dst->fields |= DATA_TIME;
					break;
				}
				default:
					break;
				}
			}
			sf = data + off;
			off += sf[1] + 4;
		}
		break;
	}
	case SONDE_M10:
		if (r.ok && r.data[4] == 0x9F) { dst->seq = r.data[103]; dst->fields |= DATA_SEQ; }
		break;
	case SONDE_MRZN1:
		if (r.ok) { dst->seq = r.data[4]; dst->fields |= DATA_SEQ; }
		break;
	default:
		break;
	}
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
		m_stage = (float *)sonde_b200_host_alloc((size_t)in.size() * max_chunk * 2 * sizeof(float));
		if (!m_stage) throw std::runtime_error("pinned staging allocation failed");
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
		if (m_stage) sonde_b200_host_free(m_stage);
		m_stage = nullptr;
		sonde_b200_destroy(m_h);
		m_h = nullptr;
	}

	int run() override
	{
		const size_t C = m_in.size();
		int count = -1;
		for (size_t c = 0; c < C; c++) {
			const int n = m_in[c]->read();
			if (n < 0) return -1;
			if (count < 0 || n < count) count = n;          /* channels are fed in lock step */
		}
		if (count > m_max_chunk) count = m_max_chunk;
		for (size_t c = 0; c < C; c++)
			memcpy(m_stage + c * (size_t)count * 2, m_in[c]->readBuf, (size_t)count * sizeof(dsp::complex_t));

		if (count > 0) {
			if (sonde_b200_process_iq(m_h, m_stage, (size_t)count) != SONDE_OK ||
			    sonde_b200_fetch(m_h, m_recs.data(), m_counts.data()) != SONDE_OK)
				throw std::runtime_error(std::string("sonde_b200: ") + sonde_b200_last_error(m_h));
			for (size_t c = 0; c < C; c++) {
				for (int k = 0; k < m_counts[c]; k++) {
					const sonde_frame_rec &r = m_recs[c * m_max_frames + k];
					if (m_fcb) m_fcb((int)c, &r, m_fctx);
					SondeData fragment;
					fragment_from_record(r, &fragment);
					merge_fragment(m_data[c], fragment);
					if (fragment.fields && m_cb) m_cb(&m_data[c], m_ctx);
				}
			}
		}
		for (auto *s : m_in) s->flush();
		return 0;
	}

	sonde_b200 *handle() { return m_h; }

private:
	std::vector<dsp::stream<dsp::complex_t> *> m_in;
	std::vector<int32_t> m_types;
	SondeCallback m_cb = nullptr;
	FrameCallback m_fcb = nullptr;
	void *m_ctx = nullptr, *m_fctx = nullptr;
	sonde_b200 *m_h = nullptr;
	int m_max_frames = 0, m_max_chunk = 0;
	std::vector<sonde_frame_rec> m_recs;
	std::vector<int32_t> m_counts;
	std::vector<SondeFullData> m_data;
	float *m_stage = nullptr;
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
