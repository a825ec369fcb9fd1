/*
 * telemetry.hpp — frame bytes -> SondeData, host side.
 *
 * One radiosonde::Telemetry object per channel carries the state the reference keeps in its decoder
 * structs (calibration fragments, partially assembled output, previous position).  parse() takes the
 * frame record the GPU path produced (sonde_frame_rec: post-FEC frame bytes + gate) and fills one
 * SondeData exactly where the reference's xxx_decode() would (fields == 0: nothing decodable).
 *
 * Covered: RS41 incl. PTU, XDATA ozone and the 51-fragment calibration image (rs41.c:126-323, rs41/parser.c),
 * DFM06/09/17 (dfm09.c:69-237, dfm09/parser.c), M10 and M20 (m10.c:104-183, m10/parser.c), iMS-100 and RS-11G
 * (ims100.c:81-346, ims100/parser.c), MRZ-N1 (mrzn1.c:93-142, mrz-n1/parser.c), iMet-1/4 (imet4.c:91-145,158-226,
 * imet4/parser.c), SRS-C50 (c50.c:82-137, c50/parser.c).  Where the reference indexes its state with an unchecked
 * value from the frame (calibration fragment numbers) this bounds the index instead of writing out of range.
 *
 * Float expressions keep the reference's evaluation types (float vs double constants) so that the values
 * agree to rounding; tests/test_telemetry.py compares against the compiled reference.
 */
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "../../include/sonde_b200.h"
#include "sonde_data.hpp"

namespace radiosonde {

namespace tl {

inline uint16_t crc16_msb(uint16_t crc, const uint8_t *p, size_t n)
{
	for (; n; n--) {
		crc ^= (uint16_t)(*p++ << 8);
		for (int i = 0; i < 8; i++) crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x1021) : (uint16_t)(crc << 1);
	}
	return crc;
}

inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline int32_t le32(const uint8_t *p) { return (int32_t)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
inline int16_t le16(const uint8_t *p) { return (int16_t)(p[0] | p[1] << 8); }
inline int16_t be16(const uint8_t *p) { return (int16_t)(p[0] << 8 | p[1]); }
inline float f32le(const uint8_t *p) { float f; memcpy(&f, p, 4); return f; }
inline float f32be(const uint8_t *p) { const uint32_t u = be32(p); float f; memcpy(&f, &u, 4); return f; }

constexpr double PI = 3.14159265358979323846;
constexpr double WA = 6378137.0, WF = 1 / 298.257223563, WB = WA * (1 - WF);
constexpr double WE2 = (WA * WA - WB * WB) / (WA * WA), WEP2 = (WA * WA - WB * WB) / (WB * WB);

/* WGS-84 ECEF -> geodetic, Bowring's one-step method (SD/gps/ecef.c:6-33) */
inline void ecef_to_lla(float *lat, float *lon, float *alt, float x, float y, float z)
{
	const float lambda = atan2f(y, x);
	const float p = sqrtf(x * x + y * y);
	const float theta = atan2f((float)(z * WA), (float)(p * WB));
	const float st = sinf(theta), ct = cosf(theta);
	if (x == 0 || y == 0 || z == 0) {
		*lat = *lon = *alt = NAN;
		return;
	}
	const float phi = atan2f((float)(z + WEP2 * WB * (st * st * st)), (float)(p - WE2 * WA * (ct * ct * ct)));
	const float sp = sinf(phi);
	const float n = (float)(WA / sqrtf((float)(1 - WE2 * sp * sp)));
	*lat = (float)(phi * 180 / PI);
	*lon = (float)(lambda * 180 / PI);
	*alt = p / cosf(phi) - n;
}

/* ECEF velocity -> horizontal speed, heading, climb (SD/gps/ecef.c:35-57) */
inline void ecef_velocity(float *speed, float *heading, float *climb, float lat, float lon, float dx, float dy, float dz)
{
	lat = (float)(lat * (PI / 180));
	lon = (float)(lon * (PI / 180));
	if (dx == 0 && dy == 0 && dz == 0) {
		*speed = *heading = *climb = 0;
		return;
	}
	*climb = dx * cosf(lat) * cosf(lon) + dy * cosf(lat) * sinf(lon) + dz * sinf(lat);
	const float vn = -dx * sinf(lat) * cosf(lon) - dy * sinf(lat) * sinf(lon) + dz * cosf(lat);
	const float ve = -dx * sinf(lon) + dy * cosf(lon);
	*speed = sqrtf(vn * vn + ve * ve);
	*heading = (float)(atan2f(ve, vn) * 180 / PI);
	if (*heading < 0) *heading += 360;
}

/* geodetic -> ECEF (SD/gps/ecef.c:59-75) */
inline void lla_to_ecef(float *x, float *y, float *z, float lat, float lon, float alt)
{
	lat = (float)(lat / (180 / PI));
	lon = (float)(lon / (180 / PI));
	const float sp = sinf(lat);
	const float n = (float)(WA / sqrtf((float)(1 - WE2 * sp * sp)));
	*x = (n + alt) * cosf(lat) * cosf(lon);
	*y = (n + alt) * cosf(lat) * sinf(lon);
	*z = (float)((1 - WE2) * (n + alt) * sinf(lat));
}

/* US standard atmosphere, layer-wise barometric formula, hPa (SD/physics.c:5-39) */
inline float altitude_to_pressure(float alt)
{
	static const float hb[] = {0.0f, 11000.0f, 20000.0f, 32000.0f, 47000.0f, 51000.0f, 77000.0f};
	static const float Lb[] = {-0.0065f, 0.0f, 0.001f, 0.0028f, 0.0f, -0.0028f, -0.002f};
	static const float Pb[] = {101325.0f, 22632.1f, 5474.89f, 868.02f, 110.91f, 66.94f, 3.96f};
	static const float Tb[] = {288.15f, 216.65f, 216.65f, 228.65f, 270.65f, 270.65f, 214.65f};
	const float g0 = 9.80665f, M = 0.0289644f, R = 8.3144598f;
	int b = 0;
	while (b < 6 && !(alt < hb[b + 1])) b++;
	if (Lb[b] != 0) return (float)(1e-2 * Pb[b] * powf((Tb[b] + Lb[b] * (alt - hb[b])) / Tb[b], -(g0 * M) / (R * Lb[b])));
	return (float)(1e-2 * Pb[b] * expf(-g0 * M * (alt - hb[b]) / (R * Tb[b])));
}

/* saturation vapour pressure over water, hPa (SD/physics.c:63-86): Hyland-Wexler with a temperature correction */
inline float wv_sat_pressure(float temp)
{
	const float c[] = {(float)-0.493158, (float)(1.0 + 4.6094296e-3), (float)-1.3746454e-5, (float)1.2743214e-8};
	temp = (float)(temp + 273.15);
	float T = 0;
	for (int i = 3; i >= 0; i--) {
		T *= temp;
		T += c[i];
	}
	const float p = expf((float)(-5800.2206 / T + 1.3914993 + 6.5459673 * logf(T) - 4.8640239e-2 * T + 4.1764768e-5 * T * T
	                             - 1.4452093e-8 * T * T * T));
	return (float)(p / 100.0);
}

/* MSB-first merge of nbits starting at a byte boundary (SD/bitops.c:32-43).  Like the reference it always ORs in the
 * top (nbits % 8) + 1 bits of the byte that follows the whole bytes, i.e. for nbits = 8k it reads one byte more and
 * takes its most significant bit into bit 0 of the result. */
inline uint64_t merge_bits(const uint8_t *p, int nbits)
{
	uint64_t v = 0;
	for (; nbits >= 8; nbits -= 8) v = v << 8 | *p++;
	return v << nbits | (uint64_t)(*p >> (7 - nbits));
}

/* Microsoft binary format, little endian -> float (SD/bitops.c:132-152) */
inline float mbf_le(const uint8_t *p)
{
	const uint32_t u = (uint32_t)(p[2] >> 7) << 31 | (uint32_t)(uint8_t)(p[3] - 2) << 23 | (uint32_t)(p[2] & 0x7F) << 16 | (uint32_t)p[1] << 8 | p[0];
	float f;
	memcpy(&f, &u, 4);
	return f;
}

/* cubic Hermite spline through (xs, ys) with finite-difference tangents, evaluated at x; -1 outside the table or in its
 * first interval (SD/utils.c:41-65,88-101) */
inline float spline_slope(const float *xs, const float *ys, int k)
{
	if (k == 0) return (ys[1] - ys[0]) / (xs[1] - xs[0]);
	return (float)(0.5 * ((ys[k + 1] - ys[k]) / (xs[k + 1] - xs[k]) + (ys[k] - ys[k - 1]) / (xs[k] - xs[k - 1])));
}
inline float cubic_spline(const float *xs, const float *ys, float count, float x)
{
	for (int i = 1; i < count - 1; i++) {
		if ((x - xs[i + 1]) * (x - xs[i]) <= 0) {
			const float m0 = spline_slope(xs, ys, i), m1 = spline_slope(xs, ys, i + 1);
			const float t = (x - xs[i]) / (xs[i + 1] - xs[i]);
			const float h00 = (1 + 2 * t) * (1 - t) * (1 - t), h10 = t * (1 - t) * (1 - t);
			const float h01 = t * t * (3 - 2 * t), h11 = t * t * (t - 1);
			return h00 * ys[i] + h10 * (xs[i + 1] - xs[i]) * m0 + h01 * ys[i + 1] + h11 * (xs[i + 1] - xs[i]) * m1;
		}
	}
	return -1;
}

/* days since the epoch used by the reference's own timegm (SD/utils.c:42-52,68-95) */
inline unsigned day_number(unsigned year, unsigned month, unsigned day)
{
	static const unsigned cum[2][12] = {{0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334},
	                                    {0, 31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335}};
	int leap = (!(year % 4) && year % 100) || !(year % 400);
	unsigned n = cum[leap][month % 12] + day;
	while (n >= 365U + leap) {
		year++;
		n -= 365U + leap;
		leap = (!(year % 4) && year % 100) || !(year % 400);
	}
	n += 365 * year + (year / 4) - (year / 100) + (year / 400);
	return n - 1;
}

inline time_t utc_seconds(const struct tm *tm)
{
	time_t t = tm->tm_sec + tm->tm_min * 60 + tm->tm_hour * 3600;
	t += 86400UL * (day_number(1900 + tm->tm_year, tm->tm_mon, tm->tm_mday) - day_number(1970, 0, 1));
	return t;
}

inline time_t gps_to_utc(uint16_t week, uint32_t ms) { return (time_t)(ms / 1000UL) + (86400UL * 7) * week + 315964800UL; }

/* NTC on the Meteomodem boards: divider with range-switched bias / parallel resistors (m10/parser.c:62-113,188-233) */
inline float meteomodem_ntc(float beta, unsigned adc_val, unsigned range)
{
	const float r0 = 15000.0f, t0 = 273.15f;
	const float rinf = r0 * expf(-beta / t0);
	const float bias[] = {12.1e3f, 36.5e3f, 475e3f};
	const float par[] = {3.402823466e+38F, 330e3f, 2e6f};
	const float pct = adc_val / (float)((1 << 12) - 1);
	float r;
	switch (range) {
	case 0: r = pct * bias[0] / (1 - pct); break;
	case 1:
	case 2: r = pct * bias[range] * par[range] / (par[range] - pct * (bias[range] + par[range])); break;
	default: r = rinf; break;
	}
	return (float)(beta / logf(r / rinf) - 273.15);
}

}  // namespace tl

class Telemetry {
public:
	explicit Telemetry(int type = SONDE_RS41) { reset(type); }
	int type() const { return m_type; }

	void reset(int type)
	{
		m_type = type;
		const time_t zero = 0;
		memset(&m_partial, 0, sizeof(m_partial));
		memset(m_mrz_calib, 0, sizeof(m_mrz_calib));
		static const uint8_t rs41_default[816] = {
#include "rs41_default_calib.inc"
		};
		memcpy(m_rs41_calib, rs41_default, sizeof(m_rs41_calib));
		memset(m_rs41_have, 0, sizeof(m_rs41_have));
		memset(m_dfm_raw, 0, sizeof(m_dfm_raw));
		m_dfm_serial_ch = 0;
		m_dfm_serial = 0;
		m_dfm_tm = *gmtime(&zero);
		memset(&m_dfm_tm, 0, sizeof(m_dfm_tm));
		memset(m_ims_calib, 0, sizeof(m_ims_calib));
		m_ims_mask = 0;
		m_ims_date = 0;
		m_ims_prev_alt = 0;
		m_ims_prev_time = -1;
		m_adc_ref = m_adc_temp = m_adc_rh = m_adc_rh_temp = 0;
		m_mrz_mask = 0;
		m_c50_tm = *gmtime(&zero);
		m_imet_prev[0] = m_imet_prev[1] = m_imet_prev[2] = 0;
		m_imet_prev_time = 0;
	}

	/* one framer window -> one SondeData (what xxx_decode() leaves in *dst on PARSED) */
	void parse(const sonde_frame_rec &r, SondeData *dst)
	{
		memset(dst, 0, sizeof(*dst));
		switch (r.type) {
		case SONDE_RS41:  rs41(r, dst); break;
		case SONDE_M10:   if (r.ok) m10(r.data, dst); break;
		case SONDE_MRZN1: if (r.ok) mrzn1(r.data, dst); break;
		case SONDE_IMET4: imet4(r, dst); break;
		case SONDE_C50:   if (r.ok) c50(r.data, dst); break;
		case SONDE_DFM09: if (r.ok) dfm09(r.data + 64, dst); break;
		case SONDE_IMS100:
			if (r.ok) {
				if (r.data[80 + 15] == 0xA2) rs11g(r.data + 80, dst);
				else if (r.data[80 + 15] == 0xC1) ims100(r.data + 80, dst);
			}
			break;
		default: break;
		}
	}

private:
	/* ---- RS41: CRC-valid subframes are parsed even if RS failed (rs41.c:157-175) ------------------------ */
	void rs41(const sonde_frame_rec &r, SondeData *dst)
	{
		const uint8_t *data = r.data + 57;
		const int data_len = 263 + (r.data[56] == 0xF0 ? 198 : 0);
		const uint8_t *sf = data;
		int off = sf[1] + 4;
		while (off < data_len && sf[1]) {
			const uint16_t want = (uint16_t)(sf[2 + sf[1]] | sf[3 + sf[1]] << 8);
			if (tl::crc16_msb(0xFFFF, sf + 2, sf[1]) == want) rs41_subframe(sf, dst);
			sf = data + off;
			off += sf[1] + 4;
		}
	}

	/* calibration image accessors: packed little-endian floats at fixed offsets (rs41/protocol.h:155-181) */
	enum { RS41_T_REF = 61, RS41_RH_REF = 69, RS41_T_POLY = 77, RS41_T_COEFF = 89, RS41_RH_CAP = 117, RS41_RH_COEFF = 125,
	       RS41_TH_POLY = 293, RS41_TH_COEFF = 305, RS41_P_COEFF = 606, RS41_BURSTKILL = 800, RS41_NFRAG = 51 };
	float cal(int base, int i) const { return tl::f32le(m_rs41_calib + base + 4 * i); }
	static float adc24(const uint8_t *p) { return (float)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16); }

	void rs41_subframe(const uint8_t *sf, SondeData *dst)
	{
		const uint8_t *d = sf + 2;
		switch (sf[0]) {
		case 0x79: {                                         /* status: sequence number, serial, one calibration fragment */
			const unsigned frag = d[23];
			if (frag < RS41_NFRAG) {                             /* (the reference does not bound this index) */
				memcpy(m_rs41_calib + 16 * frag, d + 24, 16);
				m_rs41_have[frag / 8] |= (uint8_t)(1 << (7 - frag % 8));
			}
			dst->fields |= DATA_SERIAL | DATA_SEQ;
			memset(dst->serial, 0, sizeof(dst->serial));
			for (int i = 0; i < 8 && d[2 + i]; i++) dst->serial[i] = (char)d[2 + i];
			dst->seq = d[0] | d[1] << 8;
			const uint16_t burstkill = (uint16_t)(m_rs41_calib[RS41_BURSTKILL] | m_rs41_calib[RS41_BURSTKILL + 1] << 8);
			if (burstkill != 0xFFFF) {
				dst->fields |= DATA_SHUTDOWN;
				dst->shutdown = burstkill;
			}
			break;
		}
		case 0x7A: {                                         /* PTU (rs41/parser.c:10-194) */
			dst->fields |= DATA_PTU;
			dst->temp = rs41_temp(d);
			dst->rh = rs41_humidity(d);
			dst->pressure = rs41_pressure(d);
			int have = 0;
			for (int i = 0; i < 7; i++) have += __builtin_popcount(m_rs41_have[i]);
			dst->calib_percent = (float)(((float)have * 100.0) / RS41_NFRAG);
			break;
		}
		case 0x7B: {                                         /* GPS position: ECEF cm, cm/s */
			const float x = (float)(tl::le32(d) / 100.0), y = (float)(tl::le32(d + 4) / 100.0), z = (float)(tl::le32(d + 8) / 100.0);
			const float dx = (float)(tl::le16(d + 12) / 100.0), dy = (float)(tl::le16(d + 14) / 100.0), dz = (float)(tl::le16(d + 16) / 100.0);
			dst->fields |= DATA_POS | DATA_SPEED;
			tl::ecef_to_lla(&dst->lat, &dst->lon, &dst->alt, x, y, z);
			tl::ecef_velocity(&dst->speed, &dst->heading, &dst->climb, dst->lat, dst->lon, dx, dy, dz);
			break;
		}
		case 0x7C:                                           /* GPS week + ms of week */
			dst->time = tl::gps_to_utc((uint16_t)(d[0] | d[1] << 8), (uint32_t)tl::le32(d + 2));
			dst->fields |= DATA_TIME;
			break;
		case 0x7E:                                           /* XDATA: ASCII hex after one unknown byte */
			if (!(dst->pressure > 0)) dst->pressure = tl::altitude_to_pressure(dst->alt);
			rs41_xdata(dst, (const char *)d + 1, sf[1] - 1);
			break;
		default:
			break;
		}
	}

	/* thermistor channel: ratiometric ADC reading -> resistance -> 2nd-degree polynomial (parser.c:10-48,110-140) */
	float rs41_resist_temp(const uint8_t *m, int poly, int coeff0) const
	{
		const float a = adc24(m), r1 = adc24(m + 3), r2 = adc24(m + 6);
		const float ratio = (a - r1) / (r2 - r1);
		const float ohm = (cal(RS41_T_REF, 0) + (cal(RS41_T_REF, 1) - cal(RS41_T_REF, 0)) * ratio) * cal(coeff0, 0);
		return cal(poly, 0) + cal(poly, 1) * ohm + cal(poly, 2) * ohm * ohm;
	}
	float rs41_temp(const uint8_t *d) const
	{
		if (adc24(d + 6) - adc24(d + 3) == 0) return NAN;
		const float tu = rs41_resist_temp(d, RS41_T_POLY, RS41_T_COEFF);
		float tc = 0;
		for (int i = 6; i > 0; i--) {
			tc *= tu;
			tc += cal(RS41_T_COEFF, i);
		}
		return tc + tu;
	}
	float rs41_temp_humidity(const uint8_t *d) const
	{
		if (adc24(d + 24) - adc24(d + 21) == 0) return NAN;
		if (!cal(RS41_T_REF, 0) || !cal(RS41_T_REF, 1)) return NAN;
		return rs41_resist_temp(d + 18, RS41_TH_POLY, RS41_TH_COEFF);
	}
	float rs41_humidity(const uint8_t *d) const
	{
		const float a = adc24(d + 9), r1 = adc24(d + 12), r2 = adc24(d + 15);
		if (r2 - r1 == 0) return NAN;
		const float th_raw = rs41_temp_humidity(d);
		const float t_air = rs41_temp(d);
		float th = 0;
		for (int i = 6; i > 0; i--) {
			th *= th_raw;
			th += cal(RS41_TH_COEFF, i);
		}
		th += th_raw;
		const float ratio = (a - r1) / (r2 - r1);
		const float cap = cal(RS41_RH_REF, 0) + ratio * (cal(RS41_RH_REF, 1) - cal(RS41_RH_REF, 0));
		const float cc = (cap / cal(RS41_RH_CAP, 0) - 1) * cal(RS41_RH_CAP, 1);
		th = (th - 20) / 180;
		float rh = 0, f1 = 1;
		for (int i = 0; i < 7; i++) {
			float f2 = 1;
			for (int j = 0; j < 6; j++) {
				rh += f1 * f2 * cal(RS41_RH_COEFF, 6 * i + j);
				f2 *= th;
			}
			f1 *= cc;
		}
		const float out = rh * tl::wv_sat_pressure(th_raw) / tl::wv_sat_pressure(t_air);
		const float hi = (100 < out) ? 100 : out;
		return (0 > hi) ? 0 : hi;
	}
	float rs41_pressure(const uint8_t *d) const
	{
		const float a = adc24(d + 27), r1 = adc24(d + 30), r2 = adc24(d + 33);
		const float pt = (float)(tl::le16(d + 39) / 100.0);
		if (r2 - r1 == 0) return NAN;
		if (a == 0) return NAN;
		float q = (a - r1) / (r2 - r1);
		float poly[6];
		for (int k = 0; k < 3; k++)
			poly[k] = cal(RS41_P_COEFF, k) + cal(RS41_P_COEFF, 7 + k) * pt + cal(RS41_P_COEFF, 11 + k) * pt * pt
			          + cal(RS41_P_COEFF, 15 + k) * pt * pt * pt;
		poly[3] = cal(RS41_P_COEFF, 3) + cal(RS41_P_COEFF, 10) * pt + cal(RS41_P_COEFF, 14) * pt * pt;
		poly[4] = cal(RS41_P_COEFF, 4);
		poly[5] = cal(RS41_P_COEFF, 5);
		q = cal(RS41_P_COEFF, 6) / q;
		return poly[0] + poly[1] * q + poly[2] * q * q + poly[3] * q * q * q + poly[4] * q * q * q * q
		       + poly[5] * q * q * q * q * q;
	}

	/* XDATA instrument chain (rs41.c:291-323).  Mirrors the reference's scan, including that only the 4-character
	 * instrument header is charged against `len`, so the scan may run into the bytes that follow the subframe. */
	static void rs41_xdata(SondeData *dst, const char *ascii, int len)
	{
		unsigned pump_t = 0, cur = 0, batt = 0, pump_i = 0, ext = 0, id = 0, num = 0;
		while (len > 0) {
			sscanf(ascii, "%02X%02X", &id, &num);
			ascii += 4;
			len -= 4;
			if (id != 0x05) continue;                            /* ENSCI ozone */
			if (sscanf(ascii, "%04X%05X%02X%03X%02X", &pump_t, &cur, &batt, &pump_i, &ext) == 5) {
				ascii += 16;
				const float kelvin = (float)((pump_t & 0x8000 ? -1 : 1) * 0.001 * (pump_t & 0x7FFF) + 273.15);
				const float ua = (float)(cur * 1e-5);
				dst->fields |= DATA_OZONE;
				dst->o3_mpa = (float)(4.307e-3 * ua * kelvin * 30.0f);
			} else {
				ascii += 17;
			}
		}
	}

	/* ---- M10 (type 0x9F) and M20 (type 0x20): big-endian fields (m10/protocol.h:27-82) ------------------- */
	void m10(const uint8_t *f, SondeData *dst)
	{
		float dx, dy, dz;
		if (f[4] == 0x9F) {
			/* M10Frame_9f: dlat 7, dlon 9, dalt 11, time 13, lat 17, lon 21, alt 25, week 35, rh_ref 53,
			 * rh_counts 56, adc_temp_range 65, adc_temp_val 66, serial 96 */
			const uint32_t s0 = (f[98] >> 4) * 100 + (f[98] & 0xF), s1 = f[96], s2 = f[99] | f[100] << 8;
			sprintf(dst->serial, "M10-%03d-%d-%1d%04d", s0, s1, s2 >> 13, s2 & 0x1FFF);
			dst->time = tl::gps_to_utc((uint16_t)(f[35] << 8 | f[36]), tl::be32(f + 13));
			dst->lat = (float)((int32_t)tl::be32(f + 17) * 360.0 / ((uint64_t)1UL << 32));
			dst->lon = (float)((int32_t)tl::be32(f + 21) * 360.0 / ((uint64_t)1UL << 32));
			dst->alt = (float)((int32_t)tl::be32(f + 25) / 1e3);
			dx = (float)(tl::be16(f + 9) / 200.0);       /* dlon */
			dy = (float)(tl::be16(f + 7) / 200.0);       /* dlat */
			dz = (float)(tl::be16(f + 11) / 200.0);
			dst->fields |= DATA_SERIAL | DATA_TIME | DATA_POS | DATA_SPEED | DATA_PTU;
			dst->speed = sqrtf(dx * dx + dy * dy);
			dst->climb = dz;
			dst->heading = (float)(atan2f(dy, dx) * 180.0 / tl::PI);
			if (dst->heading < 0) dst->heading = (float)(dst->heading + 360.0);
			dst->calib_percent = 100.0f;
			dst->pressure = tl::altitude_to_pressure(dst->alt);
			const float temp = tl::meteomodem_ntc(3100.0f, (f[66] | f[67] << 8) & 0xFFF, f[65]);
			dst->temp = temp;
			const float rh_counts = (float)(f[58] << 16 | f[57] << 8 | f[56]);
			const float rh_ref = (float)(f[55] << 16 | f[54] << 8 | f[53]);
			const float corr = 1.0f - 400.0e-6f * temp;
			const float rh = (float)((rh_counts * corr / rh_ref - 0.8955) / 0.002);
			const float hi = (100 < rh) ? 100 : rh;
			dst->rh = (0 > hi) ? 0 : hi;
		} else if (f[4] == 0x20) {
			/* M20Frame_20: adc_temp 7, alt 11, dlat 14, dlon 16, time 18, sn 21, seq 24, dalt 27, week 29,
			 * lat 31, lon 35 */
			const uint32_t raw = f[23] << 16 | f[22] << 8 | f[21];
			const uint8_t s0 = raw & 0x3F, s1 = (raw >> 6) & 0x0F;
			const uint16_t s2 = (uint16_t)(raw >> 10);
			sprintf(dst->serial, "M20-%01d%02d-%d-%05d", s0 / 12, s0 % 12 + 1, s1, s2);
			dst->seq = f[24];
			uint32_t ms = f[18] << 16 | f[19] << 8 | f[20];
			ms *= 1000;
			dst->time = tl::gps_to_utc((uint16_t)(f[29] << 8 | f[30]), ms);
			dst->lat = (float)((int32_t)tl::be32(f + 31) / 1e6);
			dst->lon = (float)((int32_t)tl::be32(f + 35) / 1e6);
			dst->alt = (float)((int32_t)(f[11] << 16 | f[12] << 8 | f[13]) / 1e2);
			dx = (float)(tl::be16(f + 16) / 100.0);      /* dlon */
			dy = (float)(tl::be16(f + 14) / 100.0);      /* dlat */
			dz = (float)(tl::be16(f + 27) / 100.0);
			dst->fields |= DATA_SERIAL | DATA_SEQ | DATA_TIME | DATA_POS | DATA_SPEED | DATA_PTU;
			dst->climb = dz;
			dst->speed = sqrtf(dx * dx + dy * dy);
			dst->heading = (float)(atan2f(dy, dx) * 180.0 / tl::PI);
			if (dst->heading < 0) dst->heading += 360;
			dst->calib_percent = 100.0f;
			const unsigned adc = f[7] | f[8] << 8;
			dst->temp = tl::meteomodem_ntc(3450.0f, adc & 0xFFF, adc >> 12);
			dst->rh = 0;
			dst->pressure = tl::altitude_to_pressure(dst->alt);
		}
	}

	/* ---- MRZ-N1: little-endian packed frame, 16 calibration fragments of 4 bytes (mrz-n1/protocol.h:20-48) */
	void mrzn1(const uint8_t *f, SondeData *dst)
	{
		/* update_calibration(seq = calib_frag_seq - 1) */
		const int seq = f[44] - 1;
		if (seq >= 0 && seq < 16) {                              /* (the reference does not bound this index) */
			memcpy(m_mrz_calib + 4 * seq, f + 45, 4);
			m_mrz_mask |= (uint16_t)(1 << (16 - seq - 1));
		}
		/* MRZN1Calibration: 9 floats, serial @36, 4 unknown words, cal_date @56, date @60 */
		uint32_t serial, cal_date, date;
		memcpy(&serial, m_mrz_calib + 36, 4);
		memcpy(&cal_date, m_mrz_calib + 56, 4);
		memcpy(&date, m_mrz_calib + 60, 4);

		dst->seq = f[4] & 0x0F;
		if (date) {
			struct tm tm;
			memset(&tm, 0, sizeof(tm));
			tm.tm_hour = f[5]; tm.tm_min = f[6]; tm.tm_sec = f[7];
			tm.tm_year = 2000 + date % 100 - 1900;
			tm.tm_mon = (date / 100) % 100 - 1;
			tm.tm_mday = date / 10000;
			dst->time = tl::utc_seconds(&tm);
		} else {
			dst->time = 3600 * f[5] + 60 * f[6] + f[7];
		}
		const float x = tl::le32(f + 9) / 100.0f, y = tl::le32(f + 13) / 100.0f, z = tl::le32(f + 17) / 100.0f;
		const float dx = tl::le16(f + 21) / 100.0f, dy = tl::le16(f + 23) / 100.0f, dz = tl::le16(f + 25) / 100.0f;
		tl::ecef_to_lla(&dst->lat, &dst->lon, &dst->alt, x, y, z);
		tl::ecef_velocity(&dst->speed, &dst->heading, &dst->climb, dst->lat, dst->lon, dx, dy, dz);
		dst->calib_percent = 100.0f * __builtin_popcount(m_mrz_mask) / 16;
		dst->temp = tl::le16(f + 30) / 100.0f;
		dst->rh = tl::le16(f + 32) / 100.0f;
		dst->pressure = tl::altitude_to_pressure(dst->alt);
		dst->fields |= DATA_SEQ | DATA_TIME | DATA_POS | DATA_SPEED | DATA_PTU;
		if ((m_mrz_mask & (0x40 | 0x02)) == (0x40 | 0x02)) {
			sprintf(dst->serial, "MRZ-H1%02d%05d", (int)(cal_date % 100) - 10, serial);
			dst->fields |= DATA_SERIAL;
		}
	}

	/* ---- iMet-1/4: walk the SOH/type subframes, AUG-CCITT residue 0 (imet4.c:91-145) -------------------- */
	void imet4(const sonde_frame_rec &r, SondeData *dst)
	{
		const uint8_t *fr = r.data;                          /* 60 payload bytes, zero beyond */
		size_t len = 0;
		for (size_t i = 0; i < 72; i += len) {
			const uint8_t *sf = fr + i;
			if (sf[0] != 0x01) break;
			switch (sf[1]) {
			case 1: len = 14; break;
			case 2: len = 18; break;
			case 3: len = 5 + sf[2]; break;
			case 4: len = 20; break;
			case 5: len = 30; break;
			default: len = 0; break;
			}
			if (!len) break;
			if (i + len > SONDE_REC_BYTES) break;
			if (tl::crc16_msb(0x1D0F, sf, len) == 0) imet4_subframe(sf, dst);
			else if (sf[1] == 3) break;
		}
		if (!(dst->fields & DATA_SPEED) && (dst->fields & (DATA_POS | DATA_TIME)) == (DATA_POS | DATA_TIME)) {
			float x, y, z;
			dst->fields |= DATA_SPEED;
			const float dt = (float)(dst->time - m_imet_prev_time);
			tl::lla_to_ecef(&x, &y, &z, dst->lat, dst->lon, dst->alt);
			tl::ecef_velocity(&dst->speed, &dst->heading, &dst->climb, dst->lat, dst->lon, (x - m_imet_prev[0]) / dt,
			                  (y - m_imet_prev[1]) / dt, (z - m_imet_prev[2]) / dt);
			m_imet_prev[0] = x; m_imet_prev[1] = y; m_imet_prev[2] = z;
			m_imet_prev_time = (uint32_t)dst->time;
		}
		if ((dst->fields & (DATA_SEQ | DATA_TIME)) == (DATA_SEQ | DATA_TIME)) {
			time_t t = dst->time - dst->seq;                     /* estimated switch-on time */
			dst->fields |= DATA_SERIAL;
			sprintf(dst->serial, "iMet-%04X", tl::crc16_msb(0x1D0F, (const uint8_t *)&t, 4));
		}
	}

	void imet4_subframe(const uint8_t *sf, SondeData *dst)
	{
		switch (sf[1]) {
		case 1:
		case 4:                                                  /* PTU / PTUX */
			dst->calib_percent = 100.0;
			dst->temp = tl::le16(sf + 7) / 100.0f;
			dst->rh = tl::le16(sf + 9) / 100.0f;
			dst->pressure = (uint32_t)(sf[4] | sf[5] << 8 | sf[6] << 16) / 100.0f;
			dst->seq = (uint16_t)(sf[2] | sf[3] << 8);
			dst->fields |= DATA_PTU | DATA_SEQ;
			break;
		case 2:
		case 5: {                                                /* GPS / GPSX: floats, altitude + 5000 m */
			dst->lat = tl::f32le(sf + 2);
			dst->lon = tl::f32le(sf + 6);
			dst->alt = tl::le16(sf + 10) - 5000.f;
			dst->fields |= DATA_POS | DATA_TIME;
			const uint8_t *hms = sf + (sf[1] == 2 ? 13 : 25);
			dst->time = imet4_time(hms[0], hms[1], hms[2]);
			if (sf[1] == 5) {
				const float dlon = tl::f32le(sf + 13), dlat = tl::f32le(sf + 17);
				dst->speed = sqrtf(dlat * dlat + dlon * dlon);
				float h = (float)(atan2f(dlat, dlon) / 180.0f / tl::PI);
				if (h < 0) h = (float)(h + 360.0);
				dst->heading = h;
				dst->climb = tl::f32le(sf + 21);
				dst->fields |= DATA_SPEED;
			}
			break;
		}
		case 3:
			if (sf[3] == 0x01) {                                 /* ENSCI ozone */
				const uint8_t *o = sf + 5;
				const float cell = (o[0] << 8 | o[1]) / 1000.0f, pump = (o[2] << 8 | o[3]) / 100.0f;
				dst->fields |= DATA_OZONE;
				dst->o3_mpa = (float)(4.307e-3 * cell * pump * 30);
			}
			break;
		default:
			break;
		}
	}

	/* the date is not transmitted: today's date, corrected around 0Z (imet4/parser.c:43-64) */
	static time_t imet4_time(int hour, int min, int sec)
	{
		time_t now = time(nullptr);
		struct tm tm = *gmtime(&now);
		if (abs(hour - tm.tm_hour) >= 12) {
			now += (hour < tm.tm_hour) ? 86400 : -86400;
			tm = *gmtime(&now);
		}
		tm.tm_hour = hour; tm.tm_min = min; tm.tm_sec = sec;
		return tl::utc_seconds(&tm);
	}

	/* ---- SRS-C50: one value per 9-byte frame, output assembled across frames (c50.c:82-137) -------------- */
	void c50(const uint8_t *f, SondeData *dst)
	{
		const uint8_t *d = f + 3;
		const uint32_t raw = tl::be32(d);
		switch (f[2]) {
		case 0x03:
			m_partial.calib_percent = 100.0f;
			m_partial.temp = tl::f32be(d);
			m_partial.fields |= DATA_PTU;
			break;
		case 0x10: m_partial.rh = tl::f32be(d); break;
		case 0x14:
			m_c50_tm.tm_mday = raw / 10000 % 100;
			m_c50_tm.tm_mon = (raw / 100) % 100 - 1;
			m_c50_tm.tm_year = 2000 + raw % 100 - 1900;
			break;
		case 0x15:
			m_c50_tm.tm_hour = raw / 10000;
			m_c50_tm.tm_min = (raw / 100) % 100;
			m_c50_tm.tm_sec = raw % 100;
			m_partial.fields |= DATA_TIME;
			m_partial.time = tl::utc_seconds(&m_c50_tm);
			break;
		case 0x16: m_partial.lat = (float)((int)((int32_t)raw / 1e7) + ((int32_t)raw % 10000000 / 60.0 * 100.0) / 1e7); break;
		case 0x17: m_partial.lon = (float)((int)((int32_t)raw / 1e7) + ((int32_t)raw % 10000000 / 60.0 * 100.0) / 1e7); break;
		case 0x18:
			m_partial.alt = (int32_t)raw / 10.0f;
			m_partial.fields |= DATA_POS;
			m_partial.pressure = tl::altitude_to_pressure(m_partial.alt);
			break;
		case 0x64:
			sprintf(m_partial.serial, "C50-%d", raw);
			m_partial.fields |= DATA_SERIAL;
			break;
		default:
			break;
		}
		memcpy(dst, &m_partial, sizeof(*dst));
		m_partial.fields = 0;
	}

	/* ---- DFM06/09/17: 18-byte unpacked frame = PTU channel (type, 24 bits) + 2 GPS slots (type, 48 bits); the output
	 * record is assembled across frames (dfm09.c:69-237, dfm09/parser.c) ------------------------------------------ */
	void dfm09(const uint8_t *f, SondeData *dst)
	{
		uint8_t fr[20];
		memcpy(fr, f, 18);
		fr[18] = fr[19] = 0;                                   /* what merge_bits() may touch past the frame */
		dfm09_ptu(fr);
		dfm09_gps(fr + 4);
		dfm09_gps(fr + 11);
		m_partial.pressure = tl::altitude_to_pressure(m_partial.alt);
		memcpy(dst, &m_partial, sizeof(*dst));
		m_partial.fields = 0;
	}

	static float dfm09_mantissa(uint32_t raw) { return (float)(raw & 0xFFFFF) / (1 << (raw >> 20)); }

	void dfm09_ptu(const uint8_t *sf)
	{
		const uint8_t type = sf[0];
		const uint32_t ch = (uint32_t)tl::merge_bits(sf + 1, 24);
		m_dfm_raw[type & 15] = ch;
		if ((ch & 0xFFFF) == 0) m_dfm_serial_ch = (uint8_t)(type + 1);     /* the channel before the serial is all zero */
		if (type == 0) {                                       /* thermistor against the two reference channels 3, 4 */
			const uint32_t r1 = m_dfm_raw[3], r2 = m_dfm_raw[4];
			const float bb0 = 3260.0f, t0 = (float)(25 + 273.15), r0 = 5.0e3f, rf = 220e3f;
			const float g = dfm09_mantissa(r2) / rf;
			float res = (dfm09_mantissa(ch) - dfm09_mantissa(r1)) / g;
			float temp = 0;
			if (!ch || !r1 || !r2) res = 0;
			if (res > 0) temp = (float)(1.0 / (1 / t0 + 1 / bb0 * logf(res / r0)) - 273.15);
			m_partial.temp = temp;
			m_partial.fields |= DATA_PTU;
			m_partial.calib_percent = 100.0;
		} else if (type == 1) {
			m_partial.rh = 0;
		} else if (type == m_dfm_serial_ch) {
			if (type == 0x06) {                                  /* DFM06: the serial in one piece */
				m_partial.fields |= DATA_SERIAL;
				sprintf(m_partial.serial, "D%06X", ch);
			} else {                                             /* DFM09/17: four 16-bit shards, index in the low nibble */
				const int idx = 3 - (int)(ch & 0xF);
				const uint64_t shard = (ch >> 4) & 0xFFFF;
				if (idx >= 0) {                                  /* (a negative shift is undefined in the reference) */
					m_dfm_serial &= ~((uint64_t)0xFFFF << (16 * idx));
					m_dfm_serial |= shard << (16 * idx);
				}
				if ((ch & 0xF) == 0) {
					uint64_t v = m_dfm_serial;
					while (v && !(v & 0xFFFF)) v >>= 16;
					m_partial.fields |= DATA_SERIAL;
					sprintf(m_partial.serial, "D%08ld", (long)v);
				}
			}
		}
	}

	void dfm09_gps(const uint8_t *sf)
	{
		const uint8_t *d = sf + 1;
		switch (sf[0]) {
		case 0x00:
			m_partial.fields |= DATA_SEQ;
			m_partial.seq = (int)(uint32_t)tl::merge_bits(d + 3, 8);
			break;
		case 0x01:
			m_dfm_tm.tm_sec = (int)(tl::merge_bits(d + 4, 16) / 1000);
			break;
		case 0x02:
			m_partial.lat = (float)((int32_t)tl::merge_bits(d, 32) / 1e7);
			m_partial.speed = (float)(tl::merge_bits(d + 4, 16) / 1e2);
			break;
		case 0x03:
			m_partial.lon = (float)((int32_t)tl::merge_bits(d, 32) / 1e7);
			m_partial.heading = (float)(tl::merge_bits(d + 4, 16) / 1e2);
			break;
		case 0x04:
			m_partial.alt = (float)((int32_t)tl::merge_bits(d, 32) / 1e2);
			m_partial.climb = (float)((int16_t)tl::merge_bits(d + 4, 16) / 1e2);
			m_partial.fields |= DATA_POS | DATA_SPEED;
			break;
		case 0x08: {
			const uint32_t raw = (uint32_t)tl::merge_bits(d, 32);
			m_dfm_tm.tm_year = (int)((raw >> 20) & 0xFFF) - 1900;
			m_dfm_tm.tm_mon = (int)((raw >> 16) & 0xF) - 1;
			m_dfm_tm.tm_mday = (raw >> 11) & 0x1F;
			m_dfm_tm.tm_hour = (raw >> 6) & 0x1F;
			m_dfm_tm.tm_min = raw & 0x3F;
			m_partial.fields |= DATA_TIME;
			m_partial.time = tl::utc_seconds(&m_dfm_tm);
			break;
		}
		default:
			break;
		}
	}

	/* ---- Meisei iMS-100 / RS-11G: 48 payload bytes + 24-bit validity mask (one bit per 16-bit word, MSB = word 0);
	 * 64 calibration floats arrive one per frame, indexed by seq % 64 (ims100.c:129-346, ims100/parser.c) -------- */
	static bool has(uint32_t valid, uint32_t mask) { return (valid & mask) == mask; }
	static uint16_t be16u(const uint8_t *p) { return (uint16_t)(p[0] << 8 | p[1]); }
	static int32_t be32s(const uint8_t *p) { return (int32_t)tl::be32(p); }

	/* oscillator count ratio -> thermistor resistance polynomial -> spline over the calibration points */
	static float meisei_temp(float cnt, float ref, const float *poly, const float *ohms, const float *temps, size_t n)
	{
		const float fc = (float)(4.0 * cnt / ref);
		const float x = (float)(1.0 / (fc - 1.0));
		const float r = poly[0] + poly[1] * x + poly[2] * x * x + poly[3] * x * x * x;
		float lg[12];
		for (size_t i = 0; i < n; i++) lg[i] = logf(ohms[i]);
		const float t = tl::cubic_spline(lg, temps, (float)n, logf(r));
		const float hi = (100 < t) ? 100 : t;
		return (-100 > hi) ? -100 : hi;
	}
	static float meisei_rh(float cnt, float ref, const float *poly)
	{
		const float f = (float)(4.0 * cnt / ref);
		return poly[0] + poly[1] * f + poly[2] * f * f + poly[3] * f * f * f;
	}
	static float clamp_pct(float v)
	{
		const float hi = (100 < v) ? 100 : v;
		return (0 > hi) ? 0 : hi;
	}
	void meisei_common_head(const uint8_t *f, uint32_t valid, SondeData *dst, bool rs11g_frag)
	{
		if (has(valid, 0x800000)) {
			dst->fields |= DATA_SEQ;
			dst->seq = be16u(f);
		}
		if (has(valid, 0x800000 | 0x300000)) {
			const int k = dst->seq % 64;
			float coeff;
			if (rs11g_frag) {
				coeff = tl::mbf_le(f + 4);
			} else {
				const uint8_t swapped[4] = {f[6], f[7], f[4], f[5]};
				coeff = tl::f32be(swapped);
			}
			m_ims_calib[k] = coeff;
			m_ims_mask |= 1ULL << (63 - k);
		}
	}
	void meisei_serial(SondeData *dst, const char *fmt)
	{
		if (dst->fields && (m_ims_mask & 0x8000000000000000ULL)) {
			dst->fields |= DATA_SERIAL;
			sprintf(dst->serial, fmt, (int)m_ims_calib[0]);
		}
	}

	void ims100(const uint8_t *f, SondeData *dst)
	{
		uint32_t valid;
		memcpy(&valid, f + 48, 4);
		const float *cal = m_ims_calib;                        /* IMS100Calibration as 64 floats (ims100/protocol.h:155-176) */
		meisei_common_head(f, valid, dst, false);
		if (has(valid, 0x800000 | 0x460000)) {
			m_adc_temp = be16u(f + 10);
			switch (be16u(f) & 3) {                              /* two of the ADC slots are multiplexed by seq */
			case 0: m_adc_ref = be16u(f + 2); m_adc_rh = be16u(f + 12); break;
			case 1:
			case 2: m_adc_rh = be16u(f + 12); break;
			default: m_adc_rh_temp = be16u(f + 2); m_adc_ref = be16u(f + 12); break;
			}
			const float poly[4] = {cal[53] - cal[56], cal[54], cal[55], 0};
			const float air = 1 + meisei_temp(m_adc_temp, m_adc_ref, poly, cal + 33, cal + 17, 12);
			/* humidity sensor temperature: resistance -> Steinhart-Hart-like cubic in ln R */
			const float fc = (float)(4.0 * m_adc_rh_temp / m_adc_ref);
			float x = (float)(1.0 / (fc - 1.0));
			const float r = poly[0] + poly[1] * x + poly[2] * x * x + poly[3] * x * x * x;
			x = logf(r);
			const float rh_t = 1 + (float)(1.0 / (cal[57] * x * x * x + cal[58] * x + cal[59]) - 273.15);
			float rh = meisei_rh(m_adc_rh, m_adc_ref, cal + 49);
			if (air < 100 && air > -100) rh *= tl::wv_sat_pressure(rh_t) / tl::wv_sat_pressure(air);
			dst->fields |= DATA_PTU;
			dst->temp = air;
			dst->rh = clamp_pct(rh);
			dst->pressure = 0;
			dst->calib_percent = (float)(100.0 * __builtin_popcountll(m_ims_mask) / 64);
		}
		if ((f[14] << 8 | f[15]) == 0x30C1) {                    /* GPS page */
			if (has(valid, 0x006000 | 0x001000)) {
				dst->fields |= DATA_TIME;
				dst->time = ims100_time(f);
			}
			if (has(valid, 0x000C00 | 0x000300 | 0x0000C0)) {
				const int32_t la = be32s(f + 26), lo = be32s(f + 30);
				const int32_t al = (int32_t)((uint32_t)f[34] << 24 | (uint32_t)f[35] << 16 | (uint32_t)f[36] << 8);
				dst->fields |= DATA_POS;
				dst->lat = (float)((int)(la / 1e6) + (la % 1000000 / 60.0 * 100.0) / 1e6);      /* NMEA ddmm.mmmm */
				dst->lon = (float)((int)(lo / 1e6) + (lo % 1000000 / 60.0 * 100.0) / 1e6);
				dst->alt = (float)((al >> 8) / 1e2);
			}
			if (has(valid, 0x000002 | 0x000004)) {
				dst->fields |= DATA_SPEED;
				dst->speed = (float)(be16u(f + 44) / 1.943844e2);                              /* knots * 100 */
				dst->heading = (float)(abs((int16_t)be16u(f + 42)) / 1e2);
				dst->climb = NAN;
			}
		}
		if ((dst->fields & (DATA_POS | DATA_TIME)) == (DATA_POS | DATA_TIME)) {             /* climb from successive fixes */
			if (dst->time > m_ims_prev_time) dst->climb = (dst->alt - m_ims_prev_alt) / (dst->time - m_ims_prev_time);
			m_ims_prev_alt = dst->alt;
			m_ims_prev_time = dst->time;
		}
		meisei_serial(dst, "IMS%d");
	}

	/* the frame carries day/month and the last digit of the year; the decade comes from the wall clock
	 * (ims100/parser.c:18-41) */
	static time_t ims100_time(const uint8_t *f)
	{
		const uint16_t date = be16u(f + 24), ms = be16u(f + 20);
		time_t now = time(nullptr);
		struct tm tm = *gmtime(&now);
		const int unit = tm.tm_year % 10;
		tm.tm_year -= unit;
		tm.tm_year += date % 10 - (unit < date % 10 ? 10 : 0);
		tm.tm_mon = (date / 10) % 100 - 1;
		tm.tm_mday = date / 1000;
		tm.tm_hour = f[22];
		tm.tm_min = f[23];
		tm.tm_sec = ms / 1000;
		return tl::utc_seconds(&tm);
	}

	void rs11g(const uint8_t *f, SondeData *dst)
	{
		uint32_t valid;
		memcpy(&valid, f + 48, 4);
		const float *cal = m_ims_calib;                        /* RS11GCalibration as 64 floats (ims100/protocol.h:178-195) */
		meisei_common_head(f, valid, dst, true);
		if (has(valid, 0x800000 | 0x460000)) {
			if ((be16u(f) & 3) == 0) m_adc_ref = be16u(f + 2);
			m_adc_temp = be16u(f + 10);
			m_adc_rh = be16u(f + 12);
			const float air = meisei_temp(m_adc_temp, m_adc_ref, cal + 33, cal + 37, cal + 17, 11);
			const float rh = meisei_rh(m_adc_rh, m_adc_ref, cal + 49);
			/* temperature dependence of the humidity sensor (fit of the GRUAN TD-5 curves, parser.c:243-266) */
			const float k[] = {5.79231318e-02, -2.64030081e-03, -1.32089353e-05, 7.15251769e-07,
			                   -4.81000481e-04, 1.86628187e+00, -7.69600770e-01};
			float corr = k[0] + k[1] * air + k[2] * air * air + k[3] * air * air * air;
			corr *= k[4] + k[5] * rh / 100 + k[6] * rh / 100 * rh / 100;
			dst->fields |= DATA_PTU;
			dst->temp = air;
			dst->rh = clamp_pct(rh + corr * 100);
			dst->pressure = 0;
			dst->calib_percent = (float)(100.0 * __builtin_popcountll(m_ims_mask) / 64);
		}
		switch (f[14] << 8 | f[15]) {
		case 0x30A2:                                             /* position, velocity, date */
			if (has(valid, 0x000600 | 0x000180 | 0x000060)) {
				dst->fields |= DATA_POS;
				dst->lat = (float)(be32s(f + 26) / 1e7);
				dst->lon = (float)(be32s(f + 30) / 1e7);
				dst->alt = (float)(be32s(f + 34) / 1e2);
			}
			if (has(valid, 0x000010 | 0x000008 | 0x000004)) {
				dst->fields |= DATA_SPEED;
				dst->speed = (float)(be16u(f + 38) / 1e2);
				dst->heading = (float)(abs((int16_t)be16u(f + 40)) / 1e2);
				dst->climb = (float)(abs((int16_t)be16u(f + 42)) / 1e2);
			}
			if (has(valid, 0x000003)) {
				struct tm tm;
				memset(&tm, 0, sizeof(tm));
				tm.tm_year = f[45] + 0x700 - 1900;
				tm.tm_mon = f[46] - 1;
				tm.tm_mday = f[47];
				m_ims_date = tl::utc_seconds(&tm);
			}
			break;
		case 0x31A2:                                             /* time of day (ms is little endian on this page) */
			if (has(valid, 0x006000)) {
				const uint16_t ms = (uint16_t)(f[20] | f[21] << 8);
				dst->fields |= DATA_TIME;
				dst->time = m_ims_date + (f[22] * 3600 + f[23] * 60 + ms / 1000);
			}
			break;
		default:
			break;
		}
		meisei_serial(dst, "RS11G-%d");
	}

	int m_type;
	SondeData m_partial;
	uint32_t m_dfm_raw[16];
	uint8_t m_dfm_serial_ch;
	uint64_t m_dfm_serial;
	struct tm m_dfm_tm;
	float m_ims_calib[64];
	uint64_t m_ims_mask;
	time_t m_ims_date;
	float m_ims_prev_alt;
	time_t m_ims_prev_time;
	uint16_t m_adc_ref, m_adc_temp, m_adc_rh, m_adc_rh_temp;
	uint8_t m_mrz_calib[64];
	uint8_t m_rs41_calib[816];
	uint8_t m_rs41_have[7];
	uint16_t m_mrz_mask;
	struct tm m_c50_tm;
	float m_imet_prev[3];
	uint32_t m_imet_prev_time;
};

}  // namespace radiosonde
