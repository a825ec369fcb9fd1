/*
 * track_files.hpp — the output files either side of the decoder block.
 *
 * Host-side only; nothing here touches the GPU.  Two families, byte-compatible with what the reference writes from the
 * same sequence of data points (tests/cpp/track_files_test.cpp drives these and the compiled reference writers with one
 * random point sequence and compares the files byte for byte):
 *
 *   radiosonde::GPXWriter, radiosonde::PTUWriter        the SDR++ module's track and log files, same class and method
 *                                                       names (src/gpx.hpp:11-59 / src/gpx.cpp:9-120,
 *                                                       src/ptu.hpp:10-27 / src/ptu.cpp:3-35), fed from the
 *                                                       SondeFullData callback the way src/main.cpp:320-331 does
 *   radiosonde::cli::CsvFile, GpxFile, KmlFile          the command-line tool's -c / -g / -k / -l files
 *                                                       (SD/io/csv.c:6-61, SD/io/gpx.c:11-117, SD/io/kml.c:11-192), fed
 *                                                       from SondeData the way SD/main.c:347-365 does
 *   radiosonde::cli::print_data                         the tool's text line per data point and its -f format language
 *                                                       (SD/main.c:111,489-566)
 *
 * All the XML files are kept well-formed while they grow: a body that only ever grows, followed by a provisional
 * trailer that the next update overwrites.  TrailerFile below is that mechanism; the writers differ in what they put in
 * the body and in the trailer.  The reference never truncates, and neither does this: a trailer that gets shorter than
 * the bytes it replaces (a live KML whose position marker disappears when a coordinate goes negative) leaves the old
 * bytes behind it, exactly as there.
 */
#pragma once
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>

#include "sonde_data.hpp"
#include "telemetry.hpp"                                 /* tl::altitude_to_pressure */

namespace radiosonde {
namespace io {

/* An append-only body plus a rewritable trailer, on a stdio stream */
class TrailerFile {
public:
	TrailerFile() = default;
	TrailerFile(const TrailerFile &) = delete;
	TrailerFile &operator=(const TrailerFile &) = delete;
	~TrailerFile() { close(); }

	bool open(const char *fname)
	{
		close();
		m_f = fopen(fname, "wb");
		m_body_end = 0;
		return m_f != nullptr;
	}
	bool is_open() const { return m_f != nullptr; }
	void close()
	{
		if (m_f) fclose(m_f);
		m_f = nullptr;
	}
	/* the next body() call goes over the current trailer */
	void rewind_to_body() { if (m_f) fseek(m_f, m_body_end, SEEK_SET); }
	/* text at the stream position, counted into the body */
	__attribute__((format(printf, 2, 3))) void body(const char *fmt, ...)
	{
		if (!m_f) return;
		va_list ap;
		va_start(ap, fmt);
		vfprintf(m_f, fmt, ap);
		va_end(ap);
		m_body_end = ftell(m_f);
	}
	/* text at the stream position that the next rewind_to_body() + body() replaces */
	__attribute__((format(printf, 2, 3))) void trailer(const char *fmt, ...)
	{
		if (!m_f) return;
		va_list ap;
		va_start(ap, fmt);
		vfprintf(m_f, fmt, ap);
		va_end(ap);
	}
	/* what has been written behind the body so far stays: the body ends at the stream position */
	void keep_trailer() { if (m_f) m_body_end = ftell(m_f); }
	void flush() { if (m_f) fflush(m_f); }

private:
	FILE *m_f = nullptr;
	long m_body_end = 0;
};

inline void utc_stamp(char (&out)[24], time_t t)
{
	struct tm tmv;
	out[0] = 0;
	if (gmtime_r(&t, &tmv)) strftime(out, sizeof(out), "%Y-%m-%dT%H:%M:%SZ", &tmv);
}

}  // namespace io

/* ---------------------------------------------------------------- the SDR++ module's files ---------------------- */

/* src/utils.cpp:3-17: where the module puts its output files by default (src/main.cpp:41-42: "radiosonde.gpx",
 * "radiosonde_ptu.csv") — $TMP, else $TEMP, else /tmp.  The reference joins directory and name with a backslash on
 * every platform but Windows (its #ifdef is the wrong way round), which on Linux names a file "tmp\radiosonde.gpx" in
 * the root directory; this joins with '/'. */
inline std::string getTempFile(const std::string &file)
{
	const char *dir = getenv("TMP");
	if (!dir) dir = getenv("TEMP");
	if (!dir) dir = "/tmp";
	return std::string(dir) + "/" + file;
}

/* src/gpx.hpp:11-59.  One GPX 1.1 file, one <trk> per sonde serial, complete after every call. */
class GPXWriter {
public:
	GPXWriter() = default;
	~GPXWriter() { deinit(); }

	/* src/gpx.cpp:9-27 */
	bool init(const char *fname)
	{
		if (m_file.is_open()) deinit();
		if (!m_file.open(fname)) return false;
		m_lat = m_lon = m_alt = 0;
		m_time = 0;
		m_active = false;
		m_file.body("<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n"
		            "<gpx xmlns=\"http://www.topografix.com/GPX/1/1\" version=\"1.1\" creator=\"SDR++\">\n");
		finish();
		return true;
	}
	/* src/gpx.cpp:29-36: the track stays "active" across deinit(); only init() clears that */
	void deinit()
	{
		if (!m_file.is_open()) return;
		finish();
		m_file.close();
	}
	/* src/gpx.cpp:38-58: no-op for the serial already being tracked and for names with non-printing characters; any
	 * other name closes the running track and opens a new one.  Serials are remembered to 63 characters, so a longer
	 * one never compares equal and opens a new track on every call (gpx.hpp:53, gpx.cpp:42,50). */
	void startTrack(const char *name)
	{
		if (!m_file.is_open()) return;
		if (m_active && m_serial == name) return;
		for (const char *p = name; *p; p++)
			if (!isgraph((unsigned char)*p)) return;
		if (m_active) stopTrack();
		m_serial.assign(name, strnlen(name, 63));
		m_file.rewind_to_body();
		m_file.body("<trk>\n<name>%s</name>\n<trkseg>\n", name);
		m_active = true;
		finish();
	}
	/* src/gpx.cpp:61-68 */
	void stopTrack()
	{
		if (!m_file.is_open() || !m_active) return;
		m_file.rewind_to_body();
		m_file.body("</trkseg>\n</trk>\n");
		m_active = false;
		finish();
	}
	/* src/gpx.cpp:70-98: drops NaN and all-zero fixes, and any point whose time OR whose whole position repeats the
	 * last one written */
	void addTrackPoint(time_t time, float lat, float lon, float alt, float spd, float hdg)
	{
		if (!m_file.is_open() || !m_active) return;
		if (std::isnan(lat) || std::isnan(lon) || std::isnan(alt)) return;
		if (lat == 0 && lon == 0 && alt == 0) return;
		if (time == m_time || (lat == m_lat && lon == m_lon && alt == m_alt)) return;
		m_lat = lat; m_lon = lon; m_alt = alt; m_time = time;
		char stamp[24];
		io::utc_stamp(stamp, time);
		m_file.rewind_to_body();
		m_file.body("<trkpt lat=\"%f\" lon=\"%f\">\n<time>%s</time>\n<ele>%f</ele>\n<speed>%f</speed>\n<course>%f</course>\n</trkpt>\n",
		            lat, lon, stamp, alt, spd, hdg);
		finish();
	}

private:
	/* src/gpx.cpp:100-112 */
	void finish()
	{
		m_file.rewind_to_body();
		m_file.trailer("%s</gpx>\n", m_active ? "</trkseg>\n</trk>\n" : "");
		m_file.flush();
	}
	io::TrailerFile m_file;
	bool m_active = false;
	std::string m_serial;
	float m_lat = 0, m_lon = 0, m_alt = 0;
	time_t m_time = 0;
};

/* src/ptu.hpp:10-27, src/ptu.cpp:3-35: one CSV row per callback, flushed */
class PTUWriter {
public:
	PTUWriter() = default;
	PTUWriter(const PTUWriter &) = delete;
	PTUWriter &operator=(const PTUWriter &) = delete;
	~PTUWriter() { deinit(); }

	bool init(const char *fname)
	{
		if (m_f) deinit();
		if (!(m_f = fopen(fname, "wb"))) return false;
		fputs("Epoch,Temperature,Relative humidity,Dew point,Pressure,Latitude,Longitude,Altitude,Speed,Heading,Climb,XDATA\n", m_f);
		return true;
	}
	void deinit()
	{
		if (m_f) fclose(m_f);
		m_f = nullptr;
	}
	void addPoint(const SondeFullData *d)
	{
		if (!m_f) return;
		fprintf(m_f, "%ld,%.1f,%.1f,%.1f,%.1f,%.6f,%.6f,%.1f,%.1f,%.1f,%.1f,%s\n", (long)d->time, d->temp, d->rh, d->dewpt,
		        d->pressure, d->lat, d->lon, d->alt, d->spd, d->hdg, d->climb, d->auxData.c_str());
		fflush(m_f);
	}

private:
	FILE *m_f = nullptr;
};

/* ---------------------------------------------------------------- the command-line tool's files ----------------- */
namespace cli {

inline bool has_all(int fields, int mask) { return (fields & mask) == mask; }

/* SD/io/csv.c:6-61: a column group is empty when the record does not carry it */
class CsvFile {
public:
	CsvFile() = default;
	CsvFile(const CsvFile &) = delete;
	CsvFile &operator=(const CsvFile &) = delete;
	~CsvFile() { close(); }

	bool init(const char *fname)
	{
		close();
		if (!(m_f = fopen(fname, "wb"))) return false;
		fputs("Time,Temperature,RH,Pressure,Latitude,Longitude,Altitude,Speed,Heading,Climb,XDATA\n", m_f);
		return true;
	}
	void close()
	{
		if (m_f) fclose(m_f);
		m_f = nullptr;
	}
	bool is_open() const { return m_f != nullptr; }
	void add_point(const SondeData &d)
	{
		if (!m_f) return;
		if (d.fields & DATA_TIME) {
			char stamp[24];
			io::utc_stamp(stamp, d.time);
			fprintf(m_f, "%s,", stamp);
		} else {
			fputs(",", m_f);
		}
		if (d.fields & DATA_PTU) fprintf(m_f, "%f,%f,%f,", d.temp, d.rh, d.pressure); else fputs(",,,", m_f);
		if (d.fields & DATA_POS) fprintf(m_f, "%f,%f,%f,", d.lat, d.lon, d.alt); else fputs(",,,", m_f);
		if (d.fields & DATA_SPEED) fprintf(m_f, "%f,%f,%f,", d.speed, d.heading, d.climb); else fputs(",,,", m_f);
		if (d.fields & DATA_OZONE) fprintf(m_f, "O3=%fmPa", d.o3_mpa);
		fputs("\n", m_f);
	}

private:
	FILE *m_f = nullptr;
};

/* SD/io/gpx.c:11-117.  Differences to the module's writer: creator "SondeDump", names are refused only when empty or
 * containing a double quote, points need DATA_POS and DATA_SPEED and a latitude / longitude in range, the course is
 * reduced modulo 360, and repeated points are written. */
class GpxFile {
public:
	~GpxFile() { close(); }

	bool init(const char *fname)
	{
		m_tracking = false;
		if (!m_file.open(fname)) return false;
		m_file.body("<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n"
		            "<gpx xmlns=\"http://www.topografix.com/GPX/1/1\" version=\"1.1\" creator=\"SondeDump\">\n");
		m_file.trailer("</gpx>\n");
		return true;
	}
	bool is_open() const { return m_file.is_open(); }
	void close()
	{
		if (!m_file.is_open()) return;
		if (m_tracking) stop_track();
		m_file.rewind_to_body();
		m_file.trailer("</gpx>\n");
		m_file.close();
	}
	void start_track(const char *name)
	{
		if (!m_file.is_open() || !name[0] || strchr(name, '"')) return;
		if (m_tracking && m_serial == name) return;
		if (m_tracking) stop_track();
		m_serial = name;
		m_tracking = true;
		m_file.rewind_to_body();
		m_file.body("<trk>\n<name>%s</name>\n<trkseg>\n", name);
		m_file.trailer("</trkseg>\n</trk>\n</gpx>\n");
		m_file.flush();
	}
	void add_trackpoint(const SondeData &d)
	{
		if (!m_file.is_open() || !m_tracking) return;
		if (!has_all(d.fields, DATA_POS | DATA_SPEED)) return;
		if (std::isnan(d.lat) || std::isnan(d.lon) || std::isnan(d.alt)) return;
		if (d.lat == 0 && d.lon == 0 && d.alt == 0) return;
		if (d.lat > 90 || d.lat < -90 || d.lon > 180 || d.lon < -180) return;
		const float course = (float)fmod((double)d.heading, 360.0);
		char stamp[24];
		io::utc_stamp(stamp, d.time);
		m_file.rewind_to_body();
		m_file.body("<trkpt lat=\"%f\" lon=\"%f\">\n<time>%s</time>\n<ele>%f</ele>\n<speed>%f</speed>\n<course>%f</course>\n</trkpt>\n",
		            d.lat, d.lon, stamp, d.alt, d.speed, course);
		m_file.trailer("</trkseg>\n</trk>\n</gpx>\n");
		m_file.flush();
	}
	void stop_track()
	{
		if (!m_file.is_open()) return;
		m_file.rewind_to_body();
		m_file.body("</trkseg>\n</trk>\n");
		m_file.trailer("</gpx>\n");
		m_tracking = false;
	}

private:
	io::TrailerFile m_file;
	bool m_tracking = false;
	std::string m_serial;
};

/*
 * SD/io/kml.c:11-192.  `-k file`: one <Placemark> with a <LineString> per serial, completed only by close().
 * `-l file` (live): `file` is a small KML with a <NetworkLink> that makes the viewer re-read `file-live.kml` every 5 s
 * (kml.h:8), and the live file is completed after every update.
 *
 * Reproduced as the reference behaves, oddities included:
 *   - the position marker after the track is written whenever the last latitude, longitude and altitude are all >= 0
 *     (kml.c:165); they start at 0 here (the reference leaves them uninitialised, SD/main.c:99);
 *   - the marker is named after the serial, which the reference has already dropped when close() gets there, or not
 *     yet set at init(): glibc prints "(null)" for it (kml.c:84-85,146-147,183);
 *   - start_track() takes any name, the empty one too (SD/main.c:351-352 calls it for every data point);
 *   - in live mode a change of serial writes the new track's header behind the trailer of the old one rather than over
 *     it (kml.c:93-107: the seek comes first, stop_track() then leaves the position behind its trailer).
 */
class KmlFile {
public:
	~KmlFile() { close(); }

	/* 0 on success, 1 when the track file, 2 when the link file of live mode cannot be created (kml.c:11-75) */
	int init(const char *fname, bool live_update)
	{
		m_live = live_update;
		m_tracking = m_named = false;
		m_lat = m_lon = m_alt = 0;
		std::string live_name;
		if (live_update) {
			char buf[256];                                   /* kml.c:15-19: at most 254 characters of name */
			snprintf(buf, sizeof(buf) - 1, "%s-live.kml", fname);
			buf[sizeof(buf) - 1] = 0;
			live_name = buf;
		}
		if (!m_file.open(live_update ? live_name.c_str() : fname)) return 1;
		m_file.body("<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n"
		            "<kml xmlns=\"http://www.opengis.net/kml/2.2\">\n"
		            "<Document>\n"
		            "<visibility>1</visibility>\n"
		            "<open>1</open>\n"
		            "<Style id=\"sondepath\"><LineStyle><color>FFFF7800</color><width>4</width></LineStyle></Style>\n"
		            "<Placemark>\n");
		if (live_update) {
			FILE *link = fopen(fname, "wb");
			if (!link) {
				close();
				return 2;
			}
			fprintf(link,
			        "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"no\" ?>\n"
			        "<kml xmlns=\"http://www.opengis.net/kml/2.2\">\n"
			        "<Document>\n"
			        "<NetworkLink>\n"
			        "<visibility>1</visibility>\n"
			        "<open>1</open>\n"
			        "<name>SondeDump live feed</name>\n"
			        "<Link><href>%s</href>\n"
			        "<refreshMode>onInterval</refreshMode><refreshInterval>%d</refreshInterval></Link>\n"
			        "</NetworkLink>\n"
			        "</Document>\n"
			        "</kml>\n",
			        live_name.c_str(), 5);
			fclose(link);
			finish();
		}
		return 0;
	}
	bool is_open() const { return m_file.is_open(); }
	/* kml.c:77-85 */
	void close()
	{
		if (!m_file.is_open()) return;
		if (m_tracking) stop_track();
		if (!m_live) finish();
		m_file.close();
	}
	/* kml.c:87-108 */
	void start_track(const char *name)
	{
		if (!m_file.is_open()) return;
		if (m_live) m_file.rewind_to_body();
		if (m_named && m_serial == name) return;
		if (m_named) stop_track();                           /* live: leaves the position BEHIND the trailer */
		m_serial = name;
		m_named = true;
		m_file.body("<name>%s</name><styleUrl>#sondepath</styleUrl>\n<LineString>\n<tessellate>0</tessellate>\n<coordinates>\n", name);
		if (m_live) finish();                                /* still without the LineString's end tags */
		m_tracking = true;
	}
	/* kml.c:110-124 */
	void add_trackpoint(const SondeData &d)
	{
		if (!m_file.is_open() || !m_tracking) return;
		if (!has_all(d.fields, DATA_POS)) return;
		if (std::isnan(d.lat) || std::isnan(d.lon) || std::isnan(d.alt)) return;
		if (m_live) m_file.rewind_to_body();
		m_file.body("%f,%f,%f\n", d.lon, d.lat, d.alt);
		m_lat = d.lat; m_lon = d.lon; m_alt = d.alt;
		if (m_live) finish();
	}
	/* kml.c:126-141 */
	void stop_track()
	{
		if (!m_file.is_open()) return;
		if (m_live) m_file.rewind_to_body();
		m_file.body("</coordinates>\n</LineString>\n");
		m_tracking = false;
		if (m_live) finish();
		m_named = false;
	}

private:
	/* kml.c:143-161 */
	void finish()
	{
		m_file.keep_trailer();                               /* kml.c:146: the trailer starts wherever the stream is */
		if (m_tracking) m_file.trailer("</coordinates></LineString>\n");
		m_file.trailer("</Placemark>\n");
		if (m_lat >= 0 && m_lon >= 0 && m_alt >= 0)
			m_file.trailer("<Placemark>\n<name>%s</name>\n<Point>\n<altitudeMode>absolute</altitudeMode>\n"
			               "<coordinates>%f,%f,%f</coordinates>\n</Point>\n</Placemark>\n",
			               m_named ? m_serial.c_str() : "(null)", m_lon, m_lat, m_alt);
		m_file.trailer("</Document>\n</kml>\n");
		m_file.flush();
	}
	io::TrailerFile m_file;
	bool m_live = false, m_tracking = false, m_named = false;
	std::string m_serial;
	float m_lat = 0, m_lon = 0, m_alt = 0;
};

/* The tool's line format when -f is not given (SD/main.c:111) */
inline const char *default_format() { return "(%S) [%f] %t'C %r%%    %l %o %am    %sm/s %h' %cm/s\t%x"; }

/* SD/physics.c:43-48: the tool's dew point (the SDR++ module computes its own, src/decode/decoder.hpp:132-140);
 * float / double mix as there */
inline float dew_point(float temp, float rh)
{
	const float q = (float)((logf((float)(rh / 100.0)) + (17.27 * temp / (237.3 + temp))) / 17.27);
	return (float)(237.3 * q / (1 - q));
}

/*
 * One text line per data point, SD/main.c:489-566.  `%` + letter inserts a value:
 *   a altitude  b shutdown timer  c climb  d dew point  f frame counter  h heading  l latitude  o longitude  p pressure
 *   r humidity  s speed  S serial  t temperature  T time (UTC)  x ozone (XDATA)
 * any other character after a `%` is printed as it is (so `%%` is a percent sign), a lone `%` at the end prints
 * nothing, and every line of a non-empty format ends with a newline.
 */
inline void print_data(FILE *f, const char *fmt, const SondeData &d)
{
	if (!*fmt) return;
	for (const char *p = fmt; *p; p++) {
		if (*p != '%') { fputc(*p, f); continue; }
		if (!*++p) break;
		switch (*p) {
		case 'a': fprintf(f, "%6.0f", d.alt); break;
		case 'b':
			if (d.fields & DATA_SHUTDOWN) fprintf(f, "%d:%02d:%02d", d.shutdown / 3600, d.shutdown / 60 % 60, d.shutdown % 60);
			else fputs("(disabled)", f);
			break;
		case 'c': fprintf(f, "%+5.1f", d.climb); break;
		case 'd': fprintf(f, "%6.1f", dew_point(d.temp, d.rh)); break;
		case 'f': fprintf(f, "%5d", d.seq); break;
		case 'h': fprintf(f, "%3.0f", d.heading); break;
		case 'l': fprintf(f, "%8.5f%c", fabs(d.lat), d.lat >= 0 ? 'N' : 'S'); break;
		case 'o': fprintf(f, "%8.5f%c", fabs(d.lon), d.lon >= 0 ? 'E' : 'W'); break;
		case 'p': fprintf(f, "%4.1f", std::isnormal(d.pressure) ? d.pressure : tl::altitude_to_pressure(d.alt)); break;
		case 'r': fprintf(f, "%3.0f", d.rh); break;
		case 's': fprintf(f, "%4.1f", d.speed); break;
		case 'S': fputs(d.serial, f); break;
		case 't': fprintf(f, "%5.1f", d.temp); break;
		case 'T': {
			char stamp[64];
			struct tm tmv;
			stamp[0] = 0;
			if (gmtime_r(&d.time, &tmv)) strftime(stamp, sizeof(stamp), "%a %b %d %Y %H:%M:%S", &tmv);
			fputs(stamp, f);
			break;
		}
		case 'x':
			if (d.fields & DATA_OZONE) fprintf(f, "O3=%.2f mPa", d.o3_mpa);
			break;
		default: fputc(*p, f); break;
		}
	}
	fputc('\n', f);
}

}  // namespace cli
}  // namespace radiosonde
