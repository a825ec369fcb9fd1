/*
 * track_files_test.cpp — host/track_files.hpp against the compiled reference writers (oracle/_ref/libwriters_ref.so).
 *
 * Every scenario drives one of this repo's writers and the reference's writer of the same file with the same random
 * sequence of calls — serial changes, empty / quoted / over-long / non-printing names, NaN and all-zero fixes, repeated
 * times and positions, out-of-range coordinates, negative coordinates (the live KML marker disappears), stop without
 * start, re-init — and compares the files byte for byte.
 *
 *   track_files_test <libwriters_ref.so> <scratch dir> [n_rounds]
 *   track_files_test - <scratch dir> [n_rounds]      only this repo's writers, nothing compared: leaves the ours_* files of
 *                                                    the last round, whose digests tests/golden/writers.json pins
 */
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <random>
#include <string>
#include <vector>

#include "../../sdrpp_radiosonde_b200/host/track_files.hpp"

namespace {

void *g_lib;
bool g_ours_only = false;
template <typename F> struct Noop;
template <typename R, typename... A> struct Noop<R (*)(A...)> { static R fn(A...) { return R(); } };
template <typename F> F sym(const char *name)
{
	if (g_ours_only) return &Noop<F>::fn;
	void *p = dlsym(g_lib, name);
	if (!p) { fprintf(stderr, "missing %s\n", name); exit(2); }
	return (F)p;
}

std::string slurp(const std::string &path)
{
	std::ifstream f(path, std::ios::binary);
	return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int g_fail = 0, g_checked = 0;
void same(const std::string &a, const std::string &b, const char *what, int round, bool may_be_empty = false)
{
	if (g_ours_only) return;
	const std::string x = slurp(a), y = slurp(b);
	g_checked++;
	if ((x.empty() && !may_be_empty) || x != y) {
		g_fail++;
		size_t i = 0;
		while (i < x.size() && i < y.size() && x[i] == y[i]) i++;
		fprintf(stderr, "MISMATCH %s round %d: %zu vs %zu bytes, first difference at %zu\n  ours: %.80s\n  ref : %.80s\n", what, round,
		        x.size(), y.size(), i, x.substr(i > 40 ? i - 40 : 0).c_str(), y.substr(i > 40 ? i - 40 : 0).c_str());
	}
}

struct Gen {
	std::mt19937 rng;
	explicit Gen(unsigned seed) : rng(seed) {}
	int upto(int n) { return (int)(rng() % (unsigned)n); }
	float coord(float lim)
	{
		switch (upto(12)) {
		case 0: return NAN;
		case 1: return 0.0f;
		case 2: return lim * 1.5f;                               /* out of range */
		case 3: return -lim * 1.5f;
		default: return ((int)(rng() % 2000001) - 1000000) * 1e-6f * lim;
		}
	}
	float value(float lo, float hi) { return lo + (hi - lo) * (float)(rng() % 100001) / 100000.0f; }
	std::string serial(bool cli)
	{
		static const char *pool[] = {"S1234567", "T4920311", "ME0012AB", "IMS-77001", "D19012345", "", "bad name", "tab\there",
		                             "quo\"te", "x"};
		const int k = upto(cli ? 12 : 13);
		if (k < 10) return pool[k];
		if (k == 12) return std::string(70, 'L');                /* longer than the module's 63-character memory */
		return pool[upto(3)];
	}
	SondeData point(const std::string &serial, time_t &clock)
	{
		SondeData d;
		memset(&d, 0, sizeof(d));
		static const int masks[] = {DATA_POS | DATA_SPEED | DATA_TIME | DATA_PTU | DATA_SERIAL, DATA_POS | DATA_SPEED, DATA_POS, DATA_SPEED,
		                            DATA_PTU | DATA_TIME, DATA_OZONE | DATA_POS | DATA_SPEED | DATA_TIME, 0, DATA_TIME, DATA_SEQ | DATA_POS};
		d.fields = masks[upto(sizeof(masks) / sizeof(*masks))];
		if (upto(4)) clock += upto(3);
		d.time = clock;
		d.lat = coord(90);
		d.lon = coord(180);
		d.alt = upto(10) ? value(-50, 35000) : coord(30000);
		d.speed = value(0, 80);
		d.heading = value(-400, 800);
		d.climb = value(-30, 10);
		d.temp = value(-90, 40);
		d.rh = value(0, 100);
		d.pressure = value(3, 1030);
		d.o3_mpa = value(0, 20);
		strncpy(d.serial, serial.c_str(), sizeof(d.serial) - 1);
		return d;
	}
};

}  // namespace

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: %s libwriters_ref.so scratch_dir [rounds]\n", argv[0]); return 2; }
	g_ours_only = !strcmp(argv[1], "-");
	if (!g_ours_only) {
		g_lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
		if (!g_lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
	}
	const std::string dir = argv[2];
	const int rounds = argc > 3 ? atoi(argv[3]) : 40;

	auto gpxw_new = sym<void *(*)()>("refw_gpxw_new");
	auto gpxw_free = sym<void (*)(void *)>("refw_gpxw_free");
	auto gpxw_init = sym<int (*)(void *, const char *)>("refw_gpxw_init");
	auto gpxw_deinit = sym<void (*)(void *)>("refw_gpxw_deinit");
	auto gpxw_start = sym<void (*)(void *, const char *)>("refw_gpxw_start");
	auto gpxw_stop = sym<void (*)(void *)>("refw_gpxw_stop");
	auto gpxw_add = sym<void (*)(void *, long, float, float, float, float, float)>("refw_gpxw_add");
	auto ptu_new = sym<void *(*)()>("refw_ptu_new");
	auto ptu_free = sym<void (*)(void *)>("refw_ptu_free");
	auto ptu_init = sym<int (*)(void *, const char *)>("refw_ptu_init");
	auto ptu_add = sym<void (*)(void *, long, const float *, const char *)>("refw_ptu_add");
	auto csv_new = sym<void *(*)(const char *)>("refw_csv_new");
	auto csv_add = sym<void (*)(void *, const void *)>("refw_csv_add");
	auto csv_close = sym<void (*)(void *)>("refw_csv_close");
	auto gpx_new = sym<void *(*)(const char *)>("refw_gpx_new");
	auto gpx_start = sym<void (*)(void *, const char *)>("refw_gpx_start");
	auto gpx_add = sym<void (*)(void *, const void *)>("refw_gpx_add");
	auto gpx_stop = sym<void (*)(void *)>("refw_gpx_stop");
	auto gpx_close = sym<void (*)(void *)>("refw_gpx_close");
	auto kml_new = sym<void *(*)(const char *, int)>("refw_kml_new");
	auto kml_start = sym<void (*)(void *, const char *)>("refw_kml_start");
	auto kml_add = sym<void (*)(void *, const void *)>("refw_kml_add");
	auto kml_stop = sym<void (*)(void *)>("refw_kml_stop");
	auto kml_close = sym<void (*)(void *)>("refw_kml_close");

	for (int round = 0; round < rounds; round++) {
		const std::string a = dir + "/ours_", b = dir + "/ref_";
		const int n_ops = 20 + 37 * (round % 7);

		/* ---- the module's GPX track: src/main.cpp:320-331 call pattern plus stop / re-init ---- */
		{
			Gen g(1000 + round);
			radiosonde::GPXWriter mine;
			void *ref = gpxw_new();
			mine.init((a + "w.gpx").c_str());
			gpxw_init(ref, (b + "w.gpx").c_str());
			time_t clock = 1700000000 + round * 1000;
			std::string serial = "S1234567";
			for (int i = 0; i < n_ops; i++) {
				const int op = g.upto(20);
				if (op < 3) {
					serial = g.serial(false);
					mine.startTrack(serial.c_str());
					gpxw_start(ref, serial.c_str());
				} else if (op == 3) {
					mine.stopTrack();
					gpxw_stop(ref);
				} else if (op == 4 && round % 3 == 0) {
					/* the checkbox toggled off and on again: same file, started over */
					mine.deinit();
					gpxw_deinit(ref);
					same(a + "w.gpx", b + "w.gpx", "module gpx at deinit", round);
					mine.init((a + "w.gpx").c_str());
					gpxw_init(ref, (b + "w.gpx").c_str());
				} else {
					const SondeData d = g.point(serial, clock);
					if (!serial.empty() && g.upto(3)) {          /* main.cpp:325 */
						mine.startTrack(serial.c_str());
						gpxw_start(ref, serial.c_str());
					}
					mine.addTrackPoint(d.time, d.lat, d.lon, d.alt, d.speed, d.heading);
					gpxw_add(ref, (long)d.time, d.lat, d.lon, d.alt, d.speed, d.heading);
				}
				if (i % 16 == 5) same(a + "w.gpx", b + "w.gpx", "module gpx mid-run", round);   /* complete after every call */
			}
			mine.deinit();
			gpxw_deinit(ref);
			gpxw_free(ref);
			same(a + "w.gpx", b + "w.gpx", "module gpx", round);
		}
		/* ---- the module's PTU log ---- */
		{
			Gen g(2000 + round);
			radiosonde::PTUWriter mine;
			void *ref = ptu_new();
			mine.init((a + "ptu.csv").c_str());
			ptu_init(ref, (b + "ptu.csv").c_str());
			time_t clock = 1600000000;
			for (int i = 0; i < n_ops; i++) {
				const SondeData d = g.point("x", clock);
				SondeFullData f;
				f.time = d.time; f.temp = d.temp; f.rh = d.rh; f.dewpt = g.value(-100, 30); f.pressure = d.pressure;
				f.lat = d.lat; f.lon = d.lon; f.alt = d.alt; f.spd = d.speed; f.hdg = d.heading; f.climb = d.climb;
				f.auxData = g.upto(3) ? "" : "O3=" + std::to_string(d.o3_mpa) + "mPa";
				const float v[10] = {f.temp, f.rh, f.dewpt, f.pressure, f.lat, f.lon, f.alt, f.spd, f.hdg, f.climb};
				mine.addPoint(&f);
				ptu_add(ref, (long)f.time, v, f.auxData.c_str());
			}
			mine.deinit();
			ptu_free(ref);
			same(a + "ptu.csv", b + "ptu.csv", "module ptu", round);
		}
		/* ---- the tool's CSV / GPX / KML / live KML from one SondeData sequence: SD/main.c:347-365 ---- */
		{
			Gen g(3000 + round);
			radiosonde::cli::CsvFile csv;
			radiosonde::cli::GpxFile gpx;
			radiosonde::cli::KmlFile kml, live;
			csv.init((a + "c.csv").c_str());
			gpx.init((a + "c.gpx").c_str());
			const int k0 = kml.init((a + "c.kml").c_str(), false), l0 = live.init((a + "l.kml").c_str(), true);
			void *rcsv = csv_new((b + "c.csv").c_str()), *rgpx = gpx_new((b + "c.gpx").c_str());
			void *rkml = kml_new((b + "c.kml").c_str(), 0), *rlive = kml_new((b + "l.kml").c_str(), 1);
			if ((!g_ours_only && (!rcsv || !rgpx || !rkml || !rlive)) || k0 || l0) { fprintf(stderr, "cannot create files in %s\n", dir.c_str()); return 2; }
			/* the link file names the live file by path, so its text differs by the "ours_" / "ref_" prefix only */
			time_t clock = 1650000000;
			std::string serial = round % 2 ? "" : "T4920311";       /* SD/main.c:352 starts the KML track for "" too */
			for (int i = 0; i < n_ops; i++) {
				const int op = g.upto(24);
				if (op < 2) serial = g.serial(true);
				if (op == 2) {
					gpx.stop_track(); gpx_stop(rgpx);
					kml.stop_track(); kml_stop(rkml);
					live.stop_track(); kml_stop(rlive);
					continue;
				}
				const SondeData d = g.point(serial, clock);
				csv.add_point(d);
				csv_add(rcsv, &d);
				kml.start_track(d.serial); kml.add_trackpoint(d);
				kml_start(rkml, d.serial); kml_add(rkml, &d);
				live.start_track(d.serial); live.add_trackpoint(d);
				kml_start(rlive, d.serial); kml_add(rlive, &d);
				if (d.fields & DATA_SERIAL) { gpx.start_track(d.serial); gpx_start(rgpx, d.serial); }
				if ((d.fields & (DATA_POS | DATA_SPEED)) == (DATA_POS | DATA_SPEED)) { gpx.add_trackpoint(d); gpx_add(rgpx, &d); }
				if (i % 16 == 9) {
					same(a + "l.kml-live.kml", b + "l.kml-live.kml", "tool live kml mid-run", round);
					same(a + "c.gpx", b + "c.gpx", "tool gpx mid-run", round, true);   /* nothing is flushed before the first track */
				}
			}
			csv.close(); gpx.close(); kml.close(); live.close();
			csv_close(rcsv); gpx_close(rgpx); kml_close(rkml); kml_close(rlive);
			same(a + "c.csv", b + "c.csv", "tool csv", round);
			same(a + "c.gpx", b + "c.gpx", "tool gpx", round);
			same(a + "c.kml", b + "c.kml", "tool kml", round);
			same(a + "l.kml-live.kml", b + "l.kml-live.kml", "tool live kml", round);
			std::string lo = slurp(a + "l.kml"), lr = slurp(b + "l.kml");
			const size_t p = lo.find("ours_");
			if (p != std::string::npos) lo.replace(p, 5, "ref_");
			g_checked++;
			if (!g_ours_only && (lo.empty() || lo != lr)) { g_fail++; fprintf(stderr, "MISMATCH tool live link file round %d\n", round); }
		}
	}
	printf("%s %d files compared, %d differ\n", g_fail ? "FAIL" : "OK", g_checked, g_fail);
	return g_fail ? 1 : 0;
}
