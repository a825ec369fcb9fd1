/*
 * ref_writers.cpp — C entry points onto the UNMODIFIED reference file writers.  TEST INFRASTRUCTURE ONLY: only
 * tests/ may load the library this builds (oracle/_ref/libwriters_ref.so, `make -C oracle writers`).
 *
 * The library holds the reference's own objects, compiled from the sources where they lie:
 *   src/gpx.cpp, src/ptu.cpp                    the SDR++ module's GPXWriter / PTUWriter
 *   SD/io/csv.c, SD/io/gpx.c, SD/io/kml.c       the command-line tool's writers (+ SD/utils.c for my_strdup)
 * and this file, which only forwards.  tests/cpp/track_files_test.cpp drives these and
 * sdrpp_radiosonde_b200/host/track_files.hpp with the same point sequence and compares the files.
 *
 * The tool's structs are calloc'ed here: the reference leaves KMLFile.lat/lon/alt uninitialised (SD/main.c:99) and
 * tests them (SD/io/kml.c:165), so a defined starting value is needed for a reproducible comparison; zero is what
 * track_files.hpp documents.
 */
#include <cstdlib>
#include <cstring>

#include "gpx.hpp"                  /* $(REF)/src */
#include "ptu.hpp"
extern "C" {
#include "io/csv.h"                 /* $(REF)/src/decode/sondedump */
#include "io/gpx.h"
#include "io/kml.h"
}

extern "C" {

/* ---- module writers ---- */
void *refw_gpxw_new(void) { return new GPXWriter(); }
void refw_gpxw_free(void *w) { delete (GPXWriter *)w; }
int refw_gpxw_init(void *w, const char *fname) { return ((GPXWriter *)w)->init(fname) ? 1 : 0; }
void refw_gpxw_deinit(void *w) { ((GPXWriter *)w)->deinit(); }
void refw_gpxw_start(void *w, const char *name) { ((GPXWriter *)w)->startTrack(name); }
void refw_gpxw_stop(void *w) { ((GPXWriter *)w)->stopTrack(); }
void refw_gpxw_add(void *w, long t, float lat, float lon, float alt, float spd, float hdg)
{
	((GPXWriter *)w)->addTrackPoint((time_t)t, lat, lon, alt, spd, hdg);
}

void *refw_ptu_new(void) { return new PTUWriter(); }
void refw_ptu_free(void *w) { delete (PTUWriter *)w; }
int refw_ptu_init(void *w, const char *fname) { return ((PTUWriter *)w)->init(fname) ? 1 : 0; }
void refw_ptu_deinit(void *w) { ((PTUWriter *)w)->deinit(); }
void refw_ptu_add(void *w, long t, const float *v10, const char *aux)
{
	SondeFullData d;
	d.time = (time_t)t;
	d.temp = v10[0]; d.rh = v10[1]; d.dewpt = v10[2]; d.pressure = v10[3];
	d.lat = v10[4]; d.lon = v10[5]; d.alt = v10[6];
	d.spd = v10[7]; d.hdg = v10[8]; d.climb = v10[9];
	d.auxData = aux;
	((PTUWriter *)w)->addPoint(&d);
}

/* ---- command-line tool writers; `data` is a SondeData (SD/include/data.h) ---- */
void *refw_csv_new(const char *fname)
{
	CSVFile *f = (CSVFile *)calloc(1, sizeof(CSVFile));
	if (csv_init(f, fname)) { free(f); return NULL; }
	return f;
}
void refw_csv_add(void *f, const void *data) { csv_add_point((CSVFile *)f, (const SondeData *)data); }
void refw_csv_close(void *f) { csv_close((CSVFile *)f); free(f); }

void *refw_gpx_new(const char *fname)
{
	GPXFile *f = (GPXFile *)calloc(1, sizeof(GPXFile));
	if (gpx_init(f, fname)) { free(f); return NULL; }
	return f;
}
void refw_gpx_start(void *f, const char *name) { gpx_start_track((GPXFile *)f, name); }
void refw_gpx_add(void *f, const void *data) { gpx_add_trackpoint((GPXFile *)f, (const SondeData *)data); }
void refw_gpx_stop(void *f) { gpx_stop_track((GPXFile *)f); }
void refw_gpx_close(void *f) { gpx_close((GPXFile *)f); free(f); }

void *refw_kml_new(const char *fname, int live)
{
	KMLFile *f = (KMLFile *)calloc(1, sizeof(KMLFile));
	if (kml_init(f, fname, live)) { free(f); return NULL; }
	return f;
}
void refw_kml_start(void *f, const char *name) { kml_start_track((KMLFile *)f, name); }
void refw_kml_add(void *f, const void *data) { kml_add_trackpoint((KMLFile *)f, (const SondeData *)data); }
void refw_kml_stop(void *f) { kml_stop_track((KMLFile *)f); }
void refw_kml_close(void *f) { kml_close((KMLFile *)f); free(f); }

}  /* extern "C" */
