/*
 * sonde_data.hpp — the output types of the decoder block surface.
 *
 * Layout- and name-compatible with the reference's public types so that callers written
 * against the plugin keep compiling:
 *   SondeData / DataBitmask / ParserStatus   SD/include/data.h:7-50   (96-byte POD on x86-64)
 *   SondeFullData                            src/decode/common.hpp:4-28
 */
#pragma once
#include <ctime>
#include <string>

#include "../../include/sonde_b200_compat.h"      /* SondeData, DataBitmask, ParserStatus (C definitions) */
static_assert(sizeof(SondeData) == 96, "SondeData must match SD/include/data.h");

class SondeFullData {
public:
	SondeFullData() { init(); }
	void init()
	{
		serial.clear();
		auxData.clear();
		seq = burstkill = 0;
		time = 0;
		lat = lon = alt = spd = hdg = climb = temp = rh = dewpt = pressure = calib_percent = 0;
		calibrated = false;
	}

	std::string serial;
	int seq;
	time_t time;
	int burstkill;
	float lat, lon, alt;
	float spd, hdg, climb;
	float temp, rh;
	float dewpt, pressure;
	bool calibrated;
	float calib_percent;
	std::string auxData;
};
