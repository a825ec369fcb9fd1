/*
 * sonde_batch.cpp — sonde_b200_batch: a batch command-line runner on the batch C ABI (include/sonde_b200.h).
 *
 * The reference's CLI (SD/main.c:96-366) decodes ONE recording per process: read 1024 samples, decode(), print,
 * append to CSV/GPX/KML.  This runner decodes MANY recordings at once — every input file is one channel of one GPU
 * batch — and writes, per channel, the same CSV / GPX / KML files the reference's `-c`, `-g`, `-k` and `-l` options write
 * (SD/io/csv.c, gpx.c, kml.c; the writers are host/track_files.hpp) from the same aggregation of decoder fragments
 * (SD/decode.c:278-376 append_data_point: fields accumulate in one `printable` record per channel, pressure falls back
 * to the barometric formula, one row per fragment that carried data).
 *
 *   sonde_b200_batch [-t type] [-b buflen] [-f format] [-o prefix] [-c prefix] [-g prefix] [-k prefix] [-l prefix] [-i] [-q]
 *                    file0 [file1 ...]
 *   sonde_b200_batch -w rate -F f0[,f1...] [the same options] wideband.c64
 *     -t, --type     auto|c50|dfm|imet4|ims100|m10|mrzn1|rs41 (the reference's names, SD/main.c:66; default auto),
 *                    or a comma-separated list, one per file
 *     -b, --buflen   samples per channel per GPU call (default 1024 = the reference's BUFLEN, SD/main.c:32).  Larger buffers
 *                    mean fewer, fuller GPU calls; the output then equals the reference LIBRARY driven with buffers of that
 *                    length rather than the reference TOOL byte for byte, because the reference's demodulator forgets its
 *                    mid-symbol sample at the start of every buffer (SD/demod/gfsk.c:73), which the timing loop feels
 *                    while it is acquiring
 *     -f, --fmt      the reference's line format for a data point (SD/main.c:489-566: %S serial, %f frame counter, %t
 *                    temperature, %l latitude, ... ; its default when -o is given without -f)
 *     -o, --output   write <prefix><channel>.txt per channel: the lines the reference prints on stdout for that recording
 *     -c, --csv      write <prefix><channel>.csv per channel
 *     -g, --gpx      write <prefix><channel>.gpx per channel: one track per serial, points that carry position and speed
 *     -k, --kml      write <prefix><channel>.kml per channel
 *     -l, --live-kml write <prefix><channel>.kml (a network link) and <prefix><channel>.kml-live.kml, which is complete
 *                    after every point
 *     -i, --iq       the files are raw complex64 IQ at 48 kS/s instead of FM audio
 *     -w, --wideband <rate>   ONE input file: raw complex64 IQ of a wideband receiver at <rate> S/s (rate * L / M = 48000 with
 *                    1 <= L <= 16, e.g. 2304000, 2048000, 2500000).  The channels are cut out of it on the GPU
 *                    (include/sonde_b200_channelizer.h: polyphase filter bank on the tensor cores, per-type channel
 *                    bandwidths as the plugin's VFOs, src/main.hpp:45-51) and decoded without leaving the device
 *     -F, --freqs <f0,f1,...> with -w: the channel centres in Hz relative to the centre of the recording, one channel each
 * Input files are read the way the reference reads them (SD/main.c:248-266, SD/io/wavfile.c): a file that starts with a
 * 44-byte RIFF/WAVE header is a WAV recording (8 / 16 / 32-bit samples, first channel, 32 KiB blocks — a trailing partial
 * block is ignored, as there), anything else raw mono float32 at 48 kHz read from behind those 44 bytes.  All recordings
 * of a batch must have the same sample rate.
 *     -q, --quiet    no per-point lines on stdout
 * stdout (unless -q): one line per data point  "<channel> <serial> <seq> <lat> <lon> <alt> <temp>", or with -f
 *                                             "<channel> <the formatted line>"
 * and always one summary line per channel     "CH <channel> type=<decoder> frames=<n> ok=<n> points=<n>".
 * Exit code 3 when the CUDA path is unavailable: there is no CPU fallback.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/sonde_b200.h"
#include "sonde_data.hpp"
#include "../../include/sonde_b200_channelizer.h"
#include "gpu_wideband.hpp"                  /* GpuWidebandBank::typeCutoffHz: the per-type channel bandwidths */
#include "telemetry.hpp"
#include "track_files.hpp"

namespace {

/* One recording, read like the reference's wav_read_wrapper / raw_read_wrapper (SD/main.c:384-405, SD/io/wavfile.c:33-113) */
struct Input {
	FILE *f = nullptr;
	bool wav = false;
	int bps = 32, nch = 1, rate = 48000;
	std::vector<unsigned char> blk;
	size_t off = 0;                              /* position in the current block, in samples of bps/8 bytes */
	bool have_blk = false;

	bool open(const char *name, bool iq)
	{
		if (!(f = fopen(name, "rb"))) return false;
		if (iq) return true;
		unsigned char hdr[44];
		if (fread(hdr, sizeof(hdr), 1, f) == 1 && !memcmp(hdr, "RIFF", 4) && !memcmp(hdr + 8, "WAVE", 4)) {
			nch = hdr[22] | (hdr[23] << 8);
			rate = (int)(hdr[24] | (hdr[25] << 8) | (hdr[26] << 16) | ((unsigned)hdr[27] << 24));
			bps = hdr[34] | (hdr[35] << 8);
			if (bps && nch) { wav = true; blk.resize(32768); }
		}
		/* not a WAV file: raw float32 — the reference goes on reading behind the 44 bytes its header probe consumed
		 * (a file shorter than that is read from wherever the short read left it) */
		return true;
	}
	/* raw: the number of samples read (a short last read is decoded with the previous buffer's tail, SD/main.c:328-333);
	 * WAV: `count` or 0 — a read that cannot be completed ends the recording (SD/io/wavfile.c:53-113: blocks are read
	 * when the first sample of a new block is needed, a block that cannot be read whole ends the file).  A frame of
	 * interleaved channels may straddle two blocks (3 channels of 16 bits: 16384 words per block): the position of the
	 * next first-channel sample is carried into the next block.  The reference's reader never advances in that case
	 * and hangs (its copy count becomes 0 with data left, wavfile.c:66,110); with 1, 2, 4, ... channels, where frames
	 * never straddle, both read the same samples. */
	size_t read(void *dst_, size_t count, size_t esz)
	{
		if (!f) return 0;
		if (!wav) return fread(dst_, esz, count, f);
		if (bps != 8 && bps != 16 && bps != 32) return 0;
		float *dst = (float *)dst_;
		const size_t per = blk.size() / (size_t)(bps / 8), want = count;
		while (count > 0) {
			if (!have_blk) {
				if (fread(blk.data(), blk.size(), 1, f) != 1) return 0;
				have_blk = true;
			}
			for (; count > 0 && off < per; off += (size_t)nch, count--) {
				if (bps == 8)       *dst++ = (float)(blk[off] - 127);
				else if (bps == 16) { short v; memcpy(&v, &blk[2 * off], 2); *dst++ = (float)v; }
				else                { float v; memcpy(&v, &blk[4 * off], 4); *dst++ = v; }
			}
			if (off >= per) { off -= per; have_blk = false; }
		}
		return want;
	}
	void close() { if (f) fclose(f); f = nullptr; }
};

const char *kNames[] = {"auto", "c50", "dfm", "imet4", "ims100", "m10", "mrzn1", "rs41"};
const int kTypes[] = {SONDE_AUTO, SONDE_C50, SONDE_DFM09, SONDE_IMET4, SONDE_IMS100, SONDE_M10, SONDE_MRZN1, SONDE_RS41};

int type_of(const std::string &s)
{
	for (size_t i = 0; i < sizeof(kNames) / sizeof(*kNames); i++)
		if (s == kNames[i]) return kTypes[i];
	fprintf(stderr, "unknown sonde type '%s'\n", s.c_str());
	exit(2);
}
const char *name_of(int type)
{
	for (size_t i = 0; i < sizeof(kNames) / sizeof(*kNames); i++)
		if (kTypes[i] == type) return kNames[i];
	return "?";
}

/* SD/decode.c:278-376: fold one fragment into the channel's accumulated record; true if it carried data */
bool append_data_point(SondeData &printable, const SondeData &d)
{
	if (!d.fields) return false;
	if (d.fields & DATA_PTU) {
		printable.fields |= DATA_PTU;
		printable.temp = d.temp; printable.rh = d.rh; printable.pressure = d.pressure;
		printable.calib_percent = d.calib_percent;
	}
	if (d.fields & DATA_TIME) { printable.fields |= DATA_TIME; printable.time = d.time; }
	if (d.fields & DATA_POS) { printable.fields |= DATA_POS; printable.lat = d.lat; printable.lon = d.lon; printable.alt = d.alt; }
	if (d.fields & DATA_SPEED) {
		printable.fields |= DATA_SPEED;
		printable.speed = d.speed; printable.heading = d.heading; printable.climb = d.climb;
	}
	if (d.fields & DATA_SERIAL) {
		printable.fields |= DATA_SERIAL;
		strncpy(printable.serial, d.serial, sizeof(printable.serial) - 1);
		printable.serial[sizeof(printable.serial) - 1] = 0;
	}
	if (d.fields & DATA_OZONE) { printable.fields |= DATA_OZONE; printable.o3_mpa = d.o3_mpa; }
	if (d.fields & DATA_SHUTDOWN) { printable.fields |= DATA_SHUTDOWN; printable.shutdown = d.shutdown; }
	if (d.fields & DATA_SEQ) { printable.fields |= DATA_SEQ; printable.seq = d.seq; }
	if (!(printable.pressure > 0)) printable.pressure = radiosonde::tl::altitude_to_pressure(printable.alt);
	return true;
}

void usage(const char *prog)
{
	printf("usage: %s [options] file0 [file1 ...]      one GPU batch, one channel per file\n"
	       "   -t, --type <type[,type...]>  auto|c50|dfm|imet4|ims100|m10|mrzn1|rs41, one for all files or one per file (default auto)\n"
	       "   -b, --buflen <samples>       samples per channel per GPU call (default 1024)\n"
	       "   -i, --iq                     the files are raw complex64 IQ at 48 kS/s (default: WAV, or raw mono float32 FM audio)\n"
	       "   -w, --wideband <rate>        ONE file of raw complex64 wideband IQ at <rate> S/s, channels cut out on the GPU\n"
	       "   -F, --freqs <f0[,f1...]>     with -w: channel centres in Hz relative to the centre of the recording\n"
	       "   -f, --fmt <format>           format of the text line per data point\n"
	       "   -o, --output <prefix>        text lines to <prefix><channel>.txt\n"
	       "   -c, --csv <prefix>           CSV to <prefix><channel>.csv\n"
	       "   -g, --gpx <prefix>           GPX track to <prefix><channel>.gpx\n"
	       "   -k, --kml <prefix>           KML track to <prefix><channel>.kml\n"
	       "   -l, --live-kml <prefix>      live KML to <prefix><channel>.kml + <prefix><channel>.kml-live.kml\n"
	       "   -q, --quiet                  no per-point lines on stdout\n"
	       "\nFormat specifiers (default \"%s\"):\n"
	       "   %%a altitude (m)        %%b shutdown timer      %%c climb rate (m/s)    %%d dew point ('C)\n"
	       "   %%f frame counter       %%h heading (deg)       %%l latitude            %%o longitude\n"
	       "   %%p pressure (hPa)      %%r rel. humidity (%%)   %%s speed (m/s)         %%S serial number\n"
	       "   %%t temperature ('C)    %%T time stamp (UTC)    %%x decoded XDATA\n",
	       prog, radiosonde::cli::default_format());
}

}  // namespace

int main(int argc, char **argv)
{
	std::string type_arg = "auto", csv_prefix, gpx_prefix, kml_prefix, live_prefix, txt_prefix, fmt;
	bool have_fmt = false;
	double wide_rate = 0;
	std::vector<double> freqs;
	size_t buflen = 1024;
	bool iq = false, quiet = false;
	std::vector<std::string> files;
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		auto need = [&](const char *what) { if (i + 1 >= argc) { fprintf(stderr, "%s needs an argument\n", what); exit(2); } return std::string(argv[++i]); };
		if (a == "-t" || a == "--type") type_arg = need("-t");
		else if (a == "-b" || a == "--buflen") buflen = strtoul(need("-b").c_str(), nullptr, 10);
		else if (a == "-c" || a == "--csv") csv_prefix = need("-c");
		else if (a == "-f" || a == "--fmt") { fmt = need("-f"); have_fmt = true; }
		else if (a == "-o" || a == "--output") txt_prefix = need("-o");
		else if (a == "-g" || a == "--gpx") gpx_prefix = need("-g");
		else if (a == "-k" || a == "--kml") kml_prefix = need("-k");
		else if (a == "-l" || a == "--live-kml") live_prefix = need("-l");
		else if (a == "-i" || a == "--iq") iq = true;
		else if (a == "-w" || a == "--wideband") wide_rate = atof(need("-w").c_str());
		else if (a == "-F" || a == "--freqs") {
			const std::string list = need("-F");
			for (size_t pos = 0; pos <= list.size();) {
				const size_t c = list.find(',', pos);
				freqs.push_back(atof(list.substr(pos, c == std::string::npos ? std::string::npos : c - pos).c_str()));
				if (c == std::string::npos) break;
				pos = c + 1;
			}
		}
		else if (a == "-q" || a == "--quiet") quiet = true;
		else if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
		else files.push_back(a);
	}
	const bool wideband = wide_rate > 0;
	if (wideband && (files.size() != 1 || freqs.empty())) { fprintf(stderr, "-w takes one input file and -F f0[,f1...]\n"); return 2; }
	const size_t C = wideband ? freqs.size() : files.size();
	if (!C || !buflen) { fprintf(stderr, "no input files\n"); return 2; }
	/* wideband: rate * L / D = 48000 (host/gpu_wideband.hpp:56-62); one call takes n_in = buflen / L * D input samples */
	int L = 0, D = 0;
	if (wideband) {
		for (int l = 1; l <= 16 && !L; l++) {
			const long long mi = (long long)(wide_rate * l / 48000.0 + 0.5);
			if (mi >= 2 && mi > l && (double)mi * 48000.0 == wide_rate * l) { L = l; D = (int)mi; }
		}
		if (!L) { fprintf(stderr, "-w %g: rate * L / M must be 48000 with 1 <= L <= 16, M >= 2\n", wide_rate); return 2; }
		buflen = (buflen + L - 1) / L * L;
		iq = true;
	}
	const size_t n_in = wideband ? buflen / L * D : 0;
	std::vector<int32_t> types;
	{
		size_t pos = 0;
		while (pos <= type_arg.size()) {
			const size_t c = type_arg.find(',', pos);
			types.push_back(type_of(type_arg.substr(pos, c == std::string::npos ? std::string::npos : c - pos)));
			if (c == std::string::npos) break;
			pos = c + 1;
		}
		if (types.size() == 1) types.assign(C, types[0]);
		if (types.size() != C) { fprintf(stderr, "%zu types for %zu channels\n", types.size(), C); return 2; }
	}

	const size_t esz = iq ? 8 : 4;
	std::vector<Input> in(wideband ? 1 : C);
	for (size_t c = 0; c < in.size(); c++) {
		if (!in[c].open(files[c].c_str(), iq)) { fprintf(stderr, "cannot open %s\n", files[c].c_str()); return 2; }
		if (in[c].rate != in[0].rate) { fprintf(stderr, "%s: %d S/s, the batch runs at %d S/s\n", files[c].c_str(), in[c].rate, in[0].rate); return 2; }
	}

	sonde_b200_config cfg = {};
	cfg.n_channels = (int32_t)C;
	cfg.samplerate = wideband ? 48000 : in[0].rate;
	cfg.max_chunk_len = (int32_t)buflen;
	cfg.types = types.data();
	sonde_b200 *h = nullptr;
	const int rc = sonde_b200_create(&h, &cfg);
	if (rc == SONDE_ERR_NODEVICE || rc == SONDE_ERR_CUDA) {
		printf("NOGPU sonde_b200_create failed (%d): the CUDA path is required, there is no CPU fallback\n", rc);
		return 3;
	}
	if (rc != SONDE_OK) {
		fprintf(stderr, "sonde_b200_create failed (%d): check the sample rate (%d S/s) and the buffer length (%zu)\n", rc, in[0].rate, buflen);
		return 2;
	}
	sonde_chan *chan = nullptr;
	if (wideband) {
		sonde_chan_config cc = {};
		cc.n_channels = (int32_t)C;
		cc.decim = D;
		cc.fs_out = 48000;
		cc.max_in_len = (int32_t)n_in;
		cc.freq_hz = freqs.data();
		std::vector<float> cut(C);
		for (size_t c = 0; c < C; c++) cut[c] = radiosonde::GpuWidebandBank::typeCutoffHz(types[c]);
		sonde_chan_options co = {};
		co.interp = L;
		co.cutoff_hz = cut.data();
		const int crc = sonde_chan_create_ex(&chan, &cc, &co);
		if (crc != SONDE_OK) { fprintf(stderr, "sonde_chan_create failed (%d) for %zu channels, L/M = %d/%d\n", crc, C, L, D); return crc == SONDE_ERR_NODEVICE ? 3 : 2; }
	}
	const int max_frames = sonde_b200_max_frames(h);
	std::vector<sonde_frame_rec> recs(C * (size_t)max_frames);
	std::vector<int32_t> counts(C), locked(types);
	std::vector<radiosonde::Telemetry> tele;
	std::vector<SondeData> printable(C), fragment(C);
	struct Outputs {
		radiosonde::cli::CsvFile csv;
		radiosonde::cli::GpxFile gpx;
		radiosonde::cli::KmlFile kml, live;
		FILE *txt = nullptr;
	};
	if (!have_fmt) fmt = radiosonde::cli::default_format();
	std::vector<Outputs> out(C);
	std::vector<long> n_frames(C, 0), n_ok(C, 0), n_points(C, 0);
	for (size_t c = 0; c < C; c++) {
		tele.emplace_back(types[c] == SONDE_AUTO ? SONDE_RS41 : types[c]);
		memset(&printable[c], 0, sizeof(SondeData));
		memset(&fragment[c], 0, sizeof(SondeData));
		const std::string ch = std::to_string(c);
		if ((!csv_prefix.empty() && !out[c].csv.init((csv_prefix + ch + ".csv").c_str())) ||
		    (!gpx_prefix.empty() && !out[c].gpx.init((gpx_prefix + ch + ".gpx").c_str())) ||
		    (!kml_prefix.empty() && out[c].kml.init((kml_prefix + ch + ".kml").c_str(), false)) ||
		    (!live_prefix.empty() && out[c].live.init((live_prefix + ch + ".kml").c_str(), true)) ||
		    (!txt_prefix.empty() && !(out[c].txt = fopen((txt_prefix + ch + ".txt").c_str(), "wb")))) {
			fprintf(stderr, "cannot create the output files of channel %zu\n", c);
			return 2;
		}
	}
	/* two pinned staging buffers: buffer k+1 is read and submitted while buffer k decodes (sonde_b200.h, fetch()) */
	const size_t stage_bytes = wideband ? n_in * 8 : C * buflen * esz;
	char *stage[2] = {(char *)sonde_b200_host_alloc(stage_bytes), (char *)sonde_b200_host_alloc(stage_bytes)};
	if (!stage[0] || !stage[1]) { fprintf(stderr, "pinned allocation failed\n"); return 2; }

	long n_submitted = 0, n_delivered = 0;
	std::vector<long> last_call(C, -1);          /* the reference stops decoding a recording at its end: records that
	                                                later (all-zero) buffers complete for it are not reported */
	auto submit = [&](int slot) -> bool {
		/* The reference's loop (SD/main.c:328-333, raw_read_wrapper :400-405): a read that returns at least one sample
		 * is followed by decode() over the WHOLE 1024-sample buffer, so a short last read is decoded together with the
		 * tail the previous read left in the buffer; a read of zero samples ends the recording.  Here a recording that
		 * has ended contributes exact zeros (which the AGC passes through untouched, agc.c:23) until all have ended. */
		bool any = false;
		if (wideband) {
			/* the next n_in samples of the recording (zeros behind its end), channelised on the decoder's stream into
			 * [C][stride] rows that the decoder reads in place */
			const size_t got = in[0].read(stage[slot], n_in, 8);
			if (got == 0) return false;
			if (got < n_in) memset(stage[slot] + got * 8, 0, (n_in - got) * 8);
			void *d_rows = nullptr;
			size_t stride = 0;
			if (sonde_chan_process_c64(chan, (const float *)stage[slot], n_in, sonde_b200_stream(h), &d_rows, &stride) != SONDE_OK) {
				fprintf(stderr, "sonde_chan: %s\n", sonde_chan_last_error(chan));
				exit(1);
			}
			if (sonde_b200_process_iq_device(h, d_rows, buflen, stride) != SONDE_OK) { fprintf(stderr, "sonde_b200: %s\n", sonde_b200_last_error(h)); exit(1); }
			for (size_t c = 0; c < C; c++) last_call[c] = n_submitted;
			n_submitted++;
			return true;
		}
		for (size_t c = 0; c < C; c++) {
			char *row = stage[slot] + c * buflen * esz;
			const size_t got = in[c].read(row, buflen, esz);
			if (got == 0) {
				memset(row, 0, buflen * esz);
				in[c].close();
			} else {
				any = true;
				last_call[c] = n_submitted;          /* this call still carries samples of recording c */
				if (got < buflen) {
					if (n_submitted) memcpy(row + got * esz, stage[slot ^ 1] + c * buflen * esz + got * esz, (buflen - got) * esz);
					else memset(row + got * esz, 0, (buflen - got) * esz);
				}
			}
		}
		if (!any) return false;
		const int r = iq ? sonde_b200_process_iq(h, (const float *)stage[slot], buflen) : sonde_b200_process_fm(h, (const float *)stage[slot], buflen);
		if (r != SONDE_OK) { fprintf(stderr, "sonde_b200: %s\n", sonde_b200_last_error(h)); exit(1); }
		n_submitted++;
		return true;
	};
	auto deliver = [&]() {
		if (sonde_b200_fetch(h, recs.data(), counts.data()) != SONDE_OK) { fprintf(stderr, "sonde_b200: %s\n", sonde_b200_last_error(h)); exit(1); }
		sonde_b200_detected_types(h, locked.data());
		const long call = n_delivered++;
		for (size_t c = 0; c < C; c++) {
			if (call > last_call[c]) continue;
			for (int k = 0; k < counts[c]; k++) {
				const sonde_frame_rec &r = recs[c * max_frames + k];
				if (types[c] == SONDE_AUTO && tele[c].type() != r.type) {
					/* the channel locked: SD/decode.c:242-251 set_active_decoder() clears the accumulated record */
					tele[c].reset(r.type);
					memset(&printable[c], 0, sizeof(SondeData));
				}
				n_frames[c]++;
				n_ok[c] += r.ok;
				/* The reference's decode() keeps its SondeData in one stack slot that every call reuses without clearing
				 * it (SD/decode.c:128), and decoders such as the DFM's write only the members a subframe carries: a
				 * fragment therefore shows the values the previous ones left behind.  One persistent fragment per
				 * channel reproduces that (it starts zeroed; the reference's first fragments show stack garbage in
				 * members no subframe has delivered yet). */
				SondeData &frag = fragment[c];
				tele[c].parse(r, &frag);
				if (!append_data_point(printable[c], frag)) continue;
				n_points[c]++;
				const SondeData &d = printable[c];
				if (!quiet) {
					if (have_fmt) { printf("%zu ", c); radiosonde::cli::print_data(stdout, fmt.c_str(), d); }
					else printf("%zu %s %d %.5f %.5f %.1f %.1f\n", c, d.serial, d.seq, d.lat, d.lon, d.alt, d.temp);
				}
				/* SD/main.c:347-365 */
				Outputs &o = out[c];
				if (o.txt) radiosonde::cli::print_data(o.txt, fmt.c_str(), d);
				o.csv.add_point(d);
				o.kml.start_track(d.serial);
				o.kml.add_trackpoint(d);
				o.live.start_track(d.serial);
				o.live.add_trackpoint(d);
				if (d.fields & DATA_SERIAL) o.gpx.start_track(d.serial);
				o.gpx.add_trackpoint(d);
			}
		}
	};
	int slot = 0, in_flight = 0;
	while (submit(slot)) {
		if (in_flight) deliver();
		in_flight = 1;
		slot ^= 1;
	}
	if (in_flight) deliver();
	for (size_t c = 0; c < C; c++) {
		printf("CH %zu type=%s frames=%ld ok=%ld points=%ld\n", c, name_of(locked[c]), n_frames[c], n_ok[c], n_points[c]);
		out[c].kml.close();
		out[c].live.close();
		out[c].gpx.close();
		out[c].csv.close();
		if (out[c].txt) fclose(out[c].txt);
		if (c < in.size()) in[c].close();
	}
	sonde_b200_host_free(stage[0]);
	sonde_b200_host_free(stage[1]);
	sonde_b200_destroy(h);                       /* the decoder first: it reads the channelizer's buffers */
	sonde_chan_destroy(chan);
	return 0;
}
