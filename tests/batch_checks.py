"""Checks of the batch command-line runner against the reference's command-line tool (oracle/_ref/sondedump_ref), written
once and run twice: on the GPU box with the product binary (tests/test_cli_dropin.py, tests/test_zz_batch_tracks.py) and on
the CPU with a test copy of the same runner source linked to an oracle-backed stand-in of the batch ABI
(tests/test_batch_host_logic.py), which exercises everything of the runner but the GPU library.  Not a test module."""
import os
import re
import subprocess

import numpy as np

from sdrpp_radiosonde_b200 import synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "sondedump_ref")


def check_csv_per_channel(runner, tmp_path, auto=True):
    """sonde_b200_batch (the runner of its own on the batch ABI): five recordings of five sonde
    types — of different lengths, one not a multiple of the 1024-sample buffer — decoded in ONE batch; every channel's
    CSV must be byte-identical to what the reference's CLI writes for that recording alone (-t <type> -c)."""
    cases = [("rs41", synth.RS41, 48000 * 6), ("dfm", synth.DFM09, 48000 * 5 + 1024 * 3), ("m10", synth.M10, 48000 * 4 + 517),
             ("c50", synth.C50, 48000 * 4), ("mrzn1", synth.MRZN1, 48000 * 5)]
    files = []
    for i, (flag, stype, n) in enumerate(cases):
        raw = tmp_path / f"in{i}.raw"
        synth.make_fm(synth.default_spec(stype, 10 + i), n).astype(np.float32).tofile(raw)
        files.append(str(raw))
    r = subprocess.run([runner, "-q", "-t", ",".join(c[0] for c in cases), "-c", str(tmp_path / "b200_"), *files],
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for i, (flag, stype, n) in enumerate(cases):
        ref_csv = tmp_path / f"ref{i}.csv"
        a = subprocess.run([REF, "-q", "-t", flag, "-c", str(ref_csv), files[i]], capture_output=True, timeout=600)
        assert a.returncode == 0, a.stderr[-300:]
        want, got = ref_csv.read_bytes(), (tmp_path / f"b200_{i}.csv").read_bytes()
        assert want.count(b"\n") >= 3, (flag, want[:200])
        # decode() passes an UNINITIALISED SondeData to the decoder (SD/decode.c:128) and e.g. dfm09_decode returns PARSED
        # for an undecodable first window without touching it: the reference then logs a "data point" whose fields are
        # whatever its stack held — a row of nothing but commas.  Such rows carry no data and are dropped from both files.
        def rows(b):
            return [l for l in b.split(b"\n") if l.strip(b",")]
        rg, rw = rows(got), rows(want)
        assert rg[0] == rw[0]
        # start-up: the reference's per-decoder state is malloc()ed and not cleared (e.g. MRZ-N1 calibration,
        # SD/sonde/mrz-n1/mrzn1.c:13-28), so whether the very first frame already yields a data point depends on heap
        # garbage; this repo's parsers start from zeros.  At most one such leading row may differ — every
        # row after it must be identical.
        n = min(len(rg), len(rw)) - 1
        assert abs(len(rg) - len(rw)) <= 1 and n >= 3 and rg[-n:] == rw[-n:], (flag, got[:300], want[:300])
    if not auto:
        return
    # the AUTO path of the runner: same recordings, every channel autodetects its decoder
    r2 = subprocess.run([runner, "-q", "-t", "auto", *files], capture_output=True, timeout=600)
    assert r2.returncode == 0
    locked = [l.split()[2].split("=")[1] for l in r2.stdout.decode().splitlines() if l.startswith("CH ")]
    assert locked == [c[0] for c in cases], r2.stdout[-400:]




def _write_wav(path, data, rate=48000):
    """data: [n] or [n][channels], int16 or float32; the plain 44-byte header the reference's wav_parse expects"""
    import struct
    data = np.ascontiguousarray(data)
    nch = 1 if data.ndim == 1 else data.shape[1]
    bps = data.dtype.itemsize * 8
    raw = data.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1 if bps == 16 else 3, nch, rate, rate * nch * bps // 8, nch * bps // 8, bps) + b"data" + struct.pack("<I", len(raw))
    assert len(hdr) == 44
    with open(path, "wb") as f:
        f.write(hdr + raw)


def check_wav_inputs(runner, tmp_path):
    """WAV recordings through sonde_b200_batch, read the way the reference reads them (SD/io/wavfile.c: 44-byte header,
    first channel, raw sample values, 32 KiB blocks with the trailing partial block ignored): 16-bit mono, 32-bit float
    stereo (second channel is noise) and a raw float32 file in one batch; every channel's CSV equals sondedump_ref's."""
    rng = np.random.default_rng(3)
    n = 48000 * 5 + 777
    fm0 = synth.make_fm(synth.default_spec(synth.RS41, 21), n)
    fm1 = synth.make_fm(synth.default_spec(synth.M10, 22), n)
    fm2 = synth.make_fm(synth.default_spec(synth.DFM09, 23), n)
    w0, w1, r2 = tmp_path / "a.wav", tmp_path / "b.wav", tmp_path / "c.raw"
    _write_wav(w0, np.clip(np.round(fm0 * (12000.0 / np.abs(fm0).max())), -32768, 32767).astype(np.int16))
    _write_wav(w1, np.stack([fm1.astype(np.float32), rng.standard_normal(n).astype(np.float32)], axis=1))
    fm2.astype(np.float32).tofile(r2)
    files, flags = [str(w0), str(w1), str(r2)], ["rs41", "m10", "dfm"]
    r = subprocess.run([runner, "-q", "-t", ",".join(flags), "-c", str(tmp_path / "b200_"), *files], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for i, flag in enumerate(flags):
        ref_csv = tmp_path / f"ref{i}.csv"
        a = subprocess.run([REF, "-q", "-t", flag, "-c", str(ref_csv), files[i]], capture_output=True, timeout=600)
        assert a.returncode == 0, a.stderr[-300:]

        def rows(b):
            # a time value that does not fit the reference's fixed buffer (the synthetic DFM's date decodes to an 8-digit
            # year) is printed cut off, followed by whatever byte of the reference's stack comes next (SD/io/csv.c): the time
            # field of such rows is masked in both files, everything else is compared
            ok = re.compile(rb"^\d{4}-\d{2}-\d{2}T\d{2}:\d{2}:\d{2}Z,")
            out = []
            for l in b.split(b"\n"):
                if not l.strip(b","):
                    continue
                out.append(l if (ok.match(l) or l.startswith(b"Time,")) else b"<time>," + l.split(b",", 1)[-1])
            return out
        rg, rw = rows((tmp_path / f"b200_{i}.csv").read_bytes()), rows(ref_csv.read_bytes())
        assert len(rw) >= 4, (flag, rw[:2])
        m = min(len(rg), len(rw)) - 1
        assert abs(len(rg) - len(rw)) <= 1 and m >= 3 and rg[-m:] == rw[-m:], (flag, rg[:3], rw[:3])


def gpx_parts(b: bytes):
    """(everything that is not a track point, [track points])"""
    pts = re.findall(rb"<trkpt .*?</trkpt>\n", b, flags=re.S)
    return re.sub(rb"<trkpt .*?</trkpt>\n", b"", b, flags=re.S), pts


def kml_parts(b: bytes):
    """(everything but coordinate rows and the closing position marker, [coordinate rows])"""
    rows = re.findall(rb"^-?[0-9.naif]+,-?[0-9.naif]+,-?[0-9.naif]+\n", b, flags=re.M)
    rest = re.sub(rb"^-?[0-9.naif]+,-?[0-9.naif]+,-?[0-9.naif]+\n", b"", b, flags=re.M)
    rest = re.sub(rb"<Placemark>\s*<name>[^<]*</name>\s*<Point>.*?</Point>\s*</Placemark>\s*", b"", rest, flags=re.S)
    return rest, rows


def same_but_first(got, want, least):
    """The reference's decoders start from uncleared heap and stack (tests/test_cli_dropin.py), so its very first data
    point may differ from a run that starts from zeros; every later one must be identical."""
    n = min(len(got), len(want)) - 1
    return abs(len(got) - len(want)) <= 1 and n >= least and got[-n:] == want[-n:]


def check_tracks(runner, tmp_path):
    """Per-channel GPX, KML and live KML of a batch (-g / -k / -l) against what the reference's tool writes for each
    recording alone (-g / -k)."""
    cases = [("rs41", synth.RS41, 48000 * 6, 21), ("rs41", synth.RS41, 48000 * 5 + 300, 22), ("m10", synth.M10, 48000 * 4, 23)]
    files = []
    for i, (flag, stype, n, seed) in enumerate(cases):
        raw = tmp_path / f"in{i}.raw"
        synth.make_fm(synth.default_spec(stype, seed), n).astype(np.float32).tofile(raw)
        files.append(str(raw))
    r = subprocess.run([runner, "-q", "-t", ",".join(c[0] for c in cases), "-g", str(tmp_path / "g_"), "-k", str(tmp_path / "k_"),
                        "-l", str(tmp_path / "l_"), *files], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for i, (flag, stype, n, seed) in enumerate(cases):
        a = subprocess.run([REF, "-q", "-t", flag, "-g", str(tmp_path / f"ref{i}.gpx"), "-k", str(tmp_path / f"ref{i}.kml"), files[i]],
                           capture_output=True, timeout=600)
        assert a.returncode == 0, a.stderr[-300:]
        want_rest, want_pts = gpx_parts((tmp_path / f"ref{i}.gpx").read_bytes())
        got_rest, got_pts = gpx_parts((tmp_path / f"g_{i}.gpx").read_bytes())
        assert same_but_first(got_pts, want_pts, 2), (flag, len(got_pts), len(want_pts))
        want_krest, want_rows = kml_parts((tmp_path / f"ref{i}.kml").read_bytes())
        got_krest, got_rows = kml_parts((tmp_path / f"k_{i}.kml").read_bytes())
        assert same_but_first(got_rows, want_rows, 2), (flag, len(got_rows), len(want_rows))
        if flag == "rs41":
            # one serial for the whole recording: the files without their points are identical
            assert got_rest == want_rest, (got_rest[:400], want_rest[:400])
            assert got_krest == want_krest, (got_krest[-400:], want_krest[-400:])
        else:
            # the synthetic M10 / M20 frames change serial from frame to frame: one track per frame, same names in
            # the same order (the first may be the reference's start-up point, see same_but_first)
            names = lambda b: re.findall(rb"<name>[^<]*</name>", b)
            assert same_but_first(names(got_rest), names(want_rest), 2) and same_but_first(names(got_krest), names(want_krest), 2)
            assert got_rest.endswith(b"</trkseg>\n</trk>\n</gpx>\n") and got_krest.endswith(b"</Placemark>\n</Document>\n</kml>\n")
        # the live file carries the same track as the plain one and a closing trailer (it is never truncated, so bytes
        # of an older, longer trailer may follow the current one: SD/io/kml.c:143-161, host/track_files.hpp)
        live = (tmp_path / f"l_{i}.kml-live.kml").read_bytes()
        assert b"</Document>\n</kml>\n" in live and kml_parts(live)[1] == got_rows
        assert b"<NetworkLink>" in (tmp_path / f"l_{i}.kml").read_bytes()


ALL_SPECIFIERS = "%S|%f|%t|%r|%d|%p|%l|%o|%a|%s|%h|%c|%b|%T|%x|%%|%q|100%"


def check_text_output(runner, tmp_path):
    """-o / -f: the per-channel text file of a batch holds the lines the reference's tool prints on stdout for that
    recording alone — with its default format (SD/main.c:111) and with a format that uses every specifier of
    SD/main.c:489-566, an unknown one, a doubled and a trailing percent sign."""
    cases = [("rs41", synth.RS41, 48000 * 6, 31), ("m10", synth.M10, 48000 * 4 + 100, 32), ("mrzn1", synth.MRZN1, 48000 * 5, 33)]
    files = []
    for i, (flag, stype, n, seed) in enumerate(cases):
        raw = tmp_path / f"in{i}.raw"
        synth.make_fm(synth.default_spec(stype, seed), n).astype(np.float32).tofile(raw)
        files.append(str(raw))
    for tag, fmt in (("d", None), ("a", ALL_SPECIFIERS)):
        extra = ["-f", fmt] if fmt else []
        r = subprocess.run([runner, "-q", "-t", ",".join(c[0] for c in cases), "-o", str(tmp_path / f"{tag}_"), *extra, *files],
                           capture_output=True, timeout=600)
        assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
        for i, (flag, stype, n, seed) in enumerate(cases):
            a = subprocess.run([REF, "-q", "-t", flag, *extra, files[i]], capture_output=True, timeout=600)
            assert a.returncode == 0, a.stderr[-300:]
            # the tool's log lines share stdout with the data lines ("[<file>:<line>] Could not recognize input file type")
            want = [l for l in a.stdout.split(b"\n")[:-1] if not l.startswith(b"[\x1b[")]
            got = (tmp_path / f"{tag}_{i}.txt").read_bytes().split(b"\n")[:-1]
            assert same_but_first(got, want, 3), (flag, tag, got[:3], want[:3])
    # without -q and with -f the same lines go to stdout behind the channel number
    r = subprocess.run([runner, "-t", cases[0][0], "-f", "%S %f", files[0]], capture_output=True, timeout=600)
    lines = [l for l in r.stdout.decode().splitlines() if not l.startswith("CH ")]
    assert r.returncode == 0 and len(lines) >= 4 and all(re.fullmatch(r"0 \S+ +\d+", l) for l in lines), lines[:3]


def check_iq_input(runner, tmp_path):
    """-i: raw complex64 recordings.  The same recordings discriminated on the CPU (the oracle's discriminator, the stage
    the IQ entry point adds in front of the chain) and fed as FM files must give the same per-channel CSV.  The FM
    files carry 44 leading bytes because the runner, like the reference, reads raw FM files from behind its WAV header
    probe (SD/main.c:248-259); lengths are whole buffers so that no padded tail differs."""
    from tests import reflib
    orc = reflib.OracleLib()
    cases = [("rs41", synth.RS41, 1024 * 260, 41), ("m10", synth.M10, 1024 * 200, 42)]
    iq_files, fm_files = [], []
    for i, (flag, stype, n, seed) in enumerate(cases):
        iq = synth.make_iq(synth.default_spec(stype, seed), n).astype(np.complex64)
        fm = np.asarray(orc.discriminate(iq), dtype=np.float32)
        assert fm.shape == (n,)
        (tmp_path / f"in{i}.c64").write_bytes(iq.tobytes())
        (tmp_path / f"in{i}.raw").write_bytes(b"\0" * 44 + fm.tobytes())
        iq_files.append(str(tmp_path / f"in{i}.c64"))
        fm_files.append(str(tmp_path / f"in{i}.raw"))
    types = ",".join(c[0] for c in cases)
    a = subprocess.run([runner, "-q", "-i", "-t", types, "-c", str(tmp_path / "iq_"), *iq_files], capture_output=True, timeout=600)
    b = subprocess.run([runner, "-q", "-t", types, "-c", str(tmp_path / "fm_"), *fm_files], capture_output=True, timeout=600)
    assert a.returncode == 0 and b.returncode == 0, a.stderr[-300:] + b.stderr[-300:]
    assert a.stdout == b.stdout and a.stdout.count(b"CH ") == len(cases)
    for i in range(len(cases)):
        got, want = (tmp_path / f"iq_{i}.csv").read_bytes(), (tmp_path / f"fm_{i}.csv").read_bytes()
        assert want.count(b"\n") >= 5 and got == want, (cases[i][0], got[:300], want[:300])


def check_wall_clock_sondes(runner, tmp_path):
    """iMS-100 and iMet-4 in one batch: their parsers take the DATE of the time stamp from time(NULL)
    (SD/sonde/ims100/parser.c:26, imet4/parser.c:47,70), so dates are masked; everything else in the per-channel CSV
    and text output must equal the reference tool's, with the usual tolerance for its very first data point."""
    def mask(b):
        return re.sub(rb"\d{4}-\d{2}-\d{2}", b"DATE", b)
    cases = [("ims100", synth.IMS100, 48000 * 6, 51), ("imet4", synth.IMET4, 48000 * 6, 52)]
    files = []
    for i, (flag, stype, n, seed) in enumerate(cases):
        raw = tmp_path / f"in{i}.raw"
        synth.make_fm(synth.default_spec(stype, seed), n).astype(np.float32).tofile(raw)
        files.append(str(raw))
    fmt = "%S|%f|%t|%r|%p|%l|%o|%a|%s|%h|%c"
    r = subprocess.run([runner, "-q", "-t", ",".join(c[0] for c in cases), "-c", str(tmp_path / "c_"), "-o", str(tmp_path / "o_"), "-f", fmt, *files],
                       capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for i, (flag, stype, n, seed) in enumerate(cases):
        a = subprocess.run([REF, "-q", "-t", flag, "-c", str(tmp_path / f"ref{i}.csv"), "-f", fmt, files[i]], capture_output=True, timeout=600)
        assert a.returncode == 0, a.stderr[-300:]
        want = [l for l in mask((tmp_path / f"ref{i}.csv").read_bytes()).split(b"\n") if l.strip(b",")]
        got = [l for l in mask((tmp_path / f"c_{i}.csv").read_bytes()).split(b"\n") if l.strip(b",")]
        assert got[0] == want[0] and same_but_first(got[1:], want[1:], 3), (flag, got[:3], want[:3])
        want_txt = [l for l in a.stdout.split(b"\n")[:-1] if not l.startswith(b"[\x1b[")]
        got_txt = (tmp_path / f"o_{i}.txt").read_bytes().split(b"\n")[:-1]
        assert same_but_first(got_txt, want_txt, 3), (flag, got_txt[:3], want_txt[:3])


def check_wideband_input(runner, tmp_path):
    """-w / -F: ONE wideband complex64 recording (three sondes at three offsets, 48 x 48 kS/s) channelised and decoded
    in one batch.  Every channel's CSV rows must be rows the reference's tool writes for that sonde's own narrowband
    signal (discriminated on the CPU), in the same order, all of them but possibly the first and the last (the channel
    filter delays the signal by a few samples, so a frame at either end of the recording may be cut)."""
    from tests import reflib
    from tests.gpu_util import make_wideband
    orc = reflib.OracleLib()
    D, nsec = 48, 5
    types, flags = [synth.RS41, synth.M10, synth.DFM09], ["rs41", "m10", "dfm"]
    freqs = [-400e3, 250e3, 31.25e3]
    nb, wide = make_wideband(types, freqs, D, nsec)
    (tmp_path / "wide.c64").write_bytes(wide.tobytes())
    r = subprocess.run([runner, "-q", "-w", str(48000 * D), "-F", ",".join(str(f) for f in freqs), "-t", ",".join(flags),
                        "-c", str(tmp_path / "w_"), str(tmp_path / "wide.c64")], capture_output=True, timeout=900)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]

    def rows(b):
        return [l for l in b.split(b"\n")[1:] if l.strip(b",")]
    for c, flag in enumerate(flags):
        fm = np.asarray(orc.discriminate(nb[c].astype(np.complex64)), dtype=np.float32)
        (tmp_path / f"nb{c}.raw").write_bytes(b"\0" * 44 + fm.tobytes())
        a = subprocess.run([REF, "-q", "-t", flag, "-c", str(tmp_path / f"ref{c}.csv"), str(tmp_path / f"nb{c}.raw")], capture_output=True, timeout=600)
        assert a.returncode == 0, a.stderr[-300:]
        want, got = rows((tmp_path / f"ref{c}.csv").read_bytes()), rows((tmp_path / f"w_{c}.csv").read_bytes())
        assert len(want) >= 4 and len(got) >= len(want) - 2, (flag, len(got), len(want))
        # got[1:] must appear in want as one contiguous run
        core = got[1:]
        assert any(want[i:i + len(core)] == core for i in range(len(want) - len(core) + 1)), (flag, got[:3], want[:3])
