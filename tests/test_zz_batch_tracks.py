"""sonde_b200_batch -g / -k: the per-channel GPX and KML tracks of a batch against what the
reference's command-line tool writes for each recording alone.  The writers themselves are checked byte for byte on the
CPU (tests/test_track_files.py); this is the same check through the GPU decode."""
import os
import re
import subprocess

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "sondedump_ref")
BATCH = os.path.join(ROOT, "sdrpp_radiosonde_b200", "sonde_b200_batch")


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


def test_part_helpers():
    g = (b"<gpx>\n<trk>\n<name>A</name>\n<trkseg>\n<trkpt lat=\"1.0\" lon=\"2.0\">\n<time>t</time>\n</trkpt>\n"
         b"<trkpt lat=\"3.0\" lon=\"4.0\">\n<time>u</time>\n</trkpt>\n</trkseg>\n</trk>\n</gpx>\n")
    rest, pts = gpx_parts(g)
    assert len(pts) == 2 and rest == b"<gpx>\n<trk>\n<name>A</name>\n<trkseg>\n</trkseg>\n</trk>\n</gpx>\n"
    k = (b"<coordinates>\n9.5,45.25,1000.0\n-9.5,-45.25,nan\n</coordinates>\n</LineString>\n</Placemark>\n<Placemark>\n<name>(null)</name>\n"
         b"<Point>\n<altitudeMode>absolute</altitudeMode>\n<coordinates>9.5,45.25,1000.0</coordinates>\n</Point>\n</Placemark>\n</Document>\n")
    rest, rows = kml_parts(k)
    assert len(rows) == 2 and rest == b"<coordinates>\n</coordinates>\n</LineString>\n</Placemark>\n</Document>\n"
    assert same_but_first([b"x", b"b", b"c"], [b"b", b"c"], 1) and not same_but_first([b"a", b"b", b"c"], [b"a", b"x", b"c"], 1)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_gpx_kml_equal_reference_cli_per_channel(tmp_path):
    cases = [("rs41", synth.RS41, 48000 * 6, 21), ("rs41", synth.RS41, 48000 * 5 + 300, 22), ("m10", synth.M10, 48000 * 4, 23)]
    files = []
    for i, (flag, stype, n, seed) in enumerate(cases):
        raw = tmp_path / f"in{i}.raw"
        synth.make_fm(synth.default_spec(stype, seed), n).astype(np.float32).tofile(raw)
        files.append(str(raw))
    r = subprocess.run([BATCH, "-q", "-t", ",".join(c[0] for c in cases), "-g", str(tmp_path / "g_"), "-k", str(tmp_path / "k_"),
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
