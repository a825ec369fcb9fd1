"""sonde_b200_batch on the GPU, the outputs and inputs beyond the CSV of test_cli_dropin.py: per-channel
GPX / KML / live KML (-g -k -l), text lines and the format language (-o -f), complex64 recordings (-i), the two sondes
whose parsers read the wall clock, and one wideband recording through the channelizer (-w -F) — each against what the
reference's command-line tool writes for every recording alone.  The check functions live in tests/batch_checks.py and
also run on the CPU against a test copy of the runner linked to the oracle-backed stand-in of the ABI
(tests/test_batch_host_logic.py); the writers themselves are compared byte for byte with the compiled reference writers in
tests/test_track_files.py.  The file sorts last on purpose: these are the newest checks of the suite."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import batch_checks  # noqa: E402
from batch_checks import gpx_parts, kml_parts, same_but_first  # noqa: E402

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "sondedump_ref")
BATCH = os.path.join(ROOT, "sdrpp_radiosonde_b200", "sonde_b200_batch")


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
    batch_checks.check_tracks(BATCH, tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_text_output_equals_reference_cli_stdout(tmp_path):
    batch_checks.check_text_output(BATCH, tmp_path)


@pytest.mark.gpu
def test_batch_runner_iq_input_equals_fm_input(tmp_path):
    from tests import reflib
    if not reflib.have_oracle():
        pytest.skip("oracle/_build/libsonde_oracle.so not built")
    batch_checks.check_iq_input(BATCH, tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_wall_clock_sondes(tmp_path):
    batch_checks.check_wall_clock_sondes(BATCH, tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_wideband_input(tmp_path):
    from tests import reflib
    if not reflib.have_oracle():
        pytest.skip("oracle/_build/libsonde_oracle.so not built")
    batch_checks.check_wideband_input(BATCH, tmp_path)
