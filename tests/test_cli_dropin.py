"""SURVEY.md §8 f-4 as a drop-in check: the reference's own command-line tool (SD/main.c, decode.c, io/*.c — compiled
unmodified from the sources where they lie by `make -C oracle cli`) linked once with the reference's decoders
(`sondedump_ref`) and once with libsonde_b200_compat.so (`sondedump_b200`).  On the same input file the two tools must
print the same lines and write the same CSV / GPX / KML files, byte for byte: same PARSED sequence, same SondeData,
same autodetect decision."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import synth

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import batch_checks  # noqa: E402

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.path.join(ROOT, "oracle", "_ref", "sondedump_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "sondedump_b200")

needs_cli = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)),
                               reason="oracle/_ref/sondedump_* not built (needs /root/reference at build time)")


def run_tool(exe, flag, raw, outdir, tag):
    paths = {k: os.path.join(outdir, f"{tag}.{k}") for k in ("csv", "gpx", "kml")}
    cmd = [exe, "-t", flag, "-c", paths["csv"], "-g", paths["gpx"], "-k", paths["kml"], raw]
    r = subprocess.run(cmd, capture_output=True, timeout=600)
    files = {k: open(p, "rb").read() if os.path.exists(p) else None for k, p in paths.items()}
    return r, files


def normalise_kml(kml: bytes) -> bytes:
    """The tool's closing <Placemark><Point> block prints KMLFile.lat/lon/alt — and is only written at all if they are
    >= 0 — but kml_init() never initialises them (SD/io/kml.c:12-27,138-160): when no track point was ever written they
    are whatever the stack held, which differs between two binaries and between runs.  In that one case the block is
    dropped from both files; everything else is compared verbatim."""
    if re.search(rb"^-?\d+\.\d+,-?\d+\.\d+,-?\d+\.\d+$", kml, flags=re.M):
        return kml
    return re.sub(rb"<Placemark>\s*<name>[^<]*</name>\s*<Point>.*?</Point>\s*</Placemark>\s*", b"", kml, flags=re.S)


@needs_cli
def test_reference_cli_relinked_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    raw = tmp_path / "x.raw"
    np.zeros(4096, np.float32).tofile(raw)
    r, _ = run_tool(B200, "rs41", str(raw), str(tmp_path), "b200")
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("flag,stype,nsec", [("rs41", synth.RS41, 6), ("dfm", synth.DFM09, 5), ("m10", synth.M10, 5),
                                             ("mrzn1", synth.MRZN1, 5), ("c50", synth.C50, 4), ("auto", synth.RS41, 6),
                                             ("auto", synth.M10, 4)])
def test_reference_cli_on_gpu_library_matches_reference_cli(tmp_path, flag, stype, nsec):
    fm = synth.make_fm(synth.default_spec(stype, 3), 48000 * nsec)
    raw = tmp_path / "in.raw"
    fm.astype(np.float32).tofile(raw)
    a, fa = run_tool(REF, flag, str(raw), str(tmp_path), "ref")
    b, fb = run_tool(B200, flag, str(raw), str(tmp_path), "b200")
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-300:], b.stderr[-300:])
    assert a.stdout.count(b"\n") >= 3, a.stdout[:300]           # the reference itself decoded something
    assert b.stdout == a.stdout
    for k in ("csv", "gpx"):
        assert fa[k] is not None and fb[k] == fa[k], k
    assert fa["kml"] is not None and normalise_kml(fb["kml"]) == normalise_kml(fa["kml"])


def mask_clock(b: bytes) -> bytes:
    """iMS-100 and iMet-4 take the DATE of their time stamps from time(NULL) (SD/sonde/ims100/parser.c:26,
    imet4/parser.c:47,70 — SURVEY.md H5): two runs that straddle midnight UTC would differ in it, so dates are masked."""
    return re.sub(rb"\d{4}-\d{2}-\d{2}", b"DATE", b)


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("flag,stype,nsec", [("ims100", synth.IMS100, 6), ("imet4", synth.IMET4, 6)])
def test_reference_cli_wall_clock_sondes(tmp_path, flag, stype, nsec):
    """The two decoders whose parsers read the wall clock: same drop-in comparison with the dates masked."""
    fm = synth.make_fm(synth.default_spec(stype, 3), 48000 * nsec)
    raw = tmp_path / "in.raw"
    fm.astype(np.float32).tofile(raw)
    a, fa = run_tool(REF, flag, str(raw), str(tmp_path), "ref")
    b, fb = run_tool(B200, flag, str(raw), str(tmp_path), "b200")
    assert a.returncode == 0 and b.returncode == 0, (a.stderr[-300:], b.stderr[-300:])
    assert a.stdout.count(b"\n") >= 2, a.stdout[:300]
    assert mask_clock(b.stdout) == mask_clock(a.stdout)
    assert fa["csv"] is not None and mask_clock(fb["csv"]) == mask_clock(fa["csv"])


BATCH = os.path.join(ROOT, "sdrpp_radiosonde_b200", "sonde_b200_batch")


def test_batch_runner_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    raw = tmp_path / "x.raw"
    np.zeros(4096, np.float32).tofile(raw)
    r = subprocess.run([BATCH, "-t", "rs41", str(raw)], capture_output=True, timeout=60)
    assert r.returncode == 3 and b"no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_csv_equals_reference_cli_per_channel(tmp_path):
    """five recordings of five sonde types in ONE batch, every channel's CSV equal to the reference CLI's; then the AUTO
    path of the runner (tests/batch_checks.py)"""
    batch_checks.check_csv_per_channel(BATCH, tmp_path, auto=True)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/sondedump_ref not built")
def test_batch_runner_wav_inputs_equal_reference_cli(tmp_path):
    """16-bit mono WAV, 32-bit float stereo WAV and a raw float32 file in one batch (tests/batch_checks.py)"""
    batch_checks.check_wav_inputs(BATCH, tmp_path)
