"""The HOST logic of the batch runner (sdrpp_radiosonde_b200/host/sonde_batch.cpp) on the CPU: recordings read the way the
reference reads them (raw / WAV, ragged lengths, the padded last buffer), the two-deep submit / fetch pipeline, fragment
aggregation (SD/decode.c:278-376) and the CSV / GPX / KML / live-KML writers, against the reference's command-line tool.

The runner source is compiled a second time, for this test only, against tests/cpp/stub_sonde_b200.c — a stand-in for
the ten batch-ABI entry points the runner calls, served by the CPU oracle.  That copy lives in build/stub/ and is test
infrastructure: the product runner links the CUDA library, has no CPU path and exits with code 3 without a GPU
(test_cli_dropin.py::test_batch_runner_fails_loudly_without_gpu).  The same checks run against the product binary on the
GPU box (test_cli_dropin.py, test_zz_batch_tracks.py)."""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import batch_checks  # noqa: E402

ROOT = batch_checks.ROOT
ORACLE = os.path.join(ROOT, "oracle", "_build", "libsonde_oracle.so")
STUBDIR = os.path.join(ROOT, "build", "stub")
RUNNER = os.path.join(STUBDIR, "sonde_b200_batch_stub")

pytestmark = pytest.mark.skipif(not os.path.exists(batch_checks.REF), reason="oracle/_ref/sondedump_ref not built")


@pytest.fixture(scope="module")
def runner():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    os.makedirs(STUBDIR, exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", "stub_sonde_b200.c"),
                    "-o", os.path.join(STUBDIR, "libbatch_abi_stub.so"), f"-L{os.path.dirname(ORACLE)}", "-lsonde_oracle",
                    f"-Wl,-rpath,{os.path.dirname(ORACLE)}"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", os.path.join(ROOT, "sdrpp_radiosonde_b200", "host", "sonde_batch.cpp"),
                    "-o", RUNNER, f"-L{STUBDIR}", "-lbatch_abi_stub", f"-Wl,-rpath,{STUBDIR}"], check=True)
    return RUNNER


def test_csv_per_channel(runner, tmp_path):
    batch_checks.check_csv_per_channel(runner, tmp_path, auto=False)       # the stand-in has no AUTO


def test_wav_inputs(runner, tmp_path):
    batch_checks.check_wav_inputs(runner, tmp_path)


def test_gpx_kml_tracks(runner, tmp_path):
    batch_checks.check_tracks(runner, tmp_path)


def test_text_output_and_format(runner, tmp_path):
    batch_checks.check_text_output(runner, tmp_path)
