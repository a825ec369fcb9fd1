"""The HOST logic on the CPU: the batch runner (sdrpp_radiosonde_b200/host/sonde_batch.cpp — recordings read the way the
reference reads them, the padded last buffer, the two-deep submit / fetch pipeline, fragment aggregation
SD/decode.c:278-376, the text / CSV / GPX / KML / live-KML outputs), the decoder block and the reference-signature compat
layer (host/gpu_decoder.hpp GpuDecoder, csrc/compat.cpp), the channel bank (GpuChannelBank: backlogs of streams that
deliver unequal lengths) and the wideband bank (host/gpu_wideband.hpp GpuWidebandBank: staging and the n mod D carry),
each against the reference's command-line tool, the compiled reference library or the oracle.

The host sources are compiled a second time, for this test only, against tests/cpp/stub_sonde_b200.c — a stand-in for
the batch-ABI entry points they call, served by the CPU oracle's streaming form — and tests/cpp/stub_sonde_chan.c, a
plain C polyphase filter behind the channelizer entry points.  Those copies live in build/stub/
and are test infrastructure: the product binaries link the CUDA library, have no CPU path and exit with code 3 without a
GPU (test_cli_dropin.py / test_host_cpp.py: *_fails_loudly_without_gpu).  The same checks — literally the same functions,
tests/batch_checks.py and the check_* functions of test_host_cpp.py — run against the product binaries on the GPU box."""
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
                    os.path.join(ROOT, "tests", "cpp", "stub_sonde_chan.c"), "-lm", "-o", os.path.join(STUBDIR, "libbatch_abi_stub.so"), f"-L{os.path.dirname(ORACLE)}", "-lsonde_oracle",
                    f"-Wl,-rpath,{os.path.dirname(ORACLE)}"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", os.path.join(ROOT, "sdrpp_radiosonde_b200", "host", "sonde_batch.cpp"),
                    "-o", RUNNER, f"-L{STUBDIR}", "-lbatch_abi_stub", f"-Wl,-rpath,{STUBDIR}"], check=True)
    return RUNNER


def _build_host_program(name, extra_sources=()):
    """a test copy of tests/cpp/<name>.cpp (and of product host sources it needs) linked to the stand-in library"""
    exe = os.path.join(STUBDIR, name + "_stub")
    subprocess.run(["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "cpp", name + ".cpp"), *extra_sources,
                    "-o", exe, f"-L{STUBDIR}", "-lbatch_abi_stub", f"-Wl,-rpath,{STUBDIR}", "-lpthread"], check=True)
    return exe


def test_decoder_block_and_compat_api(runner, tmp_path):
    """radiosonde::GpuDecoder (the dsp::block drop-in: worker thread, merge_fragment, callbacks) and the
    reference-signature compat layer (csrc/compat.cpp, compiled into the test program) on the stand-in library: the
    golden RS41 frame's serial / sequence number / position / time, PARSED counts and field values against the compiled
    reference — the assertions of test_host_cpp.py::test_host_block_and_compat_api_on_gpu."""
    import test_host_cpp
    exe = _build_host_program("host_block_test", [os.path.join(ROOT, "sdrpp_radiosonde_b200", "csrc", "compat.cpp")])
    test_host_cpp.check_host_block_and_compat_api(exe, tmp_path)
    # buffers longer than the compat layer's GPU call size (compat.cpp kMaxChunk = 65536) are decoded in pieces while
    # the caller repeats xxx_decode(): same PARSED count and telemetry as the reference's loop over the same buffers
    test_host_cpp.check_host_block_and_compat_api(exe, tmp_path, chunk=100000)
    # and very short ones (the reference's results depend on where buffers end, SD/demod/gfsk.c:73: the expected values
    # come from the compiled reference run with the same buffers)
    for chunk in (7, 1000):
        test_host_cpp.check_host_block_and_compat_api(exe, tmp_path, chunk=chunk)


def test_channel_bank_unequal_streams_lose_nothing(runner, tmp_path):
    """radiosonde::GpuChannelBank with streams that deliver different lengths per pass, some beyond max_chunk: no
    backlog left, the frames of a straight run all there, the barometric pressure fallback — the assertions of
    test_host_cpp.py::test_channel_bank_unequal_streams_lose_nothing, with the oracle as the straight run."""
    import numpy as np
    import test_host_cpp
    from tests import reflib
    exe = _build_host_program("host_bank_test")
    orc = reflib.OracleLib()

    def straight(types, nb, n):
        recs = [orc.frames_run_iq(t, nb[c], 48000) for c, t in enumerate(types)]
        return np.array([len(r) for r in recs]), np.array([sum(int(x.ok) for x in r) for r in recs])
    test_host_cpp.check_channel_bank(exe, tmp_path, straight)


def test_csv_per_channel(runner, tmp_path):
    batch_checks.check_csv_per_channel(runner, tmp_path, auto=True)


def test_wav_inputs(runner, tmp_path):
    batch_checks.check_wav_inputs(runner, tmp_path)


def test_gpx_kml_tracks(runner, tmp_path):
    batch_checks.check_tracks(runner, tmp_path)


def test_text_output_and_format(runner, tmp_path):
    batch_checks.check_text_output(runner, tmp_path)


def test_iq_input(runner, tmp_path):
    batch_checks.check_iq_input(runner, tmp_path)


def test_wall_clock_sondes(runner, tmp_path):
    batch_checks.check_wall_clock_sondes(runner, tmp_path)


def test_channel_bank_random_buffer_lengths_deliver_every_sample_in_order(runner, tmp_path):
    """GpuChannelBank fuzzed: per pass every stream hands over a buffer of a random length (1 .. 40000 samples, a third
    below 64, max_chunk 5000), so backlogs grow and drain unevenly.  With the deterministic stand-in the frames are a
    function of the sample sequence alone (test_oracle_stream_api_equals_whole_recording_calls: same frames whatever
    the buffering), so the digest over every channel's records must equal the digest of one straight oracle run — any
    sample lost, duplicated or reordered by the bank changes it.  (The first record of a stream is left out of the
    digest: the reference's demodulator forgets its mid-symbol sample at the start of every call, SD/demod/gfsk.c:73,
    so where a buffer ends matters while the timing loop is still acquiring — a first buffer of 55-57 samples flips the
    third bit of this RS41 stream in the compiled reference and the oracle alike — and the bank chooses the call
    lengths.)"""
    import numpy as np
    from sdrpp_radiosonde_b200 import synth
    from tests import reflib
    exe = _build_host_program("host_bank_test")
    orc = reflib.OracleLib()
    types = [synth.RS41, synth.M10, synth.DFM09, synth.C50]
    n = 48000 * 3
    nb = np.stack([synth.make_iq(synth.default_spec(t, 70 + c), n) for c, t in enumerate(types)])

    def fnv(h, b):
        for x in b:
            h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h
    want = []
    for c, t in enumerate(types):
        h = 1469598103934665603
        recs = orc.frames_run_iq(t, nb[c], 4096)
        for r in recs[1:]:
            h = fnv(h, int(r.status).to_bytes(4, "little", signed=True))
            h = fnv(h, bytes(r.data)[:r.data_len])
        want.append((len(recs), f"{h:016x}"))
        assert len(recs) >= 3
    args = [exe, str(len(types)), str(n), "5000"]
    for c, t in enumerate(types):
        (tmp_path / f"iq{c}").write_bytes(nb[c].tobytes())
        args += [str(t), str(tmp_path / f"iq{c}")]
    for seed in range(1, 11):
        r = subprocess.run(args, capture_output=True, text=True, timeout=300, env=dict(os.environ, BANK_TEST_SEED=str(seed)))
        assert r.returncode == 0, r.stdout + r.stderr
        lines = r.stdout.strip().splitlines()
        assert [l for l in lines if l.startswith("BACKLOG")][0].split()[1] == "0"
        ch = [dict(kv.split("=") for kv in l.split()[2:]) for l in lines if l.startswith("CH ")]
        got = [(int(x["frames"]), x["digest"]) for x in ch]
        assert got == want, (seed, got, want)


def test_wideband_bank(runner, tmp_path):
    """radiosonde::GpuWidebandBank (one wideband dsp::stream in, buffers of an odd length so the n mod D carry is
    exercised, three sondes out with telemetry callbacks) on the stand-in: its channelizer part is a plain C polyphase
    filter (tests/cpp/stub_sonde_chan.c) — the assertions of test_host_cpp.py::test_wideband_block_on_gpu with the
    oracle as the straight run on the narrowband signals."""
    import numpy as np
    import test_host_cpp
    from tests import reflib
    exe = _build_host_program("host_wideband_test")
    orc = reflib.OracleLib()

    def straight(types, nb):
        return np.array([sum(int(x.ok) for x in orc.frames_run_iq(t, nb[c], 48000)) for c, t in enumerate(types)])
    test_host_cpp.check_wideband_block(exe, tmp_path, straight)


def test_wideband_input(runner, tmp_path):
    batch_checks.check_wideband_input(runner, tmp_path)


def test_wav_8bit_and_three_channel_inputs(runner, tmp_path):
    """The remaining WAV layouts of SD/io/wavfile.c:53-113.  8-bit samples (value - 127), stereo: CSV against the
    reference tool.  Three interleaved channels: a 32 KiB block does not hold a whole number of 3-channel frames and the
    reference's reader hangs at the first block boundary (its copy count becomes 0 with data left, wavfile.c:66,110), so
    there is nothing to compare with; the runner carries the frame position across blocks and must give exactly what it
    gives for a mono file of the same first channel."""
    import numpy as np
    from sdrpp_radiosonde_b200 import synth
    rng = np.random.default_rng(8)
    n = 48000 * 5
    fm0 = synth.make_fm(synth.default_spec(synth.RS41, 61), n)
    fm1 = synth.make_fm(synth.default_spec(synth.DFM09, 62), n)
    w0, w3, w1 = tmp_path / "a8.wav", tmp_path / "b3.wav", tmp_path / "b1.wav"
    u8 = np.clip(np.round(fm0 * (100.0 / np.abs(fm0).max())) + 127, 0, 255).astype(np.uint8)
    batch_checks._write_wav(w0, np.stack([u8, rng.integers(0, 256, n).astype(np.uint8)], axis=1))
    s16 = np.clip(np.round(fm1 * (9000.0 / np.abs(fm1).max())), -32768, 32767).astype(np.int16)
    batch_checks._write_wav(w3, np.stack([s16, rng.integers(-3000, 3000, n).astype(np.int16), np.zeros(n, np.int16)], axis=1))
    # the mono twin holds what a reader of whole 32 KiB blocks gets from the 3-channel file: its first 6n / 32768 blocks
    batch_checks._write_wav(w1, s16[:(6 * n // 32768) * 32768 // 6 + 1])
    r = subprocess.run([runner, "-q", "-t", "rs41,dfm,dfm", "-c", str(tmp_path / "o_"), str(w0), str(w3), str(w1)], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    a = subprocess.run([batch_checks.REF, "-q", "-t", "rs41", "-c", str(tmp_path / "ref0.csv"), str(w0)], capture_output=True, timeout=600)
    assert a.returncode == 0, a.stderr[-300:]
    got, want = (tmp_path / "o_0.csv").read_bytes(), (tmp_path / "ref0.csv").read_bytes()
    assert want.count(b"\n") >= 5 and got == want
    three, mono = (tmp_path / "o_1.csv").read_bytes().split(b"\n"), (tmp_path / "o_2.csv").read_bytes().split(b"\n")
    assert len(mono) >= 8 and three[:len(mono) - 2] == mono[:len(mono) - 2], (len(three), len(mono))
