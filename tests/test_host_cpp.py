"""The C++ host layer: radiosonde::GpuDecoder (dsp::block drop-in, sdrpp_radiosonde_b200/host/gpu_decoder.hpp)
and the reference-signature compat API (include/sonde_b200_compat.h)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import synth
from tests import reflib

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
EXE = os.path.join(ROOT, "build", "host_block_test")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", f"{ROOT}/tests/cpp/host_block_test.cpp", "-o", EXE,
           f"-L{ROOT}/sdrpp_radiosonde_b200", "-lsonde_b200_compat", "-lsonde_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "sdrpp_radiosonde_b200"), "-lpthread"]
    subprocess.run(cmd, check=True, cwd=ROOT)


WEXE = os.path.join(ROOT, "build", "host_wideband_test")


def build_wideband_exe():
    os.makedirs(os.path.dirname(WEXE), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", f"{ROOT}/tests/cpp/host_wideband_test.cpp", "-o", WEXE,
           f"-L{ROOT}/sdrpp_radiosonde_b200", "-lsonde_b200", "-Wl,-rpath," + os.path.join(ROOT, "sdrpp_radiosonde_b200"),
           "-lpthread"]
    subprocess.run(cmd, check=True, cwd=ROOT)


def test_wideband_block_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    build_wideband_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    (tmp_path / "w").write_bytes(np.zeros(4800, np.complex64).tobytes())
    r = subprocess.run([WEXE, str(tmp_path / "w"), "4800", "48", "1000", "0", "0"], capture_output=True, text=True)
    assert r.returncode == 3 and "NOGPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_wideband_block_on_gpu(tmp_path):
    """radiosonde::GpuWidebandBank: one wideband dsp::stream in (buffers of an odd length, so the n mod D carry is
    exercised), three sondes out with telemetry callbacks."""
    from sdrpp_radiosonde_b200 import capi
    build_wideband_exe()

    def straight(types, nb):
        dec = capi.BatchDecoder(types, 48000)
        want_ok = np.zeros(len(types), dtype=int)
        for pos in range(0, nb.shape[1], 48000):
            dec.process_iq(np.ascontiguousarray(nb[:, pos:pos + 48000]))
            recs, counts = dec.fetch()
            want_ok += [sum(int(r_["ok"]) for r_ in recs[c, :counts[c]]) for c in range(len(types))]
        dec.close()
        return want_ok
    check_wideband_block(WEXE, tmp_path, straight)


def check_wideband_block(exe, tmp_path, straight):
    """shared with tests/test_batch_host_logic.py (CPU, stand-in library); `straight(types, nb[C][n])` returns the
    per-channel ok counts of a plain run over the narrowband signals"""
    from tests.gpu_util import make_wideband
    D, nsec = 48, 4
    types = [synth.RS41, synth.M10, synth.RS41]
    freqs = [-400e3, 250e3, 31.25e3]
    nb, wide = make_wideband(types, freqs, D, nsec)
    (tmp_path / "w").write_bytes(wide.tobytes())
    args = [exe, str(tmp_path / "w"), str(wide.size), str(D), "100003"]
    for f, t in zip(freqs, types):
        args += [str(f), str(t)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    ch = [dict(kv.split("=") for kv in l.split()[2:]) for l in lines if l.startswith("CH ")]
    want_ok = straight(types, nb)
    for c in range(len(types)):
        assert int(ch[c]["ok"]) >= want_ok[c] - 1 and want_ok[c] >= 3, (c, ch[c], want_ok[c])
    assert any(l.startswith("SERIAL R3551568:") for l in lines)          # the golden RS41 frame's serial
    assert int([l for l in lines if l.startswith("CALLBACKS")][0].split()[1]) >= 6


def test_compat_library_exports_reference_signatures():
    lib = ctypes.CDLL(os.path.join(ROOT, "sdrpp_radiosonde_b200", "libsonde_b200_compat.so"))
    for x in ("rs41", "dfm09", "m10", "ims100", "mrzn1", "imet4", "c50"):
        for fn in ("decoder_init", "decoder_deinit", "decode", "last_frame"):
            assert hasattr(lib, f"{x}_{fn}"), f"{x}_{fn}"


def test_host_block_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    build_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    (tmp_path / "iq").write_bytes(np.zeros(2048, np.complex64).tobytes())
    (tmp_path / "fm").write_bytes(np.zeros(2048, np.float32).tobytes())
    r = subprocess.run([EXE, str(tmp_path / "iq"), str(tmp_path / "fm"), "2048", "1024"], capture_output=True, text=True)
    assert r.returncode == 3 and "NOGPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_block_and_compat_api_on_gpu(tmp_path):
    build_exe()
    check_host_block_and_compat_api(EXE, tmp_path)


def check_host_block_and_compat_api(exe, tmp_path, chunk=4096):
    """shared with tests/test_batch_host_logic.py, which runs it on the CPU against a test copy of the same program
    linked to the oracle-backed stand-in of the batch ABI"""
    n = 48000 * 3
    spec = synth.default_spec(synth.RS41, 0)
    iq, fm = synth.make_iq(spec, n), synth.make_fm(spec, n)
    (tmp_path / "iq").write_bytes(iq.tobytes())
    (tmp_path / "fm").write_bytes(fm.tobytes())
    r = subprocess.run([exe, str(tmp_path / "iq"), str(tmp_path / "fm"), str(n), str(chunk)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    blk = dict(kv.split("=") for kv in out["BLOCK"].split())
    cmp_ = dict(kv.split("=") for kv in out["COMPAT"].split())
    # the golden RS41 frame: serial R3551568, sequence number 11509
    assert blk["serial"] == "R3551568" and blk["seq"] == "11509" and int(blk["callbacks"]) >= 2
    assert int(blk["frames"]) >= int(blk["ok"]) >= 2
    assert cmp_["serial"] == "R3551568" and cmp_["seq"] == "11509"
    # same number of PARSED returns as the reference's own xxx_decode loop on the same float input
    chk = reflib.RefLib() if reflib.have_ref() else reflib.OracleLib()
    want = chk.frames_run(synth.RS41, fm, chunk)
    assert int(cmp_["parsed"]) == len(want)
    tel = dict(kv.split("=") for kv in out["TELEM"].split())
    # 2021-03-17 04:16:03 UTC, 37.81514 S 145.01633 E, 7738 m
    assert abs(float(tel["lat"]) + 37.81514) < 1e-4 and abs(float(tel["lon"]) - 145.01633) < 1e-4
    assert abs(float(tel["alt"]) - 7738) < 2 and int(tel["time"]) == 1615954563
    if reflib.have_ref():
        sd, _ = reflib.RefLib().decode_run(synth.RS41, fm, chunk)
        assert int(cmp_["with_fields"]) == sum(1 for s in sd if s.fields)
        ref = [s for s in sd if (s.fields & 0x14) == 0x14][-1]      # DATA_POS | DATA_TIME
        for k in ("lat", "lon", "alt", "speed", "heading", "climb"):
            assert abs(float(tel[k]) - getattr(ref, k)) <= 1e-5 * max(1.0, abs(getattr(ref, k))), k
        assert int(tel["time"]) == ref.time


BEXE = os.path.join(ROOT, "build", "host_bank_test")


def build_bank_exe():
    os.makedirs(os.path.dirname(BEXE), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", f"{ROOT}/tests/cpp/host_bank_test.cpp", "-o", BEXE,
           f"-L{ROOT}/sdrpp_radiosonde_b200", "-lsonde_b200", "-Wl,-rpath," + os.path.join(ROOT, "sdrpp_radiosonde_b200"),
           "-lpthread"]
    subprocess.run(cmd, check=True, cwd=ROOT)


def test_channel_bank_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    build_bank_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    (tmp_path / "iq").write_bytes(np.zeros(2048, np.complex64).tobytes())
    r = subprocess.run([BEXE, "1", "2048", "1024", "0", str(tmp_path / "iq")], capture_output=True, text=True)
    assert r.returncode == 3 and "NOGPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_channel_bank_unequal_streams_lose_nothing(tmp_path):
    """GpuChannelBank fed by three streams whose buffers differ in length per pass and exceed max_chunk: every sample
    is decoded (no backlog left, the frames of a straight run are all there) and SondeFullData.pressure carries the
    plugin's barometric fallback (src/decode/decoder.hpp:108-110) for the sonde without a pressure sensor."""
    from sdrpp_radiosonde_b200 import capi
    build_bank_exe()

    def straight(types, nb, n):
        dec = capi.BatchDecoder(types, 48000)
        want_frames = np.zeros(len(types), dtype=int)
        want_ok = np.zeros(len(types), dtype=int)
        for pos in range(0, n, 48000):
            dec.process_iq(np.ascontiguousarray(nb[:, pos:pos + 48000]))
            recs, counts = dec.fetch()
            want_frames += counts
            want_ok += [sum(int(r_["ok"]) for r_ in recs[c, :counts[c]]) for c in range(len(types))]
        dec.close()
        return want_frames, want_ok
    check_channel_bank(BEXE, tmp_path, straight)


def check_channel_bank(exe, tmp_path, straight):
    """shared with tests/test_batch_host_logic.py (CPU, stand-in library); `straight(types, iq[C][n], n)` returns the
    per-channel frame and ok counts of a plain run in 48000-sample calls"""
    types = [synth.RS41, synth.M10, synth.DFM09]
    n = 48000 * 4
    nb = np.stack([synth.make_iq(synth.default_spec(t, 60 + c), n) for c, t in enumerate(types)])
    args = [exe, str(len(types)), str(n), "5000"]
    for c, t in enumerate(types):
        (tmp_path / f"iq{c}").write_bytes(nb[c].tobytes())
        args += [str(t), str(tmp_path / f"iq{c}")]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    ch = [dict(kv.split("=") for kv in l.split()[2:]) for l in lines if l.startswith("CH ")]
    assert [l for l in lines if l.startswith("BACKLOG")][0].split()[1] == "0"
    want_frames, want_ok = straight(types, nb, n)
    for c in range(len(types)):
        # the framer emits one window per frame length of bits whatever the buffering; the FEC gate is robust to it
        assert abs(int(ch[c]["frames"]) - want_frames[c]) <= 1, (c, ch[c], want_frames[c])
        assert int(ch[c]["ok"]) >= want_ok[c] - 1 and want_ok[c] >= 3, (c, ch[c], want_ok[c])
        assert int(ch[c]["callbacks"]) >= 1 and ch[c]["pressure_ok"] == "1", (c, ch[c])


MEXE = os.path.join(ROOT, "build", "host_multibank_test")


def build_multibank_exe():
    os.makedirs(os.path.dirname(MEXE), exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-O2", "-std=c++17", f"-I{ROOT}/include", f"-I{cuda}/include", f"{ROOT}/tests/cpp/host_multibank_test.cpp",
           "-o", MEXE, f"-L{ROOT}/sdrpp_radiosonde_b200", "-lsonde_b200", f"-L{cuda}/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "sdrpp_radiosonde_b200"), "-lpthread"]
    subprocess.run(cmd, check=True, cwd=ROOT)


def test_multibank_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    build_multibank_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    (tmp_path / "iq").write_bytes(np.zeros(2048, np.complex64).tobytes())
    r = subprocess.run([MEXE, "1", "2048", "1024", "2", "0", str(tmp_path / "iq")], capture_output=True, text=True)
    assert r.returncode == 3 and "NOGPU" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_multibank_shards_equal_single_handle(tmp_path):
    """radiosonde::GpuMultiBank (C++ host, one process): 9 channels of four types in 4 shards over the devices of the
    box (all on device 0 when there is one GPU), fed from a host buffer and from a buffer on device 0 that the shards
    pull with the copy engine (sonde_b200_process_iq_peer), two buffers in flight: the gathered records are
    byte-identical to one handle decoding everything."""
    build_multibank_exe()
    types = [synth.RS41, synth.M10, synth.DFM09, synth.RS41, synth.C50, synth.IMS100, synth.RS41, synth.M10, synth.MRZN1]
    n, chunk = 48000 * 3, 12000
    args = [MEXE, str(len(types)), str(n), str(chunk), "4"]
    for c, t in enumerate(types):
        (tmp_path / f"iq{c}").write_bytes(synth.make_iq(synth.default_spec(t, 80 + c), n).tobytes())
        args += [str(t), str(tmp_path / f"iq{c}")]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = dict(kv.split("=") for kv in r.stdout.strip().splitlines()[-1].split()[1:])
    assert out["host_equal"] == "1" and out["peer_equal"] == "1", r.stdout
    assert int(out["frames"]) >= int(out["ok"]) >= 20, r.stdout
