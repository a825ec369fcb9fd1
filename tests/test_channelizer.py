"""Wideband channelizer: C ABI exports, the numpy oracle's own sanity, and — on the GPU —
the tcgen05 GEMM kernel against the oracle, stream continuity across calls, and the chain
wideband IQ -> channelizer -> decoder against decoding the narrowband signals directly.

Tolerances (floating point; "parity unpinned", see oracle/channelizer_oracle.py):
  * vs the operand-faithful oracle (bf16 samples and weights, double accumulation): 2e-4 of the largest output
    magnitude — what is left is fp32 accumulation order and the fast sincos of the final rotation;
  * vs the ideal double-precision formula: 1.5e-2 of the output RMS — the bf16 rounding of samples and weights
    (2^-9 relative per operand, i.e. a noise floor about 45 dB below the signal)."""
import ctypes
import os
import sys

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import capi, synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import channelizer_oracle as orc  # noqa: E402


def test_channelizer_symbols_exported():
    """every symbol include/sonde_b200_channelizer.h declares is exported, and nothing else is claimed"""
    import re
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "sonde_b200_channelizer.h")).read()
    declared = set(re.findall(r"SONDE_API\s+[\w\s\*]+?\b(sonde_chan_\w+)\s*\(", hdr))
    assert declared == set(capi.CHAN_EXPORTS), declared ^ set(capi.CHAN_EXPORTS)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_channelizer_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SondeError) as e:
        capi.Channelizer([0.0], 8, 8 * 100)
    assert e.value.code == capi.ERR_NODEVICE


def test_oracle_tone_goes_to_dc_and_bf16_rounding():
    D, K = 8, 64
    fs_in = 48000.0 * D
    t = np.arange(K, dtype=np.float64) - 0.5 * (K - 1)
    taps = np.sinc(2 * 0.42 / D * t) * np.hamming(K)
    taps /= taps.sum()
    f = 37000.0
    step = np.uint32(round(f / fs_in * 2 ** 32))
    n = np.arange(D * 400)
    x = np.exp(2j * np.pi * (int(step) / 2 ** 32) * n)
    y = orc.channelize(x, taps, [step], D)[0]
    assert np.allclose(y[K // D + 1:], 1.0, atol=1e-6)               # a tone at the channel centre lands on DC, gain 1
    y2 = np.concatenate([orc.channelize(x[:D * 100], taps, [step], D)[0],
                         orc.channelize(x[D * 100:], taps, [step], D, n_start=D * 100, history=x[:D * 100])[0]])
    assert np.allclose(y, y2, atol=1e-12)                            # chunked == one shot
    a = np.array([1.0, 1.00390625, 1.005859375, -3.14159], dtype=np.float32)
    assert orc.to_bf16(a).tolist() == [1.0, 1.0, 1.0078125, -3.140625]


def _wideband(rng, n, freqs, fs_in):
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.05
    t = np.arange(n)
    for i, f in enumerate(freqs):
        # a slowly frequency-modulated carrier near each channel centre
        x += (0.3 + 0.1 * i) * np.exp(2j * np.pi * ((f + 900.0) / fs_in * t + 0.3 * np.sin(2 * np.pi * 2400.0 / fs_in * t)))
    return x.astype(np.complex64)


@pytest.mark.gpu
@pytest.mark.parametrize("D,C", [(8, 5), (48, 20), (128, 130), (50, 7), (25, 3)])
def test_channelizer_matches_oracle(D, C):
    import torch
    rng = np.random.default_rng(5 + D)
    fs_in = 48000.0 * D
    freqs = rng.uniform(-0.45, 0.45, C) * fs_in
    chunks = [D * 300, D * 128, D * 7, D * 1000, D * 129]
    n = sum(chunks)
    x = _wideband(rng, n, freqs[:4], fs_in)
    ch = capi.Channelizer(freqs, D, max(chunks))
    try:
        got = []
        pos = 0
        for cl in chunks:
            ptr, stride, m = ch.process_c64(x[pos:pos + cl])
            torch.cuda.synchronize()
            out = capi.device_view(ptr, (C, stride, 2))[:, :m, :].cpu().numpy()
            got.append(out[..., 0] + 1j * out[..., 1])
            pos += cl
        got = np.concatenate(got, axis=1)
        want_f = orc.channelize(x, ch.taps, ch.steps, D, bf16=True)
        want_i = orc.channelize(x, ch.taps, ch.steps, D, bf16=False)
        scale = np.abs(want_f).max()
        err_f = np.abs(got - want_f).max() / scale
        rms = np.sqrt(np.mean(np.abs(want_i) ** 2))
        err_i = np.sqrt(np.mean(np.abs(got - want_i) ** 2)) / rms
        print(f"D={D} C={C}: max err vs bf16-faithful oracle {err_f:.2e} of max |y|; rms err vs ideal {err_i:.2e} of rms")
        assert err_f < 2e-4
        assert err_i < 1.5e-2
    finally:
        ch.close()


@pytest.mark.gpu
def test_wideband_to_frames():
    """Three RS41 transmitters and one M10 at different offsets of a 48 x 48 kS/s wideband stream: the chain
    channelizer -> decoder (both on the decoder's stream, no host sync in between) recovers the frames that decoding
    each narrowband signal directly gives."""
    from tests.gpu_util import make_wideband
    D, nsec = 48, 4
    types = [synth.RS41, synth.RS41, synth.M10, synth.RS41]
    freqs = np.array([-500e3, 123.4e3, 420e3, -37.5e3])
    n = 48000 * nsec
    nb, wide = make_wideband(types, freqs, D, nsec)

    def decode_direct():
        dec = capi.BatchDecoder(types, 48000)
        out = [[] for _ in types]
        for pos in range(0, n, 48000):
            dec.process_iq(np.ascontiguousarray(nb[:, pos:pos + 48000]))
            recs, counts = dec.fetch()
            for c in range(len(types)):
                out[c] += [bytes(r["data"][:int(r["data_len"])]) for r in recs[c, :counts[c]] if r["ok"]]
        dec.close()
        return out

    want = decode_direct()
    dec = capi.BatchDecoder(types, 48000)
    ch = capi.Channelizer(freqs, D, 48000 * D)
    got = [[] for _ in types]
    for pos in range(0, n * D, 48000 * D):
        ptr, stride, m = ch.process_c64(wide[pos:pos + 48000 * D], stream=dec.stream)
        dec.process_iq_device(ptr, m, stride)
        recs, counts = dec.fetch()
        for c in range(len(types)):
            got[c] += [bytes(r["data"][:int(r["data_len"])]) for r in recs[c, :counts[c]] if r["ok"]]
    ch.close()
    dec.close()
    for c in range(len(types)):
        print(f"channel {c}: {len(want[c])} frames direct, {len(got[c])} through the channelizer")
        assert len(want[c]) >= 3
        common = set(want[c]) & set(got[c])
        assert len(common) >= len(set(want[c])) - 1, (c, len(common), len(want[c]))


@pytest.mark.gpu
def test_channelizer_entry_points_agree_and_reject_bad_arguments():
    """int16, host-float and device-float entry points give bit-identical outputs on the same samples; argument
    errors are reported, not absorbed."""
    import torch
    D, C = 16, 9
    rng = np.random.default_rng(3)
    freqs = rng.uniform(-0.4, 0.4, C) * 48000.0 * D
    n_in = D * 1000
    q = rng.integers(-20000, 20000, size=(n_in, 2)).astype(np.int16)
    xf = (q[:, 0].astype(np.float32) / 32768 + 1j * (q[:, 1].astype(np.float32) / 32768)).astype(np.complex64)
    outs = []
    for kind in ("s16", "host", "device"):
        ch = capi.Channelizer(freqs, D, n_in)
        if kind == "s16":
            ptr, stride, m = ch.process_s16(q)
        elif kind == "host":
            ptr, stride, m = ch.process_c64(xf)
        else:
            xd = torch.view_as_real(torch.from_numpy(xf)).cuda()
            ptr, stride, m = ch.process_c64_device(xd.data_ptr(), n_in)
        torch.cuda.synchronize()
        outs.append(capi.device_view(ptr, (C, stride, 2))[:, :m].cpu().numpy().copy())
        ch.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(outs[1].view(np.uint32), outs[2].view(np.uint32))
    assert np.abs(outs[0]).max() > 0

    # 8-bit offset-binary IQ == the float entry point on (u8 - 127.5) / 128
    u8 = rng.integers(0, 256, size=(n_in, 2)).astype(np.uint8)
    xf8 = ((u8[:, 0].astype(np.float32) - 127.5) / 128 + 1j * ((u8[:, 1].astype(np.float32) - 127.5) / 128)).astype(np.complex64)
    res = []
    for kind in ("u8", "host"):
        ch = capi.Channelizer(freqs, D, n_in)
        ptr, stride, m = ch.process_u8(u8) if kind == "u8" else ch.process_c64(xf8)
        torch.cuda.synchronize()
        res.append(capi.device_view(ptr, (C, stride, 2))[:, :m].cpu().numpy().copy())
        ch.close()
    assert np.array_equal(res[0].view(np.uint32), res[1].view(np.uint32))

    # a handle whose largest call is shorter than one 128-row tile
    ch = capi.Channelizer(freqs, D, D * 10)
    ptr, stride, m = ch.process_c64(xf[:D * 10])
    torch.cuda.synchronize()
    tiny = capi.device_view(ptr, (C, stride, 2))[:, :m].cpu().numpy()
    ch.close()
    assert m == 10 and np.array_equal(tiny.view(np.uint32), outs[1][:, :10].view(np.uint32))

    with pytest.raises(capi.SondeError):
        capi.Channelizer(freqs, 1, 100)                        # no decimation
    with pytest.raises(capi.SondeError):
        capi.Channelizer([0.6 * 48000.0 * D], D, n_in)         # centre outside the wideband
    ch = capi.Channelizer(freqs, D, n_in)
    with pytest.raises(capi.SondeError):
        ch.process_c64(xf[:D * 10 + 3])                        # not a whole number of output samples
    with pytest.raises(capi.SondeError):
        ch.process_c64(np.concatenate([xf, xf]))               # longer than max_in_len
    ch.close()
