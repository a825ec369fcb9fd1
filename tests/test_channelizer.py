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


def test_oracle_rational_resampling_is_upsample_filter_decimate():
    """interp = L: the polyphase form of the oracle equals the textbook chain (mix to DC, zero-stuff by L, FIR with the
    L*K-tap prototype, keep every M-th sample) and a tone at the channel centre lands on DC with unit gain."""
    L, M_, K = 3, 8, 24
    Kg = L * K
    t = np.arange(Kg, dtype=np.float64) - 0.5 * (Kg - 1)
    g = np.sinc(2 * 0.4 / M_ * t) * np.hamming(Kg)            # cut-off 0.4 fs_out at the rate L fs_in
    g *= L / g.sum()
    rng = np.random.default_rng(1)
    step = np.uint32(round(0.1234 * 2 ** 32))
    n = M_ * 200
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = x.astype(np.complex64)
    y = orc.channelize(x, g, [step], M_, interp=L)[0]
    mixed = x.astype(np.complex128) * np.exp(-2j * np.pi * (int(step) / 2 ** 32) * np.arange(n))
    up = np.zeros(n * L, dtype=np.complex128)
    up[::L] = mixed
    full = np.convolve(up, g)                                     # full[j] = sum_i g[j - i] up[i]
    want = full[np.arange(len(y)) * M_ + M_ - 1]
    assert len(y) == n // M_ * L and np.allclose(y, want, atol=1e-9)
    tone = np.exp(2j * np.pi * (int(step) / 2 ** 32) * np.arange(n))
    yt = orc.channelize(tone, g, [step], M_, interp=L)[0]
    assert np.allclose(yt[K * L // M_ + L:], 1.0, atol=2e-3)
    y2 = np.concatenate([orc.channelize(x[:M_ * 50], g, [step], M_, interp=L)[0],
                         orc.channelize(x[M_ * 50:], g, [step], M_, n_start=M_ * 50, history=x[:M_ * 50], interp=L)[0]])
    assert np.allclose(y, y2, atol=1e-12)


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
    # the peer entry point (source buffer on "another" GPU, here device 0 itself: same copy-engine path)
    ch = capi.Channelizer(freqs, D, n_in)
    xd = torch.view_as_real(torch.from_numpy(xf)).cuda()
    ptr, stride, m = ch.process_c64_peer(0, xd.data_ptr(), n_in)
    torch.cuda.synchronize()
    peer = capi.device_view(ptr, (C, stride, 2))[:, :m].cpu().numpy().copy()
    ch.close()
    assert np.array_equal(peer.view(np.uint32), outs[2].view(np.uint32))

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


def _run_chunks(ch, x, chunks, C):
    import torch
    got, pos = [], 0
    for cl in chunks:
        ptr, stride, m = ch.process_c64(x[pos:pos + cl])
        torch.cuda.synchronize()
        out = capi.device_view(ptr, (C, stride, 2))[:, :m, :].cpu().numpy()
        got.append(out[..., 0] + 1j * out[..., 1])
        pos += cl
    return np.concatenate(got, axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("D,C", [(48, 20), (50, 7), (128, 130)])
def test_split_precision_reaches_fp32_grade(D, C):
    """sonde_chan_options.precision = SONDE_CHAN_SPLIT_BF16: hi + lo bf16 operands, three tensor passes.  Tolerance: rms
    error <= 1e-5 of the output rms against the exact formula in double precision for windows of up to 512 taps (the bench
    shape, D = 48, measures 4.6e-6; the plain bf16 mode sits at ~1.5e-3), <= 1.5e-5 for the 1024-tap window of D = 128,
    where the fp32 accumulation over 2048 products inside the tensor core adds its share (measured 1.04e-5)."""
    rng = np.random.default_rng(50 + D)
    fs_in = 48000.0 * D
    freqs = rng.uniform(-0.45, 0.45, C) * fs_in
    chunks = [D * 300, D * 128, D * 7, D * 600]
    x = _wideband(rng, sum(chunks), freqs[:4], fs_in)
    ch = capi.Channelizer(freqs, D, max(chunks), precision=1)
    try:
        got = _run_chunks(ch, x, chunks, C)
        want = orc.channelize(x, ch.taps, ch.steps, D)
        rms = np.sqrt(np.mean(np.abs(want) ** 2))
        err = np.sqrt(np.mean(np.abs(got - want) ** 2)) / rms
        worst = np.abs(got - want).max() / np.abs(want).max()
        print(f"split bf16 D={D} C={C}: rms err {err:.2e} of rms, max err {worst:.2e} of max |y|")
        assert err < (1e-5 if ch.K <= 512 else 1.5e-5)
        assert worst < 3e-5
    finally:
        ch.close()


@pytest.mark.gpu
@pytest.mark.parametrize("L,M_,C,prec", [(3, 128, 9, 0), (3, 128, 9, 1), (12, 625, 5, 1), (5, 52, 6, 0)])
def test_rational_resampling_matches_oracle(L, M_, C, prec):
    """sonde_chan_options.interp: fs_out = fs_in L / M (2.048 MS/s -> 3/128, 2.5 MS/s -> 12/625, and an even, non-multiple-
    of-4 decimation that uses the zero-slot layout).  Outputs are interleaved branch by branch; chunk boundaries
    continue the stream."""
    rng = np.random.default_rng(7 * L + M_)
    fs_in = 48000.0 * M_ / L
    freqs = rng.uniform(-0.45, 0.45, C) * fs_in
    chunks = [M_ * 140, M_ * 3, M_ * 257]
    x = _wideband(rng, sum(chunks), freqs[:3], fs_in)
    ch = capi.Channelizer(freqs, M_, max(chunks), precision=prec, interp=L)
    try:
        got = _run_chunks(ch, x, chunks, C)
        assert got.shape[1] == sum(chunks) // M_ * L
        want = orc.channelize(x, ch.taps, ch.steps, M_, interp=L)
        rms = np.sqrt(np.mean(np.abs(want) ** 2))
        err = np.sqrt(np.mean(np.abs(got - want) ** 2)) / rms
        print(f"L/M = {L}/{M_} precision {prec}: rms err {err:.2e} of rms")
        # split precision: 1e-5 while the window (taps per branch x slot stride of the zero-slot layout) stays within 512
        # slots, 2e-5 beyond (fp32 accumulation inside the tensor core over thousands of products; measured 1.6e-5 at
        # 12/625, whose window is 1696 slots)
        slots = ch.K // L * (1 if M_ % 4 == 0 else 2 if M_ % 2 == 0 else 4)
        assert err < ((1e-5 if slots <= 512 else 2e-5) if prec else 1.5e-2)
        if not prec:
            want_f = orc.channelize(x, ch.taps, ch.steps, M_, interp=L, bf16=True)
            assert np.abs(got - want_f).max() / np.abs(want_f).max() < 2e-4
    finally:
        ch.close()


@pytest.mark.gpu
def test_per_channel_cutoffs():
    """sonde_chan_options.cutoff_hz: every channel gets the prototype of its own bandwidth (half the per-type VFO
    bandwidths of src/main.hpp:45-51); the result equals the oracle run with each channel's own taps, and a narrow
    channel rejects a tone 8 kHz off its centre that a wide one passes."""
    D = 48
    fs_in = 48000.0 * D
    freqs = np.array([-300e3, -300e3, 150e3, 410e3])
    cut = np.array([5e3, 21.6e3, 7.5e3, 10e3], dtype=np.float32)
    n = D * 2000
    t = np.arange(n)
    x = (0.5 * np.exp(2j * np.pi * (-300e3 + 8e3) / fs_in * t)).astype(np.complex64)
    ch = capi.Channelizer(freqs, D, n, taps_per_phase=32, precision=1, cutoffs=cut)
    try:
        got = _run_chunks(ch, x, [n], 4)
        tp = [ch.taps_of(c) for c in range(4)]
        assert not np.allclose(tp[0], tp[1]) and np.allclose(tp[0], ch.taps)
        want = orc.channelize(x, ch.taps, ch.steps, D, taps_per_channel=tp)
        assert np.sqrt(np.mean(np.abs(got - want) ** 2)) < 1e-5 * 0.5
        p_narrow = np.mean(np.abs(got[0, 100:]) ** 2)
        p_wide = np.mean(np.abs(got[1, 100:]) ** 2)
        print(f"tone 8 kHz off centre: 5 kHz channel {10 * np.log10(p_narrow / 0.25):.1f} dB, 21.6 kHz channel {10 * np.log10(p_wide / 0.25):.1f} dB")
        assert p_wide > 0.9 * 0.25 and p_narrow < 1e-3 * 0.25
    finally:
        ch.close()
