"""Oracle of the wideband channelizer — TEST INFRASTRUCTURE ONLY.

Parity unpinned: the stage replaces the SDR++ VFO + RationalResampler in front of the plugin
(src/main.cpp:55-60 of the reference), whose sources are not part of the reference repository and are
unpinned upstream (every build clones SDR++ HEAD, docker/debian_bullseye/do_build.sh:12).  The oracle is
therefore the defining formula of include/sonde_b200_channelizer.h evaluated in double precision:

    y_c[m] = sum_{k<K} h[k] x[mD + D-1 - k] exp(-j w_c (mD + D-1 - k)),    w_c = 2 pi step_c / 2^32

with x = 0 before the start of the stream.  `h` and `step` are parameters read back from the library
(sonde_chan_taps / sonde_chan_steps), not results.  Two variants:

  channelize(x, ...)                ideal: exact inputs and weights
  channelize(x, ..., bf16=True)     operand-faithful: samples and the weights h[k] e^{j w k} rounded to bfloat16
                                    (round to nearest even) as the tensor-core kernel stores them, products and
                                    sums still in double — differs from the kernel only by fp32 accumulation
                                    order and the fast sincos of the final rotation
"""
import numpy as np


def to_bf16(a):
    """float32 -> bfloat16 (round to nearest even) -> float32, elementwise."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def channelize(x, taps, steps, decim, n_start=0, history=None, bf16=False, interp=1, taps_per_channel=None):
    """x: complex wideband chunk (len multiple of decim), history: the samples before it (zeros if None).

    General form (include/sonde_b200_channelizer.h, sonde_chan_options.interp = L, decim = M):

        y_c[m] = sum_n g[m M + M-1 - n L] x[n] exp(-j w_c n)

    with g the L*K-tap prototype at the rate L fs_in (`taps`; `taps_per_channel[c]` overrides it per channel).  For
    output m = L q + r that is branch r of the polyphase filter: h_r[k] = g[phi_r + k L] over the window ending at the
    input sample n0 = q M + e_r,  e_r = (r M + M - 1) div L,  phi_r = (r M + M - 1) mod L.  L = 1: h = g, n0 = mM + M-1.
    Returns y[C][len(x) // decim * interp] complex128."""
    x = np.asarray(x)
    M_, L = int(decim), int(interp)
    Kg = len(taps)
    K = Kg // L
    assert K * L == Kg
    hist = np.zeros(K, dtype=np.complex128) if history is None else np.asarray(history, dtype=np.complex128)[-K:]
    if len(hist) < K:
        hist = np.concatenate([np.zeros(K - len(hist), dtype=np.complex128), hist])
    xr, xi = np.real(x).astype(np.float32), np.imag(x).astype(np.float32)
    if bf16:
        xr, xi = to_bf16(xr), to_bf16(xi)
        hist = to_bf16(np.real(hist).astype(np.float32)) + 1j * to_bf16(np.imag(hist).astype(np.float32))
    full = np.concatenate([hist, xr.astype(np.float64) + 1j * xi.astype(np.float64)])
    Q = len(x) // M_
    k = np.arange(K, dtype=np.uint64)
    out = np.empty((len(steps), Q * L), dtype=np.complex128)
    for r in range(L):
        T = r * M_ + M_ - 1
        e, phi = T // L, T % L
        # windows[q][k] = x[qM + e - k]  (index into `full` shifted by K)
        idx = (np.arange(Q)[:, None] * M_ + e + K) - np.arange(K)[None, :]
        win = full[idx]                                               # [Q][K]
        n0 = (np.uint64(n_start) + np.arange(Q, dtype=np.uint64) * np.uint64(M_) + np.uint64(e))
        for c, st in enumerate(np.asarray(steps, dtype=np.uint64)):
            g = np.asarray(taps if taps_per_channel is None else taps_per_channel[c], dtype=np.float64)
            ph = (st * k) & 0xFFFFFFFF
            a = 2.0 * np.pi * ph.astype(np.int64).astype(np.uint32).view(np.int32).astype(np.float64) / 4294967296.0
            w = g[phi::L] * np.exp(1j * a)                            # h_r[k] e^{+j w k}
            if bf16:
                w = to_bf16(np.real(w).astype(np.float32)).astype(np.float64) + 1j * to_bf16(np.imag(w).astype(np.float32)).astype(np.float64)
            p0 = (st * n0) & 0xFFFFFFFF
            a0 = 2.0 * np.pi * p0.astype(np.uint32).view(np.int32).astype(np.float64) / 4294967296.0
            out[c, r::L] = (win @ w) * np.exp(-1j * a0)
    return out
