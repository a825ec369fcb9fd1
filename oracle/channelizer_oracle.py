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


def channelize(x, taps, steps, decim, n_start=0, history=None, bf16=False):
    """x: complex wideband chunk (len multiple of decim), history: the samples before it (zeros if None).

    Returns y[C][len(x) // decim] complex128."""
    x = np.asarray(x)
    K, D = len(taps), int(decim)
    hist = np.zeros(K, dtype=np.complex128) if history is None else np.asarray(history, dtype=np.complex128)[-K:]
    if len(hist) < K:
        hist = np.concatenate([np.zeros(K - len(hist), dtype=np.complex128), hist])
    xr, xi = np.real(x).astype(np.float32), np.imag(x).astype(np.float32)
    if bf16:
        xr, xi = to_bf16(xr), to_bf16(xi)
        hist = to_bf16(np.real(hist).astype(np.float32)) + 1j * to_bf16(np.imag(hist).astype(np.float32))
    full = np.concatenate([hist, xr.astype(np.float64) + 1j * xi.astype(np.float64)])
    M = len(x) // D
    # windows[m][k] = x[mD + D-1 - k]  (index into `full` shifted by K)
    idx = (np.arange(M)[:, None] * D + D - 1 + K) - np.arange(K)[None, :]
    win = full[idx]                                                   # [M][K]
    k = np.arange(K, dtype=np.uint64)
    out = np.empty((len(steps), M), dtype=np.complex128)
    for c, st in enumerate(np.asarray(steps, dtype=np.uint64)):
        ph = (st * k) & 0xFFFFFFFF
        a = 2.0 * np.pi * ph.astype(np.int64).astype(np.uint32).view(np.int32).astype(np.float64) / 4294967296.0
        w = np.asarray(taps, dtype=np.float64) * np.exp(1j * a)       # h[k] e^{+j w k}
        if bf16:
            w = to_bf16(np.real(w).astype(np.float32)).astype(np.float64) + 1j * to_bf16(np.imag(w).astype(np.float32)).astype(np.float64)
        n0 = (np.uint64(n_start) + np.arange(M, dtype=np.uint64) * np.uint64(D) + np.uint64(D - 1))
        p0 = (st * n0) & 0xFFFFFFFF
        a0 = 2.0 * np.pi * p0.astype(np.uint32).view(np.int32).astype(np.float64) / 4294967296.0
        out[c] = (win @ w) * np.exp(-1j * a0)
    return out
