"""Helpers for the -m gpu parity tests: drive the C ABI over a [C][n] batch in chunks."""
import numpy as np

from sdrpp_radiosonde_b200 import capi


def rec_key(r, raw_bytes):
    """Comparable view of one frame record (GPU structured-array row or ctypes FrameRec)."""
    if isinstance(r, np.void):
        n = max(int(r["data_len"]), 132)
        return (int(r["chunk"]), int(r["sync_offset"]), int(r["inverted"]), int(r["status"]), int(r["ok"]),
                int(r["aux"]), bytes(r["data"][:n]), bytes(r["raw"][:raw_bytes]))
    n = max(int(r.data_len), 132)
    return (int(r.chunk), int(r.sync_offset), int(r.inverted), int(r.status), int(r.ok), int(r.aux),
            bytes(r.data[:n]), bytes(r.raw[:raw_bytes]))


def run_gpu(types, batch, chunk, kind="fm", keep_soft=False, want_bits=False, legacy_kernel=False, afsk_layout=0, no_tma=False):
    """Feed batch[C][n] through the GPU path in buffers of `chunk` samples.

    Returns dict(frames=[list of record rows per channel], bits=[...], soft=[...], state=...).
    """
    C, n = batch.shape
    dec = capi.BatchDecoder(types, min(chunk, n), keep_soft=keep_soft, legacy_kernel=legacy_kernel, afsk_layout=afsk_layout, no_tma=no_tma)
    frames = [[] for _ in range(C)]
    bits = [[] for _ in range(C)]
    soft = [[] for _ in range(C)]
    try:
        for pos in range(0, n, chunk):
            part = np.ascontiguousarray(batch[:, pos:pos + chunk])
            if kind == "fm":
                dec.process_fm(part)
            else:
                dec.process_iq(part)
            recs, counts = dec.fetch()
            for c in range(C):
                frames[c].extend(recs[c, :counts[c]].copy())
            if want_bits:
                for c, b in enumerate(dec.fetch_bits()):
                    bits[c].append(b)
            if keep_soft:
                for c, s in enumerate(dec.fetch_soft()):
                    soft[c].append(s)
        state = dec.fetch_state()
        launches = dec.launch_count
    finally:
        dec.close()
    return {
        "frames": frames,
        "bits": [np.concatenate(b) if b else np.zeros(0, np.uint8) for b in bits],
        "soft": [np.concatenate(s) if s else np.zeros(0, np.float32) for s in soft],
        "state": state,
        "launches": launches,
    }


def make_wideband(types, freqs, D, nsec, seed=1, amp=0.2, noise=0.01):
    """Narrowband synthetic sondes (synth.make_iq) resampled to D x 48 kS/s, shifted to their channel centres and
    summed.  Returns (narrowband[C][n], wideband[n * D] complex64)."""
    from scipy.signal import resample_poly
    from sdrpp_radiosonde_b200 import synth
    n = 48000 * nsec
    fs_in = 48000.0 * D
    nb = np.stack([synth.make_iq(synth.default_spec(t, c), n) for c, t in enumerate(types)])
    t = np.arange(n * D)
    wide = np.zeros(n * D, dtype=np.complex128)
    for c in range(len(types)):
        wide += amp * resample_poly(nb[c].astype(np.complex128), D, 1) * np.exp(2j * np.pi * freqs[c] / fs_in * t)
    rng = np.random.default_rng(seed)
    wide += noise * (rng.standard_normal(wide.size) + 1j * rng.standard_normal(wide.size))
    return nb, wide.astype(np.complex64)
