"""ctypes binding of the C-ABI library (include/sonde_b200.h -> libsonde_b200.so).

This is the only way the Python side reaches the kernels; there is no CPU fallback.
If the library is missing, or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsonde_b200.so")

REC_BYTES = 520

SONDE_OK, ERR_ARG, ERR_CUDA, ERR_NODEVICE, ERR_TOOLONG, ERR_STATE = 0, -1, -2, -3, -4, -5
_ERR_NAMES = {ERR_ARG: "SONDE_ERR_ARG", ERR_CUDA: "SONDE_ERR_CUDA", ERR_NODEVICE: "SONDE_ERR_NODEVICE",
              ERR_TOOLONG: "SONDE_ERR_TOOLONG", ERR_STATE: "SONDE_ERR_STATE"}

# every symbol include/sonde_b200.h declares (tests/test_abi.py checks the .so exports all of them)
EXPORTS = [
    "sonde_b200_create", "sonde_b200_destroy", "sonde_b200_process_iq", "sonde_b200_process_fm",
    "sonde_b200_process_iq_s16", "sonde_b200_process_iq_device", "sonde_b200_process_fm_device", "sonde_b200_max_frames",
    "sonde_b200_fetch", "sonde_b200_fetch_counts", "sonde_b200_fetch_totals", "sonde_b200_detected_types", "sonde_b200_auto_plausible", "sonde_b200_bits_stride", "sonde_b200_fetch_bits",
    "sonde_b200_soft_stride", "sonde_b200_fetch_soft", "sonde_b200_fetch_state", "sonde_b200_modem_info",
    "sonde_b200_host_alloc", "sonde_b200_host_alloc_wc", "sonde_b200_host_free", "sonde_b200_debug_plan", "sonde_b200_debug_stalls", "sonde_b200_debug_demod_state", "sonde_b200_process_iq_peer", "sonde_b200_stream", "sonde_b200_sync", "sonde_b200_join",
    "sonde_b200_last_kernel_ms", "sonde_b200_launch_count", "sonde_b200_last_error", "sonde_b200_version",
]


class FrameRec(ctypes.Structure):
    """sonde_frame_rec"""
    _fields_ = [
        ("type", ctypes.c_int32), ("chunk", ctypes.c_int32), ("sync_offset", ctypes.c_int32),
        ("inverted", ctypes.c_int32), ("status", ctypes.c_int32), ("ok", ctypes.c_int32),
        ("aux", ctypes.c_int32), ("data_len", ctypes.c_int32), ("bit_pos", ctypes.c_uint64),
        ("raw", ctypes.c_uint8 * REC_BYTES), ("data", ctypes.c_uint8 * REC_BYTES),
    ]


REC_DTYPE = np.dtype([
    ("type", "<i4"), ("chunk", "<i4"), ("sync_offset", "<i4"), ("inverted", "<i4"), ("status", "<i4"),
    ("ok", "<i4"), ("aux", "<i4"), ("data_len", "<i4"), ("bit_pos", "<u8"),
    ("raw", "u1", (REC_BYTES,)), ("data", "u1", (REC_BYTES,)),
])
assert REC_DTYPE.itemsize == ctypes.sizeof(FrameRec) == 40 + 2 * REC_BYTES


class Config(ctypes.Structure):
    """sonde_b200_config"""
    _fields_ = [
        ("n_channels", ctypes.c_int32), ("samplerate", ctypes.c_int32), ("max_chunk_len", ctypes.c_int32),
        ("device", ctypes.c_int32), ("types", ctypes.POINTER(ctypes.c_int32)), ("fm_gain", ctypes.c_float),
        ("keep_soft", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class SondeError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")


_lib = None


def load():
    """Load libsonde_b200.so (raises if it has not been built: run `python __graft_entry__.py`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built; run __graft_entry__.build()")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError:
        # libcudart.so.12 not on the loader path: take the one PyTorch ships
        import torch  # noqa: F401  (loads libcudart into the process)
        lib = ctypes.CDLL(LIB_PATH)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    i32p, f32p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
    sig = {
        "sonde_b200_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.POINTER(Config)]),
        "sonde_b200_destroy": (None, [vp]),
        "sonde_b200_process_iq": (ctypes.c_int, [vp, vp, sz]),
        "sonde_b200_process_fm": (ctypes.c_int, [vp, vp, sz]),
        "sonde_b200_process_iq_s16": (ctypes.c_int, [vp, vp, sz, ctypes.c_float]),
        "sonde_b200_process_iq_device": (ctypes.c_int, [vp, vp, sz, sz]),
        "sonde_b200_process_fm_device": (ctypes.c_int, [vp, vp, sz, sz]),
        "sonde_b200_max_frames": (ctypes.c_int, [vp]),
        "sonde_b200_fetch": (ctypes.c_int, [vp, vp, i32p]),
        "sonde_b200_fetch_counts": (ctypes.c_int, [vp, i32p, i32p]),
        "sonde_b200_fetch_totals": (ctypes.c_int, [vp, vp, vp, vp]),
        "sonde_b200_detected_types": (ctypes.c_int, [vp, i32p]),
        "sonde_b200_auto_plausible": (ctypes.c_int, [vp, vp]),
        "sonde_b200_bits_stride": (ctypes.c_int, [vp]),
        "sonde_b200_fetch_bits": (ctypes.c_int, [vp, vp, i32p]),
        "sonde_b200_soft_stride": (ctypes.c_int, [vp]),
        "sonde_b200_fetch_soft": (ctypes.c_int, [vp, f32p, i32p]),
        "sonde_b200_fetch_state": (ctypes.c_int, [vp, f32p]),
        "sonde_b200_modem_info": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, f32p, ctypes.c_int, f32p]),
        "sonde_b200_host_alloc": (vp, [sz]),
        "sonde_b200_host_alloc_wc": (vp, [sz]),
        "sonde_b200_host_free": (None, [vp]),
        "sonde_b200_stream": (vp, [vp]),
        "sonde_b200_debug_stalls": (ctypes.c_int, [vp, vp, ctypes.c_int]),
        "sonde_b200_debug_demod_state": (ctypes.c_int, [vp, vp, sz]),
        "sonde_b200_process_iq_peer": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz]),
        "sonde_b200_sync": (ctypes.c_int, [vp]),
        "sonde_b200_join": (ctypes.c_int, [vp]),
        "sonde_b200_last_kernel_ms": (ctypes.c_int, [vp, f32p, f32p]),
        "sonde_b200_launch_count": (ctypes.c_long, [vp]),
        "sonde_b200_last_error": (ctypes.c_char_p, [vp]),
        "sonde_b200_version": (ctypes.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def modem_info(stype: int, samplerate: int = 48000):
    """(taps[P*49], consts[8]) the kernels use — host-only, no GPU needed."""
    lib = load()
    taps = np.zeros(128, dtype=np.float32)
    consts = np.zeros(8, dtype=np.float32)
    n = lib.sonde_b200_modem_info(stype, samplerate, taps.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 128,
                                  consts.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    if n < 0:
        raise SondeError(n, "modem_info")
    return taps[:n].copy(), consts


def _f32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class PinnedBuffer:
    """Page-locked host memory from sonde_b200_host_alloc() (or, write_combined=True, sonde_b200_host_alloc_wc():
    write-only staging — never read it back on the CPU), viewed as a numpy array."""

    def __init__(self, shape, dtype, write_combined=False):
        self.lib = load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = (self.lib.sonde_b200_host_alloc_wc if write_combined else self.lib.sonde_b200_host_alloc)(self.nbytes)
        if not self.ptr:
            raise MemoryError("sonde_b200_host_alloc failed")
        buf = (ctypes.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.sonde_b200_host_free(self.ptr)
            self.ptr = None


class BatchDecoder:
    """One handle of the C ABI: C channels, one sonde type per channel.

    Mirrors the reference call protocol batched over channels: one process_*() call is one
    reference buffer of `len` samples per channel (the `while (xxx_decode(...) != PROCEED)`
    loop of src/decode/decoder.hpp:61); fetch() returns one record per PARSED return.
    """

    def __init__(self, types, max_chunk_len, samplerate=48000, device=0, fm_gain=0.0, keep_soft=False,
                 legacy_kernel=False, no_tma=False, afsk_layout=0, auto_preclassify=False):
        self.lib = load()
        self.types = np.ascontiguousarray(types, dtype=np.int32)
        self.C = int(self.types.size)
        self.max_chunk_len = int(max_chunk_len)
        cfg = Config(self.C, samplerate, self.max_chunk_len, device, _i32p(self.types), fm_gain,
                     1 if keep_soft else 0, (1 if legacy_kernel else 0) | (2 if no_tma else 0) | (4 if afsk_layout else 0) |
                     (8 if auto_preclassify else 0))
        h = ctypes.c_void_p()
        rc = self.lib.sonde_b200_create(ctypes.byref(h), ctypes.byref(cfg))
        if rc != SONDE_OK:
            raise SondeError(rc, "sonde_b200_create")
        self.h = h
        self.max_frames = self.lib.sonde_b200_max_frames(h)
        self.bits_stride = self.lib.sonde_b200_bits_stride(h)
        self.soft_stride = self.lib.sonde_b200_soft_stride(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sonde_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != SONDE_OK:
            raise SondeError(rc, (self.lib.sonde_b200_last_error(self.h) or b"").decode())

    # -- host buffers ---------------------------------------------------------------------
    def process_fm(self, fm: np.ndarray):
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        assert fm.ndim == 2 and fm.shape[0] == self.C, fm.shape
        self._ck(self.lib.sonde_b200_process_fm(self.h, fm.ctypes.data, fm.shape[1]))

    def process_iq(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        assert iq.ndim == 2 and iq.shape[0] == self.C, iq.shape
        self._ck(self.lib.sonde_b200_process_iq(self.h, iq.ctypes.data, iq.shape[1]))

    def process_iq_s16(self, iq16: np.ndarray, scale: float = 1.0 / 32768.0):
        """iq16: [C][len][2] int16 (interleaved I, Q); sample = int16 * scale."""
        iq16 = np.ascontiguousarray(iq16, dtype=np.int16)
        assert iq16.ndim == 3 and iq16.shape[0] == self.C and iq16.shape[2] == 2
        self._ck(self.lib.sonde_b200_process_iq_s16(self.h, iq16.ctypes.data, iq16.shape[1], ctypes.c_float(scale)))

    def process_s16_host_ptr(self, ptr: int, length: int, scale: float = 1.0 / 32768.0):
        self._ck(self.lib.sonde_b200_process_iq_s16(self.h, ptr, length, ctypes.c_float(scale)))

    def process_host_ptr(self, ptr: int, length: int, is_iq=True):
        fn = self.lib.sonde_b200_process_iq if is_iq else self.lib.sonde_b200_process_fm
        self._ck(fn(self.h, ptr, length))

    # -- device buffers (raw pointers, e.g. torch.Tensor.data_ptr()) ------------------------
    def process_iq_device(self, ptr: int, length: int, row_stride: int | None = None):
        self._ck(self.lib.sonde_b200_process_iq_device(self.h, ptr, length, row_stride or length))

    def process_fm_device(self, ptr: int, length: int, row_stride: int | None = None):
        self._ck(self.lib.sonde_b200_process_fm_device(self.h, ptr, length, row_stride or length))

    # -- results ------------------------------------------------------------------------------
    def fetch(self):
        """-> (recs[C][max_frames] structured array, counts[C])"""
        recs = np.zeros((self.C, self.max_frames), dtype=REC_DTYPE)
        counts = np.zeros(self.C, dtype=np.int32)
        self._ck(self.lib.sonde_b200_fetch(self.h, recs.ctypes.data, _i32p(counts)))
        return recs, counts

    def fetch_counts(self):
        frames = np.zeros(self.C, dtype=np.int32)
        ok = np.zeros(self.C, dtype=np.int32)
        self._ck(self.lib.sonde_b200_fetch_counts(self.h, _i32p(frames), _i32p(ok)))
        return frames, ok

    def auto_plausible(self):
        """Bit mask of decoder types currently run per channel (see sonde_b200_auto_plausible)."""
        m = np.zeros(self.C, dtype=np.uint32)
        self._ck(self.lib.sonde_b200_auto_plausible(self.h, m.ctypes.data))
        return m

    def detected_types(self):
        """Decoder each channel reports (-1 = AUTO channel not locked yet)."""
        t = np.zeros(self.C, dtype=np.int32)
        self._ck(self.lib.sonde_b200_detected_types(self.h, _i32p(t)))
        return t

    def fetch_totals(self):
        """-> (frames, ok, bits) int64[C] running totals since create"""
        out = [np.zeros(self.C, dtype=np.int64) for _ in range(3)]
        self._ck(self.lib.sonde_b200_fetch_totals(self.h, *[o.ctypes.data for o in out]))
        return tuple(out)

    def fetch_bits(self):
        """-> list of per-channel 0/1 arrays demodulated by the last call"""
        buf = np.zeros((self.C, self.bits_stride), dtype=np.uint8)
        n = np.zeros(self.C, dtype=np.int32)
        self._ck(self.lib.sonde_b200_fetch_bits(self.h, buf.ctypes.data, _i32p(n)))
        return [np.unpackbits(buf[c])[: n[c]] for c in range(self.C)]

    def fetch_soft(self):
        buf = np.zeros((self.C, self.soft_stride), dtype=np.float32)
        n = np.zeros(self.C, dtype=np.int32)
        self._ck(self.lib.sonde_b200_fetch_soft(self.h, _f32p(buf), _i32p(n)))
        return [buf[c, : n[c]].copy() for c in range(self.C)]

    def fetch_state(self):
        st = np.zeros((self.C, 8), dtype=np.float32)
        self._ck(self.lib.sonde_b200_fetch_state(self.h, _f32p(st)))
        return st

    def process_iq_peer(self, src_device, dev_ptr, length, row_stride=None):
        """[C][length] complex64 resident on GPU `src_device`: pulled by the copy engine, then decoded."""
        self._ck(self.lib.sonde_b200_process_iq_peer(self.h, int(src_device), ctypes.c_void_p(dev_ptr), length,
                                                     length if row_stride is None else row_stride))

    def debug_demod_state(self):
        """Raw demodulator state per (virtual) channel as float32 words [C][64] (diagnostics)."""
        out = np.zeros((self.C * 8, 64), dtype=np.float32)
        rc = self.lib.sonde_b200_debug_demod_state(self.h, out.ctypes.data, out.nbytes)
        self._ck(rc if rc < 0 else SONDE_OK)
        return out

    def debug_stalls(self):
        """First call enables the pipeline kernel's stall counters; later calls return [groups][4 roles][4]."""
        n = self.lib.sonde_b200_debug_stalls(self.h, None, 0)
        if n < 0:
            self._ck(n)
        out = np.zeros((max(n, 1), 4, 4), dtype=np.int64)
        n = self.lib.sonde_b200_debug_stalls(self.h, out.ctypes.data, out.shape[0])
        return out[:n]

    def join(self):
        """Device-side: the handle's main stream waits for all framer kernels issued so far."""
        self._ck(self.lib.sonde_b200_join(self.h))

    def sync(self):
        self._ck(self.lib.sonde_b200_sync(self.h))

    def last_kernel_ms(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._ck(self.lib.sonde_b200_last_kernel_ms(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    @property
    def launch_count(self):
        return int(self.lib.sonde_b200_launch_count(self.h))

    @property
    def stream(self):
        return self.lib.sonde_b200_stream(self.h)


# ------------------------------------------------------------------------------------------------
# wideband channelizer (include/sonde_b200_channelizer.h)
CHAN_EXPORTS = [
    "sonde_chan_create", "sonde_chan_create_ex", "sonde_chan_taps_of", "sonde_chan_destroy", "sonde_chan_process_c64", "sonde_chan_process_c64_device", "sonde_chan_process_c64_peer",
    "sonde_chan_process_s16", "sonde_chan_process_u8", "sonde_chan_num_taps", "sonde_chan_taps", "sonde_chan_steps",
    "sonde_chan_last_kernel_ms", "sonde_chan_last_error",
]


class ChanConfig(ctypes.Structure):
    """sonde_chan_config"""
    _fields_ = [("n_channels", ctypes.c_int32), ("decim", ctypes.c_int32), ("fs_out", ctypes.c_int32),
                ("taps_per_phase", ctypes.c_int32), ("cutoff_hz", ctypes.c_float), ("max_in_len", ctypes.c_int32),
                ("device", ctypes.c_int32), ("freq_hz", ctypes.POINTER(ctypes.c_double))]


class ChanOptions(ctypes.Structure):
    """sonde_chan_options"""
    _fields_ = [("precision", ctypes.c_int32), ("interp", ctypes.c_int32), ("cutoff_hz", ctypes.POINTER(ctypes.c_float)),
                ("reserved", ctypes.c_int32 * 4)]


def _chan_lib():
    lib = load()
    if getattr(lib, "_chan_ready", False):
        return lib
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    pvp, psz = ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)
    sig = {
        "sonde_chan_create": (ctypes.c_int, [pvp, ctypes.POINTER(ChanConfig)]),
        "sonde_chan_create_ex": (ctypes.c_int, [pvp, ctypes.POINTER(ChanConfig), ctypes.POINTER(ChanOptions)]),
        "sonde_chan_taps_of": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_int]),
        "sonde_chan_destroy": (None, [vp]),
        "sonde_chan_process_c64": (ctypes.c_int, [vp, vp, sz, vp, pvp, psz]),
        "sonde_chan_process_c64_device": (ctypes.c_int, [vp, vp, sz, vp, pvp, psz]),
        "sonde_chan_process_c64_peer": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, vp, pvp, psz]),
        "sonde_chan_process_s16": (ctypes.c_int, [vp, vp, sz, ctypes.c_float, vp, pvp, psz]),
        "sonde_chan_process_u8": (ctypes.c_int, [vp, vp, sz, vp, pvp, psz]),
        "sonde_chan_num_taps": (ctypes.c_int, [vp]),
        "sonde_chan_taps": (ctypes.c_int, [vp, vp, ctypes.c_int]),
        "sonde_chan_steps": (ctypes.c_int, [vp, vp, ctypes.c_int]),
        "sonde_chan_last_kernel_ms": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_float)]),
        "sonde_chan_last_error": (ctypes.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._chan_ready = True
    return lib


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr: int, shape, typestr="<f4"):
    """Zero-copy torch view of library-owned device memory (e.g. the channelizer's output)."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, shape, typestr), device="cuda")


class Channelizer:
    """Wideband IQ at fs_in = decim * fs_out -> C channels of complex64 at fs_out, on the GPU (tcgen05 GEMM).

    process_*() return (device pointer, row stride in samples, samples per channel); feed them to
    BatchDecoder.process_iq_device().  `stream` is a raw cudaStream_t (int), e.g. BatchDecoder.stream."""

    def __init__(self, freq_hz, decim, max_in_len, fs_out=48000, taps_per_phase=0, cutoff_hz=0.0, device=0,
                 precision=0, interp=1, cutoffs=None):
        """precision: 0 = bf16 operands, 1 = split bf16 (hi + lo, three tensor passes); interp: L of the rational
        resampler fs_out = fs_in * L / decim; cutoffs: per-channel -6 dB points in Hz (sonde_chan_options)"""
        self.lib = _chan_lib()
        self.freq = np.ascontiguousarray(freq_hz, dtype=np.float64)
        self.C, self.D, self.L = int(self.freq.size), int(decim), int(interp)
        cfg = ChanConfig(self.C, self.D, fs_out, taps_per_phase, cutoff_hz, int(max_in_len), device,
                         self.freq.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        self._cut = None if cutoffs is None else np.ascontiguousarray(cutoffs, dtype=np.float32)
        opt = ChanOptions(int(precision), self.L,
                          None if self._cut is None else self._cut.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        h = ctypes.c_void_p()
        rc = self.lib.sonde_chan_create_ex(ctypes.byref(h), ctypes.byref(cfg), ctypes.byref(opt))
        if rc != SONDE_OK:
            raise SondeError(rc, "sonde_chan_create_ex")
        self.h = h
        self.K = self.lib.sonde_chan_num_taps(h)
        self.taps = np.zeros(self.K, dtype=np.float64)
        self.lib.sonde_chan_taps(h, self.taps.ctypes.data, self.K)
        self.steps = np.zeros(self.C, dtype=np.uint32)
        self.lib.sonde_chan_steps(h, self.steps.ctypes.data, self.C)

    def _ck(self, rc, what):
        if rc != SONDE_OK:
            raise SondeError(rc, what + ": " + self.lib.sonde_chan_last_error(self.h).decode())

    def _run(self, fn, ptr, n_in, stream, *extra):
        out, stride = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(fn(self.h, ptr, n_in, *extra, ctypes.c_void_p(stream or 0), ctypes.byref(out), ctypes.byref(stride)), fn.__name__)
        return out.value, int(stride.value), n_in // self.D * self.L

    def taps_of(self, channel):
        t = np.zeros(self.K, dtype=np.float64)
        self.lib.sonde_chan_taps_of(self.h, int(channel), t.ctypes.data, self.K)
        return t

    def process_c64(self, wide: np.ndarray, stream=0):
        wide = np.ascontiguousarray(wide, dtype=np.complex64)
        self._keep = wide
        return self._run(self.lib.sonde_chan_process_c64, wide.ctypes.data, wide.size, stream)

    def process_c64_device(self, ptr: int, n_in: int, stream=0):
        return self._run(self.lib.sonde_chan_process_c64_device, ptr, n_in, stream)

    def process_c64_peer(self, src_device: int, ptr: int, n_in: int, stream=0):
        """the chunk is on another GPU (mapped with CUDA IPC across processes): copy-engine pull, then channelise"""
        out, stride = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(self.lib.sonde_chan_process_c64_peer(self.h, int(src_device), ptr, n_in, ctypes.c_void_p(stream or 0),
                                                      ctypes.byref(out), ctypes.byref(stride)), "sonde_chan_process_c64_peer")
        return out.value, int(stride.value), n_in // self.D * self.L

    def process_c64_host_ptr(self, ptr: int, n_in: int, stream=0):
        """raw host pointer (pinned memory for an asynchronous copy)"""
        return self._run(self.lib.sonde_chan_process_c64, ptr, n_in, stream)

    def process_s16(self, wide16: np.ndarray, scale=1.0 / 32768.0, stream=0):
        wide16 = np.ascontiguousarray(wide16, dtype=np.int16)
        assert wide16.ndim == 2 and wide16.shape[1] == 2
        self._keep = wide16
        return self._run(self.lib.sonde_chan_process_s16, wide16.ctypes.data, wide16.shape[0], stream, ctypes.c_float(scale))

    def process_u8(self, wide8: np.ndarray, stream=0):
        wide8 = np.ascontiguousarray(wide8, dtype=np.uint8)
        assert wide8.ndim == 2 and wide8.shape[1] == 2
        self._keep = wide8
        return self._run(self.lib.sonde_chan_process_u8, wide8.ctypes.data, wide8.shape[0], stream)

    def last_kernel_ms(self):
        a = ctypes.c_float()
        self._ck(self.lib.sonde_chan_last_kernel_ms(self.h, ctypes.byref(a)), "last_kernel_ms")
        return a.value

    def close(self):
        if self.h:
            self.lib.sonde_chan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
