"""ctypes views of the two CPU checkers (TEST INFRASTRUCTURE):

  RefLib    -> oracle/_ref/libsonde_ref.so     the unmodified reference + oracle/ref_harness.c
  OracleLib -> oracle/_build/libsonde_oracle.so this repo's C restatement (oracle/sonde_oracle.c)

Neither is ever imported by the product package.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsonde_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libsonde_oracle.so")

REC_BYTES = 520


class FrameRec(ctypes.Structure):
    """sonde_frame_rec (include/sonde_b200.h)."""
    _fields_ = [
        ("type", ctypes.c_int32), ("chunk", ctypes.c_int32), ("sync_offset", ctypes.c_int32),
        ("inverted", ctypes.c_int32), ("status", ctypes.c_int32), ("ok", ctypes.c_int32),
        ("aux", ctypes.c_int32), ("data_len", ctypes.c_int32), ("bit_pos", ctypes.c_uint64),
        ("raw", ctypes.c_uint8 * REC_BYTES), ("data", ctypes.c_uint8 * REC_BYTES),
    ]

    def key(self, raw_bytes=None):
        n = self.data_len if self.data_len > 0 else 0
        n = max(n, 132)
        return (self.chunk, self.sync_offset, self.inverted, self.status, self.ok, self.aux,
                bytes(self.data[:n]), bytes(self.raw[:raw_bytes]) if raw_bytes else b"")


assert ctypes.sizeof(FrameRec) == 40 + 2 * REC_BYTES


class SondeData(ctypes.Structure):
    """SondeData (SD/include/data.h:28-50), 96 bytes on x86-64."""
    _fields_ = [
        ("fields", ctypes.c_int32), ("seq", ctypes.c_int32), ("serial", ctypes.c_char * 32),
        ("lat", ctypes.c_float), ("lon", ctypes.c_float), ("alt", ctypes.c_float),
        ("speed", ctypes.c_float), ("climb", ctypes.c_float), ("heading", ctypes.c_float),
        ("time", ctypes.c_int64),
        ("calib_percent", ctypes.c_float), ("temp", ctypes.c_float), ("rh", ctypes.c_float),
        ("pressure", ctypes.c_float), ("o3_mpa", ctypes.c_float), ("shutdown", ctypes.c_int32),
    ]


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _u8ptr(a):
    return (ctypes.c_uint8 * len(a)).from_buffer(a)


class _FrameRunner:
    """Shared wrapper for xxx_frames_run(type, fs, fm, n, chunk, recs, max)."""
    prefix = ""

    def frames_run(self, stype, fm, chunk, samplerate=48000, max_recs=None):
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        if max_recs is None:
            max_recs = fm.size // 400 + 64
        recs = (FrameRec * max_recs)()
        fn = getattr(self.lib, self.prefix + "_frames_run")
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                       ctypes.c_size_t, ctypes.POINTER(FrameRec), ctypes.c_int]
        n = fn(stype, samplerate, _fptr(fm), fm.size, chunk, recs, max_recs)
        assert 0 <= n <= max_recs, n
        return [recs[i] for i in range(n)]

    def demod_bits(self, stype, fm, chunk, samplerate=48000):
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        cap = fm.size // 8 + 64
        buf = np.zeros(cap, dtype=np.uint8)
        fn = getattr(self.lib, self.prefix + "_demod_bits")
        fn.restype = ctypes.c_long
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                       ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t]
        n = fn(stype, samplerate, _fptr(fm), fm.size, chunk,
               buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), cap)
        assert n >= 0
        return np.unpackbits(buf)[:n]

    def gfsk_soft(self, baud, fm, chunk, samplerate=48000):
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        cap = fm.size // 2 + 64
        soft = np.zeros(cap, dtype=np.float32)
        state = np.zeros(8, dtype=np.float32)
        fn = getattr(self.lib, self.prefix + "_gfsk_soft")
        fn.restype = ctypes.c_long
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                       ctypes.c_size_t, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                       ctypes.POINTER(ctypes.c_float)]
        n = fn(samplerate, baud, _fptr(fm), fm.size, chunk, _fptr(soft), cap, _fptr(state))
        assert 0 <= n <= cap
        return soft[:n].copy(), state

    def afsk_soft(self, stype, fm, chunk, samplerate=48000):
        """Soft symbols + final loop state of the reference's AFSK chain (compiled reference only)."""
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        cap = fm.size // 8 + 64
        soft = np.zeros(cap, dtype=np.float32)
        state = np.zeros(8, dtype=np.float32)
        fn = getattr(self.lib, self.prefix + "_afsk_soft")
        fn.restype = ctypes.c_long
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t, ctypes.c_size_t,
                       ctypes.POINTER(ctypes.c_float), ctypes.c_size_t, ctypes.POINTER(ctypes.c_float)]
        n = fn(int(stype), samplerate, _fptr(fm), fm.size, chunk, _fptr(soft), cap, _fptr(state))
        assert 0 <= n <= cap
        return soft[:n].copy(), state

    def gfsk_taps(self, baud, samplerate=48000):
        taps = np.zeros(256, dtype=np.float32)
        fn = getattr(self.lib, self.prefix + "_gfsk_taps")
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
        n = fn(samplerate, baud, _fptr(taps), 256)
        return taps[:n].copy()

    def gfsk_timing(self, baud, samplerate=48000):
        out = np.zeros(5, dtype=np.float32)
        fn = getattr(self.lib, self.prefix + "_gfsk_timing")
        fn.restype = None
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        fn(samplerate, baud, _fptr(out))
        return out

    # FEC / checksum known-answer entry points -------------------------------------------
    def _call_u8(self, name, buf, *extra, restype=ctypes.c_int):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        return fn(_u8ptr(buf), *extra)

    def rs41_correct(self, frame518: bytearray) -> int:
        assert len(frame518) == 518
        return self._call_u8("_rs41_correct", frame518)

    def bch_fix(self, message64: bytearray) -> int:
        assert len(message64) == 64
        return self._call_u8("_bch_fix", message64)

    def rs255_fix(self, block255: bytearray) -> int:
        assert len(block255) == 255
        return self._call_u8("_rs255_fix", block255)

    def rs41_descramble(self, src518: bytes) -> bytes:
        dst = bytearray(518)
        s = bytearray(src518)
        fn = getattr(self.lib, self.prefix + "_rs41_descramble")
        fn.restype = None
        fn(_u8ptr(dst), _u8ptr(s))
        return bytes(dst)

    def correlate(self, syncword, sync_len, bits: bytes, len_bytes):
        inv = ctypes.c_int(-1)
        b = bytearray(bits)
        fn = getattr(self.lib, self.prefix + "_correlate")
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        off = fn(syncword, sync_len, ctypes.addressof(_u8ptr(b)), len_bytes, ctypes.byref(inv))
        return off, inv.value

    def checksum(self, kind, data: bytes) -> int:
        b = bytearray(data) if len(data) else bytearray(1)
        fn = getattr(self.lib, f"{self.prefix}_{kind}")
        fn.restype = ctypes.c_uint
        fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        return fn(ctypes.addressof(_u8ptr(b)), len(data)) & 0xFFFF


class RefLib(_FrameRunner):
    prefix = "ref"

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)

    def decode_run(self, stype, fm, chunk, samplerate=48000, max_out=None):
        fm = np.ascontiguousarray(fm, dtype=np.float32)
        if max_out is None:
            max_out = fm.size // 400 + 64
        out = (SondeData * max_out)()
        chunks = (ctypes.c_int32 * max_out)()
        fn = self.lib.ref_decode_run
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                       ctypes.c_size_t, ctypes.POINTER(SondeData), ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        n = fn(stype, samplerate, _fptr(fm), fm.size, chunk, out, chunks, max_out)
        assert 0 <= n <= max_out
        return [out[i] for i in range(n)], [chunks[i] for i in range(n)]


class OracleLib(_FrameRunner):
    prefix = "orc"

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)


def _frames_run_iq(self, stype, iq, chunk, samplerate=48000, gain=0.0, max_recs=None):
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    if max_recs is None:
        max_recs = iq.size // 400 + 64
    recs = (FrameRec * max_recs)()
    fn = self.lib.orc_frames_run_iq
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_float,
                   ctypes.POINTER(FrameRec), ctypes.c_int]
    n = fn(stype, samplerate, iq.ctypes.data, iq.size, chunk, gain, recs, max_recs)
    assert 0 <= n <= max_recs, n
    return [recs[i] for i in range(n)]


OracleLib.frames_run_iq = _frames_run_iq


def _discriminate(self, iq, gain=0.636619747, libm=False):
    """The restated FM discriminator over one channel's IQ (prev phase 0 at the start): float FM stream.
    libm=True: the glibc atan2f variant kept to report how far an upstream-like discriminator drifts."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    out = np.zeros(iq.size, dtype=np.float32)
    prev = ctypes.c_float(0.0)
    fn = self.lib.orc_discriminate_libm if libm else self.lib.orc_discriminate
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]
    fn(iq.ctypes.data, iq.size, ctypes.c_float(gain), ctypes.byref(prev), out.ctypes.data)
    return out


OracleLib.discriminate = _discriminate


def _frames_run_ragged(self, stype, fm, chunks, samplerate=48000, max_recs=None):
    fm = np.ascontiguousarray(fm, dtype=np.float32)
    if max_recs is None:
        max_recs = fm.size // 80 + 64
    recs = (FrameRec * max_recs)()
    arr = (ctypes.c_size_t * len(chunks))(*chunks)
    fn = self.lib.orc_frames_run_ragged
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_size_t,
                   ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.POINTER(FrameRec), ctypes.c_int]
    n = fn(stype, samplerate, _fptr(fm), fm.size, arr, len(chunks), recs, max_recs)
    assert 0 <= n <= max_recs, n
    return [recs[i] for i in range(n)]


OracleLib.frames_run_ragged = _frames_run_ragged


def have_ref():
    return os.path.exists(REF_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)
