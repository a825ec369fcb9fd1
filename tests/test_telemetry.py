"""Host telemetry parsers (sdrpp_radiosonde_b200/host/telemetry.hpp):
frame record -> SondeData, against what the unmodified reference's xxx_decode() produced for the same framer
windows (tests/golden/telemetry.json, made by tests/golden/make_telemetry_golden.py) and, when oracle/_ref is
present, against the reference run live on another channel.  Host-only code: runs without a GPU."""
import ctypes
import json
import os

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import synth
from tests import reflib

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
GOLD = os.path.join(ROOT, "tests", "golden", "telemetry.json")

# types whose telemetry is converted (the others deliver fields == 0 and the raw frame callback)
COVERED = {"rs41": synth.RS41, "m10": synth.M10, "mrzn1": synth.MRZN1, "imet4": synth.IMET4, "c50": synth.C50,
           "dfm09": synth.DFM09, "ims100": synth.IMS100}
# "flights": consecutive frames with running counters and a complete calibration (synth.*_flight_bits)
FLIGHTS = {"dfm09_flight": synth.DFM09, "dfm06_flight": synth.DFM09, "ims100_flight": synth.IMS100,
           "rs11g_flight": synth.IMS100}
# values derived from time(NULL) in the reference (imet4/parser.c:43-64, ims100.c date handling): masked
WALLCLOCK = {"imet4": ("time", "serial", "speed", "heading", "climb"), "ims100": ("time", "climb"),
             "ims100_flight": ("time", "climb")}
# iMS-100 / RS-11G compute temperature and humidity from calibration memory that the reference never initialises
# (malloc, ims100.c:47-62): compared once all 64 fragments have been received (calib_percent == 100)
NEEDS_CALIB = ("ims100", "ims100_flight", "rs11g_flight")
FLOATS = ("lat", "lon", "alt", "speed", "climb", "heading", "calib_percent", "temp", "rh", "pressure", "o3_mpa")


def _lib():
    lib = ctypes.CDLL(os.path.join(ROOT, "sdrpp_radiosonde_b200", "libsonde_b200_compat.so"))
    lib.sonde_telemetry_create.restype = ctypes.c_void_p
    lib.sonde_telemetry_create.argtypes = [ctypes.c_int]
    lib.sonde_telemetry_destroy.argtypes = [ctypes.c_void_p]
    lib.sonde_telemetry_parse.argtypes = [ctypes.c_void_p, ctypes.POINTER(reflib.FrameRec), ctypes.POINTER(reflib.SondeData)]
    return lib


class DfmSeen:
    """The reference assembles DFM output in a struct it never initialises (dfm09.c:43-49: malloc, only .fields = 0), so a
    value is defined only once the subframe that carries it has been received: track that from the unpacked frames."""
    OWNER = {"lat": ("g", 2), "speed": ("g", 2), "lon": ("g", 3), "heading": ("g", 3), "alt": ("g", 4),
             "climb": ("g", 4), "pressure": ("g", 4), "temp": ("p", 0), "rh": ("p", 1), "calib_percent": ("p", 0)}

    def __init__(self):
        self.seen = set()

    def update(self, rec):
        if rec.ok:
            d = bytes(rec.data[64:82])
            self.seen |= {("p", d[0]), ("g", d[4]), ("g", d[11])}

    def skip(self):
        return tuple(k for k, o in self.OWNER.items() if o not in self.seen)


def _same(got, want, name, idx, extra_skip=()):
    assert got.fields == want.fields, (name, idx, hex(got.fields), hex(want.fields))
    skip = WALLCLOCK.get(name, ()) + tuple(extra_skip)
    f = want.fields
    if "serial" not in skip and f & 0x02:
        assert got.serial == want.serial, (name, idx)
    if f & 0x01:
        assert got.seq == want.seq, (name, idx)
    if "time" not in skip and f & 0x10:
        assert got.time == want.time, (name, idx, got.time, want.time)
    used = []
    if f & 0x04: used += ["lat", "lon", "alt"]
    if f & 0x08: used += ["speed", "climb", "heading"]
    if f & 0x20:
        used += ["calib_percent", "pressure"]
        if name not in NEEDS_CALIB or want.calib_percent == 100.0:
            used += ["temp", "rh"]
    if f & 0x40: used += ["o3_mpa"]
    for k in used:
        if k in skip:
            continue
        a, b = np.float32(getattr(got, k)), np.float32(getattr(want, k))
        # same expressions, same libm: bit-identical (NaN compares by pattern)
        assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (name, idx, k, a, b)
    if f & 0x80:
        assert got.shutdown == want.shutdown, (name, idx)


def _run(lib, stype, recs):
    t = lib.sonde_telemetry_create(stype)
    assert t
    out = []
    for r in recs:
        sd = reflib.SondeData()
        assert lib.sonde_telemetry_parse(t, ctypes.byref(r), ctypes.byref(sd)) == 0
        out.append(sd)
    lib.sonde_telemetry_destroy(t)
    return out


@pytest.mark.parametrize("name", sorted(COVERED) + sorted(FLIGHTS))
def test_telemetry_matches_reference_fixture(name):
    lib = _lib()
    stype = COVERED.get(name, FLIGHTS.get(name))
    cases = json.load(open(GOLD))[name]
    recs = []
    for c in cases:
        r = reflib.FrameRec()
        r.type, r.ok, r.status, r.aux, r.data_len = stype, c["ok"], c["status"], c["aux"], c["data_len"]
        d = bytes.fromhex(c["data_hex"])
        ctypes.memmove(r.data, d, len(d))
        recs.append(r)
    got = _run(lib, stype, recs)
    n_fields = 0
    seen = DfmSeen()
    for i, (g, c) in enumerate(zip(got, cases)):
        want = reflib.SondeData.from_buffer_copy(bytes.fromhex(c["sonde_data_hex"]))
        seen.update(recs[i])
        _same(g, want, name, i, seen.skip() if stype == synth.DFM09 else ())
        n_fields += want.fields != 0
    assert n_fields > 0
    if name in NEEDS_CALIB and name != "ims100":
        assert sum(1 for c in cases if reflib.SondeData.from_buffer_copy(bytes.fromhex(c["sonde_data_hex"])).calib_percent == 100.0) >= 4


@pytest.mark.skipif(not reflib.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(COVERED))
def test_telemetry_matches_reference_live(name):
    lib = _lib()
    stype = COVERED[name]
    ref = reflib.RefLib()
    for channel, chunk in ((11, 48000), (12, 1024)):
        fm = synth.make_fm(synth.default_spec(stype, channel), 48000 * 4)
        recs = ref.frames_run(stype, fm, chunk)
        want, _ = ref.decode_run(stype, fm, chunk)
        assert len(recs) == len(want)
        got = _run(lib, stype, recs)
        seen = DfmSeen()
        for i, (g, w) in enumerate(zip(got, want)):
            seen.update(recs[i])
            _same(g, w, name, i, seen.skip() if stype == synth.DFM09 else ())


def test_telemetry_argument_errors():
    lib = _lib()
    sd = reflib.SondeData()
    r = reflib.FrameRec()
    assert lib.sonde_telemetry_parse(None, ctypes.byref(r), ctypes.byref(sd)) != 0
