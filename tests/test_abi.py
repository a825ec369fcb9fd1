"""CPU: the C-ABI library loads, exports every symbol include/sonde_b200.h declares, fails loudly
without a GPU, and its host-side modem tables equal the reference's."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from sdrpp_radiosonde_b200 import capi, synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sonde_b200.h")).read()
    declared = set(re.findall(r"SONDE_API\s+[\w\s\*]+?\b(sonde_b200_\w+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.sonde_b200_version()


def test_record_layout_matches_header():
    assert ctypes.sizeof(capi.FrameRec) == 1080
    assert capi.REC_DTYPE.fields["raw"][1] == 40 and capi.REC_DTYPE.fields["data"][1] == 560
    assert ctypes.sizeof(capi.Config) == 40


def test_modem_tables_bit_exact():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "modem_tables.json")))
    for name, t in gold.items():
        stype = synth.TYPE_NAMES.index(name)
        taps, consts = capi.modem_info(stype)
        assert [int(v) for v in taps.view(np.uint32)] == t["taps_u32"], name
        assert [int(v) for v in consts[:5].view(np.uint32)] == t["timing_u32"], name
    # known answers quoted in SURVEY.md H3
    assert abs(capi.modem_info(synth.RS41)[0][24] - 0.288000017) < 1e-9
    assert abs(capi.modem_info(synth.DFM09)[0][24] - 0.149999991) < 1e-9
    assert capi.modem_info(synth.IMET4)[1][5] == 39 and capi.modem_info(synth.C50)[1][5] == 20


def test_bad_arguments_rejected():
    lib = capi.load()
    assert lib.sonde_b200_create(None, None) == capi.ERR_ARG
    h = ctypes.c_void_p()
    types = (ctypes.c_int32 * 1)(9)
    cfg = capi.Config(1, 48000, 1024, 0, types, 0.0, 0, 0)
    assert lib.sonde_b200_create(ctypes.byref(h), ctypes.byref(cfg)) == capi.ERR_ARG
    assert lib.sonde_b200_process_fm(None, None, 0) == capi.ERR_ARG
    assert lib.sonde_b200_modem_info(99, 48000, None, 0, None) == capi.ERR_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SondeError) as e:
        capi.BatchDecoder([synth.RS41], 1024)
    assert e.value.code == capi.ERR_NODEVICE


def test_batch_layout_plan_host_only():
    """sonde_b200_debug_plan (no GPU): how build_groups lays a batch out on 148 SMs.  Per kernel variant: groups, channels
    per group, rows of the TMA box (0 = one bulk copy per row), input-row step of a group's channels; total CTAs."""
    import ctypes
    import numpy as np
    from sdrpp_radiosonde_b200 import capi, synth
    lib = ctypes.CDLL(capi.LIB_PATH)
    lib.sonde_b200_debug_plan.restype = ctypes.c_int

    def plan(types, sms=148):
        t = np.ascontiguousarray(types, dtype=np.int32)
        out = np.zeros(17, dtype=np.int32)
        rc = lib.sonde_b200_debug_plan(t.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(t.size), ctypes.c_int(48000),
                                       ctypes.c_int(sms), out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        return out[:16].reshape(4, 4), int(out[16])

    # BASELINE config 2: 1024 x RS41 -> 147 CTAs of 7 channels on the 148 SMs, one tensor box of 7 consecutive rows per tile
    v, total = plan(np.full(1024, synth.RS41))
    assert v[0].tolist() == [147, 7, 7, 1] and total == 147 and v[1:, 0].sum() == 0
    # fewer channels than SMs: one channel per CTA
    v, total = plan(np.full(146, synth.RS41))
    assert v[0].tolist() == [146, 1, 1, 1]
    # config 3: DFM and M10 on alternating channels -> two waves, each group's channels two rows apart (3-D tensor map)
    v, total = plan([synth.DFM09 if c % 2 == 0 else synth.M10 for c in range(2048)])
    assert v[1].tolist() == [147, 7, 7, 2] and v[2].tolist() == [147, 7, 7, 2] and total == 296
    # config 5, typed: seven types interleaved -> every variant padded to an even CTA count (clusters of two), all of it
    # within one wave, the AFSK kernel gets the smaller groups, rows 7 apart
    v, total = plan([c % 7 for c in range(1024)])
    assert total <= 148
    assert v[3][1] < v[0][1] and v[3][1] <= v[2][1]
    assert all(int(r[3]) == 7 for r in v if r[0] > 0) and all(int(r[2]) == int(r[1]) for r in v if r[0] > 0)
    padded = sum((int(g) + 1) & ~1 for g in v[:, 0])
    assert padded == total
    # AUTO channels: the virtual channels of one decoder are consecutive user channels -> consecutive rows
    v, total = plan(np.full(64, -1))
    assert all(int(r[3]) == 1 and int(r[2]) == int(r[1]) for r in v if r[0] > 0)
    # a batch whose groups are not evenly spaced falls back to row copies for that variant
    v, total = plan([synth.RS41, synth.M10, synth.RS41, synth.RS41, synth.M10, synth.RS41, synth.RS41] * 64)
    assert v[0][2] == 0


def test_product_binaries_do_not_link_test_infrastructure():
    """The shipped library, the compat layer and the batch runner depend on the CUDA runtime and on each other — never on
    the CPU oracle, the compiled reference or the test stand-in of the ABI (tests/cpp/stub_*.c); and no product source
    includes anything from oracle/ or tests/."""
    import re
    import subprocess
    pkg = os.path.join(ROOT, "sdrpp_radiosonde_b200")
    for name in ("libsonde_b200.so", "libsonde_b200_compat.so", "sonde_b200_batch"):
        path = os.path.join(pkg, name)
        assert os.path.exists(path), path
        needed = re.findall(r"\(NEEDED\)\s+Shared library: \[([^\]]+)\]", subprocess.run(["readelf", "-d", path], capture_output=True, text=True, check=True).stdout)
        assert needed and not [n for n in needed if re.search(r"oracle|_ref|stub|writers", n)], (name, needed)
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".c", ".py", ".inc")):
                text = open(os.path.join(d, f), errors="replace").read()
                if re.search(r'#include\s+"[^"]*(oracle|tests)/', text) or re.search(r"^\s*(from|import)\s+(oracle|tests)\b", text, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
