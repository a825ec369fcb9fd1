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
