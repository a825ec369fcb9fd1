#!/usr/bin/env python3
"""Harvest the reference's known-answer vectors into small committed fixtures.

Run in the build container (needs /root/reference and oracle/_ref/libsonde_ref.so):
    python tests/golden/make_golden.py

Sources (read, never copied as code):
  * SD/scripts/rs_bruteforce.py:6      a real 510-byte descrambled RS41 extended frame
  * SD/scripts/rs_bruteforce.py:8-26   six 46-bit iMS-100 BCH(63,51) messages (3 live + 3 commented)
Outputs:
  * rs41_frame.hex            the 510 bytes
  * bch_messages.json         the six messages
  * fec_known_answers.json    results of the UNMODIFIED reference on them (oracle/_ref)
  * ref_<type>.json           frame bytes the reference decodes from seeded synthetic FM signals
"""
import ctypes
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF_SCRIPT = "/root/reference/src/decode/sondedump/scripts/rs_bruteforce.py"


def harvest():
    src = open(REF_SCRIPT).read()
    lists = re.findall(r"^(#\s*)?data = \[(.*?)\]", src, flags=re.S | re.M)
    rs41 = None
    bch = []
    for _, body in lists:
        body = re.sub(r"#", "", body)
        vals = [int(v, 0) for v in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]
        if len(vals) == 510:
            rs41 = vals
        elif len(vals) == 46:
            bch.append(vals)
    assert rs41 is not None and len(bch) == 6, (rs41 is not None, len(bch))
    return bytes(rs41), bch


def main():
    from tests.reflib import RefLib          # ctypes wrapper around oracle/_ref
    from sdrpp_radiosonde_b200 import synth

    rs41, bch = harvest()
    with open(os.path.join(HERE, "rs41_frame.hex"), "w") as f:
        f.write(rs41.hex() + "\n")
    with open(os.path.join(HERE, "bch_messages.json"), "w") as f:
        json.dump(bch, f)

    ref = RefLib()
    ka = {}
    frame = bytearray(synth.RS41_HEADER + rs41)
    ka["rs41_first_pass"] = ref.rs41_correct(frame)
    ka["rs41_corrected_hex"] = bytes(frame).hex()
    ka["rs41_second_pass"] = ref.rs41_correct(frame)
    ka["rs41_changed_bytes"] = [i for i in range(518) if frame[i] != (synth.RS41_HEADER + rs41)[i]]
    ka["bch"] = []
    rng = np.random.default_rng(1234)
    for msg in bch:
        for nerr in (0, 1, 2, 3):
            m = bytearray(64)
            m[17:63] = bytes(msg)
            pos = sorted(int(p) for p in rng.choice(46, nerr, replace=False)) if nerr else []
            for p in pos:
                m[17 + p] ^= 1
            before = bytes(m)
            ret = ref.bch_fix(m)
            ka["bch"].append({"msg": msg, "flip": pos, "ret": ret, "after_hex": bytes(m).hex(),
                              "before_hex": before.hex()})
    with open(os.path.join(HERE, "fec_known_answers.json"), "w") as f:
        json.dump(ka, f, indent=0)

    # demodulator constants as the unmodified reference derives them (gfsk_init)
    tabs = {}
    for stype in range(5):
        baud = synth.MODEMS[stype].baud
        tabs[synth.TYPE_NAMES[stype]] = {
            "baud": baud,
            "taps_u32": [int(v) for v in ref.gfsk_taps(baud).view(np.uint32)],
            "timing_u32": [int(v) for v in ref.gfsk_timing(baud).view(np.uint32)],
        }
    with open(os.path.join(HERE, "modem_tables.json"), "w") as f:
        json.dump(tabs, f)

    # soft symbols / loop state of the reference's own primitives on a seeded RS41 + M10 signal
    soft = {}
    for stype in (synth.RS41, synth.M10, synth.DFM09):
        fm = synth.make_fm(synth.default_spec(stype, 7), 24000)
        s, st = ref.gfsk_soft(synth.MODEMS[stype].baud, fm, 1024)
        bits = ref.demod_bits(stype, fm, 1024)
        soft[synth.TYPE_NAMES[stype]] = {
            "n": 24000, "chunk": 1024, "channel": 7,
            "fm_crc32": int(np.uint32(__import__("zlib").crc32(fm.tobytes()))),
            "soft_u32": [int(v) for v in s.view(np.uint32)],
            "state_u32": [int(v) for v in st[:6].view(np.uint32)],
            "bits_hex": np.packbits(bits).tobytes().hex(), "nbits": int(bits.size)}
    with open(os.path.join(HERE, "soft_symbols.json"), "w") as f:
        json.dump(soft, f)

    # frame bytes decoded by the unmodified reference from seeded synthetic signals
    for stype in range(7):
        name = synth.TYPE_NAMES[stype]
        out = {"type": stype, "cases": []}
        for ch, (chunk, n_s) in enumerate([(1024, 3.0), (48000, 3.0)]):
            spec = synth.default_spec(stype, ch)
            if stype == synth.RS41:
                spec.bit_errors = 20
            n = int(48000 * n_s)
            fm = synth.make_fm(spec, n)
            recs = ref.frames_run(stype, fm, chunk)
            out["cases"].append({
                "channel": ch, "chunk": chunk, "n": n,
                "fm_crc32": int(np.uint32(__import__("zlib").crc32(fm.tobytes()))),
                "frames": [{"chunk": int(r.chunk), "sync_offset": int(r.sync_offset), "inverted": int(r.inverted),
                            "status": int(r.status), "ok": int(r.ok), "aux": int(r.aux),
                            "data_hex": bytes(r.data[:max(r.data_len, 132)]).hex(),
                            "raw_hex": bytes(r.raw[:(synth.MODEMS[stype].frame_bits + 7) // 8]).hex()}
                           for r in recs]})
        with open(os.path.join(HERE, f"ref_{name}.json"), "w") as f:
            json.dump(out, f)
        print(name, [len(c["frames"]) for c in out["cases"]],
              [sum(fr["ok"] for fr in c["frames"]) for c in out["cases"]])


if __name__ == "__main__":
    main()
