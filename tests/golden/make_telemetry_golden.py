#!/usr/bin/env python3
"""Telemetry fixtures: what the UNMODIFIED reference's xxx_decode() leaves in SondeData for every framer window
of seeded synthetic signals, next to the frame record (post-FEC bytes + gate) of the same window.

Run in the build container (needs oracle/_ref/libsonde_ref.so):  python tests/golden/make_telemetry_golden.py
Output: tests/golden/telemetry.json  {type name: [{rec fields..., "sonde_data_hex": 96 bytes}, ...]}
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

SECONDS = {0: 8.0, 1: 6.0, 2: 5.0, 3: 8.0, 4: 8.0, 5: 8.0, 6: 4.0}


def flight_spec(synth, name):
    """Consecutive frames with running counters and complete calibration (synth.*_flight_bits)."""
    import numpy as np
    rng = np.random.default_rng(0xF11687)
    if name == "dfm09_flight":
        stype, bits, frames = synth.DFM09, synth.dfm_flight_bits(48, rng), 48
    elif name == "dfm06_flight":
        stype, bits, frames = synth.DFM09, synth.dfm_flight_bits(24, rng, dfm06=True), 24
    elif name == "ims100_flight":
        stype, bits, frames = synth.IMS100, synth.meisei_flight_bits(False, 72, rng, seq0=60), 72
    else:
        stype, bits, frames = synth.IMS100, synth.meisei_flight_bits(True, 72, rng, seq0=3), 72
    spec = synth.ChannelSpec(stype, 0xF1, snr_db=30.0, custom_bits=bits)
    n = int(48000 * (frames + 1.5) * synth.MODEMS[stype].frame_bits / synth.MODEMS[stype].baud)
    return stype, spec, n


FLIGHTS = ("dfm09_flight", "dfm06_flight", "ims100_flight", "rs11g_flight")


def cases(ref, synth, stype, channel=3, chunk=48000, flight=None):
    import numpy as np
    if flight:
        stype, spec, n = flight_spec(synth, flight)
        fm = synth.make_fm(spec, n)
    else:
        n = int(48000 * SECONDS[stype])
        fm = synth.make_fm(synth.default_spec(stype, channel), n)
    recs = ref.frames_run(stype, fm, chunk)
    sd, _ = ref.decode_run(stype, fm, chunk)
    assert len(recs) == len(sd), (stype, len(recs), len(sd))
    out = []
    for r, s in zip(recs, sd):
        out.append({"ok": int(r.ok), "status": int(r.status), "aux": int(r.aux), "data_len": int(r.data_len),
                    "data_hex": bytes(r.data[:max(r.data_len, 132)]).hex(),
                    "sonde_data_hex": bytes(s).hex()})
    return out


def main():
    from tests.reflib import RefLib
    from sdrpp_radiosonde_b200 import synth
    ref = RefLib()
    out = {}
    for stype in range(7):
        out[synth.TYPE_NAMES[stype]] = cases(ref, synth, stype)
        c = out[synth.TYPE_NAMES[stype]]
        print(synth.TYPE_NAMES[stype], len(c), sum(1 for x in c if int.from_bytes(bytes.fromhex(x["sonde_data_hex"])[:4], "little")))
    for name in FLIGHTS:
        out[name] = cases(ref, synth, None, flight=name)
        c = out[name]
        print(name, len(c), sum(1 for x in c if int.from_bytes(bytes.fromhex(x["sonde_data_hex"])[:4], "little")))
    with open(os.path.join(HERE, "telemetry.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
