"""Small run of one sonde type through the C ABI (for compute-sanitizer / ncu): python tools/mini_run.py <type> [channels] [len] [chunk]"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from sdrpp_radiosonde_b200 import capi, synth
stype = int(sys.argv[1]) if len(sys.argv) > 1 else 0
C = int(sys.argv[2]) if len(sys.argv) > 2 else 11
n = int(sys.argv[3]) if len(sys.argv) > 3 else 24000
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
batch = np.stack([synth.make_iq(synth.default_spec(stype, c), n) for c in range(C)])
dec = capi.BatchDecoder(np.full(C, stype, np.int32), chunk)
tot = ok = 0
for pos in range(0, n, chunk):
    dec.process_iq(np.ascontiguousarray(batch[:, pos:pos + chunk]))
    recs, counts = dec.fetch()
    tot += int(counts.sum()); ok += sum(int(recs[c, i]["ok"]) for c in range(C) for i in range(counts[c]))
dec.close()
print(f"type {stype}: {C} channels x {n} samples in chunks of {chunk}: {tot} frames, {ok} ok")
