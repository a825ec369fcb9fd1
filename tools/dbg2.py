"""debug: per-chunk comparison of soft symbols: new pipeline kernel vs the phase-by-phase kernel (legacy flag)"""
import sys, numpy as np
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi, synth
stype = int(sys.argv[1]); chunk = int(sys.argv[2]); no_tma = "notma" in sys.argv
n = 48000
batch = np.stack([synth.make_fm(synth.default_spec(stype, 0), n)])
def run(legacy):
    dec = capi.BatchDecoder([stype], chunk, keep_soft=True, legacy_kernel=legacy, no_tma=no_tma)
    out = []
    for pos in range(0, n, chunk):
        part = np.ascontiguousarray(batch[:, pos:pos + chunk])
        dec.process_fm(part); dec.fetch()
        out.append((dec.fetch_soft()[0].copy(), dec.fetch_state()[0].copy(), dec.debug_demod_state()[0].copy()))
    dec.close()
    return out
a = run(False); b = run(True)
shown = 0
for ci, ((sa, sta, ra), (sb, stb, rb)) in enumerate(zip(a, b)):
    if not np.array_equal(ra[16:].view(np.uint32), rb[16:].view(np.uint32)) and shown < 4:
        bad = np.nonzero(ra[16:].view(np.uint32) != rb[16:].view(np.uint32))[0]
        print(f'chunk {ci}: hist differs at', bad, 'new', ra[16:][bad[:6]], 'legacy', rb[16:][bad[:6]])
    if len(sa) != len(sb) or not np.array_equal(sa.view(np.uint32), sb.view(np.uint32)):
        m = min(len(sa), len(sb))
        bad = np.nonzero(sa[:m].view(np.uint32) != sb[:m].view(np.uint32))[0]
        print(f"chunk {ci}: nsoft {len(sa)} vs {len(sb)}; mismatching idx {bad[:8]} of {m}; new {sa[bad[:3]]} legacy {sb[bad[:3]]}")
        print("   state new", sta[:6], "legacy", stb[:6])
        shown += 1
        if shown >= 4: break
print("done, chunks", len(a))
