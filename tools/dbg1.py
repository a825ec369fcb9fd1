"""debug: first mismatch of bits / soft symbols vs the compiled reference, RS41 FM input"""
import sys, numpy as np
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi, synth
from tests import reflib
from tests.gpu_util import run_gpu
stype = int(sys.argv[1]) if len(sys.argv) > 1 else 0
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
kind = sys.argv[3] if len(sys.argv) > 3 else "fm"
no_tma = "notma" in sys.argv
n = 48000
spec = synth.default_spec(stype, 0)
batch = np.stack([synth.make_fm(spec, n)]) if kind == "fm" else np.stack([synth.make_iq(spec, n)])
got = run_gpu([stype], batch, chunk, kind=kind, keep_soft=True, want_bits=True, no_tma=no_tma)
chk = reflib.RefLib() if reflib.have_ref() else reflib.OracleLib()
if kind == "fm":
    soft, state = chk.gfsk_soft(synth.MODEMS[stype].baud, batch[0], chunk)
    g = got["soft"][0]
    print("nsoft", len(g), len(soft))
    m = min(len(g), len(soft))
    bad = np.nonzero(g[:m].view(np.uint32) != soft[:m].view(np.uint32))[0]
    print("soft mismatches", len(bad), "first", bad[:10])
    if len(bad):
        i = bad[0]
        print("around first:", g[max(0, i-2):i+3], soft[max(0, i-2):i+3])
    bits = chk.demod_bits(stype, batch[0], chunk)
    gb = got["bits"][0]
    m = min(len(gb), len(bits))
    badb = np.nonzero(gb[:m] != bits[:m])[0]
    print("bit bytes", len(gb), len(bits), "mismatching bytes", len(badb), badb[:10])
    print("state", got["state"][0, :6], state[:6])
