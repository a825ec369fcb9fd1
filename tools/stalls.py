"""Diagnostics: per-role stall split of the pipeline kernel on the bench workload (reduced)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from sdrpp_radiosonde_b200 import capi, synth
import bench
stype = int(sys.argv[1]) if len(sys.argv) > 1 else 0
C, L = (int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1024), 48000
iq = bench.gen_batch(stype, 0, C, L, 16)
d = torch.from_numpy(iq).cuda()
dec = capi.BatchDecoder(np.full(C, stype, np.int32), L, no_tma=('notma' in sys.argv[2:]), afsk_layout=(1 if 'layout1' in sys.argv[2:] else 0))
dec.process_iq_device(d.data_ptr(), L); dec.sync()
dec.debug_stalls()
dec.process_iq_device(d.data_ptr(), L); dec.sync()
st = dec.debug_stalls().astype(np.float64)
print("demod ms", dec.last_kernel_ms())
names = (["PW", "A1", "BX", "TM"] if stype >= 5 else ["PW first", "AG", "PW last", "TM"])
m = st.mean(axis=0)
for r in range(4):
    tot = m[r, 2]
    print(f"{names[r]}: total {tot/L:.1f} cyc/sample  wait_in {m[r,0]/L:.1f}  wait_out {m[r,1]/L:.1f}  busy {(tot-m[r,0]-m[r,1])/L:.1f}")
rs = st[:, 3, 3].astype(np.int64)
print("TM rounds/CTA", (rs >> 32).mean(), "slow rounds/CTA", (rs & 0xffffffff).mean())
