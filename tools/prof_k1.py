"""Drive K1 a few times on the bench workload (for ncu / nsys captures)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi
import bench
stype = int(sys.argv[1]) if len(sys.argv) > 1 else 0
C, L = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024), 48000
iq = bench.gen_batch(stype, 0, C, L, 16)
d = torch.from_numpy(iq).cuda()
dec = capi.BatchDecoder(np.full(C, stype, np.int32), L)
for _ in range(4):
    dec.process_iq_device(d.data_ptr(), L); dec.sync()
print("demod ms", dec.last_kernel_ms())
