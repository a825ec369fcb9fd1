"""One bench-sized RS41 step for an ncu capture of demod_pipe_kernel: python tools/ncu_rs41.py [channels]"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi, synth
import bench
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = 48000
base = torch.from_numpy(bench.gen_batch(synth.RS41, 0, 16, L * 3, 16)).cuda()
iq = base.repeat(C // 16, 1).contiguous()
dec = capi.BatchDecoder(np.full(C, synth.RS41, np.int32), L)
for k in range(3):
    chunk = iq[:, k * L:(k + 1) * L].contiguous()
    torch.cuda.synchronize()
    dec.process_iq_device(chunk.data_ptr(), L)
    dec.sync()
print("kernel ms", dec.last_kernel_ms())
dec.close()
