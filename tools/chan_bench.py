"""Channelizer benchmark: python tools/chan_bench.py [C] [D] [seconds_per_chunk] [precision 0|1]
Device-resident wideband chunk -> [C][M] complex64; reports the GEMM kernel time (CUDA events on the stream),
achieved tensor TFLOP/s (2 * M * 2C * 2Kp flops) and the output write rate (8 B per output sample)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D = int(sys.argv[2]) if len(sys.argv) > 2 else 48
sec = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
M = int(48000 * sec)
n_in = M * D
rng = np.random.default_rng(0)
freqs = rng.uniform(-0.45, 0.45, C) * 48000.0 * D
x = torch.randn((n_in, 2), device="cuda", dtype=torch.float32) * 0.1
ch = capi.Channelizer(freqs, D, n_in, precision=prec)
st = torch.cuda.Stream()
ms = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    ch.process_c64_device(x.data_ptr(), n_in, stream=st.cuda_stream)
st.synchronize()
with torch.cuda.stream(st):
    e0.record()
N = 10
for i in range(N):
    ch.process_c64_device(x.data_ptr(), n_in, stream=st.cuda_stream)
    st.synchronize()
    ms.append(ch.last_kernel_ms())
with torch.cuda.stream(st):
    e1.record()
st.synchronize()
Kp = (ch.K + 31) // 32 * 32
Np = (2 * C + 255) // 256 * 256
gemm = float(np.mean(ms))
flops = (3 if prec else 1) * 2.0 * (M + 127) // 128 * 128 * Np * 2 * Kp
print(json.dumps({"precision": prec, "channels": C, "decim": D, "taps": ch.K, "out_samples_per_channel": M, "gemm_ms": gemm,
                  "call_ms": e0.elapsed_time(e1) / N, "tensor_tflops": flops / (gemm * 1e-3) / 1e12,
                  "useful_tflops": 2.0 * M * 2 * C * 2 * ch.K / (gemm * 1e-3) / 1e12,
                  "out_gbs": 8.0 * C * M / (gemm * 1e-3) / 1e9, "out_msamples_s": C * M / (gemm * 1e-3) / 1e6}))
ch.close()
