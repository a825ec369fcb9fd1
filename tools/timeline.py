"""Kernel timeline of a few pipelined steps (CUPTI activity trace through torch.profiler; kernels keep their concurrency).

    python tools/timeline.py <config 2|3|4|5> [channels] [steps]
Prints start / duration / stream / name of every kernel of the last steps, relative to the first one shown.
"""
import sys, json, os, tempfile
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from sdrpp_radiosonde_b200 import capi
import bench

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONFIGS[cfg]["channels"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
L, NCH = 48000, 4
types = bench.config_types(cfg, C)
host = bench.gen_batch(types, 0, C, L * NCH, os.cpu_count() or 1)
iq = torch.empty((NCH, C, L), dtype=torch.complex64, device="cuda")
for k in range(NCH):
    iq[k].copy_(torch.from_numpy(np.ascontiguousarray(host[:, k * L:(k + 1) * L])))
dec = capi.BatchDecoder(types, L)
for i in range(4):
    dec.process_iq_device(iq[i % NCH].data_ptr(), L)
dec.join(); dec.sync()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        dec.process_iq_device(iq[i % NCH].data_ptr(), L)
    dec.join(); dec.sync()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "tl.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    a = e.get("args", {})
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:8.1f}  stream {a.get('stream')}  grid {a.get('grid')}  {e['name'][:70]}")
