"""Kernel durations of arbitrary channel mixes running concurrently (CUPTI trace through torch.profiler).

    python tools/mixprobe.py [--split] type:channels [type:channels ...]
One handle decodes the whole mix (its variants forked on streams); with --split every group gets its own handle and
stream.  Prints start / duration / grid of each kernel of the last two steps."""
import sys, json, os, tempfile
import numpy as np, torch
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi
import bench

args = sys.argv[1:]
split = "--split" in args
groups = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in args if ":" in a]
L, NCH = 48000, 3
base = {}
for t, _ in groups:
    if t not in base:
        base[t] = torch.from_numpy(bench.gen_batch(np.full(16, t, np.int32), 1000 * t, 16, L * NCH, 16)).cuda()
decs, iqs = [], []
def make(tlist):
    C = len(tlist)
    iq = torch.empty((NCH, C, L), dtype=torch.complex64, device="cuda")
    for c, t in enumerate(tlist):
        iq[:, c, :] = base[t][c % 16].view(NCH, L)
    return capi.BatchDecoder(np.array(tlist, np.int32), L), iq
if split:
    for t, n in groups:
        d, iq = make([t] * n)
        decs.append(d); iqs.append(iq)
else:
    tl = []
    for t, n in groups:
        tl += [t] * n
    d, iq = make(tl)
    decs.append(d); iqs.append(iq)
for i in range(3):
    for d, iq in zip(decs, iqs):
        d.process_iq_device(iq[i % NCH].data_ptr(), L)
for d in decs:
    d.join(); d.sync()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        for d, iq in zip(decs, iqs):
            d.process_iq_device(iq[i % NCH].data_ptr(), L)
    for d in decs:
        d.join(); d.sync()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "mix.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev[len(ev) * 1 // 3:]:
    a = e.get("args", {})
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:8.1f}  stream {a.get('stream')}  grid {a.get('grid')}  {e['name'][38:100]}")
