"""2+ ranks: where does the time go when the scatter of chunk i+1 runs beside the decode of chunk i?"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from sdrpp_radiosonde_b200 import capi, shard, synth
import bench
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_MAX_CTAS", sys.argv[1] if len(sys.argv) > 1 else "16")
C, L = 1024, 48000
host = bench.gen_batch(synth.RS41, 0, 16, L, 8)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
base = torch.from_numpy(host).cuda().repeat(C // 16, 1)
full = base.repeat(world, 1) if rank == 0 else None
bufs = [torch.empty((C, L), dtype=torch.complex64, device="cuda") for _ in range(2)]
dec = capi.BatchDecoder(np.full(C, synth.RS41, np.int32), L, device=lr)
shard.scatter_channels(full, bufs[0], world, rank); shard.scatter_channels(full, bufs[1], world, rank)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
side = torch.cuda.Stream()
for mode in ("decode_only", "scatter_only", "both_default_stream", "both_side_stream", "decode_first"):
    ts = []
    for i in range(6):
        dec.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        works = []
        if mode == "decode_first":
            dec.process_iq_device(bufs[i % 2].data_ptr(), L)
        if mode in ("scatter_only", "both_default_stream", "decode_first"):
            works = shard.scatter_channels(full, bufs[(i + 1) % 2], world, rank, async_op=True)
        elif mode == "both_side_stream":
            with torch.cuda.stream(side):
                works = shard.scatter_channels(full, bufs[(i + 1) % 2], world, rank, async_op=True)
        t1 = time.perf_counter()
        if mode in ("decode_only", "both_default_stream", "both_side_stream"):
            dec.process_iq_device(bufs[i % 2].data_ptr(), L)
        t2 = time.perf_counter()
        if mode == "both_side_stream":
            with torch.cuda.stream(side):
                for w in works: w.wait()
            side.synchronize()
        else:
            for w in works: w.wait()
            torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        dec.sync()
        t4 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
    m = np.array(ts[2:]).mean(axis=0) * 1e3
    print(f"rank {rank} {mode:22s} issue_scatter {m[0]:.3f} issue_decode {m[1]:.3f} wait_scatter {m[2]:.3f} wait_decode {m[3]:.3f} total {m[4]:.3f} ms", flush=True)
dec.close(); dist.destroy_process_group()
