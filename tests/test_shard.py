"""CPU: channel sharding (world_size 2 over gloo) and the AUTO plan."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from sdrpp_radiosonde_b200 import shard, synth


def test_shard_ranges_partition_everything():
    for C in (0, 1, 7, 1024, 8192, 1000):
        for W in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(C, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard.shard_range(8192, 8, 3) == (3072, 4096)          # config 5: 1024 channels per GPU
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def _worker(rank, world, port, C, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    types = [c % 7 for c in range(C)]
    mine = shard.shard_types(types, world, rank)
    lo, hi = shard.shard_range(C, world, rank)
    assert list(mine) == types[lo:hi]
    # every rank "decodes" its own shard: count = global channel index * 10 + type
    local = np.array([c * 10 + types[c] for c in range(lo, hi)], dtype=np.int64)
    allc = shard.gather_counts(local, C, world, rank)
    # the batch originates on rank 0: scatter [C][L] complex64, every rank must end up with its own rows
    import torch
    L = 5
    full = (torch.arange(C * L, dtype=torch.float32).reshape(C, L) * (1 + 2j)).to(torch.complex64) if rank == 0 else None
    mine_iq = torch.zeros((hi - lo, L), dtype=torch.complex64)
    shard.scatter_channels(full, mine_iq, world, rank)
    want_iq = (torch.arange(C * L, dtype=torch.float32).reshape(C, L) * (1 + 2j)).to(torch.complex64)[lo:hi]
    assert torch.equal(mine_iq, want_iq)
    works = shard.scatter_channels(full, mine_iq.zero_(), world, rank, async_op=True)
    for w in works:
        w.wait()
    assert torch.equal(mine_iq, want_iq)
    # frame records of every rank in global channel order with one all_gather
    from sdrpp_radiosonde_b200 import capi
    mf = 3
    recs = np.zeros((hi - lo, mf), dtype=capi.REC_DTYPE)
    cnt = np.array([(c % mf) + 1 for c in range(lo, hi)], dtype=np.int32)
    for i, c in enumerate(range(lo, hi)):
        for k in range(cnt[i]):
            recs[i, k]["chunk"] = 1000 * c + k
            recs[i, k]["ok"] = (c + k) & 1
            recs[i, k]["data"][:4] = [c, k, 0xB2, 0x00]
    allr, allcnt = shard.gather_records(recs, cnt, C, world, rank)
    assert allr.shape == (C, mf) and allcnt.tolist() == [(c % mf) + 1 for c in range(C)]
    for c in range(C):
        for k in range(allcnt[c]):
            assert int(allr[c, k]["chunk"]) == 1000 * c + k and int(allr[c, k]["ok"]) == ((c + k) & 1)
            assert allr[c, k]["data"][:4].tolist() == [c, k, 0xB2, 0x00]
    q.put((rank, allc.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    C, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, C, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [c * 10 + c % 7 for c in range(C)]
    for _, got in results:
        assert got == want


def test_auto_plan_locks_in_reference_order():
    plan = shard.AutoPlan([synth.RS41, shard.AUTO, synth.DFM09, shard.AUTO])
    assert plan.virtual_types == [0, 0, 2, 3, 1, 5, 6, 4, 1, 0, 2, 3, 1, 5, 6, 4]
    batch = np.arange(4)[:, None] * np.ones((1, 5))
    assert plan.expand(batch)[:, 0].tolist() == [0] + [1] * 7 + [2] + [3] * 7
    assert plan.active_slot(1) is None
    ok = np.zeros(16, dtype=int)
    ok[3] = 1      # ims100 of channel 1
    ok[4] = 2      # dfm09 of channel 1 — ims100 comes first in the reference's order
    ok[15] = 1     # mrzn1 of channel 3
    plan.update(ok)
    assert plan.locked == [0, synth.IMS100, 1, synth.MRZN1]
    assert plan.active_slot(1) == 3 and plan.active_slot(3) == 15 and plan.active_slot(0) == 0
    ok[:] = 0
    ok[1] = 5      # a later rs41 hit does not move a locked channel
    plan.update(ok)
    assert plan.locked[1] == synth.IMS100
