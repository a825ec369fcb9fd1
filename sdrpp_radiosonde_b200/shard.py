"""Channel sharding across the GPUs of one box, and the AUTO (try-all-decoders) policy.

Channels share nothing (one decoder object per channel in the reference: src/main.hpp:36-42,
SD/decode.c:24-30), so the only multi-GPU step is the partition itself: rank r of W owns the contiguous
block [r*C/W, (r+1)*C/W).  No collective sits on the data path; frame counts are
gathered with one small all_gather for reporting.  When the whole channel batch originates on one rank
(one SDR front end feeding the box), scatter_channels() hands every rank its block with point-to-point
sends (NCCL over NVLink on the GPU box) — the only exchange step the path has.
"""
from __future__ import annotations

import numpy as np

# reference AUTO order: SD/decode.c:174-224
AUTO_ORDER = (0, 2, 3, 1, 5, 6, 4)      # rs41, m10, ims100, dfm09, imet4, c50, mrzn1
AUTO = -1


def shard_range(n_channels: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of channels owned by `rank` (balanced to within one channel)."""
    if not (0 <= rank < world) or n_channels < 0:
        raise ValueError((n_channels, world, rank))
    return n_channels * rank // world, n_channels * (rank + 1) // world


def shard_types(types, world: int, rank: int) -> np.ndarray:
    lo, hi = shard_range(len(types), world, rank)
    return np.asarray(types[lo:hi], dtype=np.int32)


def gather_counts(local_counts: np.ndarray, n_channels: int, world: int, rank: int) -> np.ndarray:
    """All ranks' per-channel counters in global channel order (torch.distributed, any backend)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return np.asarray(local_counts)
    sizes = [shard_range(n_channels, world, r)[1] - shard_range(n_channels, world, r)[0] for r in range(world)]
    cap = max(sizes)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.zeros(cap, dtype=torch.int64, device=dev)
    mine[: len(local_counts)] = torch.as_tensor(np.asarray(local_counts, dtype=np.int64), device=dev)
    out = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, mine)
    return np.concatenate([o[: sizes[r]].cpu().numpy() for r, o in enumerate(out)])


def gather_records(local_recs: np.ndarray, local_counts: np.ndarray, n_channels: int, world: int, rank: int):
    """All ranks' frame records in global channel order with ONE all_gather (the output side of the
    multi-GPU path; records are fixed-size, a few MB per rank).  local_recs: [C_rank][max_frames] structured array
    (capi.REC_DTYPE), local_counts: [C_rank].  Returns (recs[C][max_frames], counts[C]) on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_recs, np.asarray(local_counts)
    sizes = [shard_range(n_channels, world, r)[1] - shard_range(n_channels, world, r)[0] for r in range(world)]
    cap, mf, isz = max(sizes), local_recs.shape[1], local_recs.dtype.itemsize
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    # one message per rank: [cap] int32 counts (as bytes) followed by [cap][max_frames] records
    msg = np.zeros(cap * 4 + cap * mf * isz, dtype=np.uint8)
    msg[: 4 * len(local_counts)] = np.asarray(local_counts, dtype=np.int32).view(np.uint8)
    msg[cap * 4: cap * 4 + local_recs.size * isz] = np.ascontiguousarray(local_recs).view(np.uint8).ravel()
    mine = torch.from_numpy(msg).to(dev)
    out = torch.empty(world * msg.size, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine)
    buf = out.cpu().numpy().reshape(world, msg.size)
    counts = np.concatenate([buf[r, : 4 * sizes[r]].view(np.int32) for r in range(world)])
    recs = np.concatenate([buf[r, cap * 4: cap * 4 + sizes[r] * mf * isz].view(local_recs.dtype).reshape(sizes[r], mf)
                           for r in range(world)])
    return recs, counts


def scatter_channels(full, out, world: int, rank: int, src: int = 0, async_op: bool = False):
    """Scatter the channel batch `full[C][L]` held by rank `src` so that every rank gets its shard_range()
    block in `out[(hi-lo)][L]` (torch tensors on the backend's device; complex64 is sent as float pairs).

    Blocks are contiguous row ranges, so each transfer is one send of a view, no staging copy; shards may
    differ by one channel, hence point-to-point sends rather than dist.scatter (equal sizes only).
    Returns the list of outstanding work handles when async_op is set (wait on them before using `out`),
    which lets the caller overlap the scatter of chunk i+1 with the decode of chunk i."""
    import torch
    import torch.distributed as dist

    def as_real(t):
        return torch.view_as_real(t) if t.is_complex() else t

    if world == 1:
        out.copy_(full)
        return []
    works = []
    if rank == src:
        C = full.shape[0]
        for r in range(world):
            lo, hi = shard_range(C, world, r)
            if r == src:
                out.copy_(full[lo:hi])
            elif hi > lo:
                works.append(dist.isend(as_real(full[lo:hi]), dst=r))
    elif out.shape[0] > 0:
        works.append(dist.irecv(as_real(out), src=src))
    if async_op:
        return works
    for w in works:
        w.wait()
    return []


class AutoPlan:
    """AUTO channels are expanded into one virtual channel per decoder type (all seven chains consume every
    sample until lock, SD/decode.c:174-224); the first decoder, in the reference's order, that yields a
    decodable frame wins and the channel is locked to it."""

    def __init__(self, types):
        self.types = [int(t) for t in types]
        self.virtual_types: list[int] = []
        self.source: list[int] = []           # virtual channel -> input channel
        self.slots: list[list[int]] = []      # input channel -> virtual channels (AUTO_ORDER for AUTO channels)
        for c, t in enumerate(self.types):
            if t == AUTO:
                base = len(self.virtual_types)
                self.virtual_types += list(AUTO_ORDER)
                self.source += [c] * len(AUTO_ORDER)
                self.slots.append(list(range(base, base + len(AUTO_ORDER))))
            else:
                self.slots.append([len(self.virtual_types)])
                self.virtual_types.append(t)
                self.source.append(c)
        self.locked = [None if t == AUTO else t for t in self.types]

    def expand(self, batch: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(batch[self.source])

    def update(self, ok_counts) -> None:
        """ok_counts[v] = decodable frames of virtual channel v in the last buffer."""
        for c, t in enumerate(self.types):
            if t != AUTO or self.locked[c] is not None:
                continue
            for v in self.slots[c]:
                if ok_counts[v] > 0:
                    self.locked[c] = self.virtual_types[v]
                    break

    def active_slot(self, c: int):
        """Virtual channel whose records are reported for input channel c (None while an AUTO channel is unlocked)."""
        if self.types[c] != AUTO:
            return self.slots[c][0]
        if self.locked[c] is None:
            return None
        return self.slots[c][AUTO_ORDER.index(self.locked[c])]
