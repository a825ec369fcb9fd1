#!/usr/bin/env python3
"""bench.py — IQ Msamples/s and decoded frames/s of the demod + sync + FEC hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's own CPU path, host cores)

    python bench.py --config {2,3,4,5} ...                    (the other BASELINE configs, same JSON contract)

Workload (default --config 2 = BASELINE.json configs[1], per GPU): 1024 RS41 channels, 48 kS/s synthetic GFSK complex64 IQ with
per-channel CFO / timing offset / clock error / AWGN, 10 s per channel, processed in
chunks of L = 48000 samples (1 s).  One "step" = one process call = one chunk of every channel
(49.15 M samples, 393 MB of IQ > L2, so every step streams from HBM; steps cycle through the 10 chunks).

Prints ONE JSON line:  value = whole-job IQ Msamples/s with inputs resident in HBM (CUDA events on the
library's stream, max over ranks); e2e = the same metric through the C ABI with pinned HOST buffers
(H2D of the chunk + D2H of the frame records inside the timed region); roofline for the demod kernel
(8 B per complex sample) against MEASURED_PEAKS.json; cpu_baseline = the compiled
reference (oracle/_ref, + the restated discriminator) on all host cores over a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sdrpp_radiosonde_b200 import synth  # noqa: E402

FS = 48000

# BASELINE.json configs[1..4] (1-based 2..5): sonde mix, channels per GPU, what the workload string says
CONFIGS = {
    2: dict(mix=[synth.RS41], channels=1024, seconds=10, auto=False,
            name="RS41 4800-baud GFSK + RS(255,231)", ref="BASELINE configs[1]"),
    3: dict(mix=[synth.DFM09, synth.M10], channels=2048, seconds=4, auto=False,
            name="DFM06/09/17 Manchester-FSK + M10/M20 mixed (alternating channels)", ref="BASELINE configs[2]"),
    4: dict(mix=[synth.IMS100], channels=1024, seconds=4, auto=False,
            name="iMS-100/RS-11G Manchester-FSK + 12 x BCH(63,51)", ref="BASELINE configs[3], per-GPU share of 4096 over 4 GPUs"),
    5: dict(mix=[synth.RS41, synth.DFM09, synth.M10, synth.IMS100, synth.MRZN1, synth.IMET4, synth.C50], channels=1024,
            seconds=3, auto=True, name="all seven decoders (9 sonde types), every channel AUTO (preamble autodetect)",
            ref="BASELINE configs[4], per-GPU share of 8192 over 8 GPUs"),
}


def config_types(cfg, n_ch, ch0=0):
    mix = CONFIGS[cfg]["mix"]
    return np.array([mix[(ch0 + c) % len(mix)] for c in range(n_ch)], dtype=np.int32)


METRIC = "IQ Msamples/s (decoded frames/s in config) across N channels vs reference CPU"
UNIT = "Msamples/s"


def _gen_channel(args):
    stype, ch, n = args
    return synth.make_iq(synth.default_spec(stype, ch), n)


def gen_batch(stype, ch0, n_ch, n, procs):
    """[n_ch][n] complex64, channel c seeded 0xB200 + ch0 + c; stype: one type or one per channel."""
    import multiprocessing as mp
    tl = [int(stype)] * n_ch if np.isscalar(stype) else [int(t) for t in stype]
    jobs = [(tl[c], ch0 + c, n) for c in range(n_ch)]
    if procs <= 1 or n_ch < 4:
        rows = [_gen_channel(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            rows = pool.map(_gen_channel, jobs, chunksize=max(1, n_ch // (procs * 4)))
    return np.stack(rows)


# ------------------------------------------------------------------------------------------
# CPU legs (the only place besides tests/ and smoke() that touches oracle/)
# ------------------------------------------------------------------------------------------
class CpuPath:
    """The reference's own CPU implementation of the path: restated discriminator (oracle/) +
    the UNMODIFIED xxx_decode() loop compiled into oracle/_ref; falls back to the C port."""

    def __init__(self):
        from tests import reflib
        self.orc = ctypes.CDLL(reflib.ORACLE_SO) if reflib.have_oracle() else None
        self.ref = ctypes.CDLL(reflib.REF_SO) if reflib.have_ref() else None
        if self.orc is None:
            raise RuntimeError("oracle/_build/libsonde_oracle.so missing: run __graft_entry__.build()")
        self.kind = "reference" if self.ref is not None else "port"
        fp = ctypes.POINTER(ctypes.c_float)
        self.orc.orc_discriminate.argtypes = [fp, ctypes.c_size_t, ctypes.c_float, fp, fp]
        self.orc.orc_discriminate.restype = None
        self.orc.orc_batch_run.restype = ctypes.c_int
        if self.ref is not None:
            self.ref.ref_decode_count.argtypes = [ctypes.c_int, ctypes.c_int, fp, ctypes.c_size_t, ctypes.c_size_t,
                                                  ctypes.POINTER(ctypes.c_int)]
            self.ref.ref_decode_count.restype = ctypes.c_int

    def run(self, stype, iq, chunk, threads):
        """iq [C][n] complex64 -> (seconds, framer windows, frames with fields != 0 / passing the gate).
        stype: one decoder type or one per channel."""
        C, n = iq.shape
        fp = ctypes.POINTER(ctypes.c_float)
        flat = np.ascontiguousarray(iq).view(np.float32).reshape(C, 2 * n)
        tl = np.full(C, stype, dtype=np.int32) if np.isscalar(stype) else np.asarray(stype, dtype=np.int32)
        if self.ref is None:
            types = tl
            frames = np.zeros(C, dtype=np.int32)
            ok = np.zeros(C, dtype=np.int32)
            t0 = time.perf_counter()
            self.orc.orc_batch_run(types.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), C, FS,
                                   flat.ctypes.data_as(fp), 1, ctypes.c_size_t(n), ctypes.c_size_t(chunk),
                                   ctypes.c_float(0.0), threads,
                                   frames.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   ok.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
            return time.perf_counter() - t0, int(frames.sum()), int(ok.sum())

        frames = [0] * C
        ok = [0] * C

        def work(lo, hi):
            fm = np.empty(n, dtype=np.float32)
            for c in range(lo, hi):
                prev = ctypes.c_float(0.0)
                # the discriminator is stateless across chunks apart from prev: one pass over the row
                self.orc.orc_discriminate(flat[c].ctypes.data_as(fp), n, ctypes.c_float(0.636619747),
                                          ctypes.byref(prev), fm.ctypes.data_as(fp))
                nf = ctypes.c_int(0)
                frames[c] = self.ref.ref_decode_count(int(tl[c]), FS, fm.ctypes.data_as(fp), n, chunk, ctypes.byref(nf))
                ok[c] = nf.value

        ths = [threading.Thread(target=work, args=(C * i // threads, C * (i + 1) // threads)) for i in range(threads)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0, sum(frames), sum(ok)


def cpu_sample(cfg, n_ch, n, procs):
    return gen_batch(config_types(cfg, n_ch), 0, n_ch, n, procs)


# ------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                out.append(f)
    finally:
        p.terminate()


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max((int(s[1]) for s in samples if s[1].isdigit()), default=None)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(samples)}


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 (e.g. NCCL's version
    banner) has been rerouted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE config (1-based); 2 = the headline")
    ap.add_argument("--channels", type=int, default=0, help="channels per GPU (default: the config's)")
    ap.add_argument("--chunk", type=int, default=48000)
    ap.add_argument("--seconds", type=int, default=0, help="seconds of signal per channel (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scatter", action="store_true", help="N > 1: skip the rank-0 -> all scatter measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    numa_note = None
    if world > 1 and args.impl == "b200":
        # one rank per GPU: run (and therefore first-touch / pin host buffers) on the CPU cores local to this rank's GPU,
        # otherwise all ranks' pinned staging may sit on one socket and the e2e copies share its memory controllers
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = (ncores + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
            cpus = {64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1 and 64 * w + b < ncores}
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = f"rank pinned to {len(cpus)} GPU-local cores"
        except Exception as e:                               # affinity is an optimisation, never a requirement
            numa_note = f"no GPU-local affinity ({type(e).__name__})"
    cfg = CONFIGS[args.config]
    L, C = args.chunk, args.channels or cfg["channels"]
    n_chunks = max(1, (args.seconds or cfg["seconds"]) * FS // L)
    n_total = n_chunks * L
    # `config` is the static description of the workload, identical in both arms; measured by-products go to `result`
    config = {"workload": f"{cfg['name']}, {C} channels/GPU, 48 kS/s complex64 IQ, chunk L={L}, "
                          f"{n_chunks} distinct chunks/channel ({cfg['ref']})",
              "baseline_config": args.config, "channels_per_gpu": C, "chunk_len": L,
              "cache": f"each step reads a different {C * L * 8 / 1e6:.0f} MB chunk (> 126 MB L2)"}

    # ---------------------------------------------------------------- reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return
        cpu = CpuPath()
        n_ch = 2 * ncores
        n_ref = min(n_total, 4 * L) if args.config != 2 else n_total
        ctypes_ = config_types(args.config, n_ch)
        sample = cpu_sample(args.config, n_ch, n_ref, ncores)
        for _ in range(min(args.warmup, 1)):
            cpu.run(ctypes_[:ncores], sample[:ncores, :L * 2], L, ncores)
        t = 0.0
        frames = ok = 0
        for _ in range(args.steps):
            dt, f, k = cpu.run(ctypes_, sample, L, ncores)
            t += dt
            frames += f
            ok += k
        samples = args.steps * n_ch * n_ref
        val = samples / t / 1e6
        desc = (f"a bounded sample of the workload: its first {n_ch} channels x {n_ref} samples per step on {ncores} host "
                f"threads (the CPU path is one independent decoder per channel, so samples/s does not depend on the "
                f"channel count); discriminator = this repo's restated polynomial atan2 (SDR++ is not vendored), "
                f"everything after it the unmodified reference")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "result": {"frames_per_s": frames / t, "ok_frames_per_s": ok / t, "channels_run": n_ch,
                           "samples_per_channel": n_ref},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": cpu.kind, "sample": desc},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ---------------------------------------------------------------- B200 arm
    procs = max(1, ncores // max(1, world))
    t_gen = time.time()
    sig_types = config_types(args.config, C, rank * C)
    host_iq = gen_batch(sig_types, rank * C, C, n_total, procs)           # before CUDA init (fork pool)
    t_gen = time.time() - t_gen

    import torch
    import torch.distributed as dist
    from sdrpp_radiosonde_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL only carries barriers, the timing / count reductions and (shard.gather_records) the record all-gather here; the
        # single-source legs below move their data with the copy engines.  Small CTA budget and no clusters, so that a
        # collective never has to wait for SMs the decode kernel (one large CTA on 147 of the 148 SMs) holds.
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        os.environ.setdefault("NCCL_CGA_CLUSTER_SIZE", "1")
        os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # [n_chunks][C][L] complex64 in HBM
    dev_iq = torch.empty((n_chunks, C, L), dtype=torch.complex64, device="cuda")
    for k in range(n_chunks):
        dev_iq[k].copy_(torch.from_numpy(np.ascontiguousarray(host_iq[:, k * L:(k + 1) * L])))
    torch.cuda.synchronize()

    # config 5: every channel is created as AUTO (SD/decode.c:174-224) and locks at the first fetch after a frame
    types = np.full(C, -1, dtype=np.int32) if cfg["auto"] else sig_types
    dec = capi.BatchDecoder(types, L, device=local_rank)
    ext = torch.cuda.ExternalStream(dec.stream)
    auto_info = None
    if cfg["auto"]:
        # acquisition: buffers with all seven decoders on every channel until the channels have locked
        acq = []
        for i in range(min(n_chunks, 3)):
            dec.process_iq_device(dev_iq[i % n_chunks].data_ptr(), L)
            dec.fetch_counts()
            acq.append((float(dec.last_kernel_ms()[0]), int((dec.detected_types() >= 0).sum())))
        auto_info = {"acquisition_buffers": [{"demod_kernel_ms": a, "channels_locked_after": b} for a, b in acq],
                     "channels": C, "note": "the timed steps below are the locked steady state"}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        dec.process_iq_device(dev_iq[i % n_chunks].data_ptr(), L)

    for i in range(args.warmup):
        step(i)
    dec.sync()
    f0, k0, _ = dec.fetch_totals()
    launches0 = dec.launch_count
    dec.join()

    # clocks during the timed region
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ext):
        e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    dec.join()                      # the framer kernels run on an internal stream: fold them into the timed stream
    with torch.cuda.stream(ext):
        e1.record()
    dec.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = dec.launch_count - launches0
    f1, k1, _ = dec.fetch_totals()
    frames, ok = int((f1 - f0).sum()), int((k1 - k0).sum())

    # demod kernel duration for the roofline (per-launch CUDA events inside the library)
    dem_ms, frm_ms = [], []
    for i in range(min(args.steps, 10)):
        step(args.warmup + args.steps + i)
        a, b = dec.last_kernel_ms()
        dem_ms.append(a)
        frm_ms.append(b)
    dem = float(np.mean(dem_ms))

    # e2e: pinned host chunk -> process_iq() -> fetch() of the frame records, every step
    e2e = None
    if not args.no_e2e:
        nbuf = min(n_chunks, 3)
        pins = [capi.PinnedBuffer((C, L), np.complex64) for _ in range(nbuf)]
        for k, pb in enumerate(pins):
            pb.array[...] = host_iq[:, k * L:(k + 1) * L]
        dec2 = capi.BatchDecoder(sig_types, L, device=local_rank)
        for i in range(2):
            dec2.process_host_ptr(pins[i % nbuf].ptr, L, is_iq=True)
            dec2.fetch()
        barrier()
        t0 = time.perf_counter()
        e2e_ok = 0
        # software pipeline of depth 2 (include/sonde_b200.h, fetch()): the H2D copy of step i+1 and the D2H of
        # step i's records overlap the kernels; every step's input still crosses PCIe inside the timed region
        dec2.process_host_ptr(pins[0].ptr, L, is_iq=True)
        for i in range(args.steps):
            if i + 1 < args.steps:
                dec2.process_host_ptr(pins[(i + 1) % nbuf].ptr, L, is_iq=True)
            recs, counts = dec2.fetch()
            e2e_ok += int(counts.sum())
        torch.cuda.synchronize()
        dec2.sync()
        t_e2e = time.perf_counter() - t0
        d2h = C * dec2.max_frames * capi.REC_DTYPE.itemsize + C * 8
        e2e = {"t": t_e2e, "h2d": C * L * 8, "d2h": d2h}
        # what the link itself delivers: the same pinned buffer copied to the device with nothing else running
        dst = torch.empty((C, L), dtype=torch.complex64, device="cuda")
        src_t = torch.from_numpy(pins[0].array)
        for _ in range(2):
            dst.copy_(src_t, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            dst.copy_(src_t, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        e2e["link_gbs"] = 5 * C * L * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        if world > 1:
            # what the HOST can deliver when every rank copies at once: all ranks start the same bare copy together, from
            # ordinary pinned memory and from write-combined pinned memory (no snoop traffic); gathered per rank below
            from cuda.bindings import runtime as cudart_
            cur_stream = torch.cuda.current_stream().cuda_stream

            def all_at_once(host_ptr):
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(5):
                    cudart_.cudaMemcpyAsync(dst.data_ptr(), host_ptr, C * L * 8, cudart_.cudaMemcpyKind.cudaMemcpyHostToDevice, cur_stream)
                a1.record()
                torch.cuda.synchronize()
                return 5 * C * L * 8 / (a0.elapsed_time(a1) * 1e-3) / 1e9
            try:
                e2e["link_all"] = all_at_once(pins[0].ptr)
            except Exception as ex:                      # informational probe: never lose the main line over it
                print(f"all-ranks link probe failed: {type(ex).__name__}: {ex}", file=sys.stderr)
                e2e["link_all"] = 0.0
            e2e["link_all_wc"] = 0.0
            try:
                err, wc_ptr = cudart_.cudaHostAlloc(C * L * 8, cudart_.cudaHostAllocWriteCombined)
                if int(err) == 0:
                    e2e["link_all_wc"] = all_at_once(int(wc_ptr))
                    cudart_.cudaFreeHost(wc_ptr)
                else:
                    barrier()
            except Exception as ex:
                print(f"write-combined link probe failed: {type(ex).__name__}: {ex}", file=sys.stderr)
        del dst
        dec2.close()
        for pb in pins:
            pb.free()
        try:
            # the same loop through the int16 entry point (half the PCIe bytes; informational, `e2e` stays complex64)
            pins16 = [capi.PinnedBuffer((C, L, 2), np.int16) for _ in range(nbuf)]
            for k, pb in enumerate(pins16):
                blk = host_iq[:, k * L:(k + 1) * L]
                pb.array[..., 0] = np.clip(np.round(blk.real * 16384.0), -32768, 32767)
                pb.array[..., 1] = np.clip(np.round(blk.imag * 16384.0), -32768, 32767)
            dec4 = capi.BatchDecoder(sig_types, L, device=local_rank)
            for i in range(2):
                dec4.process_s16_host_ptr(pins16[i % nbuf].ptr, L)
                dec4.fetch()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ok16 = 0
            dec4.process_s16_host_ptr(pins16[0].ptr, L)
            for i in range(args.steps):
                if i + 1 < args.steps:
                    dec4.process_s16_host_ptr(pins16[(i + 1) % nbuf].ptr, L)
                recs, counts = dec4.fetch()
                ok16 += int(sum(int(recs[c, j]["ok"]) for c in range(0, C, 64) for j in range(counts[c])))
            torch.cuda.synchronize()
            dec4.sync()
            e2e["t16"] = time.perf_counter() - t0
            e2e["ok16"] = ok16
            dec4.close()
            for pb in pins16:
                pb.free()

        except Exception as ex:                      # informational extra: never lose the main line over it
            e2e["t16"], e2e["ok16"], e2e["err16"] = 0.0, 0, f"{type(ex).__name__}: {ex}"

    # wideband front end: one host buffer of wideband IQ per step -> tcgen05 channelizer -> the same
    # C-channel decode, chained on the decoder's stream.  H2D is D*48 kS/s * 8 B instead of C*48 kS/s * 8 B.
    wide = None
    if not args.no_e2e:
        try:
            Dw = 48
            n_in = L * Dw
            rngw = np.random.default_rng(7 + rank)
            freqs = rngw.uniform(-0.45, 0.45, C) * FS * Dw
            pinw = [capi.PinnedBuffer((n_in,), np.complex64) for _ in range(2)]
            for pb in pinw:
                pb.array[...] = (0.05 * (rngw.standard_normal(n_in) + 1j * rngw.standard_normal(n_in))).astype(np.complex64)
            chz = capi.Channelizer(freqs, Dw, n_in, device=local_rank)
            dec5 = capi.BatchDecoder(sig_types, L, device=local_rank)
            # the copy and the channelizer of step i+1 run on a side stream beside the decode of step i
            side, ext5 = torch.cuda.Stream(), torch.cuda.ExternalStream(dec5.stream)
            done = [None, None]
            def wstep(i):
                if done[i & 1] is not None:
                    side.wait_event(done[i & 1])          # decode(i-2) has read the output buffer this call overwrites
                ptr, stride, m = chz.process_c64_host_ptr(pinw[i % 2].ptr, n_in, stream=side.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(side)
                ext5.wait_event(ev)
                dec5.process_iq_device(ptr, m, stride)
                done[i & 1] = torch.cuda.Event()
                done[i & 1].record(ext5)
            for i in range(3):
                wstep(i)
                dec5.fetch_counts()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            wstep(0)
            for i in range(args.steps):
                if i + 1 < args.steps:
                    wstep(i + 1)
                dec5.fetch_counts()
            dec5.sync()
            torch.cuda.synchronize()
            wide = {"t": time.perf_counter() - t0, "gemm_ms": chz.last_kernel_ms(), "h2d": n_in * 8, "D": Dw, "taps": chz.K}
            chz.close()
            dec5.close()
            for pb in pinw:
                pb.free()
        except Exception as ex:                      # informational extra
            wide = {"error": f"{type(ex).__name__}: {ex}"}

    # N > 1: the one exchange step the path can have — the whole batch originates on rank 0 (one front
    # end feeding the box).  Two forms, both reported beside the main number (which has every shard resident):
    #  (a) pre-channelised complex64: every rank PULLS its [C][L] block out of rank 0's buffer with the copy engine
    #      (CUDA IPC + sonde_b200_process_iq_peer = cudaMemcpyPeerAsync: no SMs, so the pull of chunk i+1 runs beside
    #      the decode of chunk i); rank 0's NVLink egress is (N-1) * C * L * 8 bytes per step — the physical bound
    #  (b) wideband: rank 0 broadcasts the 18 MB wideband buffer (NCCL) and every rank channelises its own block
    scatter = None
    wide_src = None
    if world > 1 and not args.no_scatter:
        from cuda.bindings import runtime as cudart

        def ck(r):
            if isinstance(r, tuple):
                err, rest = r[0], r[1:]
            else:
                err, rest = r, ()
            if int(err) != 0:
                raise RuntimeError(f"cudart error {err}")
            return rest[0] if len(rest) == 1 else rest

        try:
            n_sc = 8
            blk = C * L * 8
            handles = [None, None]
            srcs = [0, 0]
            if rank == 0:
                for j in range(2):
                    srcs[j] = int(ck(cudart.cudaMalloc(world * blk)))
                    for r in range(world):
                        ck(cudart.cudaMemcpy(srcs[j] + r * blk, dev_iq[j % n_chunks].data_ptr(), blk,
                                             cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice))
                    handles[j] = bytes(ck(cudart.cudaIpcGetMemHandle(srcs[j])).reserved)
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                for j in range(2):
                    hnd = cudart.cudaIpcMemHandle_t()
                    hnd.reserved = handles[j]
                    srcs[j] = int(ck(cudart.cudaIpcOpenMemHandle(hnd, cudart.cudaIpcMemLazyEnablePeerAccess)))
            dec3 = capi.BatchDecoder(sig_types, L, device=local_rank)

            def pull(i):
                ptr = srcs[i % 2] + rank * blk
                if rank == 0:
                    dec3.process_iq_device(ptr, L)
                else:
                    dec3.process_iq_peer(0, ptr, L)
            for i in range(2):
                pull(i)
                dec3.fetch_counts()
            barrier()
            t0 = time.perf_counter()
            pull(0)
            for i in range(n_sc):
                if i + 1 < n_sc:
                    pull(i + 1)
                dec3.fetch_counts()
            dec3.sync()
            barrier()
            t_sc = time.perf_counter() - t0
            scatter = {"t": t_sc, "n": n_sc}
            dec3.close()
            if rank != 0:
                for j in range(2):
                    cudart.cudaIpcCloseMemHandle(srcs[j])
            barrier()
            if rank == 0:
                for j in range(2):
                    cudart.cudaFree(srcs[j])
        except Exception as ex:                      # informational extra
            scatter = None
            print(f"scatter measurement failed: {type(ex).__name__}: {ex}", file=sys.stderr)
        try:
            Dw = 48
            n_in = L * Dw
            rngw = np.random.default_rng(7)
            freqs = np.random.default_rng(11 + rank).uniform(-0.45, 0.45, C) * FS * Dw
            wsrc, whandles = [0, 0], [None, None]
            if rank == 0:
                for j in range(2):
                    wsrc[j] = int(ck(cudart.cudaMalloc(n_in * 8)))
                    hostw = (0.05 * (rngw.standard_normal(n_in) + 1j * rngw.standard_normal(n_in))).astype(np.complex64)
                    ck(cudart.cudaMemcpy(wsrc[j], hostw.ctypes.data, n_in * 8, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice))
                    whandles[j] = bytes(ck(cudart.cudaIpcGetMemHandle(wsrc[j])).reserved)
            dist.broadcast_object_list(whandles, src=0)
            if rank != 0:
                for j in range(2):
                    hnd = cudart.cudaIpcMemHandle_t()
                    hnd.reserved = whandles[j]
                    wsrc[j] = int(ck(cudart.cudaIpcOpenMemHandle(hnd, cudart.cudaIpcMemLazyEnablePeerAccess)))
            chz = capi.Channelizer(freqs, Dw, n_in, device=local_rank)
            dec6 = capi.BatchDecoder(sig_types, L, device=local_rank)
            ext6 = torch.cuda.ExternalStream(dec6.stream)
            cur = torch.cuda.Stream()
            done6 = [None, None]

            def wstep6(i):
                # pull (copy engine) + channelizer of step i+1 on a side stream beside the decode of step i
                if done6[i & 1] is not None:
                    cur.wait_event(done6[i & 1])
                if rank == 0:
                    ptr, stride, m = chz.process_c64_device(wsrc[i & 1], n_in, stream=cur.cuda_stream)
                else:
                    ptr, stride, m = chz.process_c64_peer(0, wsrc[i & 1], n_in, stream=cur.cuda_stream)
                ev = torch.cuda.Event()
                ev.record(cur)
                ext6.wait_event(ev)
                dec6.process_iq_device(ptr, m, stride)
                done6[i & 1] = torch.cuda.Event()
                done6[i & 1].record(ext6)
            for i in range(3):
                wstep6(i)
                dec6.fetch_counts()
            barrier()
            t0 = time.perf_counter()
            wstep6(0)
            for i in range(args.steps):
                if i + 1 < args.steps:
                    wstep6(i + 1)
                dec6.fetch_counts()
            dec6.sync()
            barrier()
            wide_src = {"t": time.perf_counter() - t0, "bytes": n_in * 8, "D": Dw}
            chz.close()
            dec6.close()
            if rank != 0:
                for j in range(2):
                    cudart.cudaIpcCloseMemHandle(wsrc[j])
            barrier()
            if rank == 0:
                for j in range(2):
                    cudart.cudaFree(wsrc[j])
        except Exception as ex:                      # informational extra
            wide_src = None
            print(f"wideband single-source measurement failed: {type(ex).__name__}: {ex}", file=sys.stderr)

    stop.set()
    th.join(timeout=2)

    t_all = torch.tensor([ms, (e2e["t"] * 1e3) if e2e else 0.0, scatter["t"] * 1e3 if scatter else 0.0,
                          wide_src["t"] * 1e3 if wide_src else 0.0, (e2e["t16"] * 1e3) if e2e else 0.0,
                          (wide["t"] * 1e3) if wide and "t" in wide else 0.0],
                         dtype=torch.float64, device="cuda")
    tot = torch.tensor([frames, ok], dtype=torch.float64, device="cuda")
    link_all = None
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if e2e and "link_all" in e2e:
            mine = torch.tensor([e2e["link_all"], e2e["link_all_wc"]], dtype=torch.float64, device="cuda")
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            link_all = [[round(float(v[0]), 1) for v in every], [round(float(v[1]), 1) for v in every]]
    ms_max, e2e_ms_max = float(t_all[0]), float(t_all[1])
    frames, ok = int(tot[0]), int(tot[1])

    if rank == 0:
        samples_total = world * args.steps * C * L
        value = samples_total / (ms_max * 1e-3) / 1e6
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = 8.0 * C * L / (dem * 1e-3) / 1e9
        traffic = None                       # dram bytes per launch from the committed ncu --set full capture
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "demod_pipe_ncu_summary.json")))
            if prof.get("channels") == C and prof.get("chunk_len") == L:
                traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "result": dict(frames_per_s=frames / (ms_max * 1e-3), ok_frames_per_s=ok / (ms_max * 1e-3),
                           realtime_channels_equiv=value * 1e6 / FS, gen_seconds=round(t_gen, 1),
                           **({"auto": auto_info} if auto_info else {}),
                           **({"host_affinity": numa_note} if numa_note else {})),
            "gpu_launches": launches,
            "clocks": summarize_clocks(samples),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "demod_pipe_kernel (+ demod_pipe_afsk_kernel for AFSK channels): all demod "
                         "kernels of one step, forked on streams", "kernel_ms": dem,
                         "frame_kernel_ms": float(np.mean(frm_ms)),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "note": "8 B per complex sample; the chain is serial per channel (AGC IIR + Gardner NCO), "
                                 "so at 1024 channels it is latency-bound, not HBM-bound"},
        }
        if e2e:
            line["e2e"] = {"value": world * args.steps * C * L / (e2e_ms_max * 1e-3) / 1e6, "unit": UNIT,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e_ms_max / args.steps,
                           "h2d_gbs_achieved": e2e["h2d"] / (e2e_ms_max / args.steps * 1e-3) / 1e9,
                           "h2d_link_gbs_measured": e2e["link_gbs"],
                           "note": "bounded by the host link: every step moves C*L*8 bytes of complex64 IQ over PCIe; "
                                   "h2d_link_gbs_measured is a bare pinned-memory copy of the same buffer on rank 0"}
            if link_all:
                line["e2e"]["h2d_link_gbs_all_ranks_at_once"] = {"pinned": link_all[0], "write_combined": link_all[1],
                    "note": "every rank runs the same bare H2D copy at the same time: what the host side of this box delivers "
                            "per GPU when all GPUs are fed at once (the e2e figure cannot exceed it)"}
            t16 = float(t_all[4])
            line["e2e_s16"] = {"error": e2e["err16"]} if e2e.get("err16") or t16 <= 0 else {"value": world * args.steps * C * L / (t16 * 1e-3) / 1e6, "unit": UNIT,
                               "h2d_bytes_per_step": C * L * 4, "ms_per_step": t16 / args.steps,
                               "note": "same loop through sonde_b200_process_iq_s16 (int16 IQ quantised from the same "
                                       "signals, converted on the GPU); informational — `e2e` is the complex64 entry point",
                               "decodable_frames_sampled": e2e["ok16"]}
        if wide and "error" in wide:
            line["e2e_wideband"] = wide
        elif wide:
            tw = float(t_all[5])
            line["e2e_wideband"] = {"value": world * args.steps * C * L / (tw * 1e-3) / 1e6, "unit": UNIT,
                                    "ms_per_step": tw / args.steps, "h2d_bytes_per_step": wide["h2d"],
                                    "channelizer_gemm_ms": wide["gemm_ms"],
                                    "note": f"host wideband complex64 IQ at {wide['D']} x 48 kS/s (noise) -> tcgen05 channelizer "
                                            f"({wide['taps']} taps, {C} channel centres) -> the same {C}-channel decode, per step; "
                                            "channel samples per second; informational (SURVEY §8 f-2)"}
        if scatter:
            sc_ms = float(t_all[2]) / scatter["n"]
            line["scatter"] = {"what": "batch resident on rank 0; every rank pulls its [C][L] complex64 block over NVLink with the "
                                       "copy engine (CUDA IPC + cudaMemcpyPeerAsync, sonde_b200_process_iq_peer), the pull of "
                                       "chunk i+1 beside the decode of chunk i",
                               "ms_per_step_with_scatter": sc_ms, "value_with_scatter": world * C * L / (sc_ms * 1e-3) / 1e6,
                               "rank0_egress_bytes_per_step": (world - 1) * C * L * 8,
                               "rank0_egress_gbs": (world - 1) * C * L * 8 / (sc_ms * 1e-3) / 1e9,
                               "note": "bounded by rank 0's NVLink egress: (N-1) x 393 MB per step"}
        if wide_src:
            tw6 = float(t_all[3])
            line["single_source_wideband"] = {
                "what": f"rank 0 holds the wideband complex64 IQ ({wide_src['D']} x 48 kS/s, {wide_src['bytes'] / 1e6:.1f} MB per step); every "
                        "rank pulls it over NVLink with the copy engine (CUDA IPC + sonde_chan_process_c64_peer), channelises "
                        "its own channel block (tcgen05 GEMM) and decodes it; pull + GEMM of step i+1 are enqueued beside the "
                        "decode of step i",
                "value": world * args.steps * C * L / (tw6 * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": tw6 / args.steps,
                "efficiency_vs_resident_shards": (ms_max / args.steps) / (tw6 / args.steps),
                "rank0_egress_bytes_per_step": (world - 1) * wide_src["bytes"],
                "rank0_egress_gbs": (world - 1) * wide_src["bytes"] / (tw6 / args.steps * 1e-3) / 1e9}
        if not args.no_cpu_baseline and world == 1:
            cpu = CpuPath()
            n_ch = min(C, 2 * ncores)
            t, f, k = cpu.run(sig_types[:n_ch], host_iq[:n_ch], L, ncores)
            reps = 1
            while t < 8.0 and reps < 64:                      # ~10 s of CPU work
                dt, _, _ = cpu.run(sig_types[:n_ch], host_iq[:n_ch], L, ncores)
                t += dt
                reps += 1
            line["cpu_baseline"] = {"value": reps * n_ch * n_total / t / 1e6, "unit": UNIT, "cores": ncores,
                                    "kind": cpu.kind,
                                    "sample": f"first {n_ch} channels of the workload x {n_total} samples, {reps} passes; "
                                              "discriminator = this repo's restated polynomial atan2 (SDR++ is not vendored)",
                                    "ok_frames_per_s": k * reps / t}
        emit(line)

    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
