#!/usr/bin/env python3
"""bench.py -- LZS compress + decompress throughput on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch: every 64 KiB chunk of a
1 GiB-per-GPU synthetic corpus is compressed into its own LZS stream (kernels K1 match
finder, K2+K3 parse/pack) and then decompressed again (K4).  `value` is uncompressed
GB/s (1e9 B/s) through that round trip with all buffers resident in HBM, summed over
ranks; `compress_gbs` / `decompress_gbs` / `ratio` break it down.  `e2e` is the same
round trip through the host-pointer C ABI (lzs_b200_*_batch_host) with pinned host
buffers, copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm
  python bench.py --impl reference ...                          # reference C code on host cores

Multi-GPU: one process per GPU (torchrun); chunks are independent streams, so ranks take
disjoint chunk ranges and the data path has no collective (weak scaling: 1 GiB per GPU).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))

METRIC = "LZS compress+decompress round-trip throughput, uncompressed GB/s (64 KiB chunks)"
SEED = 0x5EED0000 + 2
CHUNK = 65536
CPU_SAMPLE_CHUNKS = 2048         # 128 MiB of the same corpus for the CPU legs


def profiled_traffic():
    """DRAM bytes of the dominant kernel (K1) for one launch on this workload, from the committed
    ncu --set full summary (profiles/r2_k1_match.txt); None if that file is absent."""
    try:
        text = open(os.path.join(ROOT, "profiles", "r1_k1_match.txt")).read()
        gb = float(text.split("dram traffic (read+write):")[1].split("GB")[0])
        return gb * 1e9
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ CPU legs (oracle side)

def cpu_codec():
    """The unmodified reference build when it travelled with the repo, else the port."""
    import helpers
    ref = helpers.reference()
    if ref is not None:
        return ref, "reference"
    return helpers.oracle(), "port"


def cpu_roundtrip(codec, raw, n_chunks, chunk, threads):
    """Compress then decompress n_chunks chunks of `raw` (numpy uint8, one spare byte at the
    end) on `threads` host threads; returns (t_compress, t_decompress, compressed_bytes)."""
    import numpy as np
    stride = (chunk + (chunk + 7) // 8 + 3 + 15) // 16 * 16
    idx = np.arange(n_chunks, dtype=np.uint64)
    in_off, in_len = idx * np.uint64(chunk), np.full(n_chunks, chunk, dtype=np.uint32)
    c_off, c_cap = idx * np.uint64(stride), np.full(n_chunks, stride, dtype=np.uint32)
    comp = np.zeros(n_chunks * stride + 16, dtype=np.uint8)
    dec = np.zeros(n_chunks * chunk + 16, dtype=np.uint8)
    c_len, t_c = codec.run_streams(False, raw, in_off, in_len, comp, c_off, c_cap, threads)
    d_len, t_d = codec.run_streams(True, comp, c_off, c_len, dec, in_off, in_len, threads)
    assert (d_len == in_len).all() and (dec[:n_chunks * chunk] == raw[:n_chunks * chunk]).all()
    cpu_roundtrip.last = (comp, c_off, c_len)           # for the parity check of the measured run
    return t_c, t_d, int(c_len.sum())


def host_corpus(n_chunks, chunk, first_index):
    import helpers
    import numpy as np
    buf = np.zeros(n_chunks * chunk + 16, dtype=np.uint8)
    buf[:n_chunks * chunk] = helpers.corpus(helpers.CORPUS_MIXED, n_chunks, chunk, seed=SEED, first_index=first_index)
    return buf


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    codec, kind = cpu_codec()
    threads = os.cpu_count() or 1
    n_chunks = CPU_SAMPLE_CHUNKS
    raw = host_corpus(n_chunks, CHUNK, 0)
    times = []
    comp_bytes = 0
    for it in range(args.warmup + args.steps):
        t_c, t_d, comp_bytes = cpu_roundtrip(codec, raw, n_chunks, CHUNK, threads)
        if it >= args.warmup:
            times.append((t_c, t_d))
    nbytes = n_chunks * CHUNK
    tc = sum(t[0] for t in times) / len(times)
    td = sum(t[1] for t in times) / len(times)
    value = nbytes / (tc + td) / 1e9
    sample = "%d x %d B chunks (%d MiB) of the bench corpus per step, all host threads" % (
        n_chunks, CHUNK, nbytes >> 20)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (tc + td) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(),
        "compress_gbs": nbytes / tc / 1e9, "decompress_gbs": nbytes / td / 1e9, "ratio": nbytes / comp_bytes,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(bytes_per_gpu=1 << 30, total_gib=None):
    if total_gib:
        what = ("%d GiB synthetic mixed corpus (text / binary records / incompressible by chunk) sharded over the GPUs "
                "in contiguous chunk ranges, 64 KiB independent LZS streams; compress then decompress "
                "(BASELINE configs[3])" % total_gib)
    else:
        what = ("1 GiB synthetic mixed corpus (text / binary records / incompressible by chunk) per GPU, "
                "split into 64 KiB independent LZS streams; compress then decompress (BASELINE configs[1])")
    return {"workload": what, "chunk_bytes": CHUNK, "bytes_per_gpu": bytes_per_gpu, "seed": SEED,
            "l2_policy": "inputs (>= 1 GiB per GPU) larger than L2 (126 MB); no flush needed",
            "parallelism": "independent chunk ranges per GPU, no data-path collective"}


# ------------------------------------------------------------------------------ GPU arm

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's threads to the CPUs NVML names as closest to its GPU, so that the pinned host
    buffers of the end-to-end leg are allocated (first touch) on that NUMA node and the ranks of one
    box do not all stage through node 0.  Best effort: returns what was done for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return "nvml ideal cpus: %d of %d allowed (%d..%d)" % (len(after), before, after[0], after[-1])
    except Exception as e:                                  # no NVML, a cpuset that excludes them, ...
        return "unchanged (%s)" % type(e).__name__


def run_gpu_arm(args):
    import numpy as np
    import torch
    import lzs_b200 as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the LZS codec has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    affinity = bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    B.lib()

    total = 1 << 30
    if args.total_gib:
        if (args.total_gib << 30) % (world * CHUNK):
            raise SystemExit("bench.py: --total-gib must split into whole chunks per GPU")
        total = (args.total_gib << 30) // world         # strong scaling: the corpus is fixed, the shards shrink
    n_chunks = total // CHUNK
    db = B.DeviceBatch(total, CHUNK, device=dev)
    db.fill(B.CORPUS_MIXED, SEED, first_index=rank * n_chunks)     # this rank's shard of the corpus
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(ev=None):
        if ev:
            ev[0].record()
        db.match_only()
        if ev:
            ev[1].record()
        db.parse_pack_only()
        if ev:
            ev[2].record()
        db.decompress()
        if ev:
            ev[3].record()

    for _ in range(args.warmup):
        one_step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = B.lib().lzs_b200_kernel_launches()
    barrier()
    for k in range(args.steps):
        one_step(events[k])
    barrier()
    launches = B.lib().lzs_b200_kernel_launches() - launches0
    clocks = sampler.stop()

    t_k1 = sum(e[0].elapsed_time(e[1]) for e in events) / args.steps
    t_k23 = sum(e[1].elapsed_time(e[2]) for e in events) / args.steps
    t_k4 = sum(e[2].elapsed_time(e[3]) for e in events) / args.steps
    t_total = events[0][0].elapsed_time(events[-1][3]) / args.steps       # ms per step, device clock
    assert db.roundtrip_ok(), "round trip mismatch inside the timed region"
    comp_bytes = db.compressed_bytes()

    # ---- the exchange step (N > 1): all-gather-v of the variable-size outputs over NCCL, timed on
    #      the device.  It is not part of the data path, so `value` excludes
    #      it and `value_with_gather` includes it.
    gather = None
    if dist is not None:
        import lzs_dist
        g = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        best = None
        for it in range(3):
            barrier()
            g[0].record()
            payload, lens, offs, ginfo = lzs_dist.gather_compressed(db.comp, db.comp_off, db.comp_len)
            g[2].record()
            torch.cuda.synchronize()
            t = (g[0].elapsed_time(g[2]), 0.0, g[0].elapsed_time(g[2]))
            if it and (best is None or t[0] < best[0]):
                best = t
        assert int(lens.numel()) == n_chunks * world
        # every rank holds every stream: check this rank's own streams, and one from each peer by its end marker
        mine = slice(rank * n_chunks, (rank + 1) * n_chunks)
        assert torch.equal(lens[mine], db.comp_len.to(torch.int64))
        probe = int(offs[mine][n_chunks // 2])
        own_off = int(db.comp_off[n_chunks // 2])
        ln = int(db.comp_len[n_chunks // 2])
        assert torch.equal(payload[probe:probe + ln], db.comp[own_off:own_off + ln]), "gathered stream differs"
        # ... and every rank must hold the SAME bytes for every stream: a checksum over all streams' words
        # (position weighted) has to come out equal on all ranks
        words = ((lens + 15) // 16 * 2)                                   # 8-byte words per stream, padding included
        first = offs // 8
        idx = torch.repeat_interleave(first, words) + (torch.arange(int(words.sum()), device=dev) -
                                                       torch.repeat_interleave(torch.cumsum(words, 0) - words, words))
        w64 = payload[:payload.numel() // 8 * 8].view(torch.int64)
        chk = (w64[idx] * (torch.arange(idx.numel(), device=dev) % 1021 + 1)).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert int(lo) == int(hi), "the ranks gathered different bytes"
        own_bytes = int(((db.comp_len.to(torch.int64) + 15) // 16 * 16).sum().item())
        recv = int(ginfo["payload_bytes"]) - own_bytes
        gather = {"ms": best[0], "mode": ginfo["mode"],
                  "bytes_per_rank_received": recv, "payload_bytes": int(ginfo["payload_bytes"]), "streams": int(lens.numel()),
                  "bus_gbs_per_rank_in": recv / (best[0] * 1e-3) / 1e9 if best[0] > 0 else None,
                  "what": "device pack kernel + all-gather-v of all compressed streams to every rank (NCCL); "
                          "pack and exchange timed together, best of 2 after a warm-up"}

    # ---- end to end through the host-pointer C ABI, pinned host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(B, db, total, n_chunks, max(1, min(args.steps, 3)), barrier)

    stats = torch.tensor([t_total, t_k1, t_k23, t_k4, e2e["ms"] if e2e else 0.0, gather["ms"] if gather else 0.0,
                          e2e["copy_only_ms"] if e2e else 0.0, e2e["pageable_ms"] if e2e and e2e.get("pageable_ms") else 0.0],
                         dtype=torch.float64, device=dev)
    sums = torch.tensor([float(comp_bytes)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t_total, t_k1, t_k23, t_k4, t_e2e, t_gather, t_copy, t_pageable = [float(x) for x in stats.tolist()]
    comp_all = float(sums.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        job_bytes = float(total) * world
        alg_bytes_k1 = float(total) + comp_all / world           # n + c per launch, this GPU
        roof = {"bound": "hbm", "kernel": "k1_match (dominant; followed by k23_parse_pack)",
                "achieved": alg_bytes_k1 / (t_k1 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes_k1 / (t_k1 * 1e-3) / 1e9 / peak, "traffic": profiled_traffic(),
                "traffic_source": "ncu --set full, profiles/r2_k1_match.txt (input + 2 B/position of match records)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_k1,
                "per_kernel_ms": {"k1_match": t_k1, "k23_parse_pack": t_k23, "k4_decode": t_k4},
                "compress_path_frac": alg_bytes_k1 / ((t_k1 + t_k23) * 1e-3) / 1e9 / peak,
                "decode_frac": alg_bytes_k1 / (t_k4 * 1e-3) / 1e9 / peak}
        line = {
            "metric": METRIC, "value": job_bytes / (t_total * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total, "higher_is_better": True,
            "scaling": "strong" if args.total_gib else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(total, args.total_gib),
            "compress_gbs": job_bytes / ((t_k1 + t_k23) * 1e-3) / 1e9,
            "decompress_gbs": job_bytes / (t_k4 * 1e-3) / 1e9,
            "ratio": job_bytes / comp_all,
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof,
        }
        if e2e:
            line["e2e"] = {"value": job_bytes / (t_e2e * 1e-3) / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": e2e["steps"], "api": "lzs_b200_compress_packed_host + lzs_b200_decompress_batch_host",
                           "host_buffers": "pinned", "cpu_affinity": affinity,
                           # the same bytes over PCIe in the same order with no kernel at all: the ceiling of this box
                           "copy_only": {"value": job_bytes / (t_copy * 1e-3) / 1e9, "unit": "GB/s", "ms": t_copy,
                                         "what": "H2D input, D2H packed streams, then H2D packed streams, D2H output; "
                                                 "pinned, two copy engines, max over ranks"},
                           "frac_of_copy_ceiling": t_copy / t_e2e if t_e2e > 0 else None}
            if t_pageable > 0:
                line["e2e"]["pageable"] = {"value": job_bytes / (t_pageable * 1e-3) / 1e9, "unit": "GB/s",
                                           "what": "the same calls on plain malloc'ed (pageable) caller buffers"}
        if gather:
            gather["ms"] = t_gather
            line["gather"] = gather
            line["value_with_gather"] = job_bytes / ((t_total + t_gather) * 1e-3) / 1e9
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)              # the CPU leg gets every core the process was given
            line["cpu_baseline"] = cpu_baseline_leg(db)
            line["parity"] = line["cpu_baseline"].pop("parity")
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(B, db, total, n_chunks, steps, barrier):
    import numpy as np
    import torch
    L = B.lib()
    stride = db.comp_stride
    raw = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
    comp = torch.empty(n_chunks * stride + 64, dtype=torch.uint8).pin_memory()
    dec = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
    raw[:total].copy_(db.raw[:total])
    idx = np.arange(n_chunks, dtype=np.uint64)
    in_off, in_len = idx * np.uint64(CHUNK), np.full(n_chunks, CHUNK, dtype=np.uint32)
    c_off, c_cap = idx * np.uint64(stride), np.full(n_chunks, stride, dtype=np.uint32)
    c_len = np.zeros(n_chunks, dtype=np.uint32)
    d_len = np.zeros(n_chunks, dtype=np.uint32)
    u8, u32, u64 = B.u8p, B.u32p, B.u64p

    def p(t):
        return ctypes.cast(t.data_ptr(), u8)

    c_offp = np.zeros(n_chunks, dtype=np.uint64)
    used = np.zeros(1, dtype=np.uint64)

    def step():
        # compress to packed streams (only compressed bytes cross PCIe), then decompress from them
        B.check(L.lzs_b200_compress_packed_host(p(raw), B._p(in_off, u64), B._p(in_len, u32), total, p(comp),
                                                n_chunks * stride, B._p(c_offp, u64), B._p(c_len, u32), n_chunks,
                                                B._p(used, u64)))
        B.check(L.lzs_b200_decompress_batch_host(p(comp), B._p(c_offp, u64), B._p(c_len, u32), int(used[0]),
                                                 p(dec), B._p(in_off, u64), B._p(in_len, u32), B._p(d_len, u32),
                                                 total, n_chunks))

    step()                                            # warm-up: allocations inside the library
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    assert torch.equal(dec[:total], raw[:total]), "e2e round trip mismatch"
    packed = int(used[0])

    # ---- the same bytes over PCIe with no kernels (what this box's copy engines can do at best)
    d_a = db.dec                      # any device buffers of the right size: contents do not matter
    d_b = db.comp
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_only():
        with torch.cuda.stream(s_up):
            d_a[:total].copy_(raw[:total], non_blocking=True)
        with torch.cuda.stream(s_dn):
            comp[:packed].copy_(d_b[:packed], non_blocking=True)
        torch.cuda.synchronize()       # the compress call returns before the decompress call starts
        with torch.cuda.stream(s_up):
            d_b[:packed].copy_(comp[:packed], non_blocking=True)
        with torch.cuda.stream(s_dn):
            dec[:total].copy_(d_a[:total], non_blocking=True)
        torch.cuda.synchronize()

    copy_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        copy_only()
    barrier()
    copy_ms = (time.perf_counter() - t0) * 1e3 / steps

    # ---- the same calls on pageable caller memory (what a plain lzs.h-style caller passes)
    pageable_ms = None
    if not getattr(run_e2e, "skip_pageable", False):
        raw_p = np.empty(total + 64, dtype=np.uint8)
        raw_p[:total] = raw[:total].numpy()
        comp_p = np.empty(n_chunks * stride + 64, dtype=np.uint8)
        dec_p = np.empty(total + 64, dtype=np.uint8)

        def step_pageable():
            B.check(L.lzs_b200_compress_packed_host(B._p(raw_p), B._p(in_off, u64), B._p(in_len, u32), total,
                                                    B._p(comp_p), n_chunks * stride, B._p(c_offp, u64), B._p(c_len, u32),
                                                    n_chunks, B._p(used, u64)))
            B.check(L.lzs_b200_decompress_batch_host(B._p(comp_p), B._p(c_offp, u64), B._p(c_len, u32), int(used[0]),
                                                     B._p(dec_p), B._p(in_off, u64), B._p(in_len, u32), B._p(d_len, u32),
                                                     total, n_chunks))

        step_pageable()
        barrier()
        t0 = time.perf_counter()
        step_pageable()
        barrier()
        pageable_ms = (time.perf_counter() - t0) * 1e3
        assert (dec_p[:total] == raw_p[:total]).all(), "e2e (pageable) round trip mismatch"
    arrays = n_chunks * 24
    h2d = total + arrays + 8 * n_chunks + packed + arrays  # compress input (+ tables), decompress input = packed streams
    d2h = packed + 4 * n_chunks + total + 4 * n_chunks
    return {"ms": ms, "h2d": int(h2d), "d2h": int(d2h), "steps": steps, "copy_only_ms": copy_ms,
            "pageable_ms": pageable_ms}


def cpu_baseline_leg(db):
    import numpy as np
    codec, kind = cpu_codec()
    threads = os.cpu_count() or 1
    n = CPU_SAMPLE_CHUNKS
    raw = np.zeros(n * CHUNK + 16, dtype=np.uint8)
    raw[:n * CHUNK] = db.raw[:n * CHUNK].cpu().numpy()
    cpu_roundtrip(codec, raw, min(n, 256), CHUNK, threads)                     # warm
    t_c, t_d, comp_bytes = cpu_roundtrip(codec, raw, n, CHUNK, threads)
    # parity inside the measured run: the streams the timed GPU steps left in HBM for these chunks
    # against the CPU codec's streams for the same bytes -- lengths and every byte
    cpu_comp, cpu_off, cpu_len = cpu_roundtrip.last
    gpu_len = db.comp_len[:n].cpu().numpy().astype(np.uint32)
    gpu_comp = db.comp[:n * db.comp_stride].cpu().numpy()
    mismatches = 0
    for s in range(n):
        a, b = int(cpu_off[s]), s * db.comp_stride
        if gpu_len[s] != cpu_len[s] or not np.array_equal(cpu_comp[a:a + int(cpu_len[s])], gpu_comp[b:b + int(gpu_len[s])]):
            mismatches += 1
    parity = {"streams": n, "mismatches": mismatches, "against": kind,
              "what": "compressed streams of the timed run (first %d chunks) byte-compared with the CPU codec's; "
                      "the full round trip of all chunks is asserted on the device" % n}
    if mismatches:
        raise SystemExit("bench.py: %d of %d GPU streams differ from the CPU %s" % (mismatches, n, kind))
    t1_c, t1_d, _ = cpu_roundtrip(codec, raw, 128, CHUNK, 1)
    nbytes = n * CHUNK
    return {"parity": parity, "value": nbytes / (t_c + t_d) / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
            "sample": "first %d chunks (%d MiB) of the timed corpus, one pass, all host threads" % (n, nbytes >> 20),
            "compress_gbs": nbytes / t_c / 1e9, "decompress_gbs": nbytes / t_d / 1e9, "ratio": nbytes / comp_bytes,
            "single_thread": {"compress_gbs": 128 * CHUNK / t1_c / 1e9, "decompress_gbs": 128 * CHUNK / t1_d / 1e9,
                              "sample": "first 128 chunks (8 MiB), 1 thread"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--total-gib", type=int, default=0,
                    help="BASELINE configs[3]: one corpus of this many GiB sharded over the GPUs (default: 1 GiB per GPU)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-memory end-to-end figure")
    args = ap.parse_args()
    run_e2e.skip_pageable = args.no_pageable
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
