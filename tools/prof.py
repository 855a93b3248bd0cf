#!/usr/bin/env python3
"""Small driver for ncu / timing experiments: one corpus, a few passes of the hot path.
  python tools/prof.py --mib 256 --kind mixed --chunk 65536 --iters 2 [--time]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch  # noqa: E402
import lzs_b200 as B  # noqa: E402

KINDS = {"text": 0, "binary": 1, "random": 2, "mixed": 3, "packet": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--kind", default="mixed")
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--compress-only", action="store_true", help="the whole compressor as one call (long streams are cut into pieces)")
    a = ap.parse_args()
    total = (a.mib << 20) // a.chunk * a.chunk
    if a.lanes:
        B.check(B.lib().lzs_b200_set_decode_lanes(a.lanes))
    db = B.DeviceBatch(total, a.chunk)
    for kind in a.kind.split(","):
        a.kind = kind
        run(a, db, total)


def run(a, db, total):
    db.fill(KINDS[a.kind], 0x5EED0002)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if a.compress_only:
        for it in range(a.iters):
            ev[0].record(); db.compress(); ev[1].record()
            torch.cuda.synchronize()
            if a.time:
                t = ev[0].elapsed_time(ev[1])
                print("%s chunk=%d iter %d: compress %.2f ms (%.1f GB/s)" % (a.kind, a.chunk, it, t, total / t / 1e6))
        return
    for it in range(a.iters):
        ev[0].record(); db.match_only()
        ev[1].record(); db.parse_pack_only()
        ev[2].record(); db.decompress()
        ev[3].record()
        torch.cuda.synchronize()
        if a.time:
            t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
            gb = total / 1e9
            print("%s chunk=%d iter %d: k1 %.2f ms (%.1f GB/s)  k23 %.2f ms  compress %.1f GB/s  k4 %.2f ms (%.1f GB/s)  ratio %.3f"
                  % (a.kind, a.chunk, it, t[0], gb / t[0] * 1e3, t[1], gb / (t[0] + t[1]) * 1e3, t[2], gb / t[2] * 1e3,
                     total / db.compressed_bytes()))
    assert db.roundtrip_ok()


if __name__ == "__main__":
    main()
