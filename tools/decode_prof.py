#!/usr/bin/env python3
"""Small driver for ncu on the decoder for long streams: python tools/decode_prof.py --mib 1024 --chunk 1048576 [--dpiece 8192]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch
import lzs_b200 as B

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--chunk", type=int, default=1 << 20)
ap.add_argument("--dpiece", type=int, default=2048)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--jump", action="store_true", help="pointer doubling instead of the replay")
a = ap.parse_args()
total = a.mib << 20
B.set_decode_piece_bytes(a.dpiece)
db = B.DeviceBatch(total, a.chunk)
B.check(B.lib().lzs_b200_corpus_fill_device(db.raw.data_ptr(), 65536, 65536, 0, total // 65536, 0x5EED0002, B.CORPUS_MIXED, db._stream()))
db.compress()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(a.iters):
    ev[0].record(); (db.decompress_jump() if a.jump else db.decompress()); ev[1].record()
    torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1])
    print(("jump " if a.jump else "") + "chunk=%d dpiece=%d iter %d: decompress %.2f ms (%.1f GB/s)" % (a.chunk, a.dpiece, it, t, total / t / 1e6))
assert db.roundtrip_ok()
