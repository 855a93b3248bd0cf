#!/usr/bin/env python3
"""Whole compress call (K1 + K2K3) on the device, 1 GiB, per kind."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch
import lzs_b200 as B
KINDS = {"text": 0, "binary": 1, "random": 2, "mixed": 3}
db = B.DeviceBatch(1 << 30, 65536)
for kind in sys.argv[1].split(","):
    db.fill(KINDS[kind], 0x5EED0002); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); db.compress(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    db.decompress(); torch.cuda.synchronize()
    assert db.roundtrip_ok()
    print("%s compress %.2f ms" % (kind, best))
