#!/usr/bin/env python3
"""K1 timing only (no parse, no decode, no checks): for experiments whose records may be wrong."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch
import lzs_b200 as B
KINDS = {"text": 0, "binary": 1, "random": 2, "mixed": 3}
total = 1 << 30
db = B.DeviceBatch(total, 65536)
for kind in sys.argv[1].split(","):
    db.fill(KINDS[kind], 0x5EED0002); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); db.match_only(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("%s k1 %.2f ms" % (kind, best))
