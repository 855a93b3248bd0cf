#!/usr/bin/env python3
"""Per-tile timeline of K1's CTA 0 from a -DLZS_K1_TIMELINE build (variants/tl.so): where a tile's
life goes -- loader, build warps, query warps -- and how long the stages wait for each other."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import numpy as np, torch
import lzs_b200 as B
KINDS = {"text": 0, "binary": 1, "random": 2, "mixed": 3}
L = B.lib()
L.lzs_b200_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
db = B.DeviceBatch(1 << 30, 65536)
for kind in sys.argv[1].split(","):
    db.fill(KINDS[kind], 0x5EED0002); torch.cuda.synchronize()
    db.match_only(); torch.cuda.synchronize()
    L.lzs_b200_debug_timeline(None, 1)
    L.lzs_b200_debug_warp_busy.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.lzs_b200_debug_warp_busy(None, 1)
    db.match_only(); torch.cuda.synchronize()
    wb = np.zeros((64, 2), dtype=np.uint64)
    L.lzs_b200_debug_warp_busy(wb.ctypes.data, 0)
    tl = np.zeros((8192, 8), dtype=np.uint64)
    L.lzs_b200_debug_timeline(tl.ctypes.data, 0)
    ok = (tl[:, 1] != 0) & (tl[:, 7] != 0) & (tl[:, 6] != np.uint64(2**64 - 1))
    t = tl[ok].astype(np.int64)
    t = t[200:-50]                                   # steady state
    n = len(t)
    period = (t[-1, 1] - t[0, 1]) / (n - 1)
    print("%s: %d tiles, period %.0f cycles per tile" % (kind, n, period))
    ntl = int(ok.sum())
    print("   build cycles per tile by warp (0-10 levels 2-12, 11 run table): " + " ".join("%d" % (int(wb[w, 0]) // max(1, ntl)) for w in range(12)))
    names = ["loader start -> published", "published -> first build start", "first -> last build start", "build: first start -> first done",
             "first build done -> last build done", "last build done -> first query enters", "query: first enters -> last leaves",
             "tile life: loader start -> last query leaves"]
    vals = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 2], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5], t[:, 7] - t[:, 6], t[:, 7] - t[:, 0]]
    for nm, v in zip(names, vals):
        print("   %-48s mean %7.0f  p50 %7.0f  p95 %7.0f" % (nm, v.mean(), np.median(v), np.percentile(v, 95)))
    # how far ahead is each stage (tiles): loader vs query
    print("   loader start of tile g+? happens before query of g leaves: mean lead %.2f tiles" % (np.mean([(t[:, 0] < t[i, 7]).sum() - i for i in range(0, n, 97)])))
