#!/usr/bin/env python3
"""One lzs_compress call on a buffer of zeros (one match through the whole stream: the worst case of the
piece path, measured by one warp in the sweep) beside the unmodified reference on one host core."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, lzs_b200 as B, helpers
ref = helpers.reference() or helpers.oracle()
for mib in (64, 512):
    data = bytes(mib << 20)
    B.lzs_compress(data[:1 << 20])
    best = 1e30
    for _ in range(2):                  # the first call of a size also grows the library's device buffers
        t0 = time.perf_counter(); got = B.lzs_compress(data); best = min(best, time.perf_counter() - t0)
    t1 = time.perf_counter(); want = ref.compress(data); t2 = time.perf_counter()
    assert got == want
    print("zeros %d MiB: lzs_compress %.1f ms (Python wrapper included), reference %.1f ms, %d bytes" % (mib, best * 1e3, (t2 - t1) * 1e3, len(got)), flush=True)
