#!/usr/bin/env python3
"""One lzs_decompress call on buffers of 32 KiB .. 1 MiB (pageable host memory, best of 5, the C call alone) beside the
unmodified reference on one host core.  LZS_B200_JUMP_MIN / LZS_B200_DPIECE=0 select the decoder's other ways."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, lzs_b200 as B, helpers
ref = helpers.reference() or helpers.oracle()
for kib in (32, 64, 128, 256, 512, 1024):
    data = helpers.corpus(helpers.CORPUS_MIXED, max(1, kib // 64), min(65536, kib << 10), first_index=7)
    n = len(data); comp = np.frombuffer(ref.compress(data.tobytes()), dtype=np.uint8).copy(); r = len(comp)
    comp = np.concatenate([comp, np.zeros(16, dtype=np.uint8)]); back = np.zeros(n + 16, dtype=np.uint8)
    best = 1e30
    for _ in range(5):
        t0 = time.perf_counter(); rd = B.lib().lzs_decompress(B._p(back), n, B._p(comp), r); best = min(best, time.perf_counter() - t0)
    assert rd == n and back[:n].tobytes() == data.tobytes()
    t0 = time.perf_counter(); ref.decompress(comp[:r].tobytes(), n); tc = time.perf_counter() - t0
    print(kib, "KiB: lzs_decompress %.2f ms, reference %.2f ms" % (best * 1e3, tc * 1e3), flush=True)
