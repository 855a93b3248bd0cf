#!/usr/bin/env python3
"""The drop-in calls on ONE buffer of pageable host memory: lzs_compress / lzs_decompress, the C calls alone
(best of 3), beside the unmodified reference on one host core.  One JSON line per size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, lzs_b200 as B, helpers
ref = helpers.reference() or helpers.oracle()
for mib in [int(x) for x in (sys.argv[1:] or ["1", "16", "256"])]:
    data = helpers.corpus(helpers.CORPUS_MIXED, mib * 16, 65536, first_index=7)
    n = len(data)
    src = np.ascontiguousarray(data)
    cap = B.compressed_max(n)
    dst = np.zeros(cap, dtype=np.uint8)
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); r = B.lib().lzs_compress(B._p(dst), cap, B._p(src), n); best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter(); want = ref.compress(data.tobytes()); t_cpu = time.perf_counter() - t0
    assert dst[:r].tobytes() == want
    comp = np.ascontiguousarray(dst[:r]); back = np.zeros(n + 16, dtype=np.uint8)
    best_d = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); rd = B.lib().lzs_decompress(B._p(back), n, B._p(comp), r); best_d = min(best_d, time.perf_counter() - t0)
    assert rd == n and back[:n].tobytes() == data.tobytes()
    t0 = time.perf_counter(); ref.decompress(want, n); t_cpu_d = time.perf_counter() - t0
    print(json.dumps({"single_call_mib": mib, "lzs_compress_ms": round(best * 1e3, 2), "gbs": round(n / best / 1e9, 3),
                      "reference_one_core_gbs": round(n / t_cpu / 1e9, 3), "lzs_decompress_ms": round(best_d * 1e3, 2),
                      "decompress_gbs": round(n / best_d / 1e9, 3), "reference_one_core_decompress_gbs": round(n / t_cpu_d / 1e9, 3),
                      "host_memory": "pageable"}), flush=True)
