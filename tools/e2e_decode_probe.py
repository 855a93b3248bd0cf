#!/usr/bin/env python3
"""Times the two host-pointer calls of the end-to-end figure separately (pinned buffers, 1 GiB mixed)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import numpy as np, torch
import lzs_b200 as B
L = B.lib()
total, chunk = 1 << 30, 65536
n = total // chunk
db = B.DeviceBatch(total, chunk); db.fill(B.CORPUS_MIXED, 0x5EED0002); torch.cuda.synchronize()
stride = db.comp_stride
raw = torch.empty(total + 64, dtype=torch.uint8).pin_memory(); raw[:total].copy_(db.raw[:total])
comp = torch.empty(n * stride + 64, dtype=torch.uint8).pin_memory()
dec = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
idx = np.arange(n, dtype=np.uint64)
in_off, in_len = idx * np.uint64(chunk), np.full(n, chunk, dtype=np.uint32)
c_off = np.zeros(n, dtype=np.uint64); c_len = np.zeros(n, dtype=np.uint32); d_len = np.zeros(n, dtype=np.uint32)
used = np.zeros(1, dtype=np.uint64)
p = lambda t: ctypes.cast(t.data_ptr(), B.u8p)
def comp_call():
    B.check(L.lzs_b200_compress_packed_host(p(raw), B._p(in_off, B.u64p), B._p(in_len, B.u32p), total, p(comp), n * stride,
                                            B._p(c_off, B.u64p), B._p(c_len, B.u32p), n, B._p(used, B.u64p)))
def dec_call():
    B.check(L.lzs_b200_decompress_batch_host(p(comp), B._p(c_off, B.u64p), B._p(c_len, B.u32p), int(used[0]), p(dec),
                                             B._p(in_off, B.u64p), B._p(in_len, B.u32p), B._p(d_len, B.u32p), total, n))
comp_call(); dec_call()
tc = td = 1e9
for _ in range(4):
    t0 = time.perf_counter(); comp_call(); t1 = time.perf_counter(); dec_call(); t2 = time.perf_counter()
    tc, td = min(tc, t1 - t0), min(td, t2 - t1)
assert torch.equal(dec[:total], raw[:total])
print("compress call %.1f ms  decompress call %.1f ms  (SLICE_MIB=%s)" % (tc * 1e3, td * 1e3, os.environ.get("LZS_B200_SLICE_MIB", "default")))
