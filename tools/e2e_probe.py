#!/usr/bin/env python3
"""Break down the host-pointer (e2e) path: raw PCIe copy rates vs the two batch_host calls."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import numpy as np, torch
import lzs_b200 as B
L = B.lib()
total, CHUNK = 1 << 30, 65536
n = total // CHUNK
db = B.DeviceBatch(total, CHUNK)
db.fill(B.CORPUS_MIXED, 0x5EED0002)
torch.cuda.synchronize()
stride = db.comp_stride
raw = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
comp = torch.empty(n * stride + 64, dtype=torch.uint8).pin_memory()
dec = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
raw[:total].copy_(db.raw[:total])
def t(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
d = torch.empty(total, dtype=torch.uint8, device="cuda")
print("H2D 1 GiB pinned: %.1f ms" % t(lambda: d.copy_(raw[:total], non_blocking=True)))
print("D2H 1 GiB pinned: %.1f ms" % t(lambda: dec[:total].copy_(d, non_blocking=True)))
idx = np.arange(n, dtype=np.uint64)
in_off, in_len = idx * np.uint64(CHUNK), np.full(n, CHUNK, dtype=np.uint32)
c_off, c_cap = idx * np.uint64(stride), np.full(n, stride, dtype=np.uint32)
c_len = np.zeros(n, dtype=np.uint32); d_len = np.zeros(n, dtype=np.uint32)
p = lambda x: ctypes.cast(x.data_ptr(), B.u8p)
def comp_call():
    B.check(L.lzs_b200_compress_batch_host(p(raw), B._p(in_off, B.u64p), B._p(in_len, B.u32p), total, p(comp), B._p(c_off, B.u64p), B._p(c_cap, B.u32p), B._p(c_len, B.u32p), n * stride, n))
def dec_call():
    B.check(L.lzs_b200_decompress_batch_host(p(comp), B._p(c_off, B.u64p), B._p(c_len, B.u32p), n * stride, p(dec), B._p(in_off, B.u64p), B._p(in_len, B.u32p), B._p(d_len, B.u32p), total, n))
print("compress_batch_host: %.1f ms" % t(comp_call))
print("decompress_batch_host: %.1f ms" % t(dec_call))
assert torch.equal(dec[:total], raw[:total])
out_off = np.zeros(n, dtype=np.uint64); used = ctypes.c_uint64(0)
def packed_call():
    B.check(L.lzs_b200_compress_packed_host(p(raw), B._p(in_off, B.u64p), B._p(in_len, B.u32p), total, p(comp), n * stride, B._p(out_off, B.u64p), B._p(c_len, B.u32p), n, ctypes.cast(ctypes.byref(used), B.u64p)))
try:
    print("compress_packed_host: %.1f ms" % t(packed_call))
    def dec_packed():
        B.check(L.lzs_b200_decompress_batch_host(p(comp), B._p(out_off, B.u64p), B._p(c_len, B.u32p), used.value, p(dec), B._p(in_off, B.u64p), B._p(in_len, B.u32p), B._p(d_len, B.u32p), total, n))
    print("decompress_batch_host (packed input): %.1f ms" % t(dec_packed))
    assert torch.equal(dec[:total], raw[:total])
except Exception as e:
    print("packed probe failed:", e)
