#!/usr/bin/env python3
"""Does K2+K3 of one half of a batch overlap with K1 of the other half (two CUDA streams)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch
import lzs_b200 as B
half, CHUNK = 1 << 29, 65536
A, Bb = B.DeviceBatch(half, CHUNK), B.DeviceBatch(half, CHUNK)
A.fill(B.CORPUS_MIXED, 0x5EED0002, first_index=0); Bb.fill(B.CORPUS_MIXED, 0x5EED0002, first_index=half // CHUNK)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def seq():
    with torch.cuda.stream(s1):
        A.match_only(); A.parse_pack_only(); Bb.match_only(); Bb.parse_pack_only()
def ovl():
    with torch.cuda.stream(s1):
        A.match_only()
        e = torch.cuda.Event(); e.record(s1)
        Bb.match_only()
    with torch.cuda.stream(s2):
        s2.wait_event(e)
        A.parse_pack_only()
    with torch.cuda.stream(s1):
        s1.wait_stream(s2)
        Bb.parse_pack_only()
for name, f in (("sequential", seq), ("overlapped", ovl)):
    for _ in range(2): f()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(s1)
    for _ in range(3): f()
    s1.wait_stream(s2); t1.record(s1); torch.cuda.synchronize()
    print("%s: %.2f ms per GiB compressed" % (name, t0.elapsed_time(t1) / 3))
