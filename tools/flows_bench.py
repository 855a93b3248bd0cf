#!/usr/bin/env python3
"""Bulk path for flows with kept history: n_flows x n_packets x 1500 B, all packets
compressed in one call with their flow's earlier packets as history, decoded one packet index per call.
Prints one JSON line (CUDA events)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import torch
import lzs_b200 as B

ap = argparse.ArgumentParser()
ap.add_argument("--flows", type=int, default=1 << 18)
ap.add_argument("--packets", type=int, default=4)
a = ap.parse_args()
tab = B.DeviceFlowTable(a.flows, a.packets, 1500)
B.check(B.lib().lzs_b200_corpus_fill_device(tab.raw.data_ptr(), 1500, 1500, 0, tab.n, 0x5EED0000 + 3, B.CORPUS_PACKET, None))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
best = [1e9, 1e9, 1e9]
for _ in range(3):
    ev[0].record(); tab.compress(); ev[1].record(); tab.decompress(); ev[2].record()
    torch.cuda.synchronize()
    best[0] = min(best[0], ev[0].elapsed_time(ev[1])); best[1] = min(best[1], ev[1].elapsed_time(ev[2]))
assert torch.equal(tab.dec[:tab.n * 1500], tab.raw[:tab.n * 1500])
with_hist = int(tab.comp_len.sum())
lens = tab.comp_len.clone()
for _ in range(3):
    ev[0].record(); tab.compress_table(); ev[1].record(); torch.cuda.synchronize()
    best[2] = min(best[2], ev[0].elapsed_time(ev[1]))
assert torch.equal(lens, tab.comp_len)
ev[0].record(); tab.compress(with_history=False); ev[1].record(); torch.cuda.synchronize()
nbytes = tab.n * 1500
print(json.dumps({"what": "flows with kept history, bulk path: %d flows x %d packets x 1500 B" % (a.flows, a.packets),
                  "compress_gbs": nbytes / best[0] / 1e6, "compress_gbs_flow_as_one_stream": nbytes / best[2] / 1e6,
                  "decompress_gbs": nbytes / best[1] / 1e6,
                  "ratio_with_history": nbytes / with_hist, "ratio_independent_packets": nbytes / int(tab.comp_len.sum()),
                  "compress_gbs_independent_packets": nbytes / ev[0].elapsed_time(ev[1]) / 1e6}))
