#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel, odd sizes,
unaligned streams, truncated capacities, damaged streams, incremental calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lzs_b200 as B
import helpers
o = helpers.oracle()
rng = np.random.default_rng(1)
data = list(helpers.edge_case_inputs().values())
data += [helpers.corpus(helpers.CORPUS_MIXED, 1, int(rng.integers(1, 5000)), first_index=i).tobytes() for i in range(40)]
data += [helpers.corpus(helpers.CORPUS_PACKET, 1, 1500, first_index=i).tobytes() for i in range(40)]
comp = B.compress_streams(data)
assert comp == [o.compress(d) for d in data]
assert B.decompress_streams(comp, [len(d) for d in data]) == data
caps = [max(0, len(c) - 3) for c in comp]
assert B.compress_streams(data, caps=caps) == [c[:k] for c, k in zip(comp, caps)]
dam = [c[:max(0, len(c) // 2)] for c in comp] + [bytes(rng.integers(0, 256, 100, dtype=np.uint8)) for _ in range(20)]
got = B.decompress_streams(dam, [700] * len(dam))
assert got == [o.decompress(s, 700) for s in dam]
packed, off, ln = B.compress_streams_packed(data)
assert [packed[int(a):int(a) + int(l)].tobytes() for a, l in zip(off, ln)] == comp
for lanes in (4, 16, 32, 8):
    B.check(B.lib().lzs_b200_set_decode_lanes(lanes))
    assert B.decompress_streams(comp[:30], [len(d) for d in data[:30]]) == data[:30]
assert B.lzs_compress(data[5]) == comp[5] and B.lzs_decompress(comp[5], len(data[5])) == data[5]
import ctypes, inc_drivers as D
ours = D.StructCodec(B.lib())
out, _ = D.drive(ours, False, data[-1], 200, 100, 4000)
assert out == comp[-1]
back, _ = D.drive(ours, True, comp[-1], 50, 70, 1600)
assert back == data[-1]
print("sanitize workload ok")
# round 2: device-resident flows through the incremental API, the decoder's launch order (>= 1024
# streams with room in scratch), the pack kernel, the forced exact match-finder launch
import torch
flows = 1200
db = B.DeviceBatch(flows * 1500, 1500)
db.fill(B.CORPUS_PACKET, 0x5EED0000 + 3)
db.compress(); db.decompress(); torch.cuda.synchronize()
assert db.roundtrip_ok()
dev = torch.device("cuda:0")
cap = (B.compressed_max(1500) + 15) // 16 * 16
out = torch.zeros(flows * cap, dtype=torch.uint8, device=dev)
idx = torch.arange(flows, dtype=torch.int64, device=dev)
f = B.DeviceFlows(flows)
full = lambda v: torch.full((flows,), v, dtype=torch.int64, device=dev)
f.offer(db.raw.data_ptr() + idx * 1500, full(1500), out.data_ptr() + idx * cap, full(cap), True)
produced = torch.zeros(flows, dtype=torch.int64, device=dev)
for _ in range(64):
    _, ou, _ = f.call()
    produced += ou
    if bool(((f.jobs[:, 5] >> 32) == 0).all()):
        break
assert torch.equal(produced.to(torch.int32), db.comp_len)
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
import lzs_dist
packed, poff = lzs_dist.pack_streams(db.comp, db.comp_off, db.comp_len)
torch.cuda.synchronize()
B.lib().lzs_b200_set_force_safe_match.argtypes = [ctypes.c_int]
B.lib().lzs_b200_set_force_safe_match(1)
small = B.DeviceBatch(40 * 3000, 3000)
small.fill(B.CORPUS_MIXED, 0x5EED0000 + 9)
small.compress(); small.decompress(); torch.cuda.synchronize()
B.lib().lzs_b200_set_force_safe_match(0)
assert small.roundtrip_ok()
print("sanitize workload (round 2 additions) ok")
# long streams cut into pieces (csrc/k23_pieces.cuh): plan, K1 with look-ahead, spec, fix, sweep, pack
B.set_piece_bytes(1024)
long_data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 50000 + 13 * i, first_index=60 + i).tobytes() for i in range(3)]
long_data += [b"\1" * 30000 + helpers.corpus(helpers.CORPUS_TEXT, 1, 9000, first_index=2).tobytes() + b"\0" * 20000, b"", b"x"]
got = B.compress_streams(long_data)
cut = B.compress_streams(long_data[:2], caps=[777, 0])
B.set_piece_bytes(65536)
assert got == [o.compress(d) for d in long_data]
assert cut == [o.compress(long_data[0])[:777], b""]
print("sanitize workload (pieces) ok")
# the decoder for long streams (csrc/k4_pieces.cuh): plan, spec, fix x2, sweep, emit, copy, dirty list, k4_decode for the dirty
B.set_decode_piece_bytes(64)
comp_long = [o.compress(d) for d in long_data]
mixed_bag = comp_long + [c[:len(c) // 2] for c in comp_long[:2]] + [bytes(rng.integers(0, 256, 500, dtype=np.uint8)) for _ in range(5)]
bag_caps = [len(d) for d in long_data] + [len(d) for d in long_data[:2]] + [3000] * 5
got = B.decompress_streams(mixed_bag, bag_caps)
B.set_decode_piece_bytes(2048)
assert got == [o.decompress(s, c) for s, c in zip(mixed_bag, bag_caps)]
print("sanitize workload (decoder pieces) ok")
# a handful of long streams through the host call: the copies resolved by pointer doubling (k4j_*), one of them damaged
big3 = [helpers.corpus(kind, 1, 420000, first_index=9).tobytes() for kind in (helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)]
big3.append(b"\0" * 300000 + b"ab" * 100000)
c3 = [o.compress(d) for d in big3]
c3.append(c3[0][:len(c3[0]) // 2])
caps3 = [len(d) for d in big3] + [len(big3[0])]
B.set_decode_piece_bytes(512)
got = B.decompress_streams(c3, caps3)
B.set_decode_piece_bytes(2048)
assert got == [o.decompress(s, c) for s, c in zip(c3, caps3)]
print("sanitize workload (pointer doubling) ok")
