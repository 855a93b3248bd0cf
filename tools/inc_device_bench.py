#!/usr/bin/env python3
"""BASELINE configs[2] "via the incremental API", device resident: n flows x one 1500-byte packet,
every flow through lzs_compress_init + lzs_compress_incremental(add_end_marker) until
LZS_C_STATUS_END_MARKER, all flows advanced together by lzs_b200_compress_incremental_batch_device
(one launch per call round).  Checks: every flow's bytes equal the batch compressor's (= the
reference's lzs_compress, tests/test_gpu_parity.py), and the call trace of a sample of flows
(bytes taken / produced and status of EVERY call) equals the unmodified reference's.  Then a second
packet per flow on the kept history, sample-checked against the reference, and the flows decoded
through the device-resident incremental decoder.  Prints one JSON line."""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lzs_b200 as B
import helpers, inc_drivers as D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flows", type=int, default=1 << 20)
    ap.add_argument("--sample", type=int, default=256)
    a = ap.parse_args()
    n, plen = a.flows, 1500
    dev = torch.device("cuda:0")
    cap = (B.compressed_max(plen) + 15) // 16 * 16
    db = B.DeviceBatch(n * plen, plen)
    db.fill(B.CORPUS_PACKET, 0x5EED0000 + 3)
    db.compress()                                   # the bulk path: what every flow's first packet must equal
    torch.cuda.synchronize()
    out = torch.zeros(n * cap, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    flows = B.DeviceFlows(n)
    full = lambda v: torch.full((n,), v, dtype=torch.int64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    traces = []
    ev[0].record()
    flows.offer(db.raw.data_ptr() + idx * plen, full(plen), out.data_ptr() + idx * cap, full(cap), True)
    produced = torch.zeros(n, dtype=torch.int64, device=dev)
    rounds = 0
    while True:
        in_used, out_used, status = flows.call()
        produced += out_used
        traces.append((in_used[:a.sample].cpu().numpy(), out_used[:a.sample].cpu().numpy(), status[:a.sample].cpu().numpy()))
        rounds += 1
        if bool(((flows.jobs[:, 5] >> 32) == 0).all()):
            break
        assert rounds < 64
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    # every flow equals the bulk path
    assert torch.equal(produced.to(torch.int32), db.comp_len), "lengths differ from the batch compressor"
    o2 = out.view(n, cap)
    c2 = db.comp[:n * db.comp_stride].view(n, db.comp_stride)[:, :cap]
    cols = torch.arange(cap, device=dev)[None, :]
    step = 1 << 16
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        live = cols < produced[lo:hi, None]
        assert not ((o2[lo:hi] != c2[lo:hi]) & live).any(), "bytes differ from the batch compressor"
    # call traces of the sample against the unmodified reference
    ref = D.StructCodec(ctypes.CDLL(helpers.REF_SO)) if os.path.exists(helpers.REF_SO) else None
    checked = 0
    if ref is not None:
        raw = db.raw[:a.sample * plen].cpu().numpy()
        for s in range(min(a.sample, n)):
            data = raw[s * plen:(s + 1) * plen].tobytes()
            _, want = D.drive(ref, False, data, plen, cap, cap)
            got = [(int(t[1][s]), int(t[2][s]), int(t[0][s])) for t in traces]
            got = got[:len(want)]
            assert got == want, (s, got, want)
            checked += 1
    # second packet per flow on the kept history (sample-checked), then decode both through one decoder state
    db2 = B.DeviceBatch(n * plen, plen)
    db2.fill(B.CORPUS_PACKET, 0x5EED0000 + 3, first_index=1)      # packet s+1's bytes: overlaps flow s+1's first packet
    out2 = torch.zeros(n * cap, dtype=torch.uint8, device=dev)
    flows.offer(db2.raw.data_ptr() + idx * plen, full(plen), out2.data_ptr() + idx * cap, full(cap), True)
    produced2 = torch.zeros(n, dtype=torch.int64, device=dev)
    for _ in range(64):
        _, out_used, _ = flows.call()
        produced2 += out_used
        if bool(((flows.jobs[:, 5] >> 32) == 0).all()):
            break
    torch.cuda.synchronize()
    if ref is not None:
        R = ref.lib
        raw1, raw2 = db.raw[:64 * plen].cpu().numpy(), db2.raw[:64 * plen].cpu().numpy()
        got2 = out2[:64 * cap].cpu().numpy()
        for s in range(min(64, n)):
            st = ref.new(False)
            outs = []
            for data in (raw1[s * plen:(s + 1) * plen].tobytes(), raw2[s * plen:(s + 1) * plen].tobytes()):
                src = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
                dst = np.zeros(cap, dtype=np.uint8)
                f = (ctypes.c_uint64 * 4).from_address(ctypes.addressof(st))
                f[0], f[1], f[2], f[3] = src.ctypes.data, dst.ctypes.data, len(data), cap
                for _ in range(64):
                    R.lzs_compress_incremental(st, True)
                    if ctypes.c_uint8.from_address(ctypes.addressof(st) + 32).value & 0x04:
                        break
                outs.append(dst[:cap - f[3]].tobytes())
            assert got2[s * cap:s * cap + int(produced2[s])].tobytes() == outs[1], s
    dec = B.DeviceFlows(n, decompress=True)
    back = torch.zeros(n * 2 * plen, dtype=torch.uint8, device=dev)
    d0 = torch.cuda.Event(enable_timing=True); d1 = torch.cuda.Event(enable_timing=True)
    d0.record()
    dec.offer(out.data_ptr() + idx * cap, produced, back.data_ptr() + idx * 2 * plen, full(2 * plen), False)
    _, ou1, st1 = dec.call()
    dec.offer(out2.data_ptr() + idx * cap, produced2, back.data_ptr() + idx * 2 * plen + plen, full(plen), False)
    _, ou2, st2 = dec.call()
    d1.record()
    torch.cuda.synchronize()
    assert bool((ou1 == plen).all()) and bool((ou2 == plen).all()) and bool(((st1 & 4) != 0).all())
    b2 = back.view(n, 2 * plen)
    assert torch.equal(b2[:, :plen], db.raw[:n * plen].view(n, plen)) and torch.equal(b2[:, plen:], db2.raw[:n * plen].view(n, plen))
    print(json.dumps({"what": "incremental API, device resident: %d flows x 1500 B, init + incremental until END_MARKER" % n,
                      "compress_gbs": n * plen / ms / 1e6, "ms": ms, "call_rounds": rounds,
                      "decompress_gbs_two_packets": 2 * n * plen / d0.elapsed_time(d1) / 1e6,
                      "flows_equal_to_batch_compressor": n, "call_traces_equal_to_reference": checked,
                      "second_packet_on_kept_history_checked": min(64, n) if ref is not None else 0,
                      "ratio_first_packet": n * plen / float(produced.sum()), "ratio_second_packet_kept_history": n * plen / float(produced2.sum())}))


if __name__ == "__main__":
    main()
