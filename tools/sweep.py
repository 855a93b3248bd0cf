#!/usr/bin/env python3
"""BASELINE configs 3 and 5: packets (1 Mi x 1500 B) and the chunk-size sweep 4 KiB - 1 MiB over
the 1 GiB mixed corpus: device-resident GB/s (CUDA events), ratio, and the unmodified reference
on the host cores over a 64 MiB sample of the same bytes.  Prints one JSON line per row."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lzs_b200 as B
import helpers


def gpu_row(kind, chunk, total, seed, iters=3):
    db = B.DeviceBatch(total, chunk)
    db.fill(kind, seed)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best_c = best_d = 1e30
    for _ in range(iters):
        ev[0].record(); db.compress(); ev[1].record(); db.decompress(); ev[2].record()
        torch.cuda.synchronize()
        best_c = min(best_c, ev[0].elapsed_time(ev[1])); best_d = min(best_d, ev[1].elapsed_time(ev[2]))
    assert db.roundtrip_ok()
    return db, total / best_c / 1e6, total / best_d / 1e6, total / db.compressed_bytes()


def cpu_row(db, chunk, sample_bytes):
    codec = helpers.reference() or helpers.oracle()
    n = max(1, sample_bytes // chunk)
    raw = np.zeros(n * chunk + 16, dtype=np.uint8)
    raw[:n * chunk] = db.raw[:n * chunk].cpu().numpy()
    stride = (helpers.compressed_max(chunk) + 15) // 16 * 16
    idx = np.arange(n, dtype=np.uint64)
    in_off, in_len = idx * np.uint64(chunk), np.full(n, chunk, dtype=np.uint32)
    c_off, c_cap = idx * np.uint64(stride), np.full(n, stride, dtype=np.uint32)
    comp = np.zeros(n * stride + 16, dtype=np.uint8); dec = np.zeros(n * chunk + 16, dtype=np.uint8)
    th = os.cpu_count() or 1
    c_len, t_c = codec.run_streams(False, raw, in_off, in_len, comp, c_off, c_cap, th)
    d_len, t_d = codec.run_streams(True, comp, c_off, c_len, dec, in_off, in_len, th)
    # parity on the sample: GPU streams must equal the CPU streams byte for byte
    g_len = db.comp_len[:n].cpu().numpy()
    assert (g_len == c_len).all()
    g = db.comp[:n * db.comp_stride].cpu().numpy().reshape(n, db.comp_stride)
    c = comp[:n * stride].reshape(n, stride)
    for s in range(0, n, max(1, n // 64)):
        assert (g[s, :g_len[s]] == c[s, :c_len[s]]).all()
    return n * chunk / t_c / 1e9, n * chunk / t_d / 1e9, th


def main():
    rows = [("packets 1Mi x 1500 B (config 3)", B.CORPUS_PACKET, 1500, (1 << 20) * 1500, 0x5EED0003)]
    for kib in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
        rows.append(("chunk %d KiB (config 5)" % kib, B.CORPUS_MIXED, kib << 10, 1 << 30, 0x5EED0002))
    for name, kind, chunk, total, seed in rows:
        db, c, d, r = gpu_row(kind, chunk, total, seed)
        cc, cd, th = cpu_row(db, chunk, 64 << 20)
        print(json.dumps({"row": name, "gpu_compress_gbs": round(c, 2), "gpu_decompress_gbs": round(d, 2),
                          "ratio": round(r, 4), "cpu_compress_gbs": round(cc, 3), "cpu_decompress_gbs": round(cd, 3),
                          "cpu_threads": th, "cpu_kind": "reference" if helpers.reference() else "port",
                          "bit_exact_vs_cpu_on_sample": True}), flush=True)
        del db
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
