#!/usr/bin/env python3
"""Long streams (csrc/k23_pieces.cuh): device-resident compress GB/s of the 1 GiB mixed corpus in chunks
of 128 KiB .. 1 GiB (ONE stream), with the streams cut into pieces of 32 / 64 / 128 KiB and uncut, and
the drop-in single call lzs_compress() on one host buffer beside the unmodified reference on one
host core.  Every cut result is compared with the uncut one (lengths and bytes).  One JSON line per row."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lzs-compression_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lzs_b200 as B
import helpers


def timed(db, iters):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e30
    for _ in range(iters):
        ev[0].record(); db.compress(); ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return best


def main():
    total = int(os.environ.get("PIECES_TOTAL_MIB", "1024")) << 20
    quick = os.environ.get("PIECES_QUICK") == "1"
    for kib in ((1024, total >> 10) if quick else (128, 256, 512, 1024, 16384, total >> 10)):
        chunk = kib << 10
        row = {"chunk_kib": kib, "streams": total // chunk}
        want = None
        for piece in ((0, 65536) if quick else (0, 32768, 65536, 131072)):
            if piece == 0 and total // chunk < 64:
                continue                      # uncut, one warp parses a whole stream: minutes
            iters = 1 if piece == 0 and total // chunk < 1024 else 3
            B.set_piece_bytes(piece)
            db = B.DeviceBatch(total, chunk)
            # the same bytes whatever the chunking: the 1 GiB mixed corpus of 64 KiB units
            B.check(B.lib().lzs_b200_corpus_fill_device(db.raw.data_ptr(), 65536, 65536, 0, total // 65536, 0x5EED0002,
                                                        B.CORPUS_MIXED, db._stream()))
            before = B.lib().lzs_b200_kernel_launches()
            ms = timed(db, iters)
            cut = (B.lib().lzs_b200_kernel_launches() - before) // iters == 8
            lens = db.comp_len.cpu().numpy().copy()
            if want is None:
                want = (lens, db.comp[:db.n * db.comp_stride].clone())
            else:
                assert (lens == want[0]).all(), "lengths differ with pieces of %d" % piece
                a = db.comp[:db.n * db.comp_stride].view(db.n, db.comp_stride)
                b = want[1].view(db.n, db.comp_stride)
                if db.n <= 64:
                    same = all(bool(torch.equal(a[k, :int(lens[k])], b[k, :int(lens[k])])) for k in range(db.n))
                else:
                    live = torch.arange(db.comp_stride, device=db.device)[None, :] < db.comp_len[:, None]
                    same = not bool(((a != b) & live).any())
                    del live
                assert same, "bytes differ with pieces of %d" % piece
                del a, b
            row["piece_%d" % piece if piece else "uncut"] = {"ms": round(ms, 2), "gbs": round(total / ms / 1e6, 2), "cut": cut}
            del db
            torch.cuda.empty_cache()
        del want
        torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)
    B.set_piece_bytes(65536)

    # the decoder: pieces of the compressed stream (csrc/k4_pieces.cuh) against one group of lanes per stream
    for kib in ((1024,) if quick else (64, 128, 256, 512, 1024, 16384, 262144)):
        chunk = kib << 10
        row = {"decode_chunk_kib": kib, "streams": total // chunk}
        db = B.DeviceBatch(total, chunk)
        B.check(B.lib().lzs_b200_corpus_fill_device(db.raw.data_ptr(), 65536, 65536, 0, total // 65536, 0x5EED0002,
                                                    B.CORPUS_MIXED, db._stream()))
        db.compress()
        torch.cuda.synchronize()
        for dpiece in (0, 2048, 4096):
            if dpiece == 0 and total // chunk < 64:
                continue                      # one group of lanes per stream: minutes
            B.set_decode_piece_bytes(dpiece)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            best = 1e30
            before = B.lib().lzs_b200_kernel_launches()
            iters = 1 if dpiece == 0 and total // chunk < 1024 else 3
            for _ in range(iters):
                db.dec.zero_()
                ev[0].record(); db.decompress(); ev[1].record()
                torch.cuda.synchronize()
                best = min(best, ev[0].elapsed_time(ev[1]))
            cut = (B.lib().lzs_b200_kernel_launches() - before) // iters == 11
            assert db.roundtrip_ok(), "decode differs with pieces of %d" % dpiece
            row["dpiece_%d" % dpiece if dpiece else "uncut"] = {"ms": round(best, 2), "gbs": round(total / best / 1e6, 2), "cut": cut}
        B.set_decode_piece_bytes(2048)
        if total // chunk <= 1024:
            # a handful of streams: pointer doubling instead of the replay (lzs_b200_decompress_long_batch_device)
            best = 1e30
            for _ in range(3):
                db.dec.zero_()
                ev[0].record(); db.decompress_jump(); ev[1].record()
                torch.cuda.synchronize()
                best = min(best, ev[0].elapsed_time(ev[1]))
            assert db.roundtrip_ok(), "decode by pointer doubling differs"
            row["jump"] = {"ms": round(best, 2), "gbs": round(total / best / 1e6, 2)}
        del db
        torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)

    # the drop-in call on one large host buffer
    ref = helpers.reference() or helpers.oracle()
    for mib in (1, 16, 256):
        data = helpers.corpus(helpers.CORPUS_MIXED, mib * 16, 65536, first_index=7)
        n = len(data)
        src = np.ascontiguousarray(data)
        cap = B.compressed_max(n)
        dst = np.zeros(cap, dtype=np.uint8)
        best = 1e30
        for _ in range(3):                       # the C call alone; the first call also sizes the library's device arena
            t0 = time.perf_counter(); r = B.lib().lzs_compress(B._p(dst), cap, B._p(src), n); best = min(best, time.perf_counter() - t0)
        t0 = time.perf_counter(); want = ref.compress(data.tobytes()); t_cpu = time.perf_counter() - t0
        assert dst[:r].tobytes() == want
        comp = np.ascontiguousarray(dst[:r]); back = np.zeros(n + 16, dtype=np.uint8)
        best_d = 1e30
        for _ in range(2):
            t0 = time.perf_counter(); rd = B.lib().lzs_decompress(B._p(back), n, B._p(comp), r); best_d = min(best_d, time.perf_counter() - t0)
        assert rd == n and back[:n].tobytes() == data.tobytes()
        t0 = time.perf_counter(); ref.decompress(want, n); t_cpu_d = time.perf_counter() - t0
        print(json.dumps({"single_call_mib": mib, "lzs_compress_ms": round(best * 1e3, 2), "gbs": round(n / best / 1e9, 3),
                          "reference_one_core_gbs": round(n / t_cpu / 1e9, 3), "lzs_decompress_ms": round(best_d * 1e3, 2),
                          "decompress_gbs": round(n / best_d / 1e9, 3), "reference_one_core_decompress_gbs": round(n / t_cpu_d / 1e9, 3),
                          "host_memory": "pageable"}), flush=True)


if __name__ == "__main__":
    main()
