"""N>1 host logic on CPU: two gloo ranks shard a corpus by stream index, each produces its
streams (with the oracle standing in for the GPU, which this container does not have), and
the all-gather-v of variable-size outputs must reassemble, on every rank, exactly the
streams a single process would produce -- and they must decode back to the corpus."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "lzs-compression_b200", "python"))

import helpers  # noqa: E402

CHUNK, N_CHUNKS = 3000, 37


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import lzs_dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = helpers.oracle()
    lo, hi = lzs_dist.shard_range(N_CHUNKS, rank, world)
    raw = helpers.corpus(helpers.CORPUS_MIXED, hi - lo, CHUNK, seed=0x5EED0000 + 4, first_index=lo)
    stride = (helpers.compressed_max(CHUNK) + 15) // 16 * 16
    comp = np.zeros((hi - lo) * stride, dtype=np.uint8)
    lens = np.zeros(hi - lo, dtype=np.int64)
    for s in range(hi - lo):
        c = o.compress(raw[s * CHUNK:(s + 1) * CHUNK].tobytes())
        comp[s * stride:s * stride + len(c)] = np.frombuffer(c, dtype=np.uint8)
        lens[s] = len(c)
    off = torch.arange(hi - lo, dtype=torch.int64) * stride
    packed, packed_off = lzs_dist.pack_streams(torch.from_numpy(comp), off, torch.from_numpy(lens))
    assert packed.numel() % 16 == 0 and all(int(x) % 16 == 0 for x in packed_off)
    payload, all_len, all_off = lzs_dist.all_gather_streams(packed, torch.from_numpy(lens))
    # the one-call form (packs and gathers; on CPU tensors it takes the exact-size path) must agree
    p2, l2, o2, info = lzs_dist.gather_compressed(torch.from_numpy(comp), off, torch.from_numpy(lens))
    assert torch.equal(p2, payload) and torch.equal(l2, all_len) and torch.equal(o2, all_off) and "exact" in info["mode"]
    torch.save({"payload": payload, "len": all_len, "off": all_off, "range": (lo, hi)}, out_path % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather(tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / "rank%d.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    o = helpers.oracle()
    whole = helpers.corpus(helpers.CORPUS_MIXED, N_CHUNKS, CHUNK, seed=0x5EED0000 + 4)
    want = [o.compress(whole[s * CHUNK:(s + 1) * CHUNK].tobytes()) for s in range(N_CHUNKS)]
    ranges = []
    for r in range(world):
        d = torch.load(out % r)
        ranges.append(d["range"])
        assert d["len"].numel() == N_CHUNKS
        pay = d["payload"].numpy()
        for s in range(N_CHUNKS):
            a, l = int(d["off"][s]), int(d["len"][s])
            assert pay[a:a + l].tobytes() == want[s], (r, s)
        # the gathered payload is a plain concatenation of streams: the reference's incremental
        # decoder walks it marker by marker (lzs-decompression.c:564-576); spot-check one stream
        a, l = int(d["off"][5]), int(d["len"][5])
        assert o.decompress(pay[a:a + l].tobytes(), CHUNK) == whole[5 * CHUNK:6 * CHUNK].tobytes()
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == N_CHUNKS
