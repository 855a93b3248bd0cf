"""Parity tests proper: the sm_100a kernels, called through the C ABI (liblzs.so),
against the oracle / committed reference outputs.  Bit-exact, no tolerance: everything
on this path is integer and byte work."""
import os

import numpy as np
import pytest

import helpers
from gpu_common import binding

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    b = binding()
    b.lib()
    assert b.lib().lzs_b200_device_count() > 0, "no CUDA device visible to liblzs.so"
    return b


def _golden(name):
    return open(os.path.join(helpers.GOLDEN_DIR, name), "rb").read()


def length_bits(rep):
    if rep <= 1:
        return 9 * rep
    if rep <= 4:
        return 2
    if rep <= 7:
        return 4
    return ((rep + 22) // 15) * 4


def test_reference_golden_vector_single_call(B):
    """c/src/test/test-lzs-decompression.c:34-96 through lzs_decompress AND lzs_compress."""
    comp, plain = _golden("golden1_compressed.bin"), _golden("golden1_plain.bin")
    assert B.lzs_decompress(comp, len(plain) + 520) == plain
    assert B.lzs_compress(plain) == comp
    assert B.lzs_simple_compress(plain) == comp
    assert B.lzs_compress(b"") == bytes([0xC0, 0x00])
    assert B.lzs_compress(b"a") == bytes([0x30, 0xE0, 0x00])


def test_reference_size_law_uncompressible(B):
    """c/src/test/test-lzs.c:93-119 for every prefix length, as one batch."""
    seq = _golden("uncompressible.bin")
    data = [seq[:n] for n in range(len(seq) + 1)]
    comp = B.compress_streams(data, caps=[1000] * len(data))
    for n, c in enumerate(comp):
        assert len(c) == (n * 9 + 9 + 7) // 8
    assert B.decompress_streams(comp, [1000] * len(data)) == data


def test_reference_size_law_repeated_byte(B):
    """c/src/test/test-lzs.c:121-167 for n = 0..1000, as one batch."""
    data = [b"X" * n for n in range(1001)]
    comp = B.compress_streams(data, caps=[1000] * len(data))
    for n, c in enumerate(comp):
        bits = 0 if n == 0 else 9 if n == 1 else 18 if n == 2 else 9 + 2 + 7 + length_bits(n - 1)
        assert len(c) == (bits + 9 + 7) // 8, n
    assert B.decompress_streams(comp, [1000] * len(data)) == data


def test_committed_reference_outputs(B):
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "ref_cases.npz"))
    names = [k[4:] for k in z.files if k.startswith("in__")]
    data = [z["in__" + k].tobytes() for k in names]
    want = [z["out__" + k].tobytes() for k in names]
    assert B.compress_streams(data) == want
    assert B.decompress_streams(want, [len(d) + 9 for d in data]) == data
    keys = [k for k in z.files if k.startswith("dec_in__")]
    got = B.decompress_streams([z[k].tobytes() for k in keys], [int(k.split("__")[2]) for k in keys])
    for k, g in zip(keys, got):
        assert g == z["dec_out__" + k[len("dec_in__"):]].tobytes(), k


def test_edge_cases_and_truncated_output(B):
    o = helpers.oracle()
    cases = helpers.edge_case_inputs()
    data = list(cases.values())
    full = [o.compress(d) for d in data]
    assert B.compress_streams(data) == full
    caps = [max(0, len(f) - 1 - (i % 7)) for i, f in enumerate(full)]
    assert B.compress_streams(data, caps=caps) == [f[:c] for f, c in zip(full, caps)]
    half = [len(d) // 2 for d in data]
    assert B.decompress_streams(full, half) == [d[:h] for d, h in zip(data, half)]


def test_random_bitstrings_decode_like_the_reference(B):
    rng = np.random.default_rng(99)
    o = helpers.oracle()
    streams = [rng.integers(0, 256, int(rng.integers(0, 500)), dtype=np.uint8).tobytes() for _ in range(400)]
    caps = [int(rng.integers(0, 6000)) for _ in streams]
    got = B.decompress_streams(streams, caps)
    for s, c, g in zip(streams, caps, got):
        assert g == o.decompress(s, c)


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_decoder_lane_widths(B, lanes):
    o = helpers.oracle()
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, 20000, first_index=i).tobytes() for i in range(9)]
    comp = [o.compress(d) for d in data]
    B.check(B.lib().lzs_b200_set_decode_lanes(lanes))
    try:
        assert B.decompress_streams(comp, [len(d) for d in data]) == data
    finally:
        B.lib().lzs_b200_set_decode_lanes(8)


def _cpu_compress_all(chunks):
    """Reference build when present (fast), else the oracle restatement."""
    c = helpers.reference() or helpers.oracle()
    return [c.compress(x) for x in chunks]


@pytest.mark.parametrize("kind,chunk,count", [
    (helpers.CORPUS_TEXT, 65536, 6), (helpers.CORPUS_BINARY, 65536, 6), (helpers.CORPUS_RANDOM, 65536, 3),
    (helpers.CORPUS_MIXED, 4096, 48), (helpers.CORPUS_MIXED, 262144, 3), (helpers.CORPUS_PACKET, 1500, 600),
    (helpers.CORPUS_MIXED, 1048576, 1),
])
def test_corpus_streams_bit_exact(B, kind, chunk, count):
    buf = helpers.corpus(kind, count, chunk, seed=0x5EED0000 + chunk)
    chunks = [buf[i * chunk:(i + 1) * chunk].tobytes() for i in range(count)]
    want = _cpu_compress_all(chunks)
    got = B.compress_streams(chunks)
    assert got == want
    assert B.decompress_streams(got, [chunk] * count) == chunks


def test_match_table_equals_bruteforce_rule(B):
    """K1 alone: per-position (length, offset) against the oracle's statement of the rule."""
    import torch
    chunk, count = 20000, 6
    buf = helpers.corpus(helpers.CORPUS_MIXED, count, chunk, seed=0x5EED0000 + 77)
    db = B.DeviceBatch(chunk * count, chunk)
    db.raw[:chunk * count].copy_(torch.from_numpy(buf))
    db.match_only()
    torch.cuda.synchronize()
    m = db.scratch[256:256 + 2 * chunk * count].cpu().numpy().view(np.uint16)
    for i in range(count):
        ln, off = helpers.oracle_all_matches(buf[i * chunk:(i + 1) * chunk].tobytes())
        want = (ln.astype(np.uint16) << 11) | off
        assert (m[i * chunk:(i + 1) * chunk] == want).all(), i


def test_device_corpus_matches_host_corpus(B):
    import torch
    for kind in (B.CORPUS_TEXT, B.CORPUS_BINARY, B.CORPUS_RANDOM, B.CORPUS_MIXED, B.CORPUS_PACKET):
        db = B.DeviceBatch(6 * 3000, 3000)
        db.fill(kind, 0x5EED0000 + 5, first_index=7)
        torch.cuda.synchronize()
        host = helpers.corpus(kind, 6, 3000, seed=0x5EED0000 + 5, first_index=7)
        assert (db.raw[:18000].cpu().numpy() == host).all()


def _compare_all_streams(B, db, what):
    """Every compressed stream the GPU left in db.comp against the unmodified reference (all host
    threads, oracle/chunk_driver.c) over the same bytes: all lengths, then every byte."""
    ref = helpers.reference() or helpers.oracle()
    n, chunk, stride = db.n, db.chunk, db.comp_stride
    raw = np.zeros(db.total + 16, dtype=np.uint8)
    raw[:db.total] = db.raw[:db.total].cpu().numpy()
    idx = np.arange(n, dtype=np.uint64)
    in_off = idx * np.uint64(chunk)
    in_len = db.raw_len.cpu().numpy().astype(np.uint32)
    c_off, c_cap = idx * np.uint64(stride), np.full(n, stride, dtype=np.uint32)
    cpu = np.zeros(n * stride + 16, dtype=np.uint8)
    cpu_len, _ = ref.run_streams(False, raw, in_off, in_len, cpu, c_off, c_cap, os.cpu_count() or 1)
    gpu_len = db.comp_len.cpu().numpy().astype(np.uint32)
    bad = np.nonzero(gpu_len != cpu_len)[0]
    assert bad.size == 0, "%s: %d of %d stream lengths differ, first at stream %d" % (what, bad.size, n, bad[0])
    gpu = db.comp[:n * stride].cpu().numpy().reshape(n, stride)
    cpu = cpu[:n * stride].reshape(n, stride)
    cols = np.arange(stride, dtype=np.uint32)[None, :]
    step = max(1, (64 << 20) // stride)
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        live = cols < gpu_len[lo:hi, None]
        diff = (gpu[lo:hi] != cpu[lo:hi]) & live
        assert not diff.any(), "%s: stream %d differs from the reference" % (what, lo + int(np.nonzero(diff.any(axis=1))[0][0]))
    return n


def test_full_size_roundtrip_1gib_64k_chunks(B):
    """BASELINE config 2 at full size: the round trip, sizes within LZS_COMPRESSED_MAX, and ALL
    16 384 compressed streams byte for byte against the unmodified reference."""
    import torch
    total, chunk = 1 << 30, 65536
    db = B.DeviceBatch(total, chunk)
    db.fill(B.CORPUS_MIXED, 0x5EED0000 + 2)
    db.compress()
    db.decompress()
    torch.cuda.synchronize()
    assert db.roundtrip_ok()
    lens = db.comp_len.cpu().numpy()
    assert lens.max() <= B.compressed_max(chunk) and lens.min() > 2
    assert _compare_all_streams(B, db, "config 2") == 16384


def test_full_size_packets_roundtrip(B):
    """BASELINE config 3 shape: 1 Mi packets of 1500 bytes, each its own stream; the round trip and
    ALL streams byte for byte against the unmodified reference."""
    import torch
    n, plen = 1 << 20, 1500
    db = B.DeviceBatch(n * plen, plen)
    db.fill(B.CORPUS_PACKET, 0x5EED0000 + 3)
    db.compress()
    db.decompress()
    torch.cuda.synchronize()
    assert db.roundtrip_ok()
    assert _compare_all_streams(B, db, "config 3") == n


def test_exact_match_finder_launch_runs_on_hardware(B):
    """K1's fast insert assumes sm_100a serves same-address shared-memory exchanges in ascending
    lane order and is backed by a launch that is exact for any order (k1_match<true>), which
    normally returns at once.  Forced here, that launch recomputes every record on the GPU: the
    records must equal the fast launch's, and the streams the reference's."""
    import torch
    L = B.lib()
    L.lzs_b200_set_force_safe_match.argtypes = [__import__("ctypes").c_int]
    total, chunk = 24 << 20, 65536
    db = B.DeviceBatch(total, chunk)
    db.fill(B.CORPUS_MIXED, 0x5EED0000 + 9)
    db.match_only()
    torch.cuda.synchronize()
    fast = db.scratch[256:256 + 2 * total].clone()
    launches = L.lzs_b200_kernel_launches()
    try:
        L.lzs_b200_set_force_safe_match(1)
        db.scratch[256:256 + 2 * total].zero_()
        db.match_only()
        torch.cuda.synchronize()
        assert torch.equal(db.scratch[256:256 + 2 * total], fast), "the exact launch disagrees with the fast launch"
        db.compress()
        db.decompress()
        torch.cuda.synchronize()
    finally:
        L.lzs_b200_set_force_safe_match(0)
    assert L.lzs_b200_kernel_launches() > launches
    assert db.roundtrip_ok()
    _compare_all_streams(B, db, "forced exact launch")
    packets = B.DeviceBatch(3000 * 1500, 1500)
    packets.fill(B.CORPUS_PACKET, 0x5EED0000 + 10)
    try:
        L.lzs_b200_set_force_safe_match(1)
        packets.compress()
        torch.cuda.synchronize()
    finally:
        L.lzs_b200_set_force_safe_match(0)
    _compare_all_streams(B, packets, "forced exact launch, packets")


def test_packed_host_compression(B):
    """lzs_b200_compress_packed_host: same streams as the slot layout, back to back at multiples
    of 16, and the packed buffer feeds lzs_b200_decompress_batch_host directly."""
    o = helpers.oracle()
    rng = np.random.default_rng(8)
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, int(rng.integers(0, 9000)), first_index=i).tobytes()
            for i in range(300)]
    packed, off, ln = B.compress_streams_packed(data)
    want = [o.compress(d) for d in data]
    assert [packed[int(a):int(a) + int(l)].tobytes() for a, l in zip(off, ln)] == want
    assert all(int(a) % 16 == 0 for a in off) and int(off[-1]) + int(ln[-1]) <= len(packed)
    assert all(int(off[i + 1]) - int(off[i]) < int(ln[i]) + 16 for i in range(len(off) - 1))      # really packed
    dst = np.zeros(sum(len(d) for d in data) + 64, dtype=np.uint8)
    raw_off, raw_len, raw_span = B.layout([len(d) for d in data], align=1)
    d_len = np.zeros(len(data), dtype=np.uint32)
    B.check(B.lib().lzs_b200_decompress_batch_host(B._p(packed), B._p(off, B.u64p), B._p(ln, B.u32p), len(packed),
                                                   B._p(dst), B._p(raw_off, B.u64p), B._p(raw_len, B.u32p),
                                                   B._p(d_len, B.u32p), raw_span, len(data)))
    assert (d_len == raw_len).all()
    assert dst[:raw_span].tobytes() == b"".join(data)


def test_structured_fuzz_against_cpu_codec(B):
    """2000 structured random streams in one batch (alphabets of 2/3/20/256 symbols, copies planted
    at distances around the window edge, runs, sizes 0..12000) -- bit-exact compress, and the
    decoder on the compressed streams, on truncated streams and on bit-flipped streams."""
    cpu = helpers.reference() or helpers.oracle()
    rng = np.random.default_rng(4242)
    data = []
    for it in range(2000):
        n = int(rng.integers(0, 12000)) if it % 10 else int(rng.integers(0, 40))
        alpha = int(rng.choice([2, 3, 20, 256]))
        buf = bytearray(rng.integers(0, alpha, n, dtype=np.uint8).tobytes())
        for _ in range(int(rng.integers(0, 8))):
            if n < 2200:
                break
            d = int(rng.choice([1, 2, 3, 7, 127, 128, 2040, 2046, 2047, 2048, 2060]))
            p = int(rng.integers(d, n - 1))
            ln = int(rng.integers(2, 300))
            buf[p:p + ln] = buf[p - d:p - d + ln]
        if it % 7 == 0 and n > 100:
            p = int(rng.integers(0, n - 50))
            buf[p:p + 40] = bytes([int(rng.integers(0, 256))]) * 40
        data.append(bytes(buf[:n]))
    want = [cpu.compress(d) for d in data]
    got = B.compress_streams(data)
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if g != w]
    assert not bad, bad[:10]
    assert B.decompress_streams(got, [len(d) for d in data]) == data
    # damaged streams: truncate, flip a bit; capacity sometimes too small
    damaged, caps = [], []
    for i, w in enumerate(want[:600]):
        s = bytearray(w)
        if i % 3 == 0 and len(s) > 3:
            s = s[:int(rng.integers(1, len(s)))]
        elif i % 3 == 1 and len(s) > 0:
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        damaged.append(bytes(s))
        caps.append(int(rng.integers(0, 2 * len(data[i]) + 50)))
    got = B.decompress_streams(damaged, caps)
    for s, c, g in zip(damaged, caps, got):
        assert g == cpu.decompress(s, c)


def test_decode_status_words_against_oracle():
    """per-stream stop reasons of the batch decoder on valid, truncated,
    capacity-limited and random streams, against the oracle's statement of the reference loop."""
    B = binding()
    o = helpers.oracle()
    rng = np.random.default_rng(23)
    streams, caps = [], []
    for i in range(300):
        kind = [helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_PACKET, helpers.CORPUS_RANDOM][i % 4]
        d = helpers.corpus(kind, 1, int(rng.integers(0, 3000)), first_index=i).tobytes()
        c = o.compress(d)
        mode = i % 5
        if mode == 0:
            streams.append(c); caps.append(len(d) + int(rng.integers(0, 40)))
        elif mode == 1:
            streams.append(c); caps.append(int(rng.integers(0, len(d) + 1)))
        elif mode == 2:
            streams.append(c[:int(rng.integers(0, len(c) + 1))]); caps.append(len(d) + 16)
        elif mode == 3:
            streams.append(c + bytes(rng.integers(0, 256, 5, dtype=np.uint8))); caps.append(len(d))
        else:
            streams.append(bytes(rng.integers(0, 256, int(rng.integers(0, 500)), dtype=np.uint8)))
            caps.append(int(rng.integers(0, 4000)))
    got, status = B.decompress_streams_status(streams, caps)
    seen = set()
    for s, c, g, st in zip(streams, caps, got, status):
        want, why = o.decompress_status(s, c)
        assert g == want and st == why, (len(s), c, st, why)
        seen.add(why)
    assert seen == {0x01, 0x04, 0x08}


def test_torch_front_end_ragged_tensors():
    """ragged batches in torch.uint8 device tensors through lzs_torch
    (current stream, no host round trip) give the oracle's bytes and decode back."""
    import torch
    from gpu_common import torch_front_end
    lzs_torch = torch_front_end()
    o = helpers.oracle()
    rng = np.random.default_rng(31)
    data = [helpers.corpus([helpers.CORPUS_TEXT, helpers.CORPUS_PACKET, helpers.CORPUS_BINARY][i % 3], 1,
                           int(rng.integers(0, 5000)), first_index=i).tobytes() for i in range(64)]
    off, pos = [], 0
    for d in data:
        off.append(pos)
        pos += (len(d) + 15) // 16 * 16
    buf = np.zeros(pos + 64, dtype=np.uint8)
    for a, d in zip(off, data):
        buf[a:a + len(d)] = np.frombuffer(d, dtype=np.uint8)
    dev = torch.device("cuda:0")
    t_buf = torch.from_numpy(buf).to(dev)
    t_off = torch.tensor(off, dtype=torch.int64, device=dev)
    t_len = torch.tensor([len(d) for d in data], dtype=torch.int32, device=dev)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):                       # any stream: the ops follow torch's current one
        comp, comp_off, comp_len = lzs_torch.compress(t_buf, t_off, t_len)
        plain, plain_len, status = lzs_torch.decompress(comp, comp_off, comp_len, t_off, t_len, with_status=True)
    side.synchronize()
    c, co, cl = comp.cpu().numpy(), comp_off.cpu().numpy(), comp_len.cpu().numpy()
    for d, a, l in zip(data, co, cl):
        assert c[int(a):int(a) + int(l)].tobytes() == o.compress(d)
    p, pl = plain.cpu().numpy(), plain_len.cpu().numpy()
    for d, a, l in zip(data, off, pl):
        assert p[a:a + int(l)].tobytes() == d
    # a stream decoded into exactly its size stops at "output full" unless it is empty (marker seen first)
    assert all(int(s) in (0x04, 0x08) for s in status.cpu().numpy())


@pytest.mark.parametrize("n_streams,gap", [(5, 7), (40, 3), (300, 16), (300, 0)])
def test_host_batches_touch_nothing_outside_the_produced_bytes_slots(B, n_streams, gap):
    """The host entry points write the out_len[s] bytes of every stream (and at most the rest of
    its own slot) -- never a byte between slots, where a caller may keep framing.  Covers the
    few-streams path, the sliced path with one copy per run of touching slots, and the path
    through the pinned staging buffer (hundreds of slots with gaps)."""
    L = B.lib()
    o = helpers.oracle()
    rng = np.random.default_rng(n_streams * 31 + gap)
    data = [helpers.corpus(helpers.CORPUS_PACKET, 1, int(rng.integers(1, 1500)), first_index=i).tobytes()
            for i in range(n_streams)]
    want = [o.compress(d) for d in data]
    in_off, in_len, in_span = B.layout([len(d) for d in data])
    src = np.zeros(in_span + 64, dtype=np.uint8)
    for off, d in zip(in_off, data):
        src[int(off):int(off) + len(d)] = np.frombuffer(d, dtype=np.uint8)
    caps = np.array([helpers.compressed_max(len(d)) for d in data], dtype=np.uint32)
    out_off = np.zeros(n_streams, dtype=np.uint64)
    pos = 5
    for s in range(n_streams):
        out_off[s] = pos
        pos += int(caps[s]) + gap
    out_span = pos
    dst = np.full(out_span + 64, 0xA5, dtype=np.uint8)
    out_len = np.zeros(n_streams, dtype=np.uint32)
    B.check(L.lzs_b200_compress_batch_host(B._p(src), B._p(in_off, B.u64p), B._p(in_len, B.u32p), in_span, B._p(dst),
                                           B._p(out_off, B.u64p), B._p(caps, B.u32p), B._p(out_len, B.u32p), out_span,
                                           n_streams))
    covered = np.zeros(len(dst), dtype=bool)
    for s in range(n_streams):
        a, l, c = int(out_off[s]), int(out_len[s]), int(caps[s])
        assert dst[a:a + l].tobytes() == want[s], s
        covered[a:a + c] = True
    assert (dst[~covered] == 0xA5).all(), "bytes outside every slot were overwritten"
    # and back, into slots with the same gaps
    comp_off, comp_len = out_off, out_len
    dec_cap = in_len.copy()
    dec_off = np.zeros(n_streams, dtype=np.uint64)
    pos = 3
    for s in range(n_streams):
        dec_off[s] = pos
        pos += int(dec_cap[s]) + gap
    back = np.full(pos + 64, 0x5A, dtype=np.uint8)
    dec_len = np.zeros(n_streams, dtype=np.uint32)
    B.check(L.lzs_b200_decompress_batch_host(B._p(dst), B._p(comp_off, B.u64p), B._p(comp_len, B.u32p), out_span,
                                             B._p(back), B._p(dec_off, B.u64p), B._p(dec_cap, B.u32p),
                                             B._p(dec_len, B.u32p), pos, n_streams))
    covered = np.zeros(len(back), dtype=bool)
    for s in range(n_streams):
        a = int(dec_off[s])
        assert back[a:a + int(dec_len[s])].tobytes() == data[s], s
        covered[a:a + int(dec_cap[s])] = True
    assert (back[~covered] == 0x5A).all()
    assert L.lzs_b200_release() == 0


def test_decoder_writes_straight_into_pinned_caller_memory(B):
    """lzs_b200_decompress_batch_host on a pinned (device-mapped) output buffer can decode straight
    into it over PCIe, no staging copy (an option, lzs_b200_set_zero_copy_output): same bytes, same lengths, and still not a byte outside
    the produced ranges (sentinels between and inside the slots)."""
    import ctypes
    import torch
    L = B.lib()
    n = 700
    rng = np.random.default_rng(77)
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, int(rng.integers(1, 9000)), first_index=i).tobytes() for i in range(n)]
    packed, off, ln = B.compress_streams_packed(data)
    cap = np.array([len(d) + 40 for d in data], dtype=np.uint32)          # slots larger than what is produced
    out_off = np.zeros(n, dtype=np.uint64)
    pos = 11
    for s in range(n):
        out_off[s] = pos
        pos += int(cap[s]) + 5
    span = pos
    pinned = torch.full((span + 64,), 0x3C, dtype=torch.uint8).pin_memory()
    src = torch.from_numpy(packed.copy()).pin_memory()
    d_len = np.zeros(n, dtype=np.uint32)
    u8 = B.u8p
    L.lzs_b200_set_zero_copy_output.argtypes = [ctypes.c_int]
    L.lzs_b200_set_zero_copy_output(1)
    try:
        B.check(L.lzs_b200_decompress_batch_host(ctypes.cast(src.data_ptr(), u8), B._p(off, B.u64p), B._p(ln, B.u32p),
                                                 len(packed), ctypes.cast(pinned.data_ptr(), u8), B._p(out_off, B.u64p),
                                                 B._p(cap, B.u32p), B._p(d_len, B.u32p), span, n))
    finally:
        L.lzs_b200_set_zero_copy_output(0)
    got = pinned.numpy()
    written = np.zeros(len(got), dtype=bool)
    for s in range(n):
        a = int(out_off[s])
        assert int(d_len[s]) == len(data[s]) and got[a:a + len(data[s])].tobytes() == data[s], s
        written[a:a + len(data[s])] = True
    assert (got[~written] == 0x3C).all(), "the decoder wrote outside the bytes it produced"


def test_registered_torch_operators():
    """torch.ops.lzs_b200.compress / decompress (torch.library custom ops over the same device entry
    points): same bytes as the oracle, status words from the decoder."""
    import torch
    from gpu_common import torch_front_end
    T = torch_front_end()
    if T.compress_op is None:
        pytest.skip("this torch has no torch.library.custom_op")
    o = helpers.oracle()
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, n, first_index=i).tobytes() for i, n in enumerate([1, 700, 4096, 65536, 33])]
    off, pos = [], 0
    for d in data:
        off.append(pos)
        pos += (len(d) + 15) // 16 * 16
    buf = torch.zeros(pos + 64, dtype=torch.uint8)
    for a, d in zip(off, data):
        buf[a:a + len(d)] = torch.frombuffer(bytearray(d), dtype=torch.uint8)
    dev = torch.device("cuda:0")
    t_off = torch.tensor(off, dtype=torch.int64, device=dev)
    t_len = torch.tensor([len(d) for d in data], dtype=torch.int32, device=dev)
    comp, comp_off, comp_len = torch.ops.lzs_b200.compress(buf.to(dev), t_off, t_len)
    torch.cuda.synchronize()
    c, co, cl = comp.cpu().numpy(), comp_off.cpu().numpy(), comp_len.cpu().numpy()
    assert [c[int(a):int(a) + int(l)].tobytes() for a, l in zip(co, cl)] == [o.compress(d) for d in data]
    out, out_len, status = torch.ops.lzs_b200.decompress(comp, comp_off, comp_len, t_off, t_len)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert [got[a:a + len(d)].tobytes() for a, d in zip(off, data)] == data
    assert out_len.cpu().tolist() == [len(d) for d in data]
    assert all(s in (0x04, 0x08) for s in status.cpu().tolist())      # end marker seen, or capacity reached exactly at it


# ---------------------------------------------------------------- long streams cut into pieces

def _reference_streams(data):
    ref = helpers.reference() or helpers.oracle()
    return [ref.compress(d) for d in data]


def test_long_streams_small_pieces_all_streams_against_reference(B):
    """csrc/k23_pieces.cuh with a piece size far below the stream size, so that 16 borders fall into
    every stream of a 64 MiB batch of 64 KiB chunks: all 1024 streams byte for byte against the
    unmodified reference, and the round trip.  The launch counter says the piece path really ran."""
    import torch
    total, chunk = 64 << 20, 65536
    B.set_piece_bytes(4096)
    try:
        db = B.DeviceBatch(total, chunk)
        db.fill(B.CORPUS_MIXED, 0x5EED0000 + 21)
        before = B.lib().lzs_b200_kernel_launches()
        db.compress()
        assert B.lib().lzs_b200_kernel_launches() - before == 8      # plan x2, K1 x2, spec, fix, sweep, pack
        db.decompress()
        torch.cuda.synchronize()
        assert db.roundtrip_ok()
        assert _compare_all_streams(B, db, "pieces of 4 KiB") == 1024
    finally:
        B.set_piece_bytes(65536)


@pytest.mark.parametrize("kind", ["mixed", "text", "packets"])
def test_long_streams_1mib_chunks_default_pieces(B, kind):
    """BASELINE config 5's large end: 1 MiB chunks are cut into 64 KiB pieces by default."""
    import torch
    total, chunk = 96 << 20, 1 << 20
    db = B.DeviceBatch(total, chunk)
    db.fill({"mixed": B.CORPUS_MIXED, "text": B.CORPUS_TEXT, "packets": B.CORPUS_PACKET}[kind], 0x5EED0000 + 22)
    before = B.lib().lzs_b200_kernel_launches()
    db.compress()
    assert B.lib().lzs_b200_kernel_launches() - before == 8
    db.decompress()
    torch.cuda.synchronize()
    assert db.roundtrip_ok()
    assert _compare_all_streams(B, db, "1 MiB chunks") == 96


def test_one_large_buffer_through_the_drop_in_calls(B):
    """What a caller of lzs.h does: ONE lzs_compress call on one large buffer (here 24 MiB with runs
    of zeros of up to 3 MiB, text, records and noise in it) -- cut into pieces inside, one stream
    outside, equal to the reference's; and a batch of three unequal long streams through the host
    batch call."""
    rng = np.random.default_rng(23)
    parts = []
    for i in range(40):
        k = i % 5
        if k == 0:
            parts.append(bytes(int(rng.integers(1000, 3 << 20))))
        elif k == 4:
            parts.append(rng.integers(0, 256, int(rng.integers(1000, 300000)), dtype=np.uint8).tobytes())
        else:
            parts.append(helpers.corpus((helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)[k - 1], 1,
                                        int(rng.integers(1000, 900000)), first_index=i).tobytes())
    big = b"".join(parts)[:24 << 20]
    want = _reference_streams([big])[0]
    got = B.lzs_compress(big)
    assert got == want
    assert B.lzs_decompress(got, len(big)) == big
    three = [big[:5_000_001], big[5_000_001:5_700_000], big[9_000_000:9_000_000 + (3 << 20) + 17]]
    assert B.compress_streams(three) == _reference_streams(three)
    # a capacity smaller than the stream: the prefix that fits
    assert B.compress_streams([three[1]], caps=[100_001]) == [_reference_streams([three[1]])[0][:100_001]]
    # around the size from which one buffer is cut (two pieces of 64 KiB), and piece borders +- 1
    for n in (131071, 131072, 131073, 196607, 196608, 196609, 262145):
        assert B.lzs_compress(big[7_000_000:7_000_000 + n]) == _reference_streams([big[7_000_000:7_000_000 + n]])[0], n


def test_long_streams_through_the_sliced_host_pipeline(B):
    """More than eight ordered long streams take the host path's slices (upload / kernels / download
    overlapped); the pieces' table is shared by the slices."""
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, (1 << 20) + 4099 * i, first_index=40 + i).tobytes() for i in range(12)]
    got = B.compress_streams(data)
    assert got == _reference_streams(data)
    assert B.decompress_streams(got, [len(d) for d in data]) == data


# ---------------------------------------------------------------- the decoder for long streams (csrc/k4_pieces.cuh)

def test_long_streams_decoder_small_pieces(B):
    """Token starts found in parallel inside the streams: pieces of 256 compressed bytes put ~180 borders
    into every stream of a 64 MiB batch of 64 KiB chunks.  The launch counter says the piece passes ran."""
    import torch
    B.set_decode_piece_bytes(256)
    try:
        db = B.DeviceBatch(64 << 20, 65536)
        db.fill(B.CORPUS_MIXED, 0x5EED0000 + 31)
        db.compress()
        before = B.lib().lzs_b200_kernel_launches()
        db.decompress()
        assert B.lib().lzs_b200_kernel_launches() - before == 11  # plan, spec, fix x4, sweep, emit, copy, dirty list, k4_decode for the dirty
        torch.cuda.synchronize()
        assert db.roundtrip_ok()
    finally:
        B.set_decode_piece_bytes(2048)


@pytest.mark.parametrize("kind", ["mixed", "text", "random"])
def test_long_streams_decoder_1mib_chunks(B, kind):
    import torch
    db = B.DeviceBatch(96 << 20, 1 << 20)
    db.fill({"mixed": B.CORPUS_MIXED, "text": B.CORPUS_TEXT, "random": B.CORPUS_RANDOM}[kind], 0x5EED0000 + 32)
    db.compress()
    before = B.lib().lzs_b200_kernel_launches()
    db.decompress()
    assert B.lib().lzs_b200_kernel_launches() - before == 11
    torch.cuda.synchronize()
    assert db.roundtrip_ok()


def test_long_streams_decoder_leaves_dirty_streams_to_the_serial_decoder(B):
    """Damaged streams, random bits, outputs that are too small, runs of zeros and clean streams in ONE
    batch through the piece passes (pieces of 16 bytes): bytes and lengths as the reference gives them."""
    ref = helpers.reference() or helpers.oracle()
    rng = np.random.default_rng(41)
    data = [helpers.corpus(helpers.CORPUS_MIXED, 1, int(rng.integers(200, 9000)), first_index=i).tobytes() for i in range(60)]
    data += [bytes(70000), b"ab" * 30000, b""]
    comp = [ref.compress(d) for d in data]
    streams, caps = [], []
    for i, (c, d) in enumerate(zip(comp, data)):
        k = i % 5
        if k == 0:   streams.append(c); caps.append(len(d))                     # clean, output exactly full
        elif k == 1: streams.append(c); caps.append(len(d) + 50)                # clean
        elif k == 2: streams.append(c[:len(c) // 2]); caps.append(len(d))       # cut off
        elif k == 3: streams.append(c); caps.append(len(d) // 3)                # output too small
        else:
            bad = bytearray(c)
            if bad: bad[len(bad) // 3] ^= 0x5A
            streams.append(bytes(bad)); caps.append(len(d) + 100)               # damaged
    streams += [rng.integers(0, 256, 300, dtype=np.uint8).tobytes() for _ in range(20)]
    caps += [2000] * 20
    B.set_decode_piece_bytes(16)
    try:
        got = B.decompress_streams(streams, caps)
    finally:
        B.set_decode_piece_bytes(2048)
    for i, (s, c, g) in enumerate(zip(streams, caps, got)):
        assert g == ref.decompress(s, c), i


def test_one_large_buffer_decoded_through_the_drop_in_call(B):
    """lzs_decompress on ONE stream of 24 MiB (text, records, noise, runs of zeros of up to 3 MiB)."""
    rng = np.random.default_rng(23)
    parts = []
    for i in range(40):
        k = i % 5
        if k == 0:
            parts.append(bytes(int(rng.integers(1000, 3 << 20))))
        elif k == 4:
            parts.append(rng.integers(0, 256, int(rng.integers(1000, 300000)), dtype=np.uint8).tobytes())
        else:
            parts.append(helpers.corpus((helpers.CORPUS_TEXT, helpers.CORPUS_BINARY, helpers.CORPUS_MIXED)[k - 1], 1,
                                        int(rng.integers(1000, 900000)), first_index=i).tobytes())
    big = b"".join(parts)[:24 << 20]
    comp = _reference_streams([big])[0]
    assert B.lzs_decompress(comp, len(big)) == big
    assert B.lzs_decompress(comp, len(big) + 1000) == big
    assert B.lzs_decompress(comp, 1 << 20) == big[:1 << 20]                      # too small: the serial decoder's prefix
    assert B.lzs_decompress(comp[:len(comp) // 2], len(big)) == (helpers.reference() or helpers.oracle()).decompress(comp[:len(comp) // 2], len(big))


def test_long_streams_structured_fuzz_both_directions(B):
    """Runs, periods, copies, small alphabets, noise and text in 300 streams of 2-40 KB through the piece
    compressor (pieces of 1 KiB) and the piece decoder (pieces of 64 bytes), clean and damaged, against
    the unmodified reference."""
    import test_pieces
    ref = helpers.reference() or helpers.oracle()
    rng = np.random.default_rng(91)
    data = [test_pieces._structured(rng, int(rng.integers(2000, 40000))) for _ in range(300)]
    want = [ref.compress(d) for d in data]
    B.set_piece_bytes(1024)
    B.set_decode_piece_bytes(64)
    try:
        got = B.compress_streams(data)
        assert got == want
        streams, caps = [], []
        for i, (c, d) in enumerate(zip(want, data)):
            k = i % 6
            if k == 0:   streams.append(c); caps.append(len(d))
            elif k == 1: streams.append(c); caps.append(len(d) + 33)
            elif k == 2: streams.append(c[:int(rng.integers(0, len(c) + 1))]); caps.append(len(d) + 5)
            elif k == 3: streams.append(c); caps.append(int(rng.integers(0, len(d) + 1)))
            elif k == 4:
                b = bytearray(c)
                for _ in range(3): b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                streams.append(bytes(b)); caps.append(len(d) + 1000)
            else:        streams.append(c + b"\xff" * 5); caps.append(len(d) + 9)
        back = B.decompress_streams(streams, caps)
    finally:
        B.set_piece_bytes(65536)
        B.set_decode_piece_bytes(2048)
    for i, (s, c, g) in enumerate(zip(streams, caps, back)):
        assert g == ref.decompress(s, c), i


@pytest.mark.parametrize("chunk_mib,kind", [(16, "mixed"), (4, "text"), (32, "random")])
def test_long_streams_decoder_by_pointer_doubling(B, chunk_mib, kind):
    """lzs_b200_decompress_long_batch_device: a handful of long streams, every byte fetched from the literal it
    is a copy of (csrc/k4_pieces.cuh, k4j_*); the round trip, and the lengths."""
    import torch
    db = B.DeviceBatch(96 << 20, chunk_mib << 20)
    db.fill({"mixed": B.CORPUS_MIXED, "text": B.CORPUS_TEXT, "random": B.CORPUS_RANDOM}[kind], 0x5EED0000 + 52)
    db.compress()
    before = B.lib().lzs_b200_kernel_launches()
    db.decompress_jump()
    assert B.lib().lzs_b200_kernel_launches() - before == 8 + 3 + 32 + 2   # pieces, init/fill/gather, 32 doubling rounds, dirty list + k4_decode
    torch.cuda.synchronize()
    assert db.roundtrip_ok()
