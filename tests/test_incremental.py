"""Incremental API parity: every call's (return value, status flags, input consumed) and
the concatenated bytes must equal the unmodified reference's for any slicing of input
and output (reference tests c/src/test/test-lzs-decompression.c:130-290 do this for the
decoder with 10-byte slices; the compressor's incremental API has no reference test).

CPU part: the product's device state machines (csrc/incremental.cuh) on the SIMT
emulator.  GPU part (-m gpu): the same calls through liblzs.so."""
import ctypes
import os

import numpy as np
import pytest

import emu
import helpers
import inc_drivers as D
from gpu_common import binding

SLICINGS_C = [(1 << 20, 1 << 20), (7, 5), (512, 512), (1, 3), (1500, 3), (13, 1 << 20)]
SLICINGS_D = [(1 << 20, 1 << 20), (10, 1 << 20), (1 << 20, 10), (3, 7), (1, 1)]


def _ref():
    if not os.path.exists(helpers.REF_SO):
        pytest.skip("oracle/_ref/liblzs_ref.so not built")
    return ctypes.CDLL(helpers.REF_SO)


def _inputs():
    cases = helpers.edge_case_inputs()
    keep = ["empty", "one", "two_same", "run_38", "abab", "tail_short", "period_2047", "records"]
    data = {k: cases[k] for k in keep}
    data["packet"] = helpers.corpus(helpers.CORPUS_PACKET, 1, 1500).tobytes()
    data["text_5000"] = helpers.corpus(helpers.CORPUS_TEXT, 1, 5000).tobytes()
    data["run_5000"] = b"\0" * 5000
    return data


def _check_compress(ours, ref, data, i_s, o_s, **kw):
    cap = helpers.compressed_max(len(data)) + 8
    a, ta = D.drive(ref, False, data, i_s, o_s, cap, **kw)
    b, tb = D.drive(ours, False, data, i_s, o_s, cap, **kw)
    assert tb == ta, "call traces differ"
    assert b == a
    return a


def _check_decompress(ours, ref, stream, i_s, o_s, cap):
    a, ta = D.drive(ref, True, stream, i_s, o_s, cap)
    b, tb = D.drive(ours, True, stream, i_s, o_s, cap)
    assert tb == ta, "call traces differ"
    assert b == a
    return a


def test_emulated_incremental_compress_matches_reference_calls():
    ref, ours, o = D.StructCodec(_ref()), D.EmuCodec(emu.lib()), helpers.oracle()
    for name, data in _inputs().items():
        for i_s, o_s in SLICINGS_C:
            if len(data) > 2000 and i_s == 1:
                continue
            out = _check_compress(ours, ref, data, i_s, o_s)
            assert out == o.compress(data), (name, i_s, o_s)


def test_emulated_incremental_compress_without_end_marker():
    """add_end_marker=false holds back the last <12 / <15 bytes (lzs-compression.c:641-647,
    :750-758); the emitted prefix must still agree call for call."""
    ref, ours = D.StructCodec(_ref()), D.EmuCodec(emu.lib())
    for name in ("packet", "run_38", "records"):
        _check_compress(ours, ref, _inputs()[name], 100, 64, finish_last=False)


def test_emulated_simple_variant_is_the_same_engine():
    ref_simple, ours, o = D.StructCodec(_ref(), simple=True), D.EmuCodec(emu.lib()), helpers.oracle()
    data = _inputs()["packet"]
    assert _check_compress(ours, ref_simple, data, 64, 64) == o.compress(data)


def test_emulated_incremental_decompress_matches_reference_calls():
    ref, ours, o = D.StructCodec(_ref()), D.EmuCodec(emu.lib()), helpers.oracle()
    golden = open(os.path.join(helpers.GOLDEN_DIR, "golden1_compressed.bin"), "rb").read()
    plain = open(os.path.join(helpers.GOLDEN_DIR, "golden1_plain.bin"), "rb").read()
    for i_s, o_s in SLICINGS_D:                       # the reference's own four drivers, and more
        assert _check_decompress(ours, ref, golden, i_s, o_s, len(plain) + 520) == plain
    for name, data in _inputs().items():
        stream = o.compress(data)
        for i_s, o_s in SLICINGS_D[:4]:
            assert _check_decompress(ours, ref, stream, i_s, o_s, len(data) + 16) == data, (name, i_s, o_s)
    # concatenated streams: history is kept across end markers (lzs-decompression.c:564-576)
    two = o.compress(b"hello hello hello ") + o.compress(b"world world")
    _check_decompress(ours, ref, two, 5, 1 << 20, 100)
    # truncated and noisy streams
    rng = np.random.default_rng(3)
    for k in range(6):
        noise = rng.integers(0, 256, 120, dtype=np.uint8).tobytes()
        _check_decompress(ours, ref, noise, 9, 50, 3000)


@pytest.mark.gpu
def test_gpu_incremental_matches_reference_calls():
    B = binding()
    ours, ref, o = D.StructCodec(B.lib()), D.StructCodec(_ref()), helpers.oracle()
    ours_simple = D.StructCodec(B.lib(), simple=True)
    inputs = _inputs()
    for name in ("empty", "one", "run_38", "packet", "records"):
        data = inputs[name]
        for i_s, o_s in [(1 << 20, 1 << 20), (512, 512), (100, 3)]:
            assert _check_compress(ours, ref, data, i_s, o_s) == o.compress(data)
        assert _check_compress(ours, ref, data, 1 << 20, 1 << 20, quick=True) == o.compress(data)
        assert _check_compress(ours_simple, ref, data, 300, 300) == o.compress(data)
        _check_compress(ours, ref, data, 100, 64, finish_last=False)
        stream = o.compress(data)
        for i_s, o_s in [(1 << 20, 1 << 20), (10, 1 << 20), (1 << 20, 10)]:
            assert _check_decompress(ours, ref, stream, i_s, o_s, len(data) + 16) == data


@pytest.mark.gpu
def test_gpu_incremental_batch_of_packets():
    """BASELINE config 3 through the incremental API: init + incremental(add_end_marker=true)
    until END_MARKER per packet, all packets advanced together, equals lzs_compress(packet)."""
    B = binding()
    L = B.lib()
    o = helpers.oracle()
    n, plen = 256, 1500
    pk = helpers.corpus(helpers.CORPUS_PACKET, n, plen, seed=0x5EED0000 + 3)
    cap = helpers.compressed_max(plen)
    src = np.zeros(n * plen + 16, dtype=np.uint8)
    src[:n * plen] = pk
    dst = np.zeros(n * cap, dtype=np.uint8)
    blocks = [ctypes.create_string_buffer(D.C_SIZE) for _ in range(n)]
    L.lzs_compress_init_full.argtypes = [ctypes.c_void_p]
    arr = (ctypes.c_void_p * n)()
    for s, b in enumerate(blocks):
        L.lzs_compress_init_full(b)
        f = (ctypes.c_uint64 * 4).from_address(ctypes.addressof(b))
        f[0], f[1], f[2], f[3] = src.ctypes.data + s * plen, dst.ctypes.data + s * cap, plen, cap
        arr[s] = ctypes.addressof(b)
    L.lzs_b200_compress_incremental_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
    def done(b):
        return ctypes.c_uint8.from_address(ctypes.addressof(b) + 32).value & D.END_MARKER

    todo = list(blocks)
    for _ in range(64):                               # like the caller's loop in utils/lzs-compress.c:91-134
        arr = (ctypes.c_void_p * len(todo))(*[ctypes.addressof(b) for b in todo])
        B.check(L.lzs_b200_compress_incremental_batch(arr, len(todo), 1, None))
        todo = [b for b in todo if not done(b)]       # a stream stops being called at its end marker
        if not todo:
            break
    else:
        raise AssertionError("packets never reached END_MARKER")
    for s, b in enumerate(blocks):
        f = (ctypes.c_uint64 * 4).from_address(ctypes.addressof(b))
        got = dst[s * cap:s * cap + (cap - f[3])].tobytes()
        assert got == o.compress(pk[s * plen:(s + 1) * plen].tobytes()), s
    # and the bulk path gives the same bytes (the fast way to do the same thing)
    assert B.compress_streams([pk[s * plen:(s + 1) * plen].tobytes() for s in range(8)]) == \
        [o.compress(pk[s * plen:(s + 1) * plen].tobytes()) for s in range(8)]


@pytest.mark.gpu
def test_gpu_flows_with_shared_history_across_packets():
    """SURVEY.md section 8f-2 (RFC 1974 style): every flow compresses packet after packet through ONE
    state block -- each packet runs to its END_MARKER, the history is kept (lzs-compression.c does
    not reset it at the marker), so later packets refer back into earlier ones.  All flows advance
    together through lzs_b200_compress_incremental_batch; every packet of every flow must equal
    what the unmodified reference produces for the same call sequence, and the flow's packets,
    fed one after the other to ONE decoder state, must give the plain bytes back."""
    B = binding()
    L = B.lib()
    R = _ref()
    ours, ref = D.StructCodec(L), D.StructCodec(R)
    L.lzs_b200_compress_incremental_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p]
    n_flows, n_packets = 12, 5
    rng = np.random.default_rng(5)
    vocab = helpers.corpus(helpers.CORPUS_TEXT, 1, 4000, seed=0x5EED0000 + 12).tobytes()
    flows = []
    for f in range(n_flows):
        pk = []
        for k in range(n_packets):
            a = int(rng.integers(0, 2500)); ln = int(rng.integers(40, 1400))
            pk.append(vocab[a:a + ln] + bytes(rng.integers(0, 256, 8, dtype=np.uint8)))
        flows.append(pk)

    def fields(st):
        return (ctypes.c_uint64 * 4).from_address(ctypes.addressof(st))

    def status(st):
        return ctypes.c_uint8.from_address(ctypes.addressof(st) + 32).value

    cap = 2048
    # reference: one flow at a time
    want = []
    for pk in flows:
        st = ref.new(False)
        outs = []
        for p in pk:
            src = np.frombuffer(p + b"\0" * 16, dtype=np.uint8).copy()
            dst = np.zeros(cap, dtype=np.uint8)
            f = fields(st)
            f[0], f[1], f[2], f[3] = src.ctypes.data, dst.ctypes.data, len(p), cap
            for _ in range(64):
                R.lzs_compress_incremental(st, True)
                if status(st) & D.END_MARKER:
                    break
            else:
                raise AssertionError("reference never reached END_MARKER")
            outs.append(dst[:cap - fields(st)[3]].tobytes())
        want.append(outs)
    assert sum(len(o) for o in want[0][1:]) < sum(len(helpers.oracle().compress(p)) for p in flows[0][1:]), \
        "later packets should profit from the kept history"

    # product: all flows together, packet by packet
    states = [ours.new(False) for _ in flows]
    got = [[] for _ in flows]
    for k in range(n_packets):
        srcs = [np.frombuffer(pk[k] + b"\0" * 16, dtype=np.uint8).copy() for pk in flows]
        dsts = [np.zeros(cap, dtype=np.uint8) for _ in flows]
        for st, s, d, pk in zip(states, srcs, dsts, flows):
            f = fields(st)
            f[0], f[1], f[2], f[3] = s.ctypes.data, d.ctypes.data, len(pk[k]), cap
        todo = list(range(n_flows))
        for _ in range(64):
            arr = (ctypes.c_void_p * len(todo))(*[ctypes.addressof(states[i]) for i in todo])
            B.check(L.lzs_b200_compress_incremental_batch(arr, len(todo), 1, None))
            todo = [i for i in todo if not (status(states[i]) & D.END_MARKER)]
            if not todo:
                break
        else:
            raise AssertionError("flows never reached END_MARKER")
        for i in range(n_flows):
            got[i].append(dsts[i][:cap - fields(states[i])[3]].tobytes())
    assert got == want

    # decode: one decoder state per flow, packets fed one after the other (history kept across markers)
    for pk, outs in zip(flows[:4], got[:4]):
        st = ours.new(True)
        back = b""
        for p, c in zip(pk, outs):
            src = np.frombuffer(c + b"\0" * 16, dtype=np.uint8).copy()
            dst = np.zeros(len(p) + 64, dtype=np.uint8)
            ret, stat, used = ours.call(True, st, src.ctypes.data, len(c), dst.ctypes.data, len(dst), False)
            assert stat & D.END_MARKER and used == len(c)
            back += dst[:ret].tobytes()
        assert back == b"".join(pk)


@pytest.mark.gpu
def test_gpu_device_resident_flows_through_the_incremental_api():
    """lzs_b200_*_incremental_batch_device: states, job table, input and output all on the device.
    tools/inc_device_bench.py runs init + incremental until END_MARKER for every flow, compares every
    flow with the batch compressor, the call traces of a sample with the unmodified reference, a
    second packet on the kept history with the reference, and decodes both packets of every flow
    through one device-resident decoder state."""
    import json
    import subprocess
    import sys
    tool = os.path.join(helpers.ROOT, "tools", "inc_device_bench.py")
    r = subprocess.run([sys.executable, tool, "--flows", "6000", "--sample", "96"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["flows_equal_to_batch_compressor"] == 6000
    if os.path.exists(helpers.REF_SO):
        assert line["call_traces_equal_to_reference"] == 96 and line["second_packet_on_kept_history_checked"] == 64
    assert line["ratio_second_packet_kept_history"] > line["ratio_first_packet"]


def _reference_flows(flows, cap=4096):
    """Every flow through ONE reference compressor state: each packet to its END_MARKER, history kept."""
    R = _ref()
    ref = D.StructCodec(R)
    want = []
    for pk in flows:
        st = ref.new(False)
        outs = []
        for p in pk:
            src = np.frombuffer(bytes(p) + b"\0" * 16, dtype=np.uint8).copy()
            dst = np.zeros(cap, dtype=np.uint8)
            f = (ctypes.c_uint64 * 4).from_address(ctypes.addressof(st))
            f[0], f[1], f[2], f[3] = src.ctypes.data, dst.ctypes.data, len(p), cap
            for _ in range(64):
                R.lzs_compress_incremental(st, True)
                if ctypes.c_uint8.from_address(ctypes.addressof(st) + 32).value & D.END_MARKER:
                    break
            else:
                raise AssertionError("reference never reached END_MARKER")
            outs.append(dst[:cap - f[3]].tobytes())
        want.append(outs)
    return want


def _test_flows(seed=5, n_flows=7, n_packets=6):
    rng = np.random.default_rng(seed)
    vocab = helpers.corpus(helpers.CORPUS_TEXT, 1, 6000, seed=0x5EED0000 + 12).tobytes()
    recs = helpers.corpus(helpers.CORPUS_BINARY, 1, 6000, seed=0x5EED0000 + 13).tobytes()
    flows = []
    for f in range(n_flows):
        pk = []
        for k in range(n_packets):
            base = vocab if (f + k) % 3 else recs
            a = int(rng.integers(0, 3500)); ln = int(rng.integers(1, 1600))
            pk.append(base[a:a + ln] + bytes(rng.integers(0, 256, int(rng.integers(0, 9)), dtype=np.uint8)))
        flows.append(pk)
    flows.append([b"a", b"", b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"a" * 3000, b"ab"])      # runs across packet borders, an empty packet
    return flows


def test_emulated_bulk_path_for_flows_with_kept_history():
    """lzs_b200_compress_flows_batch_device / lzs_b200_decompress_flows_batch_device (the K1-class path
    for SURVEY.md section 8f-2) on the emulator: every packet of every flow, compressed in ONE launch
    with the flow's earlier packets as history, equals what the unmodified reference produces through
    one compressor state per flow (init once, each packet to its end marker); and the decoder with
    kept history gives the packets back, one launch per packet index."""
    flows = _test_flows()
    want = _reference_flows(flows)
    got = emu.compress_flows(flows, lead=3)
    for f, (g, w) in enumerate(zip(got, want)):
        assert g == w, f
    # later packets really use the history: smaller than the same packets compressed alone
    o = helpers.oracle()
    assert sum(len(c) for c in got[0][1:]) < sum(len(o.compress(p)) for p in flows[0][1:])
    back = emu.decode_flows(got, [[len(p) for p in pk] for pk in flows])
    assert back == [b"".join(pk) for pk in flows]


@pytest.mark.gpu
def test_gpu_bulk_path_for_flows_with_kept_history():
    """The same on the GPU through the C ABI, plus a table of equal-size packets: every packet of a
    sample of flows against the unmodified reference's one-state-per-flow output, the whole table
    through the decoder with kept history (one launch per packet index) back to the input."""
    import torch
    B = binding()
    L = B.lib()
    dev = torch.device("cuda:0")
    # (1) ragged flows against the reference
    flows = _test_flows(seed=9, n_flows=40, n_packets=5)
    want = _reference_flows(flows)
    src, in_off, in_len, hist = emu.flows_layout(flows, lead=0)
    n = len(in_len)
    caps = np.array([helpers.compressed_max(int(l)) for l in in_len], dtype=np.uint32)
    out_off, _, out_span = B.layout([int(c) for c in caps])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64 if a.dtype == np.uint64 else
                                                               (np.int32 if a.dtype == np.uint32 else np.uint8))).to(dev)
    d_src, d_inoff, d_inlen, d_hist, d_outoff, d_outcap = t(src), t(in_off), t(in_len), t(hist), t(out_off), t(caps)
    d_out = torch.zeros(out_span + 64, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.lzs_b200_compress_scratch_bytes(len(src)), dtype=torch.uint8, device=dev)
    vp = ctypes.c_void_p
    L.lzs_b200_compress_flows_batch_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint64, vp, vp, vp, vp, ctypes.c_uint32, vp,
                                                       ctypes.c_size_t, vp]
    B.check(L.lzs_b200_compress_flows_batch_device(d_src.data_ptr(), d_inoff.data_ptr(), d_inlen.data_ptr(), d_hist.data_ptr(),
                                                   len(src), d_out.data_ptr(), d_outoff.data_ptr(), d_outcap.data_ptr(),
                                                   d_len.data_ptr(), n, scratch.data_ptr(), scratch.numel(), None))
    torch.cuda.synchronize()
    out, lens = d_out.cpu().numpy(), d_len.cpu().numpy()
    i = 0
    for f, pk in enumerate(flows):
        for k in range(len(pk)):
            a = int(out_off[i])
            assert out[a:a + int(lens[i])].tobytes() == want[f][k], (f, k)
            i += 1
    # (2) a table of flows: sample against the reference, everything through the decoder
    tab = B.DeviceFlowTable(3000, 4, 1500)
    B.check(L.lzs_b200_corpus_fill_device(tab.raw.data_ptr(), 1500, 1500, 0, tab.n, 0x5EED0000 + 3, B.CORPUS_PACKET, None))
    tab.compress()
    tab.decompress()
    torch.cuda.synchronize()
    assert torch.equal(tab.dec[:tab.n * 1500], tab.raw[:tab.n * 1500]) and bool((tab.dec_len == 1500).all())
    raw = tab.raw[:16 * 4 * 1500].cpu().numpy()
    sample = [[raw[(f * 4 + k) * 1500:(f * 4 + k + 1) * 1500].tobytes() for k in range(4)] for f in range(16)]
    want = _reference_flows(sample)
    comp, clen = tab.comp[:16 * 4 * tab.cap].cpu().numpy(), tab.comp_len[:64].cpu().numpy()
    for f in range(16):
        for k in range(4):
            s = f * 4 + k
            assert comp[s * tab.cap:s * tab.cap + int(clen[s])].tobytes() == want[f][k], (f, k)
    with_hist = int(tab.comp_len.sum())
    per_packet = (tab.comp.clone(), tab.comp_len.clone())
    tab.compress_table()                               # every flow as one stream: the same bytes
    torch.cuda.synchronize()
    assert torch.equal(tab.comp_len, per_packet[1])
    live = torch.arange(tab.cap, device=tab.comp.device)[None, :] < tab.comp_len[:, None]
    assert not ((tab.comp[:tab.n * tab.cap].view(tab.n, tab.cap) != per_packet[0][:tab.n * tab.cap].view(tab.n, tab.cap)) & live).any()
    tab.compress(with_history=False)
    torch.cuda.synchronize()
    assert with_hist < int(tab.comp_len.sum()), "kept history should make later packets smaller"


def test_emulated_flow_table_as_one_stream_per_flow():
    """lzs_b200_compress_flow_table_device on the emulator: flows of equal packets, every flow ONE stream
    for the match finder (look-ahead ending with each packet), equal to the reference's one state per flow."""
    rng = np.random.default_rng(21)
    vocab = helpers.corpus(helpers.CORPUS_TEXT, 1, 9000, seed=0x5EED0000 + 12).tobytes()
    recs = helpers.corpus(helpers.CORPUS_BINARY, 1, 9000, seed=0x5EED0000 + 13).tobytes()
    plen = 700
    flows = []
    for f in range(6):
        base = vocab if f % 2 else recs
        a = int(rng.integers(0, 3000))
        total = int(rng.integers(plen + 1, 5 * plen))
        data = base[a:a + total]
        flows.append([data[k:k + plen] for k in range(0, len(data), plen)])
    flows.append([b"q" * plen, b"q" * plen, b"q" * 13])                    # a run across packet borders
    want = _reference_flows(flows)
    assert emu.compress_flow_table(flows, plen) == want
    assert emu.compress_flows(flows) == want                                  # and the per-packet form agrees
