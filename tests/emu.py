"""ctypes front end of tests/simt/libemu.so: the product's kernel sources run on the
CPU SIMT emulator.  Test infrastructure only (there is no GPU in the build container)."""
import ctypes
import os
import subprocess

import numpy as np

from helpers import ROOT, c_u8p, c_u16p, c_u32p, c_u64p, _ptr

SIMT_DIR = os.path.join(ROOT, "tests", "simt")
EMU_SO = os.path.join(SIMT_DIR, "libemu.so")
CSRC = os.path.join(ROOT, "lzs-compression_b200", "csrc")

_lib = None


def _needs_build():
    if not os.path.exists(EMU_SO):
        return True
    t = os.path.getmtime(EMU_SO)
    srcs = [os.path.join(SIMT_DIR, f) for f in os.listdir(SIMT_DIR) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(s) > t for s in srcs)


def lib():
    global _lib
    if _lib is None:
        if _needs_build():
            subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I" + SIMT_DIR, "-o", EMU_SO,
                            os.path.join(SIMT_DIR, "emu_kernels.cpp"), os.path.join(SIMT_DIR, "simt.cpp")],
                           check=True)
        _lib = ctypes.CDLL(EMU_SO)
        _lib.emu_decode.argtypes = [c_u8p, c_u64p, c_u32p, c_u8p, c_u64p, c_u32p, c_u32p, ctypes.c_uint32,
                                    ctypes.c_int, ctypes.c_uint, c_u8p]
        _lib.emu_match.argtypes = [c_u8p, c_u64p, c_u32p, c_u16p, ctypes.c_uint32, ctypes.c_uint]
        _lib.emu_parse_pack.argtypes = [c_u8p, c_u64p, c_u32p, c_u16p, c_u8p, c_u64p, c_u32p, c_u32p,
                                        ctypes.c_uint32]
    return _lib


def _align(x, a):
    return (x + a - 1) // a * a


def pack_streams(streams, align=16, lead=0):
    """Lay byte strings out in one buffer; returns (buf, off, len). `lead` shifts every
    stream by that many bytes to exercise unaligned starts."""
    offs, pos = [], 0
    for s in streams:
        pos = _align(pos, align) + lead
        offs.append(pos)
        pos += len(s)
    buf = np.zeros(_align(pos, 16) + 64, dtype=np.uint8)
    for o, s in zip(offs, streams):
        buf[o:o + len(s)] = np.frombuffer(bytes(s), dtype=np.uint8)
    return buf, np.array(offs, dtype=np.uint64), np.array([len(s) for s in streams], dtype=np.uint32)


def _out_layout(caps, out_lead):
    out_off, pos = [], 0
    for c in caps:
        pos = _align(pos, 16) + out_lead
        out_off.append(pos)
        pos += c
    dst = np.full(_align(pos, 16) + 64, 0xEE, dtype=np.uint8)
    return dst, np.array(out_off, dtype=np.uint64), np.array(caps, dtype=np.uint32)


def _collect(dst, out_off, caps, out_len, what):
    res = []
    for o, c, l in zip(out_off, caps, out_len):
        o, l = int(o), int(l)
        assert l <= c
        assert (dst[o + l:o + c] == 0xEE).all(), what + " wrote past the bytes it reported"
        res.append(dst[o:o + l].tobytes())
    return res


def decode(streams, caps, lanes=8, grid=2, align=16, lead=0, out_lead=0, with_status=False):
    src, in_off, in_len = pack_streams(streams, align, lead)
    dst, out_off, out_cap = _out_layout(caps, out_lead)
    out_len = np.zeros(len(streams), dtype=np.uint32)
    status = np.full(max(len(streams), 1), 0xEE, dtype=np.uint8)
    rc = lib().emu_decode(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(dst), _ptr(out_off, c_u64p),
                          _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), len(streams), lanes, grid,
                          _ptr(status) if with_status else None)
    assert rc == 0
    got = _collect(dst, out_off, caps, out_len, "decoder")
    return (got, [int(x) for x in status[:len(streams)]]) if with_status else got


last_match_disorder = 0   # what the fast K1 launch recorded in the last match() (1: the safe launch ran)


def scramble_exchanges(on):
    """Serve the lanes of shared-memory atomic exchanges in a scrambled order (test knob)."""
    lib().emu_scramble_exchanges(1 if on else 0)


def match(streams, grid=1, align=16, lead=0):
    global last_match_disorder
    src, in_off, in_len = pack_streams(streams, align, lead)
    m = np.zeros(len(src) + 16, dtype=np.uint16)
    last_match_disorder = lib().emu_match(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(m, c_u16p),
                                          len(streams), grid)
    return [m[int(o):int(o) + int(l)].copy() for o, l in zip(in_off, in_len)], (src, in_off, in_len, m)


def compress(streams, caps=None, grid=1, align=16, lead=0, out_lead=0):
    _, (src, in_off, in_len, m) = match(streams, grid, align, lead)
    if caps is None:
        caps = [len(s) + (len(s) + 7) // 8 + 3 for s in streams]
    dst, out_off, out_cap = _out_layout(caps, out_lead)
    out_len = np.zeros(len(streams), dtype=np.uint32)
    lib().emu_parse_pack(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(m, c_u16p), _ptr(dst),
                         _ptr(out_off, c_u64p), _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), len(streams))
    return _collect(dst, out_off, caps, out_len, "packer")


def compress_pieces(streams, piece, caps=None, grid=1, align=16, lead=0, out_lead=0, cap_entries=None):
    """The compressor for long streams (csrc/k23_pieces.cuh): streams cut into pieces of `piece` bytes.
    Returns (streams, stats) with stats = [pieces, left open by spec, by fix, pieces holding a token]."""
    src, in_off, in_len = pack_streams(streams, align, lead)
    if caps is None:
        caps = [len(s) + (len(s) + 7) // 8 + 3 for s in streams]
    dst, out_off, out_cap = _out_layout(caps, out_lead)
    out_len = np.zeros(len(streams), dtype=np.uint32)
    stats = np.zeros(4, dtype=np.uint32)
    if cap_entries is None:
        cap_entries = sum(max(1, -(-len(s) // piece)) for s in streams) + 3
    L = lib()
    L.emu_compress_pieces.argtypes = [c_u8p, c_u64p, c_u32p, ctypes.c_uint64, c_u8p, c_u64p, c_u32p, c_u32p,
                                      ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint, c_u32p]
    over = L.emu_compress_pieces(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), len(src), _ptr(dst),
                                 _ptr(out_off, c_u64p), _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), len(streams),
                                 piece, cap_entries, grid, _ptr(stats, c_u32p))
    if over:
        return None, stats
    return _collect(dst, out_off, caps, out_len, "packer"), stats


def decode_pieces(streams, caps, piece, align=16, lead=0, out_lead=0, cap_entries=None, with_status=False, jump=False,
                  in_used=None):
    """The decoder for long streams (csrc/k4_pieces.cuh): compressed streams cut into pieces of `piece`
    bytes, dirty streams finished by k4_decode.  Returns (outputs, stats[, status]) with stats = [pieces,
    dirty streams, pieces fix left open, table overflow]."""
    src, in_off, in_len = pack_streams(streams, align, lead)
    dst, out_off, out_cap = _out_layout(caps, out_lead)
    out_len = np.zeros(len(streams), dtype=np.uint32)
    status = np.zeros(max(len(streams), 1), dtype=np.uint8)
    stats = np.zeros(4, dtype=np.uint32)
    if cap_entries is None:
        cap_entries = sum(max(1, -(-len(s) // piece)) for s in streams) + len(streams) + 3
    L = lib()
    L.emu_decode_pieces.argtypes = [c_u8p, c_u64p, c_u32p, c_u8p, c_u64p, c_u32p, c_u32p, ctypes.c_uint32,
                                    ctypes.c_uint32, ctypes.c_uint32, c_u8p, c_u32p, ctypes.c_uint32, c_u32p]
    L.emu_decode_pieces(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(dst), _ptr(out_off, c_u64p),
                        _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), len(streams), piece, cap_entries, _ptr(status),
                        _ptr(stats, c_u32p), len(dst) if jump else 0, _ptr(in_used, c_u32p) if in_used is not None else None)
    got = _collect(dst, out_off, caps, out_len, "decoder")
    return (got, stats, status[:len(streams)]) if with_status else (got, stats)


# ---- flows with kept history (lzs_b200_*_flows_batch_device on the emulator) ----

def flows_layout(flows, lead=0):
    """Packets of a flow contiguous, flows 16-byte aligned (+ lead).  Returns (buf, off, len, hist)."""
    offs, lens, hist, pos = [], [], [], 0
    for pk in flows:
        pos = _align(pos, 16) + lead
        done = 0
        for p in pk:
            offs.append(pos)
            lens.append(len(p))
            hist.append(min(2047, done))
            pos += len(p)
            done += len(p)
    buf = np.zeros(_align(pos, 16) + 64, dtype=np.uint8)
    i = 0
    for pk in flows:
        for p in pk:
            buf[offs[i]:offs[i] + len(p)] = np.frombuffer(bytes(p), dtype=np.uint8)
            i += 1
    return buf, np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint32), np.array(hist, dtype=np.uint32)


def compress_flows(flows, grid=1, lead=0):
    """All packets of all flows in one emulated launch; returns a list (per flow) of lists of streams."""
    L = lib()
    L.emu_set_hist.argtypes = [c_u32p]
    src, in_off, in_len, hist = flows_layout(flows, lead)
    n = len(in_len)
    m = np.zeros(len(src) + 16, dtype=np.uint16)
    L.emu_set_hist(_ptr(hist, c_u32p))
    try:
        L.emu_match(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(m, c_u16p), n, grid)
    finally:
        L.emu_set_hist(None)
    caps = [int(l) + (int(l) + 7) // 8 + 3 for l in in_len]
    dst, out_off, out_cap = _out_layout(caps, 0)
    out_len = np.zeros(n, dtype=np.uint32)
    L.emu_parse_pack(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(m, c_u16p), _ptr(dst),
                     _ptr(out_off, c_u64p), _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), n)
    flat = _collect(dst, out_off, caps, out_len, "packer")
    res, i = [], 0
    for pk in flows:
        res.append(flat[i:i + len(pk)])
        i += len(pk)
    return res


def decode_flows(comp_flows, plain_lens):
    """Decoder with kept history: one emulated launch per packet index, all flows together."""
    L = lib()
    L.emu_decode_hist.argtypes = [c_u8p, c_u64p, c_u32p, c_u8p, c_u64p, c_u32p, c_u32p, ctypes.c_uint32, ctypes.c_uint,
                                  c_u32p]
    n_flows = len(comp_flows)
    starts, pos = [], 0
    for lens in plain_lens:
        pos = _align(pos, 16)
        starts.append(pos)
        pos += sum(lens)
    out = np.full(pos + 64, 0xEE, dtype=np.uint8)
    depth = max(len(c) for c in comp_flows)
    done = [0] * n_flows
    for k in range(depth):
        who = [f for f in range(n_flows) if k < len(comp_flows[f])]
        streams = [comp_flows[f][k] for f in who]
        src, in_off, in_len = pack_streams(streams)
        out_off = np.array([starts[f] + done[f] for f in who], dtype=np.uint64)
        out_cap = np.array([plain_lens[f][k] for f in who], dtype=np.uint32)
        hist = np.array([min(2047, done[f]) for f in who], dtype=np.uint32)
        out_len = np.zeros(len(who), dtype=np.uint32)
        L.emu_decode_hist(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(out), _ptr(out_off, c_u64p),
                          _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), len(who), 2, _ptr(hist, c_u32p))
        for f, l in zip(who, out_len):
            assert int(l) == plain_lens[f][k], (f, k, int(l))
            done[f] += int(l)
    return [out[starts[f]:starts[f] + sum(plain_lens[f])].tobytes() for f in range(n_flows)]


def compress_flow_table(flows, plen, grid=1):
    """Flows of equal packets (the last may be shorter) with the match finder taking every flow as ONE
    stream whose look-ahead ends with each packet (lzs_b200_compress_flow_table_device on the emulator)."""
    L = lib()
    L.emu_set_seg.argtypes = [c_u32p]
    src, pkt_off, pkt_len, _ = flows_layout(flows)
    flow_off, flow_len, i = [], [], 0
    for pk in flows:
        flow_off.append(int(pkt_off[i]))
        flow_len.append(sum(len(p) for p in pk))
        i += len(pk)
    flow_off, flow_len = np.array(flow_off, dtype=np.uint64), np.array(flow_len, dtype=np.uint32)
    seg = np.full(len(flows), plen, dtype=np.uint32)
    m = np.zeros(len(src) + 16, dtype=np.uint16)
    L.emu_set_seg(_ptr(seg, c_u32p))
    try:
        L.emu_match(_ptr(src), _ptr(flow_off, c_u64p), _ptr(flow_len, c_u32p), _ptr(m, c_u16p), len(flows), grid)
    finally:
        L.emu_set_seg(None)
    n = len(pkt_len)
    caps = [int(l) + (int(l) + 7) // 8 + 3 for l in pkt_len]
    dst, out_off, out_cap = _out_layout(caps, 0)
    out_len = np.zeros(n, dtype=np.uint32)
    L.emu_parse_pack(_ptr(src), _ptr(pkt_off, c_u64p), _ptr(pkt_len, c_u32p), _ptr(m, c_u16p), _ptr(dst),
                     _ptr(out_off, c_u64p), _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), n)
    flat = _collect(dst, out_off, caps, out_len, "packer")
    res, i = [], 0
    for pk in flows:
        res.append(flat[i:i + len(pk)])
        i += len(pk)
    return res
