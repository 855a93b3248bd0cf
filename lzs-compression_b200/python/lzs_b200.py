"""ctypes binding of liblzs.so (include/lzs.h + include/lzs_b200.h).

This is the thin Python mirror of the C ABI used by the tests and bench.py; it adds
no codec logic.  The library is built in-tree (lzs-compression_b200/Makefile) and is
required: if it cannot be loaded, or no CUDA device is present, calls raise -- there
is no CPU fallback anywhere in this package.
"""
import ctypes
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# LZS_B200_LIB selects another build of the same library (kernel variants in tools/ experiments)
LIB_PATH = os.environ.get("LZS_B200_LIB") or os.path.join(PKG_DIR, "liblzs.so")

u8p = ctypes.POINTER(ctypes.c_uint8)
u16p = ctypes.POINTER(ctypes.c_uint16)
u32p = ctypes.POINTER(ctypes.c_uint32)
u64p = ctypes.POINTER(ctypes.c_uint64)
vp = ctypes.c_void_p

ALIGN = 16
CORPUS_TEXT, CORPUS_BINARY, CORPUS_RANDOM, CORPUS_MIXED, CORPUS_PACKET = range(5)


class LzsError(RuntimeError):
    pass


def compressed_max(n):
    """LZS_COMPRESSED_MAX (include/lzs.h; reference lzs.h:77)."""
    return n + (n + 7) // 8 + 3


def aligned_stride(nbytes, align=ALIGN):
    return (nbytes + align - 1) // align * align


def build(verbose=False):
    """Compile liblzs.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", PKG_DIR], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise LzsError("building liblzs.so failed")
    return LIB_PATH


_lib = None

_DEVICE_BATCH = [vp, vp, vp]          # in, in_off, in_len (device pointers as integers)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LzsError("liblzs.so is not built: run `make -C lzs-compression_b200` "
                       "(or __graft_entry__.build()); there is no fallback implementation")
    L = ctypes.CDLL(LIB_PATH)
    L.lzs_b200_last_error.restype = ctypes.c_char_p
    L.lzs_b200_device_count.restype = ctypes.c_int
    L.lzs_b200_kernel_launches.restype = ctypes.c_uint64
    L.lzs_b200_compress_scratch_bytes.restype = ctypes.c_size_t
    L.lzs_b200_compress_scratch_bytes.argtypes = [ctypes.c_uint64]
    L.lzs_b200_decompress_scratch_bytes.restype = ctypes.c_size_t
    L.lzs_b200_compress_batch_device.argtypes = [vp, vp, vp, ctypes.c_uint64, vp, vp, vp, vp, ctypes.c_uint32,
                                                 vp, ctypes.c_size_t, vp]
    L.lzs_b200_decompress_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp,
                                                   ctypes.c_size_t, vp]
    L.lzs_b200_match_batch_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint32, vp, vp]
    L.lzs_b200_parse_pack_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp]
    host = [u8p, u64p, u32p, ctypes.c_uint64, u8p, u64p, u32p, u32p, ctypes.c_uint64, ctypes.c_uint32]
    L.lzs_b200_compress_batch_host.argtypes = host
    L.lzs_b200_decompress_batch_host.argtypes = host
    L.lzs_b200_decompress_status_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp,
                                                          ctypes.c_size_t, vp]
    L.lzs_b200_compress_packed_host.argtypes = [u8p, u64p, u32p, ctypes.c_uint64, u8p, ctypes.c_uint64, u64p, u32p,
                                                ctypes.c_uint32, u64p]
    L.lzs_b200_corpus_fill_device.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                              ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, vp]
    L.lzs_b200_set_decode_lanes.argtypes = [ctypes.c_int]
    L.lzs_b200_set_piece_bytes.argtypes = [ctypes.c_uint32]
    L.lzs_b200_set_decode_piece_bytes.argtypes = [ctypes.c_uint32]
    L.lzs_b200_decompress_scratch_bytes_long.restype = ctypes.c_size_t
    L.lzs_b200_decompress_scratch_bytes_long.argtypes = [ctypes.c_uint64, ctypes.c_uint32]
    L.lzs_b200_decompress_long_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_uint64, vp, vp, ctypes.c_uint32, vp,
                                                        ctypes.c_size_t, vp]
    L.lzs_b200_decompress_scratch_bytes_jump.restype = ctypes.c_size_t
    L.lzs_b200_decompress_scratch_bytes_jump.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.lzs_b200_pack_streams_device.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint32, vp]
    L.lzs_b200_pack_streams_peers_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint32, ctypes.c_uint64, vp, ctypes.c_uint32, vp]
    L.lzs_b200_pack_streams_multicast_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint64, vp, ctypes.c_uint32, vp]
    L.lzs_b200_release.restype = ctypes.c_int
    L.lzs_b200_chunk_count.restype = ctypes.c_uint32
    L.lzs_b200_chunk_count.argtypes = [ctypes.c_uint64, ctypes.c_uint32]
    for name in ("lzs_compress", "lzs_simple_compress", "lzs_decompress"):
        f = getattr(L, name)
        f.restype = ctypes.c_size_t
        f.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LzsError("liblzs call failed (%d): %s" % (rc, lib().lzs_b200_last_error().decode()))


def _p(a, typ=u8p):
    return a.ctypes.data_as(typ)


# ----------------------------------------------------------------- reference-shaped calls

def lzs_compress(data, out_size=None):
    """lzs_compress(out, outSize, in, inLen) on host bytes (reference lzs.h:218)."""
    data = bytes(data)
    n = len(data)
    cap = compressed_max(n) if out_size is None else out_size
    src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
    dst = np.zeros(max(cap, 1), dtype=np.uint8)
    r = lib().lzs_compress(_p(dst), cap, _p(src), n)
    return dst[:r].tobytes()


def lzs_simple_compress(data, out_size=None):
    data = bytes(data)
    n = len(data)
    cap = compressed_max(n) if out_size is None else out_size
    src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
    dst = np.zeros(max(cap, 1), dtype=np.uint8)
    r = lib().lzs_simple_compress(_p(dst), cap, _p(src), n)
    return dst[:r].tobytes()


def lzs_decompress(data, out_size):
    """lzs_decompress(out, outSize, in, inLen) on host bytes (reference lzs.h:229)."""
    data = bytes(data)
    src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
    dst = np.zeros(max(out_size, 1), dtype=np.uint8)
    r = lib().lzs_decompress(_p(dst), out_size, _p(src), len(data))
    return dst[:r].tobytes()


# ------------------------------------------------------------------------ host batches

def layout(lengths, slot=None, align=ALIGN):
    """Offsets for streams of the given lengths, each slot `slot(len)` bytes, aligned."""
    lengths = np.asarray(lengths, dtype=np.uint32)
    slots = lengths.astype(np.uint64) if slot is None else np.array([slot(int(x)) for x in lengths], dtype=np.uint64)
    strides = (slots + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    off = np.zeros(len(lengths), dtype=np.uint64)
    if len(lengths) > 1:
        off[1:] = np.cumsum(strides[:-1])
    span = int(off[-1] + slots[-1]) if len(lengths) else 0
    return off, slots.astype(np.uint32), span


def _host_batch(fn, src, in_off, in_len, out_off, out_cap, out_span):
    n = len(in_len)
    in_span = int((in_off + in_len.astype(np.uint64)).max()) if n else 0
    dst = np.zeros(out_span + 64, dtype=np.uint8)
    out_len = np.zeros(max(n, 1), dtype=np.uint32)
    check(fn(_p(src), _p(in_off, u64p), _p(in_len, u32p), in_span, _p(dst), _p(out_off, u64p), _p(out_cap, u32p),
             _p(out_len, u32p), out_span, n))
    return dst, out_len[:n]


def compress_streams(streams, caps=None):
    """Batch of independent host byte strings -> list of LZS streams (one launch)."""
    streams = [bytes(s) for s in streams]
    in_off, in_len, in_span = layout([len(s) for s in streams])
    src = np.zeros(in_span + 64, dtype=np.uint8)
    for o, s in zip(in_off, streams):
        src[int(o):int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    caps = [compressed_max(len(s)) for s in streams] if caps is None else list(caps)
    out_off, out_cap, out_span = layout(caps)
    dst, out_len = _host_batch(lib().lzs_b200_compress_batch_host, src, in_off, in_len, out_off, out_cap, out_span)
    return [dst[int(o):int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]


def compress_streams_packed(streams):
    """Like compress_streams, through lzs_b200_compress_packed_host: returns (packed buffer,
    offsets, lengths); stream s is packed[off[s] : off[s] + len[s]]."""
    streams = [bytes(s) for s in streams]
    in_off, in_len, in_span = layout([len(s) for s in streams])
    src = np.zeros(in_span + 64, dtype=np.uint8)
    for o, s in zip(in_off, streams):
        src[int(o):int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    n = len(streams)
    cap = sum(aligned_stride(compressed_max(len(s))) for s in streams)
    dst = np.zeros(cap + 64, dtype=np.uint8)
    out_off = np.zeros(max(n, 1), dtype=np.uint64)
    out_len = np.zeros(max(n, 1), dtype=np.uint32)
    used = np.zeros(1, dtype=np.uint64)
    check(lib().lzs_b200_compress_packed_host(_p(src), _p(in_off, u64p), _p(in_len, u32p), in_span, _p(dst), cap,
                                              _p(out_off, u64p), _p(out_len, u32p), n, _p(used, u64p)))
    return dst[:int(used[0])], out_off[:n], out_len[:n]


def decompress_streams(streams, caps):
    streams = [bytes(s) for s in streams]
    in_off, in_len, in_span = layout([len(s) for s in streams])
    src = np.zeros(in_span + 64, dtype=np.uint8)
    for o, s in zip(in_off, streams):
        src[int(o):int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    out_off, out_cap, out_span = layout(list(caps))
    dst, out_len = _host_batch(lib().lzs_b200_decompress_batch_host, src, in_off, in_len, out_off, out_cap, out_span)
    return [dst[int(o):int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len)]


def decompress_streams_status(streams, caps, device="cuda:0"):
    """Like decompress_streams, through the device entry point that also says why every stream
    stopped (LzsDecompressStatus_t values).  Returns (list of bytes, list of status bytes)."""
    import torch
    L = lib()
    streams = [bytes(s) for s in streams]
    n = len(streams)
    in_off, in_len, in_span = layout([len(s) for s in streams])
    src = np.zeros(in_span + 64, dtype=np.uint8)
    for o, s in zip(in_off, streams):
        src[int(o):int(o) + len(s)] = np.frombuffer(s, dtype=np.uint8)
    out_off, out_cap, out_span = layout(list(caps))
    dev = torch.device(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64 if a.dtype == np.uint64 else
                                                               (np.int32 if a.dtype == np.uint32 else np.uint8))).to(dev)
    d_src, d_inoff, d_inlen, d_outoff, d_outcap = t(src), t(in_off), t(in_len), t(out_off), t(out_cap)
    d_dst = torch.zeros(out_span + 64, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    d_status = torch.full((max(n, 1),), 0xEE, dtype=torch.uint8, device=dev)
    scratch = torch.empty(L.lzs_b200_decompress_scratch_bytes(), dtype=torch.uint8, device=dev)
    check(L.lzs_b200_decompress_status_batch_device(
        d_src.data_ptr(), d_inoff.data_ptr(), d_inlen.data_ptr(), d_dst.data_ptr(), d_outoff.data_ptr(),
        d_outcap.data_ptr(), d_len.data_ptr(), d_status.data_ptr(), n, scratch.data_ptr(), scratch.numel(),
        torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize(dev)
    dst, out_len, status = d_dst.cpu().numpy(), d_len.cpu().numpy(), d_status.cpu().numpy()
    return ([dst[int(o):int(o) + int(l)].tobytes() for o, l in zip(out_off, out_len[:n])],
            [int(x) for x in status[:n]])


# ---------------------------------------------------------------------- device batches

def set_piece_bytes(n):
    """Piece size for long streams (include/lzs_b200.h: lzs_b200_set_piece_bytes); 0 = never cut."""
    check(lib().lzs_b200_set_piece_bytes(int(n)))


def set_decode_piece_bytes(n):
    """Compressed bytes per piece in the decoder for long streams (lzs_b200_set_decode_piece_bytes); 0 = off."""
    check(lib().lzs_b200_set_decode_piece_bytes(int(n)))


class DeviceBatch:
    """Device-resident uniform chunking of one torch.uint8 buffer (plumbing only:
    torch provides the allocations and the stream; all work is in liblzs.so)."""

    def __init__(self, total_bytes, chunk, device="cuda:0"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.total = int(total_bytes)
        self.chunk = int(chunk)
        self.n = (self.total + self.chunk - 1) // self.chunk
        self.comp_stride = aligned_stride(compressed_max(self.chunk))
        idx = torch.arange(self.n, dtype=torch.int64, device=self.device)
        self.raw_off = idx * self.chunk
        self.raw_len = torch.full((self.n,), self.chunk, dtype=torch.int32, device=self.device)
        if self.n:
            self.raw_len[-1] = self.total - (self.n - 1) * self.chunk
        self.comp_off = idx * self.comp_stride
        self.comp_cap = torch.full((self.n,), self.comp_stride, dtype=torch.int32, device=self.device)
        self.comp_len = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.dec_len = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.raw = torch.empty(self.total + 64, dtype=torch.uint8, device=self.device)
        self.comp = torch.empty(self.n * self.comp_stride + 64, dtype=torch.uint8, device=self.device)
        self.dec = torch.empty(self.total + 64, dtype=torch.uint8, device=self.device)
        nscratch = max(lib().lzs_b200_compress_scratch_bytes(self.total),
                       lib().lzs_b200_decompress_scratch_bytes_long(self.n * self.comp_stride, self.n))
        self.scratch = torch.empty(nscratch, dtype=torch.uint8, device=self.device)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def fill(self, kind, seed, first_index=0):
        check(lib().lzs_b200_corpus_fill_device(self.raw.data_ptr(), self.chunk, self.chunk, first_index, self.n,
                                                seed, kind, self._stream()))

    def compress(self):
        check(lib().lzs_b200_compress_batch_device(
            self.raw.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(), self.total,
            self.comp.data_ptr(), self.comp_off.data_ptr(), self.comp_cap.data_ptr(), self.comp_len.data_ptr(),
            self.n, self.scratch.data_ptr(), self.scratch.numel(), self._stream()))

    def match_only(self):
        check(lib().lzs_b200_match_batch_device(
            self.raw.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(),
            self.scratch.data_ptr() + 256, self.n, self.scratch.data_ptr(), self._stream()))

    def parse_pack_only(self):
        check(lib().lzs_b200_parse_pack_batch_device(
            self.raw.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(), self.scratch.data_ptr() + 256,
            self.comp.data_ptr(), self.comp_off.data_ptr(), self.comp_cap.data_ptr(), self.comp_len.data_ptr(),
            self.n, self._stream()))

    def decompress(self):
        check(lib().lzs_b200_decompress_batch_device(
            self.comp.data_ptr(), self.comp_off.data_ptr(), self.comp_len.data_ptr(),
            self.dec.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(), self.dec_len.data_ptr(),
            self.n, self.scratch.data_ptr(), self.scratch.numel(), self._stream()))

    def decompress_jump(self):
        """A handful of long streams: pointer doubling instead of the replay (lzs_b200_decompress_long_batch_device)."""
        need = lib().lzs_b200_decompress_scratch_bytes_jump(self.n * self.comp_stride, self.total, self.n)
        if self.scratch.numel() < need:
            self.scratch = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        check(lib().lzs_b200_decompress_long_batch_device(
            self.comp.data_ptr(), self.comp_off.data_ptr(), self.comp_len.data_ptr(),
            self.dec.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(), self.total, self.dec_len.data_ptr(),
            None, self.n, self.scratch.data_ptr(), self.scratch.numel(), self._stream()))

    def roundtrip_ok(self):
        t = self.torch
        return bool(t.equal(self.dec[:self.total], self.raw[:self.total])) and \
            bool(t.equal(self.dec_len, self.raw_len))

    def compressed_bytes(self):
        return int(self.comp_len.to(self.torch.int64).sum().item())


# ------------------------------------------------ device-resident incremental batches

class DeviceFlows:
    """n flows advanced together through the incremental API with everything resident on the
    device (include/lzs_b200.h: lzs_b200_*_incremental_batch_device).  Plumbing only: torch holds
    the state blocks and the job table; every call is ONE launch, nothing crosses PCIe.
    Job record (48 bytes) as six int64 columns: state, in, out, in_len | out_cap << 32,
    in_used | out_used << 32, status | add_end_marker << 32."""

    def __init__(self, n, decompress=False, device="cuda:0"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.n = int(n)
        self.decompress = bool(decompress)
        L = lib()
        L.lzs_b200_incremental_state_bytes.restype = ctypes.c_size_t
        L.lzs_b200_incremental_state_bytes.argtypes = [ctypes.c_int]
        vp = ctypes.c_void_p
        L.lzs_b200_incremental_init_device.argtypes = [vp, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int, vp]
        L.lzs_b200_compress_incremental_batch_device.argtypes = [vp, ctypes.c_uint32, vp]
        L.lzs_b200_decompress_incremental_batch_device.argtypes = [vp, ctypes.c_uint32, vp]
        self.stride = int(L.lzs_b200_incremental_state_bytes(int(self.decompress)))
        self.states = torch.empty(self.n * self.stride, dtype=torch.uint8, device=self.device)
        self.jobs = torch.zeros((self.n, 6), dtype=torch.int64, device=self.device)
        self.jobs[:, 0] = self.states.data_ptr() + torch.arange(self.n, dtype=torch.int64, device=self.device) * self.stride
        self.init()

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def init(self):
        check(lib().lzs_b200_incremental_init_device(self.states.data_ptr(), self.stride, self.n, int(self.decompress),
                                                      self._stream()))

    def offer(self, in_ptr, in_len, out_ptr, out_cap, add_end_marker=True):
        """Set inPtr/inLength/outPtr/outLength of every flow (int64 tensors of device addresses / sizes)."""
        j = self.jobs
        j[:, 1] = in_ptr
        j[:, 2] = out_ptr
        j[:, 3] = in_len.to(self.torch.int64) | (out_cap.to(self.torch.int64) << 32)
        j[:, 4] = 0
        j[:, 5] = (1 if add_end_marker else 0) << 32

    def call(self):
        """One lzs_*_incremental call for every flow; then advance pointers and lengths as the
        reference does in its parameter block.  Returns (in_used, out_used, status) tensors."""
        f = lib().lzs_b200_decompress_incremental_batch_device if self.decompress else \
            lib().lzs_b200_compress_incremental_batch_device
        check(f(self.jobs.data_ptr(), self.n, self._stream()))
        j = self.jobs
        m32 = 0xFFFFFFFF
        in_used, out_used = j[:, 4] & m32, (j[:, 4] >> 32) & m32
        status = j[:, 5] & m32
        in_len, out_cap = (j[:, 3] & m32) - in_used, ((j[:, 3] >> 32) & m32) - out_used
        j[:, 1] += in_used
        j[:, 2] += out_used
        j[:, 3] = in_len | (out_cap << 32)
        # a flow that has written its end marker is not called for again (the caller's loop of
        # c/src/utils/lzs-compress.c:91 ends there): no input, no marker request
        done = (status & 0x04) != 0
        flag = self.torch.where(done, self.torch.zeros_like(status), (j[:, 5] >> 32) & 1)
        j[:, 5] = flag << 32
        return in_used.clone(), out_used.clone(), status.clone()


# ------------------------------------------- flows with kept history, the bulk path

class DeviceFlowTable:
    """n_flows flows x n_packets packets of plen bytes, a flow's packets contiguous in device memory
    (plumbing for lzs_b200_compress_flows_batch_device / lzs_b200_decompress_flows_batch_device).
    compress(): ALL packets of all flows in one call, each with its flow's earlier packets as history;
    decompress(): one call per packet index (a flow is serial), all flows together."""

    def __init__(self, n_flows, n_packets, plen, device="cuda:0"):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.n_flows, self.n_packets, self.plen = int(n_flows), int(n_packets), int(plen)
        self.n = self.n_flows * self.n_packets
        self.flow_bytes = self.n_packets * self.plen
        self.cap = aligned_stride(compressed_max(self.plen))
        dev = self.device
        idx = torch.arange(self.n, dtype=torch.int64, device=dev)          # stream s = flow s // P, packet s % P
        self.k = idx % self.n_packets
        self.raw_off = idx * self.plen                                     # flows back to back, packets contiguous
        self.raw_len = torch.full((self.n,), self.plen, dtype=torch.int32, device=dev)
        self.hist = torch.clamp(self.k * self.plen, max=2047).to(torch.int32)
        self.comp_off = idx * self.cap
        self.comp_cap = torch.full((self.n,), self.cap, dtype=torch.int32, device=dev)
        self.comp_len = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.dec_len = torch.zeros(self.n, dtype=torch.int32, device=dev)
        total = self.n * self.plen
        self.raw = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
        self.comp = torch.zeros(self.n * self.cap + 64, dtype=torch.uint8, device=dev)
        self.dec = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
        L = lib()
        vp = ctypes.c_void_p
        L.lzs_b200_compress_flows_batch_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint64, vp, vp, vp, vp, ctypes.c_uint32,
                                                           vp, ctypes.c_size_t, vp]
        L.lzs_b200_decompress_flows_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_uint32, vp,
                                                             ctypes.c_size_t, vp]
        self.scratch = torch.empty(L.lzs_b200_compress_scratch_bytes(total), dtype=torch.uint8, device=dev)
        L.lzs_b200_compress_flow_table_device.argtypes = [vp, vp, vp, vp, ctypes.c_uint32, vp, vp, ctypes.c_uint32,
                                                          ctypes.c_uint64, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        fidx = torch.arange(self.n_flows, dtype=torch.int64, device=dev)
        self.flow_off = fidx * self.flow_bytes
        self.flow_len = torch.full((self.n_flows,), self.flow_bytes, dtype=torch.int32, device=dev)
        self.seg_len = torch.full((self.n_flows,), self.plen, dtype=torch.int32, device=dev)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def compress_table(self):
        """Every flow as ONE stream for the match finder (the look-ahead ends with each packet)."""
        check(lib().lzs_b200_compress_flow_table_device(
            self.raw.data_ptr(), self.flow_off.data_ptr(), self.flow_len.data_ptr(), self.seg_len.data_ptr(), self.n_flows,
            self.raw_off.data_ptr(), self.raw_len.data_ptr(), self.n, self.n * self.plen, self.comp.data_ptr(),
            self.comp_off.data_ptr(), self.comp_cap.data_ptr(), self.comp_len.data_ptr(), self.scratch.data_ptr(),
            self.scratch.numel(), self._stream()))

    def compress(self, with_history=True):
        hist = self.hist.data_ptr() if with_history else None
        check(lib().lzs_b200_compress_flows_batch_device(
            self.raw.data_ptr(), self.raw_off.data_ptr(), self.raw_len.data_ptr(), hist, self.n * self.plen,
            self.comp.data_ptr(), self.comp_off.data_ptr(), self.comp_cap.data_ptr(), self.comp_len.data_ptr(), self.n,
            self.scratch.data_ptr(), self.scratch.numel(), self._stream()))

    def decompress(self):
        t = self.torch
        for k in range(self.n_packets):
            sel = t.nonzero(self.k == k).flatten()
            pick = lambda a: a[sel].contiguous()
            c_off, c_len, o_off, o_cap, h = pick(self.comp_off), pick(self.comp_len), pick(self.raw_off), pick(self.raw_len), pick(self.hist)
            o_len = t.zeros(sel.numel(), dtype=t.int32, device=self.device)
            check(lib().lzs_b200_decompress_flows_batch_device(
                self.comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), self.dec.data_ptr(), o_off.data_ptr(),
                o_cap.data_ptr(), h.data_ptr(), o_len.data_ptr(), None, int(sel.numel()), self.scratch.data_ptr(),
                self.scratch.numel(), self._stream()))
            self.dec_len[sel] = o_len
