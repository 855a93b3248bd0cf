"""Shared test plumbing: ctypes loaders for the CPU checkers and corpus helpers.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/.
Nothing here reads /root/reference at run time: oracle/_ref/liblzs_ref.so is the
prebuilt, git-ignored build of the unmodified reference (oracle/Makefile) and is
optional -- tests that need it skip when it is absent.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblzs_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "liblzs_ref.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u16p = ctypes.POINTER(ctypes.c_uint16)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_u64p = ctypes.POINTER(ctypes.c_uint64)

CORPUS_TEXT, CORPUS_BINARY, CORPUS_RANDOM, CORPUS_MIXED, CORPUS_PACKET = range(5)


def compressed_max(n):
    """LZS_COMPRESSED_MAX, c/src/liblzs/lzs.h:77."""
    return n + (n + 7) // 8 + 3


def _ptr(a, typ=c_u8p):
    return a.ctypes.data_as(typ)


def build_oracle():
    """(Re)build the CPU checkers; cheap, so tests call it once per session."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


_DRV_SIG = [c_u8p, c_u64p, c_u32p, c_u8p, c_u64p, c_u32p, c_u32p, ctypes.c_uint32, ctypes.c_int]


class _Codec:
    """compress(bytes)->bytes / decompress(bytes, cap)->bytes over a C codec."""

    def __init__(self, lib, cname, dname):
        self.lib = lib
        self._c = getattr(lib, cname)
        self._d = getattr(lib, dname)
        for f in (self._c, self._d):
            f.restype = ctypes.c_size_t
            f.argtypes = [c_u8p, ctypes.c_size_t, c_u8p, ctypes.c_size_t]
        for name in ("lzsdrv_compress_streams", "lzsdrv_decompress_streams"):
            f = getattr(lib, name)
            f.restype = ctypes.c_double
            f.argtypes = _DRV_SIG

    def compress(self, data, cap=None):
        data = bytes(data)
        n = len(data)
        cap = compressed_max(n) if cap is None else cap
        # one readable byte of slack after the input: the reference hashes in[n]
        src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
        dst = np.zeros(max(cap, 1), dtype=np.uint8)
        r = self._c(_ptr(dst), cap, _ptr(src), n)
        return dst[:r].tobytes()

    def decompress(self, data, cap):
        data = bytes(data)
        src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
        dst = np.zeros(max(cap, 1), dtype=np.uint8)
        r = self._d(_ptr(dst), cap, _ptr(src), len(data))
        return dst[:r].tobytes()

    def decompress_status(self, data, cap):
        """(bytes, why) -- oracle only: which exit of the decoder's loop was taken, as a
        LzsDecompressStatus_t value (oracle/lzs_oracle.c:lzs_oracle_decompress_status)."""
        f = self.lib.lzs_oracle_decompress_status
        f.restype = ctypes.c_size_t
        f.argtypes = [c_u8p, ctypes.c_size_t, c_u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
        data = bytes(data)
        src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
        dst = np.zeros(max(cap, 1), dtype=np.uint8)
        why = ctypes.c_int(0)
        r = f(_ptr(dst), cap, _ptr(src), len(data), ctypes.byref(why))
        return dst[:r].tobytes(), why.value

    def run_streams(self, decompress, src, in_off, in_len, dst, out_off, out_cap, threads=1):
        """Batch driver (oracle/chunk_driver.c). Returns (out_len, seconds)."""
        n = len(in_len)
        out_len = np.zeros(n, dtype=np.uint32)
        f = self.lib.lzsdrv_decompress_streams if decompress else self.lib.lzsdrv_compress_streams
        sec = f(_ptr(src), _ptr(in_off, c_u64p), _ptr(in_len, c_u32p), _ptr(dst),
                _ptr(out_off, c_u64p), _ptr(out_cap, c_u32p), _ptr(out_len, c_u32p), n, threads)
        return out_len, sec


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = ctypes.CDLL(ORACLE_SO)
        _oracle = _Codec(lib, "lzs_oracle_compress", "lzs_oracle_decompress")
        lib.lzs_oracle_all_matches.restype = None
        lib.lzs_oracle_all_matches.argtypes = [c_u8p, ctypes.c_size_t, c_u8p, c_u16p]
        lib.lzs_corpus_fill_host.restype = None
        lib.lzs_corpus_fill_host.argtypes = [c_u8p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
    return _oracle


def reference():
    """The unmodified reference build, or None when oracle/_ref is absent."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        lib = ctypes.CDLL(REF_SO)
        _ref = _Codec(lib, "lzs_compress", "lzs_decompress")
    return _ref


def oracle_all_matches(data):
    data = bytes(data)
    n = len(data)
    src = np.frombuffer(data + b"\0", dtype=np.uint8).copy()
    ln = np.zeros(max(n, 1), dtype=np.uint8)
    off = np.zeros(max(n, 1), dtype=np.uint16)
    oracle().lib.lzs_oracle_all_matches(_ptr(src), n, _ptr(ln), _ptr(off, c_u16p))
    return ln[:n], off[:n]


def corpus(kind, n_streams, stream_len, seed=0x5EED0000, first_index=0, stride=None):
    """n_streams x stream_len synthetic bytes (lzs-compression_b200/csrc/corpus.h), flat uint8."""
    stride = stream_len if stride is None else stride
    buf = np.zeros(n_streams * stride + 16, dtype=np.uint8)
    oracle().lib.lzs_corpus_fill_host(_ptr(buf), stride, stream_len, first_index, n_streams, seed, kind)
    return buf[: n_streams * stride]


def edge_case_inputs():
    """Small adversarial inputs around the format's corners (hard parts)."""
    rng = np.random.default_rng(1234)
    cases = {
        "empty": b"",
        "one": b"a",
        "two_same": b"aa",
        "two_diff": b"ab",
        "three_same": b"aaa",
        "run_23": b"x" * 23,
        "run_24": b"x" * 24,
        "run_38": b"x" * 38,
        "run_5000": b"\0" * 5000,
        "abab": b"ab" * 700,
        "abc": b"abc" * 1000,
        "tail_short": b"hello world, hello worl",
    }
    block = rng.integers(0, 256, 2046, dtype=np.uint8).tobytes()
    for period in (2046, 2047, 2048, 2049):
        b = (block + bytes(rng.integers(0, 256, 8, dtype=np.uint8)))[:period]
        cases["period_%d" % period] = b * 3
    cases["random_3000"] = rng.integers(0, 256, 3000, dtype=np.uint8).tobytes()
    cases["alpha3_6000"] = rng.integers(0, 3, 6000, dtype=np.uint8).tobytes()
    cases["alpha20_6000"] = (rng.integers(0, 20, 6000, dtype=np.uint8) + 97).astype(np.uint8).tobytes()
    # far repeats around the window edge
    far = bytearray(rng.integers(0, 256, 9000, dtype=np.uint8).tobytes())
    for d in (2040, 2046, 2047, 2048, 2050):
        p = int(rng.integers(2100, 8000))
        far[p:p + 40] = far[p - d:p - d + 40]
    cases["far_repeats"] = bytes(far)
    # zero runs inside records (long futile hash chains in the reference)
    recs = bytearray()
    for i in range(200):
        recs += int(i).to_bytes(4, "little") + bytes(8) + bytes(rng.integers(0, 4, 20, dtype=np.uint8))
    cases["records"] = bytes(recs)
    return cases
