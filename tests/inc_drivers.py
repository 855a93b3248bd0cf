"""Drive an incremental LZS codec through arbitrary input/output slicing and record every
call: (bytes returned, status, input consumed).  Three back ends share the driver:
the unmodified reference (oracle/_ref), the product library (liblzs.so, GPU) and the
product's device code on the CPU emulator."""
import ctypes

import numpy as np

C_SIZE, S_SIZE, D_SIZE = 14432, 2112, 2096           # reference lzs.h struct sizes (LP64)
END_MARKER, NO_SPACE = 0x04, 0x08


class StructCodec:
    """A library that exports the reference's incremental ABI (liblzs_ref.so or liblzs.so)."""

    def __init__(self, lib, simple=False):
        self.lib = lib
        self.simple = simple
        for name in ("lzs_compress_incremental", "lzs_simple_compress_incremental", "lzs_decompress_incremental"):
            getattr(lib, name).restype = ctypes.c_size_t
        lib.lzs_compress_incremental.argtypes = [ctypes.c_void_p, ctypes.c_bool]
        lib.lzs_simple_compress_incremental.argtypes = [ctypes.c_void_p, ctypes.c_bool]
        lib.lzs_decompress_incremental.argtypes = [ctypes.c_void_p]
        for name in ("lzs_compress_init_full", "lzs_compress_init_quick", "lzs_simple_compress_init",
                     "lzs_decompress_init"):
            getattr(lib, name).argtypes = [ctypes.c_void_p]
            getattr(lib, name).restype = None

    def new(self, decompress, quick=False):
        size = D_SIZE if decompress else (S_SIZE if self.simple else C_SIZE)
        st = ctypes.create_string_buffer(b"\x5A" * size, size)     # poisoned: init must not rely on zeros
        if decompress:
            self.lib.lzs_decompress_init(st)
        elif self.simple:
            self.lib.lzs_simple_compress_init(st)
        elif quick:
            self.lib.lzs_compress_init_quick(st)
        else:
            self.lib.lzs_compress_init_full(st)
        return st

    def call(self, decompress, st, src, in_len, dst, out_cap, finish):
        """src/dst are ctypes addresses. Returns (ret, status, in_used)."""
        base = ctypes.addressof(st)
        f = (ctypes.c_uint64 * 4).from_address(base)
        f[0], f[1], f[2], f[3] = src, dst, in_len, out_cap
        if decompress:
            ret = self.lib.lzs_decompress_incremental(st)
        elif self.simple:
            ret = self.lib.lzs_simple_compress_incremental(st, finish)
        else:
            ret = self.lib.lzs_compress_incremental(st, finish)
        status = ctypes.c_uint8.from_address(base + 32).value
        assert f[1] - dst == ret and out_cap - f[3] == ret
        assert f[0] - src == in_len - f[2]
        return int(ret), status, int(in_len - f[2])


class EmuCodec:
    def __init__(self, lib):
        self.lib = lib
        lib.emu_inc_call.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                     ctypes.c_uint32, ctypes.c_int] + [ctypes.POINTER(ctypes.c_uint32)] * 3
        lib.emu_inc_init.argtypes = [ctypes.c_int, ctypes.c_void_p]

    def new(self, decompress, quick=False):
        st = ctypes.create_string_buffer(b"\x5A" * 2112, 2112)
        self.lib.emu_inc_init(int(decompress), st)
        return st

    def call(self, decompress, st, src, in_len, dst, out_cap, finish):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        self.lib.emu_inc_call(int(decompress), st, src, in_len, dst, out_cap, int(finish), a, b, c)
        return int(b.value), int(c.value), int(a.value)


def drive(codec, decompress, data, in_slice, out_slice, out_total, finish_last=True, quick=False, max_calls=200000):
    """Feed `data` in slices of in_slice bytes, offer out_slice bytes of space per call.
    Returns (output bytes, trace) where trace lists (ret, status, in_used) per call."""
    data = bytes(data)
    src = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    dst = np.zeros(out_total + 64, dtype=np.uint8)
    st = codec.new(decompress, quick)
    ipos = opos = 0
    trace = []
    pending = 0                                    # input offered but not yet taken
    for _ in range(max_calls):
        if pending == 0:
            pending = min(in_slice, len(data) - ipos)
        last_input = ipos + pending >= len(data)
        space = min(out_slice, out_total - opos)
        finish = bool(finish_last and last_input and not decompress)
        ret, status, used = codec.call(decompress, st, src.ctypes.data + ipos, pending, dst.ctypes.data + opos, space,
                                       finish)
        trace.append((ret, status, used))
        ipos += used
        pending -= used
        opos += ret
        if decompress:
            if status & END_MARKER:
                break
            if last_input and pending == 0 and ret == 0 and not (status & NO_SPACE):
                break
            if opos >= out_total and (status & NO_SPACE):
                break
        else:
            if status & END_MARKER:
                break
            if not finish_last and last_input and pending == 0 and ret == 0:
                break
            if opos >= out_total:
                break
    return dst[:opos].tobytes(), trace
