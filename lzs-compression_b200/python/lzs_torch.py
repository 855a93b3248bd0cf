"""torch.uint8 front end of the batch C ABI: ragged batches of streams held
in device tensors are compressed / decompressed in place on the caller's current CUDA stream.
Plumbing only -- torch provides the tensors and the stream, every byte of work is in liblzs.so.

    comp, comp_off, comp_len = lzs_torch.compress(data, off, length)        # all on the same GPU
    plain, plain_len         = lzs_torch.decompress(comp, comp_off, comp_len, out_off, out_cap)

`data` is a 1-D uint8 tensor holding the streams, `off` (int64) and `length` (int32) say where each
one is; offsets should be multiples of 16.  Nothing synchronises: results are ready when the
stream reaches them, like any other op.
"""
import torch

import lzs_b200 as _B


def _check_ragged(buf, off, length):
    if buf.dtype != torch.uint8 or buf.dim() != 1 or not buf.is_cuda or not buf.is_contiguous():
        raise ValueError("streams must live in a contiguous 1-D torch.uint8 CUDA tensor")
    if off.dtype != torch.int64 or length.dtype != torch.int32 or off.shape != length.shape or off.dim() != 1:
        raise ValueError("offsets must be int64 and lengths int32, one per stream")
    if off.device != buf.device or length.device != buf.device:
        raise ValueError("offsets and lengths must be on the device of the data")


def compress(data, off, length):
    """Returns (comp, comp_off, comp_len): stream s occupies comp[comp_off[s] : comp_off[s] + comp_len[s]]
    and is byte for byte what lzs_compress gives for data[off[s] : off[s] + length[s]].  Slots are
    sized for the worst case (LZS_COMPRESSED_MAX) and 16-byte aligned."""
    _check_ragged(data, off, length)
    L = _B.lib()
    n = int(off.numel())
    cap = ((length.to(torch.int64) + (length.to(torch.int64) + 7) // 8 + 3 + 15) // 16) * 16
    comp_off = torch.cumsum(cap, 0) - cap
    total_cap = int(cap.sum().item()) if n else 0
    comp = torch.empty(total_cap + 64, dtype=torch.uint8, device=data.device)
    comp_cap = cap.to(torch.int32)
    comp_len = torch.zeros(max(n, 1), dtype=torch.int32, device=data.device)
    span = int(data.numel())
    scratch = torch.empty(L.lzs_b200_compress_scratch_bytes(span), dtype=torch.uint8, device=data.device)
    _B.check(L.lzs_b200_compress_batch_device(
        data.data_ptr(), off.data_ptr(), length.data_ptr(), span, comp.data_ptr(), comp_off.data_ptr(),
        comp_cap.data_ptr(), comp_len.data_ptr(), n, scratch.data_ptr(), scratch.numel(),
        torch.cuda.current_stream(data.device).cuda_stream))
    scratch.record_stream(torch.cuda.current_stream(data.device))
    return comp, comp_off, comp_len[:n]


def decompress(comp, comp_off, comp_len, out_off, out_cap, with_status=False):
    """Decodes stream s into out[out_off[s] : out_off[s] + out_cap[s]]; returns (out, out_len) or,
    with_status, (out, out_len, status) where status holds one LzsDecompressStatus_t byte per stream."""
    _check_ragged(comp, comp_off, comp_len)
    if out_off.dtype != torch.int64 or out_cap.dtype != torch.int32 or out_off.shape != comp_off.shape:
        raise ValueError("output offsets must be int64 and capacities int32, one per stream")
    L = _B.lib()
    n = int(comp_off.numel())
    span = int((out_off + out_cap.to(torch.int64)).max().item()) if n else 0
    out = torch.empty(span + 64, dtype=torch.uint8, device=comp.device)
    out_len = torch.zeros(max(n, 1), dtype=torch.int32, device=comp.device)
    status = torch.zeros(max(n, 1), dtype=torch.uint8, device=comp.device)
    # with room for the piece table of a batch of few long streams (csrc/k4_pieces.cuh); for many
    # streams this is the launch-order scratch of lzs_b200_decompress_scratch_bytes_for
    scratch = torch.empty(L.lzs_b200_decompress_scratch_bytes_long(int(comp.numel()), n), dtype=torch.uint8,
                          device=comp.device)
    _B.check(L.lzs_b200_decompress_status_batch_device(
        comp.data_ptr(), comp_off.data_ptr(), comp_len.data_ptr(), out.data_ptr(), out_off.data_ptr(),
        out_cap.data_ptr(), out_len.data_ptr(), status.data_ptr() if with_status else None, n, scratch.data_ptr(),
        scratch.numel(), torch.cuda.current_stream(comp.device).cuda_stream))
    scratch.record_stream(torch.cuda.current_stream(comp.device))
    return (out, out_len[:n], status[:n]) if with_status else (out, out_len[:n])


# ---------------------------------------------------------------- registered operators
# The same two calls as torch.library custom ops (namespace lzs_b200), so that they show up in
# torch's dispatcher / profiler / export like any other operator:
#     comp, comp_off, comp_len = torch.ops.lzs_b200.compress(data, off, length)
#     out, out_len, status     = torch.ops.lzs_b200.decompress(comp, comp_off, comp_len, out_off, out_cap)
# CUDA only; there is no CPU kernel behind them (the codec has no CPU path).
try:
    @torch.library.custom_op("lzs_b200::compress", mutates_args=(), device_types="cuda")
    def compress_op(data: torch.Tensor, off: torch.Tensor, length: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        comp, comp_off, comp_len = compress(data, off, length)
        return comp, comp_off, comp_len.clone()

    @compress_op.register_fake
    def _(data, off, length):
        n = off.shape[0]
        return (data.new_empty(torch.library.get_ctx().new_dynamic_size()), off.new_empty(n), length.new_empty(n))

    @torch.library.custom_op("lzs_b200::decompress", mutates_args=(), device_types="cuda")
    def decompress_op(comp: torch.Tensor, comp_off: torch.Tensor, comp_len: torch.Tensor, out_off: torch.Tensor,
                      out_cap: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        out, out_len, status = decompress(comp, comp_off, comp_len, out_off, out_cap, with_status=True)
        return out, out_len.clone(), status.clone()

    @decompress_op.register_fake
    def _(comp, comp_off, comp_len, out_off, out_cap):
        n = comp_off.shape[0]
        return (comp.new_empty(torch.library.get_ctx().new_dynamic_size()), comp_len.new_empty(n),
                comp.new_empty(n))
except (AttributeError, RuntimeError):          # an older torch without torch.library.custom_op
    compress_op = decompress_op = None
