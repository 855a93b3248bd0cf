"""Multi-GPU plumbing for the LZS batch path (one process per GPU, torch.distributed).

Chunks / packets are independent streams, so the data path needs no collective: every
rank compresses a contiguous range of stream indices on its own GPU.  The only exchange
is the optional final gather of the variable-size outputs, an
all-gather-v:
  1. every rank packs its streams back to back on the device (the library's pack kernel on
     CUDA tensors: one block per stream, 16-byte vectors; a torch gather on CPU tensors);
  2. one all-gather of (stream count, byte count) and one of the per-stream lengths;
  3. the payloads travel with their EXACT sizes, every rank receiving each peer's payload
     straight into its place in one preallocated buffer (grouped send/recv -- NCCL runs the
     group as one operation over NVLink; nothing is padded to the largest rank and nothing
     is concatenated afterwards).
Works on NCCL (CUDA tensors) and gloo (CPU tensors; the multi-rank tests run that).
"""
import torch
import torch.distributed as dist

ALIGN = 16


def shard_range(n_streams, rank, world):
    """Contiguous stream range [lo, hi) of `rank`; ranges differ by at most one stream."""
    lo = (n_streams * rank) // world
    hi = (n_streams * (rank + 1)) // world
    return lo, hi


def packed_layout(length, align=ALIGN):
    """Offsets of streams put back to back, each starting at a multiple of `align`.
    Returns (offsets int64, total bytes)."""
    length = length.to(torch.int64)
    padded = (length + (align - 1)) // align * align
    off = torch.cumsum(padded, 0) - padded
    total = int(padded.sum().item()) if length.numel() else 0
    return off, total


def pack_streams(buf, off, length, align=ALIGN):
    """Streams buf[off[s] : off[s]+length[s]] put back to back (each at a multiple of `align`).
    Returns (packed uint8 tensor, offsets into it).  CUDA tensors: the library's pack kernel on
    the current stream; CPU tensors: plain slicing (small test sizes)."""
    dst_off, total = packed_layout(length, align)
    packed = torch.zeros(total, dtype=torch.uint8, device=buf.device)
    n = int(length.numel())
    if n == 0 or total == 0:
        return packed, dst_off
    if buf.is_cuda:
        import lzs_b200 as B
        src_off = off.to(torch.int64).contiguous()
        len32 = length.to(torch.int32).contiguous()
        B.check(B.lib().lzs_b200_pack_streams_device(
            buf.data_ptr(), src_off.data_ptr(), len32.data_ptr(), packed.data_ptr(), dst_off.data_ptr(), n,
            torch.cuda.current_stream(buf.device).cuda_stream))
        return packed, dst_off
    for s in range(n):
        a, l, d = int(off[s]), int(length[s]), int(dst_off[s])
        packed[d:d + l] = buf[a:a + l]
    return packed, dst_off


def gather_compressed(buf, off, length, group=None, align=ALIGN, max_pad=1.25):
    """Pack this rank's streams and all-gather-v them in one go.  Returns (payload, lengths of all
    streams in rank order, byte offset of every stream in the payload, info dict).

    When the ranks' byte counts are close (they are for shards of one corpus) the payload is laid out
    with ONE stride per rank -- the largest rank's byte count, rounded up -- and every rank packs its
    streams straight into its own part of the gathered buffer; the exchange is then a single in-place
    NCCL all-gather, which on an NVSwitch box goes out once per rank (multicast) instead of once per
    peer.  The parts' tails are padding; the returned offsets account for it.  Uneven ranks (stride more
    than max_pad times the mean) fall back to pack_streams + all_gather_streams (exact sizes, grouped
    send/recv)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = buf.device
    length64 = length.to(torch.int64)
    local_off, local_bytes = packed_layout(length64, align)
    counts = torch.tensor([length.numel(), local_bytes], dtype=torch.int64, device=dev)
    all_counts = torch.zeros(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(world, 2).cpu()
    n_of = [int(x) for x in all_counts[:, 0]]
    b_of = [int(x) for x in all_counts[:, 1]]
    stride = (max(b_of) + 255) // 256 * 256
    uniform_n = len(set(n_of)) == 1
    if not buf.is_cuda or not uniform_n or stride * world > max_pad * max(1, sum(b_of)):
        packed, _ = pack_streams(buf, off, length, align)
        payload, lens, offsets = all_gather_streams(packed, length, group, align)
        return payload, lens, offsets, {"mode": "exact sizes, grouped send/recv", "payload_bytes": int(payload.numel())}
    import lzs_b200 as B
    payload = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    mine = payload[rank * stride:(rank + 1) * stride]
    n = int(length.numel())
    if n:
        src_off = off.to(torch.int64).contiguous()
        len32 = length.to(torch.int32).contiguous()
        B.check(B.lib().lzs_b200_pack_streams_device(
            buf.data_ptr(), src_off.data_ptr(), len32.data_ptr(), mine.data_ptr(), local_off.data_ptr(), n,
            torch.cuda.current_stream(dev).cuda_stream))
    dist.all_gather_into_tensor(payload, mine, group=group)              # in place: my part is already where it belongs
    lens = torch.empty(world * n, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens, length64.contiguous(), group=group)
    offsets = torch.empty_like(lens)
    for r in range(world):
        part = lens[r * n:(r + 1) * n]
        off_r, _ = packed_layout(part, align)
        offsets[r * n:(r + 1) * n] = off_r + r * stride
    return payload, lens, offsets, {"mode": "one stride per rank, in-place all-gather", "payload_bytes": sum(b_of),
                                    "stride": stride}


def all_gather_streams(packed, lengths, group=None, align=ALIGN):
    """All-gather-v of packed payloads (as made by pack_streams with the same `align`).
    Returns (payload of all ranks in rank order, lengths of all streams in rank order,
    byte offset of every stream in the payload)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = packed.device
    lengths = lengths.to(torch.int64).contiguous()
    counts = torch.tensor([lengths.numel(), packed.numel()], dtype=torch.int64, device=dev)
    all_counts = torch.zeros(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(world, 2).cpu()
    n_of = [int(x) for x in all_counts[:, 0]]
    b_of = [int(x) for x in all_counts[:, 1]]
    n_off = [sum(n_of[:r]) for r in range(world)]
    b_off = [sum(b_of[:r]) for r in range(world)]
    lens = torch.empty(sum(n_of), dtype=torch.int64, device=dev)
    payload = torch.empty(sum(b_of), dtype=torch.uint8, device=dev)
    lens[n_off[rank]:n_off[rank] + n_of[rank]] = lengths
    payload[b_off[rank]:b_off[rank] + b_of[rank]] = packed
    ops = []
    for step in range(1, world):
        to, frm = (rank + step) % world, (rank - step) % world
        if n_of[rank]:
            ops.append(dist.P2POp(dist.isend, lengths, to, group))
        if n_of[frm]:
            ops.append(dist.P2POp(dist.irecv, lens[n_off[frm]:n_off[frm] + n_of[frm]], frm, group))
        if b_of[rank]:
            ops.append(dist.P2POp(dist.isend, packed, to, group))
        if b_of[frm]:
            ops.append(dist.P2POp(dist.irecv, payload[b_off[frm]:b_off[frm] + b_of[frm]], frm, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    # offsets of the streams inside the gathered payload: every rank's part starts where the
    # parts before it end, and inside a part streams sit at multiples of `align`
    offsets = torch.empty_like(lens)
    for r in range(world):
        part = lens[n_off[r]:n_off[r] + n_of[r]]
        off_r, _ = packed_layout(part, align)
        offsets[n_off[r]:n_off[r] + n_of[r]] = off_r + b_off[r]
    return payload, lens, offsets
