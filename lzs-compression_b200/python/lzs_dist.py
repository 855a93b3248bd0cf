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
    n = int(length.numel())
    fused = _symmetric_gather(B, buf, off, length, local_off, n, world, rank, stride, group)
    if fused is not None:
        payload, mode = fused
        lens = torch.empty(world * n, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(lens, length64.contiguous(), group=group)
        offsets = torch.empty_like(lens)
        for r in range(world):
            off_r, _ = packed_layout(lens[r * n:(r + 1) * n], align)
            offsets[r * n:(r + 1) * n] = off_r + r * stride
        return payload, lens, offsets, {"mode": mode, "payload_bytes": sum(b_of), "stride": stride}
    payload = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    mine = payload[rank * stride:(rank + 1) * stride]
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


_symm = {}          # (device, group) -> (capacity, symmetric buffer, handle)


def _symmetric_gather(B, buf, off, length, local_off, n, world, rank, stride, group):
    """The pack kernel FUSED with the all-gather: the gathered payload lives in a symmetric buffer (one
    allocation per rank, every rank's copy mapped into every GPU over NVLink:
    torch.distributed._symmetric_memory), and this rank's pack kernel stores each of its streams straight
    into all copies -- through the NVSwitch multicast address when the box has one (one store, replicated
    in the switch), else with one store per peer.  Cross-rank barriers before (the previous contents may
    still be read) and after (all stores have landed).  LZS_B200_GATHER = nccl | peers | multicast picks a
    path; default: the fused kernel with one store per peer up to 4 ranks, NCCL's in-place all-gather
    above (what measured fastest).  Returns (payload tensor, mode) or None (the caller uses NCCL)."""
    import os
    want = os.environ.get("LZS_B200_GATHER", "auto")
    if want == "auto":
        # measured on B200 / NVSwitch (profiles/README.md): 2 ranks: 1.7 ms fused against 2.2 ms NCCL;
        # 8 ranks: 9.8 ms fused (10.6 with multicast stores) against 9.2 ms for NCCL's in-place all-gather
        want = "peers" if world <= 4 else "nccl"
    if want == "nccl":
        return None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        dev = buf.device
        key = (dev.index, id(group))
        need = world * stride
        ent = _symm.get(key)
        if ent is None or ent[0] < need:
            cap = (need + (need >> 3) + (1 << 26)) >> 26 << 26          # some head room, 64 MiB granules
            # every rank must agree on the size of a symmetric allocation
            cap_t = torch.tensor([cap], dtype=torch.int64, device=dev)
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX, group=group)
            cap = int(cap_t.item())
            sbuf = symm_mem.empty(cap, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(sbuf, group if group is not None else dist.group.WORLD)
            ent = (cap, sbuf, hdl)
            _symm[key] = ent
        cap, sbuf, hdl = ent
        mc = int(hdl.multicast_ptr or 0)                                 # 0: no NVSwitch multicast object for this buffer
        mode = "multicast" if (mc and want == "multicast") else "peers"
        if want == "multicast" and not mc:
            return None
        src_off = off.to(torch.int64).contiguous()
        len32 = length.to(torch.int32).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        hdl.barrier(channel=0)                                           # nobody still reads the previous payload
        if n:
            if mode == "multicast":
                B.check(B.lib().lzs_b200_pack_streams_multicast_device(
                    buf.data_ptr(), src_off.data_ptr(), len32.data_ptr(), mc, rank * stride, local_off.data_ptr(), n, stream))
            else:
                B.check(B.lib().lzs_b200_pack_streams_peers_device(
                    buf.data_ptr(), src_off.data_ptr(), len32.data_ptr(), int(hdl.buffer_ptrs_dev), world, rank * stride,
                    local_off.data_ptr(), n, stream))
        hdl.barrier(channel=1)                                           # every rank's stores have landed everywhere
        return sbuf[:need], ("pack fused with the all-gather over NVLink peer memory, "
                             + ("NVSwitch multicast stores" if mode == "multicast" else "one store per peer"))
    except Exception as e:                                               # no symmetric memory on this box / torch
        if want in ("peers", "multicast"):
            raise
        _symm["error"] = repr(e)
        return None


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
