"""Multi-GPU plumbing for the LZS batch path (one process per GPU, torch.distributed).

Chunks / packets are independent streams, so the data path needs no collective: every
rank compresses a contiguous range of stream indices on its own GPU.  The only exchange
is the optional final gather of the variable-size outputs:
all-gather the per-stream lengths, then all-gather the packed payloads padded to the
largest rank (an all-gather-v).  Works on NCCL (CUDA tensors) and gloo (CPU tensors).
"""
import torch
import torch.distributed as dist


def shard_range(n_streams, rank, world):
    """Contiguous stream range [lo, hi) of `rank`; ranges differ by at most one stream."""
    lo = (n_streams * rank) // world
    hi = (n_streams * (rank + 1)) // world
    return lo, hi


def pack_streams(buf, off, length):
    """Concatenate buf[off[s] : off[s]+length[s]] for all s into one contiguous uint8 tensor."""
    length = length.to(torch.int64)
    total = int(length.sum().item())
    if total == 0:
        return buf.new_empty(0)
    start = torch.cumsum(length, 0) - length                      # exclusive prefix sum
    idx = torch.repeat_interleave(off.to(torch.int64) - start, length) + torch.arange(total, device=buf.device)
    return buf[idx]


def all_gather_streams(packed, lengths, group=None):
    """All-gather-v of packed payloads.  Returns (payload of all ranks in rank order,
    lengths of all streams in rank order, byte offset of every stream in the payload)."""
    world = dist.get_world_size(group)
    lengths = lengths.to(torch.int64)
    counts = torch.tensor([lengths.numel(), packed.numel()], dtype=torch.int64, device=packed.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    n_max = max(int(c[0]) for c in all_counts)
    b_max = max(int(c[1]) for c in all_counts)
    len_pad = torch.zeros(n_max, dtype=torch.int64, device=packed.device)
    len_pad[:lengths.numel()] = lengths
    pay_pad = torch.zeros(b_max, dtype=torch.uint8, device=packed.device)
    pay_pad[:packed.numel()] = packed
    all_len = [torch.zeros_like(len_pad) for _ in range(world)]
    all_pay = [torch.zeros_like(pay_pad) for _ in range(world)]
    dist.all_gather(all_len, len_pad, group=group)
    dist.all_gather(all_pay, pay_pad, group=group)
    lens = torch.cat([t[:int(c[0])] for t, c in zip(all_len, all_counts)])
    payload = torch.cat([t[:int(c[1])] for t, c in zip(all_pay, all_counts)])
    offsets = torch.cumsum(lens, 0) - lens
    return payload, lens, offsets
