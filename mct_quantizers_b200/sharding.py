"""Partitioning of the fake-quant work over the GPUs of one box.

Every output element depends only on its input element and its channel's parameters, so the path shards with
no exchange step: weights are split by layer (size-balanced) and, for one huge tensor, by channel blocks;
activations are split by batch rows.  No collective runs on the data path; NCCL is used by the harness only to
gather checksums / outputs for verification."""
from typing import List, Sequence, Tuple


def shard_layers(sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time bin packing: layer indices per rank, balanced by element count, deterministic."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    bins: List[List[int]] = [[] for _ in range(world_size)]
    load = [0] * world_size
    for idx in sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(idx)
        load[r] += int(sizes[idx])
    for b in bins:
        b.sort()
    return bins


def shard_range(n: int, world_size: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous share of n items; boundaries are multiples of `align` (except n)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    blocks = (n + align - 1) // align
    base, extra = divmod(blocks, world_size)
    start_b = rank * base + min(rank, extra)
    stop_b = start_b + base + (1 if rank < extra else 0)
    return min(start_b * align, n), min(stop_b * align, n)


def shard_batch(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Rows [start, stop) of an activation batch owned by `rank`."""
    return shard_range(batch, world_size, rank)


def shard_channel_blocks(shape: Sequence[int], channel_axis: int, world_size: int, rank: int):
    """Split one tensor along its channel axis.  Returns (index slices for the tensor, (c0, c1) for the
    per-channel parameter arrays): the rank quantizes x[slices] with scale[c0:c1], zp[c0:c1]."""
    nd = len(shape)
    axis = channel_axis % nd
    c0, c1 = shard_range(int(shape[axis]), world_size, rank)
    slices = tuple(slice(c0, c1) if d == axis else slice(None) for d in range(nd))
    return slices, (c0, c1)


def checksum64(t) -> int:
    """Order-independent 64-bit checksum of a tensor's bit patterns (sum of 32/16-bit words, mod 2^63):
    what the harness all-gathers instead of multi-GB outputs."""
    import torch
    flat = t.detach().contiguous().view(-1)
    if flat.dtype == torch.float32:
        words = flat.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    elif flat.dtype in (torch.bfloat16, torch.float16):
        words = flat.view(torch.int16).to(torch.int64) & 0xFFFF
    else:
        words = flat.to(torch.int64)
    return int(words.sum().item()) & 0x7FFFFFFFFFFFFFFF
