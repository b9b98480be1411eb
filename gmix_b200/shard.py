"""Host-side multi-GPU plumbing of the stream path: streams are independent, so
ranks own contiguous stream ranges balanced by input bytes and the only communication is one
all_gather of {compressed size, FNV-1a 64 checksum} per stream after the kernels finish."""
import numpy as np
import torch
import torch.distributed as dist

FNV_OFFSET = 0xcbf29ce484222325
FNV_PRIME = 0x100000001b3


def fnv1a64(data):
    """Same function as the device ChecksumKernel (gmix_b200/csrc/host.cu)."""
    h = FNV_OFFSET
    for x in bytes(data):
        h = ((h ^ x) * FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
    return h


def shard_ranges(lengths, world):
    """Contiguous stream ranges [lo, hi) per rank, balanced by input bytes (stream i -> the rank whose
    byte interval contains the midpoint of stream i)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if n == 0:
        return [(0, 0)] * world
    total = int(lengths.sum())
    if total == 0:
        edges = [n * r // world for r in range(world + 1)]
        return [(edges[r], edges[r + 1]) for r in range(world)]
    mid = np.cumsum(lengths) - lengths / 2.0
    owner = np.minimum((mid * world / total).astype(np.int64), world - 1)
    out, lo = [], 0
    for r in range(world):
        hi = lo + int((owner == r).sum())
        out.append((lo, hi))
        lo = hi
    return out


def gather_sizes_checksums(sizes, sums, counts, group=None):
    """all_gather of per-stream results. sizes/sums: int64 tensors of this rank's streams (same device the
    backend needs: CUDA for nccl, CPU for gloo); counts[r] = number of streams of rank r. Returns two
    int64 tensors covering all streams in global order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return sizes.clone(), sums.clone()
    width = max(counts)
    pad = torch.zeros(2 * width, dtype=torch.int64, device=sizes.device)
    pad[:sizes.numel()] = sizes
    pad[width:width + sums.numel()] = sums
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    all_sizes = torch.cat([bufs[r][:counts[r]] for r in range(world)])
    all_sums = torch.cat([bufs[r][width:width + counts[r]] for r in range(world)])
    return all_sizes, all_sums
