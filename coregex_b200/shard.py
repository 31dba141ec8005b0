"""Corpus sharding across ranks.

Records (lines) are independent once the pattern cannot match across the delimiter, so a corpus
is split into contiguous, line-aligned byte ranges, one per rank; every rank scans its range with
the unchanged single-GPU kernel and reports offsets rebased by its shard base.  The scan itself
exchanges nothing; the only collectives are a gather of per-shard (match_count, bytes) and, for
the offset-gather workload (BASELINE config 5), a padded all_gather of the match arrays.
Pure host logic (torch.distributed with NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_blocks(total_blocks, world, rank):
    """Contiguous block range [first, first+count) of `rank`; blocks end with '\\n' by construction."""
    base, rem = divmod(total_blocks, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def split_line_aligned(buf, world):
    """Byte boundaries of `world` line-aligned shards of a host buffer (numpy uint8 / bytes)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    n = a.size
    bounds = [0]
    for r in range(1, world):
        b = max(bounds[-1], n * r // world)
        while 0 < b < n and a[b - 1] != 10:
            b += 1
        bounds.append(min(b, n))
    bounds.append(n)
    return bounds


def gather_counts(dist, device, count, nbytes):
    """all_gather of (match_count, bytes_scanned) -> list of (count, bytes) per rank."""
    import torch
    mine = torch.tensor([int(count), int(nbytes)], dtype=torch.int64, device=device)
    out = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [(int(x[0].item()), int(x[1].item())) for x in out]


def gather_offsets(dist, pairs, counts):
    """Gather variable-length (start,end) int64 arrays from every rank (padded all_gather).

    `pairs`: this rank's [count, 2] int64 tensor (already rebased to global offsets);
    `counts`: per-rank match counts from gather_counts.  Returns the concatenated [sum, 2] tensor
    in rank order, i.e. in global match order when shards are contiguous."""
    import torch
    mx = max(counts) if counts else 0
    pad = torch.zeros((mx, 2), dtype=torch.int64, device=pairs.device)
    pad[: pairs.shape[0]] = pairs
    out = [torch.zeros((mx, 2), dtype=torch.int64, device=pairs.device) for _ in counts]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
