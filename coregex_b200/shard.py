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


# ---- compact offset gather (6 B/match wire format, grouped send/recv to one rank) ----
def _wire_layout(count, shard_len):
    """(nseg, total bytes, lo offset, len offset) of one shard's message; mirrors csrc/wire.cu."""
    nseg = ((shard_len + (1 << 32) - 1) >> 32) + 1
    lo_off = nseg * 8
    len_off = lo_off + count * 4
    return nseg, len_off + ((count * 2 + 7) & ~7), lo_off, len_off


def _pack_host(pairs, base, shard_len):
    """Host twin of csrc/wire.cu pack_kernel for CPU tensors (the gloo tests of this exchange)."""
    import torch
    count = pairs.shape[0]
    nseg, total, lo_off, len_off = _wire_layout(count, shard_len)
    p = pairs.numpy()
    rel = p[:, 0] - base
    ln = p[:, 1] - p[:, 0]
    if count and (ln.min() < 0 or ln.max() > 0xFFFF or rel.min() < 0 or (rel.max() >> 32) >= nseg):
        return None
    w = np.zeros(total, dtype=np.uint8)
    w[:lo_off].view(np.uint64)[:] = np.searchsorted(p[:, 0], base + (np.arange(nseg, dtype=np.int64) << 32))
    w[lo_off:len_off].view(np.uint32)[:] = (rel & 0xFFFFFFFF).astype(np.uint32)
    w[len_off:len_off + 2 * count].view(np.uint16)[:] = ln.astype(np.uint16)
    return torch.from_numpy(w)


def _unpack_host(wire, count, base, shard_len):
    import torch
    nseg, total, lo_off, len_off = _wire_layout(count, shard_len)
    w = wire.numpy()
    seg_first = w[:lo_off].view(np.uint64).astype(np.int64)
    lo = w[lo_off:len_off].view(np.uint32).astype(np.int64)
    ln = w[len_off:len_off + 2 * count].view(np.uint16).astype(np.int64)
    seg = np.searchsorted(seg_first[1:], np.arange(count, dtype=np.int64), side="right")
    start = base + (seg.astype(np.int64) << 32) + lo
    return torch.from_numpy(np.stack([start, start + ln], axis=1))


def pack_offsets(pairs, base, shard_len):
    """Sorted [count,2] int64 pairs of one shard -> uint8 wire tensor (None: a match does not fit
    the 16-bit length, send plain pairs).  CUDA kernel for device tensors."""
    import torch
    if not pairs.is_cuda:
        return _pack_host(pairs, base, shard_len)
    import coregex_b200 as cg
    count = pairs.shape[0]
    wire = torch.empty(cg.wire_bytes(count, shard_len), dtype=torch.uint8, device=pairs.device)
    bad = torch.zeros(1, dtype=torch.int64, device=pairs.device)
    st = torch.cuda.current_stream(pairs.device).cuda_stream
    cg.pack_offsets_device(pairs.data_ptr() if count else 0, count, base, shard_len, wire.data_ptr(), bad.data_ptr(), st)
    return wire, bad


def unpack_offsets(wire, count, base, shard_len, out=None):
    import torch
    if not wire.is_cuda:
        return _unpack_host(wire, count, base, shard_len)
    import coregex_b200 as cg
    if out is None:
        out = torch.empty((count, 2), dtype=torch.int64, device=wire.device)
    st = torch.cuda.current_stream(wire.device).cuda_stream
    cg.unpack_offsets_device(wire.data_ptr(), count, base, shard_len, out.data_ptr() if count else 0, st)
    return out


def gather_offsets_compact(dist, pairs, counts, base, shard_len, dst=0, shard_lens=None, bases=None):
    """Gather every rank's sorted (start,end) pairs to rank `dst` in the compact wire format:
    each rank packs its pairs (6 B/match), sends ONE message to dst (send/recv, no padding, nobody
    but dst receives anything), dst expands them into one [sum(counts), 2] int64 tensor in rank
    order (global match order for contiguous shards).  Returns that tensor on dst, None elsewhere.
    Equal-length contiguous shards are assumed unless shard_lens / bases are given."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    shard_lens = shard_lens or [shard_len] * world
    if bases is None:
        bases = [base + (r - rank) * shard_len for r in range(world)]
    packed = pack_offsets(pairs, base, shard_len)
    bad = None
    if isinstance(packed, tuple):
        wire, bad = packed
    else:
        wire = packed
    if rank != dst:
        if wire is None or (bad is not None and int(bad.item())):
            raise ValueError("a match longer than 65535 bytes does not fit the compact wire format; use gather_offsets")
        dist.send(wire, dst)
        return None
    total = int(sum(counts))
    out = torch.empty((total, 2), dtype=torch.int64, device=pairs.device)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    bufs, reqs = {}, []
    for r in range(world):
        if r == dst:
            continue
        nbytes = _wire_layout(int(counts[r]), shard_lens[r])[1]
        bufs[r] = torch.empty(nbytes, dtype=torch.uint8, device=pairs.device)
        reqs.append(dist.irecv(bufs[r], r))
    out[offs[dst]:offs[dst + 1]] = pairs
    for q in reqs:
        q.wait()
    for r, buf in bufs.items():
        c = int(counts[r])
        seg = out[offs[r]:offs[r + 1]]
        if buf.is_cuda:
            unpack_offsets(buf, c, bases[r], shard_lens[r], out=seg)
        else:
            seg.copy_(unpack_offsets(buf, c, bases[r], shard_lens[r]))
    return out
