"""World-size-2 gloo tests (CPU) of the multi-rank host logic: shard planning, offset rebasing and
the count / offset gathers.  The per-shard scan is stood in for by the CPU oracle here; the GPU
scan itself is rank-local and covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IP = r"\d+\.\d+\.\d+\.\d+"


def _worker(rank, world, port, blocks, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import coregex_b200 as cg
    from coregex_b200 import shard
    from oracle_lib import Oracle
    first, cnt = shard.shard_blocks(blocks, world, rank)
    hay = cg.synth_host(cg.SYNTH_LOG, 77, 4096 * cnt, first_block=first)
    base = 4096 * first
    pairs = torch.from_numpy(Oracle(IP).find_all(hay) + base)
    counts = shard.gather_counts(dist, torch.device("cpu"), pairs.shape[0], hay.size)
    allp = shard.gather_offsets(dist, pairs, [c for c, _ in counts])
    # the compact 6 B/match exchange (send/recv to rank 0) must rebuild the same list
    lens = [b for _, b in counts]
    bases = [4096 * shard.shard_blocks(blocks, world, r)[0] for r in range(world)]
    comp = shard.gather_offsets_compact(dist, pairs, [c for c, _ in counts], base, hay.size, dst=0,
                                        shard_lens=lens, bases=bases)
    if rank == 0:
        assert torch.equal(comp, allp)
    else:
        assert comp is None
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # the bench's max-over-ranks timing reduction
    if rank == 0:
        q.put((counts, allp.numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_scan_and_gather():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import coregex_b200 as cg
    from oracle_lib import Oracle
    blocks, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, allp, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = cg.synth_host(cg.SYNTH_LOG, 77, 4096 * blocks)
    want = Oracle(IP).find_all(whole)
    assert sum(c for c, _ in counts) == len(want)
    assert sum(b for _, b in counts) == whole.size
    assert np.array_equal(allp, want)
    assert tmax == 2.0


def test_shard_planning():
    from coregex_b200 import shard
    for total, world in [(10, 1), (10, 3), (7, 8), (4096, 8)]:
        got = [shard.shard_blocks(total, world, r) for r in range(world)]
        assert got[0][0] == 0 and sum(c for _, c in got) == total
        for (f0, c0), (f1, _) in zip(got, got[1:]):
            assert f0 + c0 == f1
    buf = b"aa\nbbbb\nc\n" * 50
    b = shard.split_line_aligned(buf, 4)
    assert b[0] == 0 and b[-1] == len(buf) and all(x <= y for x, y in zip(b, b[1:]))
    assert all(buf[x - 1:x] == b"\n" for x in b[1:-1])


def test_wire_codec_host_twin_roundtrip_across_4gib():
    """The wire layout (csrc/wire.cu) on the host twin: starts on both sides of 4 GiB segment
    boundaries, lengths up to 65535, empty shard; a longer match is refused."""
    from coregex_b200 import shard
    base, shard_len = 5 << 40, (9 << 30) + 4096
    starts = np.array([0, 1, (1 << 32) - 3, (1 << 32), (1 << 32) + 7, (2 << 32) - 1, (2 << 32) + 5, shard_len - 10],
                      dtype=np.int64)
    lens = np.array([1, 65535, 2, 3, 9, 1, 4, 10], dtype=np.int64)
    pairs = torch.from_numpy(np.stack([base + starts, base + starts + lens], axis=1))
    w = shard.pack_offsets(pairs, base, shard_len)
    assert w.numel() == shard._wire_layout(len(starts), shard_len)[1]
    assert torch.equal(shard.unpack_offsets(w, len(starts), base, shard_len), pairs)
    e = torch.zeros((0, 2), dtype=torch.int64)
    assert shard.unpack_offsets(shard.pack_offsets(e, base, shard_len), 0, base, shard_len).shape == (0, 2)
    too_long = torch.tensor([[base, base + 65536]], dtype=torch.int64)
    assert shard.pack_offsets(too_long, base, shard_len) is None
