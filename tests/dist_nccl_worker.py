"""torchrun worker of tests/test_dist_nccl.py: every rank scans its shard of one logical corpus on
its own GPU, match offsets are gathered to rank 0 over NCCL (padded all_gather and the compact
send/recv exchange); rank 0 compares both with a single-GPU scan of the whole corpus."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import coregex_b200 as cg
from coregex_b200 import shard
from gpu_util import dev_corpus


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from bench import LIT64
    cases = [("ip", r"\d+\.\d+\.\d+\.\d+", cg.SYNTH_LOG, 0xC0FFEE, None, 3 * 262144 + 5),
             ("lit64", b"|".join(LIT64).decode(), cg.SYNTH_TEXT, 0xC0FFEE + 5, LIT64, 2 * 262144 + 1)]
    report = {}
    for name, pat, kind, seed, lits, blocks in cases:
        first, cnt = shard.shard_blocks(blocks, world, rank)
        n = cnt * 4096
        t = dev_corpus(kind, seed, n, first_block=first, literals=lits)
        r = cg.Compile(pat)
        res = torch.zeros(2, dtype=torch.int64, device=dev)
        cap = n // 16
        out = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        after = (blocks - first - cnt) * 4096
        r.scan_device(t.data_ptr(), n, cg.MODE_FINDALL, out.data_ptr(), cap, res.data_ptr(), first * 4096,
                      bytes_after=after)
        torch.cuda.synchronize()
        m = int(res[0].item())
        assert m <= cap
        counts = shard.gather_counts(dist, dev, m, n)
        cl = [c for c, _ in counts]
        padded = shard.gather_offsets(dist, out[:m], cl)
        bases = [4096 * shard.shard_blocks(blocks, world, q)[0] for q in range(world)]
        comp = shard.gather_offsets_compact(dist, out[:m], cl, first * 4096, n, dst=0,
                                            shard_lens=[b for _, b in counts], bases=bases)
        if rank == 0:
            whole = dev_corpus(kind, seed, blocks * 4096, literals=lits)
            wout = torch.empty((blocks * 4096 // 16, 2), dtype=torch.int64, device=dev)
            r.scan_device(whole.data_ptr(), blocks * 4096, cg.MODE_FINDALL, wout.data_ptr(), wout.shape[0], res.data_ptr(), 0)
            torch.cuda.synchronize()
            wm = int(res[0].item())
            ok = wm == sum(cl) and torch.equal(wout[:wm], padded) and torch.equal(wout[:wm], comp)
            report[name] = {"matches": wm, "per_rank": cl, "ok": bool(ok)}
        else:
            assert comp is None
        dist.barrier()
    if rank == 0:
        print("NCCL_GATHER " + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
