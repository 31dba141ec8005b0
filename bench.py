#!/usr/bin/env python
"""bench.py — north-star benchmark: GB/s of input scanned by FindAllIndex, IP regex, 16 GB
synthetic log corpus per GPU (weak scaling: every rank owns a 16 GB shard of one logical corpus;
ranks exchange nothing during the scan, NCCL only gathers the per-shard match counts).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference      # the reference's CPU path (restated oracle), host cores
  python bench.py --config c5 --gpus N  # BASELINE config 5: 64-literal scan per shard + NCCL gather
                                        # of the match offsets to rank 0 (scan and gather timed apart)

One "step" = `--passes` (default 20) full FindAllIndex passes over the resident shard, i.e. one
batch of 20 x 16 GiB of input: a single pass takes ~5 ms, and a timed region of 20 of those is too
short to be robust against host-side launch skew between ranks.  Every pass re-reads all 16 GiB
from HBM (the input is ~130x larger than L2, so no L2 flush is needed between passes) and writes
all match pairs again.  `value` = bytes scanned by all ranks / max-over-ranks device time.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line
# when NCCL_DEBUG is set in the environment): file descriptor 1 is pointed at stderr for the whole
# run and the line goes out through a private copy of the original stdout.
_JSON_OUT = None


def claim_stdout():
    """(only when run as a program: tests import this module)"""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


PATTERN = r"\d+\.\d+\.\d+\.\d+"
SEED = 0xC0FFEE
GIB = 1 << 30
METRIC = "GB/s input scanned (FindAllIndex, IP regex, 16 GB corpus)"
LIT64 = [("k%02dz%s" % (i, "q" * (i % 4))).encode() for i in range(64)]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ns", choices=["ns", "c5"])
    ap.add_argument("--gib", type=float, default=0.0, help="corpus GiB per GPU (default: 16 for ns, 8 for c5)")
    ap.add_argument("--passes", type=int, default=20, help="full passes over the shard per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-mib", type=int, default=0, help="cpu_baseline sample size (0 = auto)")
    ap.add_argument("--parity-windows", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock / throttle reasons of one GPU, sampled in-process through NVML (no fork: spawning
    nvidia-smi from a process that maps tens of GiB stalls the launching thread) from a thread that
    is started well before the timed region; only samples taken inside [t0, t1] are reported."""

    def __init__(self, gpu_index, period=0.02):
        self.idx, self.period = gpu_index, period
        self.rows, self.err = [], None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.idx
            if vis:
                try:
                    idx = int(vis.split(",")[self.idx])
                except ValueError:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as ex:  # no NVML: report that instead of clocks
            self.err = "nvml unavailable: %s" % str(ex)[:80]
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception as ex:
                self.err = str(ex)[:80]
                return
            self._stop.wait(self.period)

    def stop(self, t0, t1):
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"], "samples": 0}
        nv = self.nv
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        inside = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-1:]
        sm = sorted(r[1] for r in inside)
        reasons = sorted({nm for _, _, rs in inside for nm, bit in names if rs & bit})
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(inside), "source": "NVML in-process, %d ms period, samples inside the timed region"
                % int(self.period * 1000)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


def cpu_arm(pattern, hay, threads, passes):
    """The reference's CPU path (oracle restatement, all host threads over line-aligned shards)."""
    from oracle_lib import scan_mt
    scan_mt(pattern, hay[: min(hay.size, 64 << 20)], threads)  # warm: compile + page in
    tot, cnt = 0.0, 0
    for _ in range(passes):
        cnt, sec = scan_mt(pattern, hay, threads)
        tot += sec
    return cnt, tot, hay.size * passes / tot / 1e9


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Go toolchain does
    not exist in this image, so this is the oracle restatement (`kind: "port"`), on all host
    threads, on a bounded sample of the same corpus (same generator and seed) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import coregex_b200 as cg
    from oracle_lib import scan_mt
    c5 = args.config == "c5"
    pattern = b"|".join(LIT64).decode() if c5 else PATTERN
    sample = (args.cpu_sample_mib << 20) if args.cpu_sample_mib else (1 << 30)
    sample -= sample % 4096
    if c5:
        hay = cg.synth_host(cg.SYNTH_TEXT, SEED + 5, sample, literals=LIT64)
    else:
        hay = cg.synth_host(cg.SYNTH_LOG, SEED, sample)
    for _ in range(max(args.warmup, 1)):
        scan_mt(pattern, hay[: min(sample, 64 << 20)], threads)
    tot, cnt = 0.0, 0
    for _ in range(args.steps):
        cnt, sec = scan_mt(pattern, hay, threads)
        tot += sec
    gbs = sample * args.steps / tot / 1e9
    full = (8 if c5 else 16)
    line = {
        "impl": "reference", "metric": METRIC if not c5 else "GB/s input scanned (FindAllIndex, 64 literals, 64 GB corpus)",
        "value": round(gbs, 4),
        "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(tot / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": ("NS: `%s` FindAllIndex over a %d GiB synthetic access-log shard per GPU "
                                "(reference arm: a bounded sample of that shard, its first %d MiB, per step — "
                                "same generator and seed; throughput metric)" % (pattern, full, sample >> 20))
                   if not c5 else "C5: 64-literal FindAllIndex (reference arm: first %d MiB of shard 0 per step)" % (sample >> 20),
                   "pattern": pattern if len(pattern) < 80 else pattern[:77] + "...",
                   "sample_bytes": sample, "matches": int(cnt)},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d MiB of the corpus per step, line-aligned shards over %d threads "
                                   "(restatement of the reference CPU path built -O3 -mavx2, not the Go binary: "
                                   "no Go toolchain)" % (sample >> 20, threads)},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def init_dist():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local if world > 1 else 0)
    return world, rank, local, dist, dev


def gather_stats(dist, dev, vals):
    """every rank's list of floats -> [world][len] on all ranks"""
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if not dist:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def parity_windows(cg, out, matches, n, first_block, k, seed=SEED, kind=None, pattern=PATTERN, literals=None,
                   window=256 << 10):
    """Compares k block-aligned windows of the timed run's output with the CPU oracle on
    regenerated input (at least half of them beyond 4 GiB into the shard when the shard is that
    large).  Returns the `parity` object; raises on a mismatch."""
    import numpy as np
    import torch
    from oracle_lib import Oracle
    kind = cg.SYNTH_LOG if kind is None else kind
    o = Oracle(pattern)
    nwin = n // window
    rng = np.random.Generator(np.random.PCG64(12345))
    four = (4 * GIB) // window
    picks = set()
    if nwin > four + 8:
        picks.update([four - 1, four, nwin - 1])      # straddling 4 GiB, and the shard's last window
        while len(picks) < max(k // 2 + 2, 3):
            picks.add(int(rng.integers(four, nwin)))
    picks.add(0)
    while len(picks) < min(k, nwin):
        picks.add(int(rng.integers(0, nwin)))
    starts = out[:matches, 0].contiguous()  # searchsorted wants a dense boundary tensor (parity check, untimed)
    base = first_block * 4096
    checked, above = 0, 0
    for w in sorted(picks):
        lo = w * window
        hay = cg.synth_host(kind, seed, window, first_block=first_block + lo // 4096, literals=literals)
        want = o.find_all(hay) + (base + lo)
        key = torch.tensor([base + lo, base + lo + window], dtype=torch.int64, device=out.device)
        i0, i1 = [int(x) for x in torch.searchsorted(starts, key).tolist()]
        got = out[i0:i1].cpu().numpy()
        if got.shape != want.shape or not np.array_equal(got, want):
            raise AssertionError("parity: window at shard offset %d differs from the oracle (%d vs %d matches)"
                                 % (lo, got.shape[0], want.shape[0]))
        checked += 1
        above += lo >= 4 * GIB
    return {"windows": checked, "window_bytes": window, "above_4gib": int(above), "ok": True,
            "checker": "oracle restatement on regenerated input, exact (start,end) equality"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c5":
        return main_c5(args)

    import numpy as np
    import torch
    import coregex_b200 as cg
    from coregex_b200 import shard

    world, rank, local, dist, dev = init_dist()
    sampler = ClockSampler(dev.index)
    sampler.start()                      # >= 1 s before the timed region (corpus generation + warm-up follow)
    n = int((args.gib or 16.0) * GIB)
    n -= n % 4096
    blocks = n // 4096
    # weak scaling: one logical corpus of world*blocks blocks, rank r owns a contiguous,
    # line-aligned range of `blocks` blocks (every 4 KB block ends with a newline)
    first_block, my_blocks = shard.shard_blocks(blocks * world, world, rank)
    assert my_blocks == blocks
    hay = torch.empty(n + 64, dtype=torch.uint8, device=dev)[:n]
    cg.synth_device(cg.SYNTH_LOG, SEED, hay.data_ptr(), n, first_block=first_block)
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    r = cg.Compile(PATTERN)
    base = first_block * 4096
    # output capacity from a counting pass (nothing about the corpus density is assumed)
    r.scan_device(hay.data_ptr(), n, cg.MODE_COUNT, 0, 0, res.data_ptr(), base)
    torch.cuda.synchronize()
    matches = int(res[0].item())
    cap = matches + 1024
    out = torch.empty((cap, 2), dtype=torch.int64, device=dev)
    P = max(1, args.passes)
    W = max(args.warmup, 3)

    def one_pass():
        r.scan_device(hay.data_ptr(), n, cg.MODE_FINDALL, out.data_ptr(), cap, res.data_ptr(), base)

    for _ in range(W):
        one_pass()
    torch.cuda.synchronize()
    assert int(res[0].item()) == matches, "FindAll and Count disagree"
    time.sleep(0.5)

    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = r.launches
    K = args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-pass kernel events on a subset of passes (every pass of the first and last step)
    kev = []
    t0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        for p in range(P):
            if i == 0 or i == K - 1:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                one_pass()
                b.record()
                kev.append((a, b))
            else:
                one_pass()
    ev1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if dist:
        dist.barrier()
    clocks = sampler.stop(t0, t1)
    ms_local = ev0.elapsed_time(ev1)
    launches = r.launches - l0
    kms = sorted(a.elapsed_time(b) for a, b in kev)
    kernel_ms = sum(kms) / len(kms)
    assert int(res[0].item()) == matches
    per_rank = gather_stats(dist, dev, [ms_local / K, ms_local / (K * P), kms[0], kms[len(kms) // 2], kms[-1],
                                        clocks.get("sm_mhz") or 0.0, float(matches)])
    ms = max(pr[0] * K for pr in per_rank)
    if dist:
        # the only collective on the path: gather per-shard (match_count, bytes) over NCCL
        allc = shard.gather_counts(dist, dev, matches, n)
        total_matches = sum(c for c, _ in allc)
        total_bytes = sum(b for _, b in allc)
    else:
        total_matches, total_bytes = matches, n
    ms_step = ms / K
    value = total_bytes * P / (ms_step * 1e-3) / 1e9

    # ---- parity of the timed output against the oracle (outside the timed region) ----
    parity = None
    if not args.no_parity:
        parity = parity_windows(cg, out, matches, n, first_block, args.parity_windows)
        oks = gather_stats(dist, dev, [1.0 if parity["ok"] else 0.0, float(parity["windows"]), float(parity["above_4gib"])])
        parity["ranks_ok"] = int(sum(o[0] for o in oks))
        parity["windows_all_ranks"] = int(sum(o[1] for o in oks))

    # ---- end-to-end through the host-buffer C-ABI call (pinned host buffers, H2D + scan + D2H) ----
    e2e = None
    if not args.no_e2e:
        try:
            import ctypes as C
            del out
            torch.cuda.empty_cache()
            hbuf = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            hbuf.copy_(hay)
            torch.cuda.synchronize()
            hout = torch.empty((matches + 16, 2), dtype=torch.int64, pin_memory=True)
            cnt = C.c_size_t(0)

            def e2e_step():
                rc = cg._lib.cgx_find_all_index(r._h, hbuf.data_ptr(), n, -1, hout.data_ptr(), hout.shape[0], C.byref(cnt))
                assert rc == 0 and cnt.value == matches

            e2e_step()  # warm (allocates the library's device staging buffers)
            if dist:
                dist.barrier()
            t0e = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0e) / args.e2e_steps
            e2e_ranks = gather_stats(dist, dev, [dt])
            dt = max(x[0] for x in e2e_ranks)
            # the host copy of the result is the same list the device-resident run produced
            hp = hout[:matches]
            e2e_ok = bool(hp[0, 0] >= base and hp[-1, 1] <= base + n and bool((hp[1:, 0] >= hp[:-1, 1]).all()))
            e2e = {"value": round(total_bytes / dt / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": n,
                   "d2h_bytes_per_step": matches * 16 + 16, "steps": args.e2e_steps,
                   "api": "cgx_find_all_index (host buffers: pinned input, pinned output; one 16 GiB pass per e2e step)",
                   "per_rank_gbs": [round(n / x[0] / 1e9, 2) for x in e2e_ranks], "ordered_nonoverlapping": e2e_ok}
            del hbuf, hout
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "GB/s", "error": str(ex)[:200]}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    algo_bytes = n + 16 * matches            # 1 B read per input byte + 16 B written per match
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = profile_traffic()
    if traffic and traffic.get("input_bytes") != n:
        traffic = None  # the committed capture is of a different corpus size
    jit_state = cg._lib.cgx_debug_jit_state(r._h)
    kname = ("cgx_flat_jit" if jit_state == 1 else
             "scan_flat_kernel" if "bitstream" in r.engine else "scan_dfa_kernel")
    roof = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak,
            "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": round(kernel_ms, 4),
            "kernel_ms_min_med_max": [round(kms[0], 4), round(kms[len(kms) // 2], 4), round(kms[-1], 4)],
            "kernel_ms_samples": len(kms),
            "input_only_frac": round(n / (kernel_ms * 1e-3) / 1e9 / peak, 4),
            "traffic": traffic.get("dram_bytes_per_launch") if traffic else None}
    if traffic:
        roof["traffic_note"] = traffic.get("note")

    cpu = None
    if not args.no_cpu and world == 1:  # reported at N=1 only (the N>1 lines scale the GPU arm)
        threads = os.cpu_count() or 1
        sample = (args.cpu_sample_mib << 20) if args.cpu_sample_mib else (1 << 30)
        sample -= sample % 4096
        hs = cg.synth_host(cg.SYNTH_LOG, SEED, sample)
        passes = 6
        cnt, sec, gbs = cpu_arm(PATTERN, hs, threads, passes)
        cpu = {"value": round(gbs, 4), "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": "first %d MiB of the corpus x %d passes, line-aligned shards over %d threads, %.1f s wall "
                         "(%.0f core-seconds); oracle restatement built -O3 -mavx2"
                         % (sample >> 20, passes, threads, sec, sec * threads)}

    line = {
        "metric": METRIC, "value": round(value, 2),
        "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "NS: `%s` FindAllIndex over a %.1f GiB synthetic access-log shard per GPU"
                               % (PATTERN, n / GIB),
                   "passes_per_step": P, "ms_per_pass": round(ms_step / P, 4),
                   "step": "one step = %d full passes over the resident %.0f GiB shard (%.0f GiB scanned per GPU per step)"
                           % (P, n / GIB, P * n / GIB),
                   "bytes_per_gpu": n, "matches_per_gpu": matches, "total_matches": total_matches,
                   "output": "int64 (start,end) pairs in global order, written to HBM every pass",
                   "l2": "inputs (16 GiB) far exceed the 126 MB L2; no flush needed",
                   "parallelism": "corpus shards, one process per GPU, NCCL all_gather of counts only",
                   "engine": r.engine, "reference_strategy": r.strategy,
                   "kernel": kname + (" (NVRTC-specialised)" if jit_state == 1 else ""),
                   "per_rank": {"ms_per_step": [round(x[0], 4) for x in per_rank],
                                "ms_per_pass": [round(x[1], 4) for x in per_rank],
                                "kernel_ms_min": [round(x[2], 4) for x in per_rank],
                                "kernel_ms_median": [round(x[3], 4) for x in per_rank],
                                "kernel_ms_max": [round(x[4], 4) for x in per_rank],
                                "sm_mhz_median": [x[5] for x in per_rank]}},
        "clocks": clocks, "gpu_launches": int(launches), "parity": parity, "e2e": e2e, "roofline": roof,
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist:
        dist.destroy_process_group()


def main_c5(args):
    """BASELINE config 5: 64-literal FindAllIndex over one 8 GB shard per GPU, match offsets gathered
    to rank 0 over NCCL.  Scan time and gather time are reported separately."""
    import numpy as np
    import torch
    import coregex_b200 as cg
    from coregex_b200 import shard
    from gpu_util import dev_corpus

    world, rank, local, dist, dev = init_dist()
    sampler = ClockSampler(dev.index)
    sampler.start()
    n = int((args.gib or 8.0) * GIB)
    n -= n % 4096
    blocks = n // 4096
    first_block, _ = shard.shard_blocks(blocks * world, world, rank)
    pattern = b"|".join(LIT64).decode()
    seed = SEED + 5
    hay = dev_corpus(cg.SYNTH_TEXT, seed, n, first_block=first_block, literals=LIT64)
    res = torch.zeros(2, dtype=torch.int64, device=dev)
    r = cg.Compile(pattern)
    base = first_block * 4096
    r.scan_device(hay.data_ptr(), n, cg.MODE_COUNT, 0, 0, res.data_ptr(), base)
    torch.cuda.synchronize()
    matches = int(res[0].item())
    cap = matches + 1024
    out = torch.empty((cap, 2), dtype=torch.int64, device=dev)
    W, K = max(args.warmup, 3), args.steps

    def scan():
        r.scan_device(hay.data_ptr(), n, cg.MODE_FINDALL, out.data_ptr(), cap, res.data_ptr(), base)

    counts = [c for c, _ in shard.gather_counts(dist, dev, matches, n)] if dist else [matches]
    gathered = {}

    def gather(fmt):
        if not dist:
            return out[:matches]
        if fmt == "compact":
            return shard.gather_offsets_compact(dist, out[:matches], counts, base, n, dst=0)
        return shard.gather_offsets(dist, out[:matches], counts)

    for _ in range(W):
        scan()
        gathered["compact"] = gather("compact")
    torch.cuda.synchronize()
    time.sleep(0.3)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = r.launches
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    t0 = time.perf_counter()
    e[0].record()
    for i in range(K):
        scan()
        e[2 * i + 1].record()
        gathered["compact"] = gather("compact")
        e[2 * i + 2].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if dist:
        dist.barrier()
    clocks = sampler.stop(t0, t1)
    scan_ms = sum(e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(K)) / K
    gath_ms = sum(e[2 * i + 1].elapsed_time(e[2 * i + 2]) for i in range(K)) / K
    tot_ms = e[0].elapsed_time(e[2 * K]) / K
    # the padded int64 all_gather (16 B/match to every rank), timed the same way for comparison
    pad_ms = None
    if dist:
        gathered["padded"] = gather("padded")
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            gathered["padded"] = gather("padded")
        b.record()
        torch.cuda.synchronize()
        pad_ms = a.elapsed_time(b) / 3
    per_rank = gather_stats(dist, dev, [scan_ms, gath_ms, tot_ms, pad_ms or 0.0, float(matches)])
    parity = None
    if not args.no_parity:
        parity = parity_windows(cg, out, matches, n, first_block, args.parity_windows, seed=seed, kind=cg.SYNTH_TEXT,
                                pattern=pattern, literals=LIT64)
        if dist and rank == 0:
            g = gathered["compact"]
            parity["gather_equals_padded"] = bool(torch.equal(g, gathered["padded"]))
            parity["gather_sorted"] = bool((g[1:, 0] >= g[:-1, 1]).all().item())
            parity["gathered_matches"] = int(g.shape[0])
            assert parity["gather_equals_padded"] and parity["gather_sorted"] and g.shape[0] == sum(counts)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    mx = lambda k: max(p[k] for p in per_rank)
    total_bytes = n * world
    line = {
        "metric": "GB/s input scanned (FindAllIndex, 64 literals, 64 GB corpus, offsets gathered to rank 0)",
        "value": round(total_bytes / (mx(2) * 1e-3) / 1e9, 2), "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(mx(2), 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C5: 64-literal alternation FindAllIndex over a %.1f GiB text shard per GPU, "
                               "match offsets gathered to rank 0 over NCCL" % (n / GIB),
                   "bytes_per_gpu": n, "matches_per_gpu": [int(p[4]) for p in per_rank],
                   "scan_ms_max": round(mx(0), 4), "gather_ms_max": round(mx(1), 4),
                   "scan_gbs": round(total_bytes / (mx(0) * 1e-3) / 1e9, 2),
                   "gather": "compact wire format (u32 low word of the shard-relative start + u16 length = 6 B/match, "
                             "per-4-GiB segment table), packed by a CUDA kernel, sent with NCCL send/recv to rank 0, "
                             "expanded there to int64 (start,end)",
                   "gather_wire_bytes_into_rank0": int(6 * (sum(counts) - counts[0])),
                   "padded_all_gather_ms": round(mx(3), 4) if dist else None,
                   "per_rank": {"scan_ms": [round(p[0], 4) for p in per_rank],
                                "gather_ms": [round(p[1], 4) for p in per_rank]},
                   "engine": r.engine, "reference_strategy": r.strategy},
        "clocks": clocks, "gpu_launches": int(r.launches - l0), "parity": parity,
        "roofline": {"bound": "hbm", "kernel": r.engine, "achieved": round((n + 16 * matches) / (per_rank[0][0] * 1e-3) / 1e9, 1),
                     "peak": peak, "unit": "GB/s", "frac": round((n + 16 * matches) / (per_rank[0][0] * 1e-3) / 1e9 / peak, 4),
                     "peak_source": peak_src, "traffic": None},
    }
    emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    claim_stdout()
    main()
