#!/usr/bin/env python
"""Times the BASELINE.json configs 2-5 on one B200 (device-resident input, CUDA events, 3 warm-up
+ 5 timed passes each; inputs are far larger than L2) and spot-checks each against the CPU oracle
on a regenerated window.  Writes one JSON object per config.  Not the bench contract (bench.py is);
this is the evidence for the other §8 rows."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import coregex_b200 as cg
from gpu_util import dev_corpus
from oracle_lib import Oracle

LIT16 = [b"error", b"warning", b"fatal", b"critical", b"timeout", b"refused", b"denied", b"panic",
         b"overflow", b"invalid", b"missing", b"corrupt", b"expired", b"blocked", b"aborted", b"unknown"]
LIT64 = [("k%02dz%s" % (i, "q" * (i % 4))).encode() for i in range(64)]
GIB = 1 << 30
scale = float(os.environ.get("CFG_SCALE", "1.0"))


def timed(fn, steps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


E2E = os.environ.get("CFG_E2E", "1") == "1"
E2E_MAX = 4 * GIB + (1 << 20)
ONLY = os.environ.get("CFG_ONLY", "")  # e.g. "C3,C5": run only the configs whose name starts so


def run(name, pattern, kind, seed, nbytes, literals=None, submatch=False, window=64 * 4096, cap_div=40):
    if ONLY and not any(name.startswith(p) for p in ONLY.split(",")):
        return None
    bs = cg.SYNTH_BLOCK[kind]
    nbytes -= nbytes % (bs * (window // bs) if window % bs == 0 else bs)
    t = dev_corpus(kind, seed, nbytes, literals=literals)
    r = cg.Compile(pattern)
    stride = 2 * (r.NumSubexp() + 1) if submatch else 2
    cap = nbytes // cap_div
    out = torch.empty((cap, stride), dtype=torch.int64, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    if submatch:
        fn = lambda: r.scan_submatch_device(t.data_ptr(), nbytes, out.data_ptr(), cap, res.data_ptr())
    else:
        fn = lambda: r.scan_device(t.data_ptr(), nbytes, cg.MODE_FINDALL, out.data_ptr(), cap, res.data_ptr())
    ms = timed(fn)
    total = int(res[0].item())
    # parity spot check on a regenerated window
    o = Oracle(pattern)
    wblocks = window // bs
    b0 = (nbytes // bs // 2 // wblocks) * wblocks
    hay = cg.synth_host(kind, seed, window, first_block=b0, literals=literals)
    lo = b0 * bs
    got = out[:total].cpu().numpy()
    i0, i1 = np.searchsorted(got[:, 0], lo), np.searchsorted(got[:, 0], lo + window)
    want = o.find_all_submatch(hay) if submatch else o.find_all(hay)
    want = np.where(want >= 0, want + lo, want)
    ok = bool(np.array_equal(got[i0:i1], want))
    line = {"config": name, "pattern": pattern if len(pattern) < 80 else pattern[:77] + "...",
            "bytes": nbytes, "matches": total, "ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1),
            "engine": r.engine, "reference_strategy": r.strategy, "oracle_window_ok": ok,
            "frac_hbm_input_only": round(nbytes / ms / 1e6 / 6570.3, 4)}
    if E2E and nbytes <= E2E_MAX:
        # the same call through the host-buffer C ABI: pinned input and output, H2D + scan (+ captures)
        # + D2H inside the timed region, pieces pipelined over three streams
        import time
        hbuf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        hbuf.copy_(t)
        hout = torch.empty((total + 16, stride), dtype=torch.int64, pin_memory=True)
        cnt = r.find_all_into(hbuf, hout, submatch=submatch)  # warm: allocates the staging buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            cnt = r.find_all_into(hbuf, hout, submatch=submatch)
        dt = (time.perf_counter() - t0) / 3
        same = cnt == total and bool(torch.equal(hout[:total], out[:total].cpu()))
        line["e2e"] = {"GBps": round(nbytes / dt / 1e9, 2), "ms": round(dt * 1e3, 2), "h2d_bytes": nbytes,
                       "d2h_bytes": total * stride * 8, "equals_device_result": same}
        del hbuf, hout
    print(json.dumps(line), flush=True)
    del t, out
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    run("C2 IP regex, 1 GB log lines", r"\d+\.\d+\.\d+\.\d+", cg.SYNTH_LOG, 0xC0FFEE + 2, int(1 * GIB * scale))
    run("C3 16-literal Slim Teddy, 4 GB text", b"|".join(LIT16).decode(), cg.SYNTH_TEXT, 0xC0FFEE + 3,
        int(4 * GIB * scale), literals=LIT16)
    run("C4 as written (no groups), 10M x 80 B", r"\w+@\w+\.\w+", cg.SYNTH_EMAIL, 0xC0FFEE + 4,
        int(80 * 10_000_000 * scale), submatch=True, window=80 * 4000)
    run("C4 captures (\\w+)@(\\w+)\\.(\\w+), 10M x 80 B", r"(\w+)@(\w+)\.(\w+)", cg.SYNTH_EMAIL, 0xC0FFEE + 4,
        int(80 * 10_000_000 * scale), submatch=True, window=80 * 4000)
    run("C5 shape: 64-literal Fat Teddy, 8 GB shard (1 of 8)", b"|".join(LIT64).decode(), cg.SYNTH_TEXT,
        0xC0FFEE + 5, int(8 * GIB * scale), literals=LIT64)
    # beyond the BASELINE configs: the other engines on the same log corpus
    run("X1 record engine `.* 404 .*`, 1 GB log lines", r".* 404 .*", cg.SYNTH_LOG, 0xC0FFEE + 2, int(1 * GIB * scale))
    run("X2 generic DFA engine `GET /\\S+ HTTP`, 1 GB log lines", r"GET /\S+ HTTP", cg.SYNTH_LOG, 0xC0FFEE + 2,
        int(1 * GIB * scale))
    run("X3 UTF-8 dot `\"[A-Z]+ .*\" 200`, 1 GB log lines", r'"[A-Z]+ .*" 200', cg.SYNTH_LOG, 0xC0FFEE + 2,
        int(1 * GIB * scale))
    # 200 prefix-free literals: the reference's Aho-Corasick strategy, here the literal engine with 16 buckets
    import random
    rnd, big = random.Random(1), []
    while len(big) < 200:
        w = "".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(6 + len(big) % 3))
        if not any(x.startswith(w) or w.startswith(x) for x in big):
            big.append(w)
    run("X5 200 literals (Aho-Corasick strategy), 1 GB text", "|".join(big), cg.SYNTH_TEXT, 0xC0FFEE + 6,
        int(1 * GIB * scale), literals=[w.encode() for w in big[:64]])
    run("X4 char-class runs `\\w+` (dense output), 1 GB log lines", r"\w+", cg.SYNTH_LOG, 0xC0FFEE + 2, int(1 * GIB * scale),
        cap_div=4)
