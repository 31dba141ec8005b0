#!/usr/bin/env python
"""A/B of the two kernels that can run a flat deterministic pattern: the bitstream kernel
(scan_bits.cu) and the candidate+DFA kernel (scan_dfa.cu), on the same device-resident corpus.
Checks that the two independent implementations produce IDENTICAL (start,end) arrays at full size
(a size-independent parity property next to the oracle window check) and times both.
  python tools/ab_flat.py [GiB ...]            (default 1 16)
"""
import ctypes as C
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

GIB = 1 << 30


def timed(fn, steps=int(os.environ.get("AB_STEPS", "5")), warm=3):
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


cg._lib.cgx_debug_scratch.argtypes = [C.c_void_p, C.c_void_p]


def scratch(r):
    buf = (C.c_uint64 * 8)()
    cg._lib.cgx_debug_scratch(r._h, buf)
    return [int(x) for x in buf]


def run(pattern, kind, seed, nbytes, cap_div=40, window=64 * 4096):
    bs = cg.SYNTH_BLOCK[kind]
    nbytes -= nbytes % (bs * (window // bs))
    t = dev_corpus(kind, seed, nbytes)
    r = cg.Compile(pattern)
    cap = nbytes // cap_div
    outs, line = [], {"pattern": pattern, "bytes": nbytes, "engine": r.engine}
    arms = (("bitstream", 1), ("bitstream_generic", 2), ("dfa", 0))
    if os.environ.get("AB_ARMS") == "jit":  # quick experiments: only the specialised kernel
        arms = (("bitstream", 1),)
    for name, on in arms:
        r.set_bitstream(on)
        out = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
        res = torch.zeros(2, dtype=torch.int64, device="cuda")
        fn = lambda: r.scan_device(t.data_ptr(), nbytes, cg.MODE_FINDALL, out.data_ptr(), cap, res.data_ptr())
        ms = timed(fn)
        total = int(res[0].item())
        sc = scratch(r)
        line[name] = {"ms": round(ms, 3), "GBps": round(nbytes / ms / 1e6, 1), "matches": total,
                      "serial_replays": sc[2] if on else None, "redo_chunks": sc[3] if on else None}
        outs.append((total, out))
    if len(outs) == 1:
        print(json.dumps(line), flush=True)
        return line
    (ta, oa), (tg, og), (tb, ob) = outs
    line["identical"] = bool(ta == tb == tg and ta <= cap and torch.equal(oa[:ta], ob[:tb]) and
                             torch.equal(oa[:ta], og[:tg]))
    # oracle on a regenerated window
    wblocks = window // bs
    b0 = (nbytes // bs // 2 // wblocks) * wblocks
    hay = cg.synth_host(kind, seed, window, first_block=b0)
    lo = b0 * bs
    got = oa[:ta].cpu().numpy()
    i0, i1 = np.searchsorted(got[:, 0], lo), np.searchsorted(got[:, 0], lo + window)
    line["oracle_window_ok"] = bool(np.array_equal(got[i0:i1], Oracle(pattern).find_all(hay) + lo))
    print(json.dumps(line), flush=True)
    del t, outs, oa, ob, og
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    sizes = [float(x) for x in sys.argv[1:]] or [1.0, 16.0]
    for g in sizes:
        run(r"\d+\.\d+\.\d+\.\d+", cg.SYNTH_LOG, 0xC0FFEE, int(g * GIB))
    npat = int(os.environ.get("AB_PATS", "4"))  # quick experiments: fewer patterns
    if npat > 1:
        run(r"\w+@\w+\.\w+", cg.SYNTH_EMAIL, 0xC0FFEE + 4, 80 * 10_000_000, window=80 * 4000)
    if npat > 2:
        run(r"[a-z]+/\d+", cg.SYNTH_LOG, 0xC0FFEE, 1 * GIB)
    dense = int(float(os.environ.get("AB_DENSE_GIB", "1")) * GIB)
    if npat > 3:
        run(r"\d+", cg.SYNTH_LOG, 0xC0FFEE, dense, cap_div=4)
    if npat > 4:
        run(r"\w+", cg.SYNTH_LOG, 0xC0FFEE, dense, cap_div=4)
