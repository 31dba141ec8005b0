#!/usr/bin/env python
"""Where a chunk's time goes: builds the kernel with -DCGX_TIMING (clock64 around each phase of the
scanning warps, accumulated per warp), runs the north-star scan once and prints the shares.
Diagnostic only — the timing build is slower and is never what bench.py or the tests load.

  python tools/phase_timing.py build     (here, no GPU)   -> coregex_b200/lib/libcoregex_b200_timing.so
  python tools/phase_timing.py run [GiB] (on the GPU box)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "coregex_b200", "lib", "libcoregex_b200_timing.so")

if sys.argv[1] == "build":
    from coregex_b200 import build as b
    cmd = ["nvcc"] + b.NVCC_FLAGS + ["-DCGX_TIMING", "-o", SO] + b._sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stderr[-400:] if r.returncode else "built " + SO)
    sys.exit(r.returncode)

os.environ["COREGEX_B200_LIB"] = SO
import ctypes as C

import torch

import coregex_b200 as cg
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import dev_corpus, scan_device

gib = float(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(gib * (1 << 30)) // 4096 * 4096
t = dev_corpus(0, 0xC0FFEE, n)
r = cg.Compile(sys.argv[3] if len(sys.argv) > 3 else r"\d+\.\d+\.\d+\.\d+")
for _ in range(2):
    total, _, _ = scan_device(r, t)
out = (C.c_uint64 * 8)()
cg._lib.cgx_debug_scratch.argtypes = [C.c_void_p, C.c_void_p]
assert cg._lib.cgx_debug_scratch(r._h, out) == 0
names = {2: "wait for TMA load", 3: "phase A (classify + filter)", 5: "line starts + next ticket",
         6: "wait: staging buffer released", 7: "barrier after phase A"}
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
scan_device(r, t)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
import subprocess as sp
mhz = float(sp.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"], capture_output=True,
                   text=True).stdout.split()[0])
warps = 148 * 3 * 8
cyc = ms * 1e-3 * mhz * 1e6  # cycles each resident scanning warp lived
print("matches", total, "engine", r.engine, "ms", round(ms, 3), "sm MHz", mhz)
tot = sum(out[k] for k in names)
for k, v in names.items():
    print("%-34s %14d cycles  %6.2f %% of the instrumented part" % (v, out[k], 100.0 * out[k] / tot))
