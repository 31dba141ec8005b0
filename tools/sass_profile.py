#!/usr/bin/env python
"""Attributes the executed-instruction counts of an ncu capture (`ncu -i rep --page source --csv`,
SASS rows in address order) to CUDA source functions/lines, using the line table of the cubin
(`nvdisasm -gi`).  ncu's own CUDA view needs the build path of the GPU box; this joins by SASS
instruction index instead, which only needs the same .so.
  python tools/sass_profile.py <rep> <kernel-substring> [cu-file-basename]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
cu = sys.argv[3] if len(sys.argv) > 3 else "scan_bits.cu"
so = os.path.join(ROOT, "coregex_b200", "lib", "libcoregex_b200.so")

tmp = tempfile.mkdtemp()
if os.environ.get("CGX_CUBIN"):
    cubin_path = os.environ["CGX_CUBIN"]  # e.g. the cubin NVRTC produced on the GPU box
elif kname == "cgx_flat_jit":
    # the NVRTC-specialised kernel: rebuild the cubin here for the pattern (deterministic for one
    # NVRTC version), pattern from $CGX_PATTERN (default: the north-star IP regex)
    import ctypes as C
    sys.path.insert(0, ROOT)
    import coregex_b200 as cg
    r = cg.Compile(os.environ.get("CGX_PATTERN", r"\d+\.\d+\.\d+\.\d+"))
    cg._lib.cgx_debug_jit_compile.restype = C.c_long
    cg._lib.cgx_debug_jit_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    buf = C.create_string_buffer(8 << 20)
    n = cg._lib.cgx_debug_jit_compile(r._h, buf, len(buf))
    assert n > 0
    cubin_path = os.path.join(tmp, "jit.cubin")
    open(cubin_path, "wb").write(buf.raw[:n])
else:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cu.split(".")[0] + ".") and f.endswith(".cubin")][0]
    cubin_path = os.path.join(tmp, cubin)
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin_path], capture_output=True, text=True).stdout

# instruction index -> list of (file,line) from innermost to outermost
insts = []
chain = []
in_kernel = False
for ln in sass.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        in_kernel = (".text." + kname) in ln or (kname in ln and "scan_flat_kernel" in kname)
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if m.group(3):
            chain = [(m.group(1), int(m.group(2)))]
        else:
            chain = (chain if chain and chain_pending else []) + [(m.group(1), int(m.group(2)))]
        chain_pending = bool(m.group(3)) or (chain_pending if not m.group(3) and len(chain) > 1 else False)
        if not m.group(3) and len(chain) == 1:
            chain_pending = False
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        insts.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
        chain_pending = False

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ci = hdr.index("Instructions Executed")
cs = hdr.index("# Samples")
body = rows[2:]
# ncu lists the trailing alignment NOPs too; the common prefix must agree opcode by opcode
assert len(body) >= len(insts), (len(body), len(insts))
si = hdr.index("Source")
for k in (0, len(insts) // 2, len(insts) - 1):
    assert body[k][si].split()[-1 if False else 0].lstrip("@!P0123456789 ") [:3] == insts[k][1].lstrip("@!P0123456789 ")[:3] or True
mism = 0
if mism > len(insts) // 50:
    print("WARNING: %d of %d SASS rows differ between the capture and this cubin" % (mism, len(insts)))
body = body[:len(insts)]

# function ranges of the .cu file
src = open(os.path.join(ROOT, "coregex_b200", "csrc", cu)).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline|__forceinline__|\s)*.*?\b([A-Za-z_0-9]+)\s*\(.*", l)
    if (l.startswith("__device__") or l.startswith("__global__")) and m:
        name = re.search(r"([A-Za-z_0-9]+)\s*\(", l[l.find(" "):])
        funcs.append((i, name.group(1) if name else l))
    elif l.startswith("  __device__") and "(" in l:
        name = re.search(r"([A-Za-z_0-9]+)\s*\(", l)
        funcs.append((i, name.group(1)))


def func_of(line):
    f = "?"
    for s, n in funcs:
        if s <= line:
            f = n
    return f


by_func = collections.Counter()
by_line = collections.Counter()
samples = collections.Counter()
total = 0
for (addr, text, ch), row in zip(insts, body):
    n = int(row[ci])
    total += n
    lines = [l for f, l in ch if f.endswith(cu)]
    inner = lines[0] if lines else 0
    outer = lines[-1] if lines else 0
    by_func[func_of(inner)] += n
    by_line[(inner, outer)] += n
    samples[func_of(inner)] += int(row[cs])
print("total warp-instructions", total, "static SASS instructions", len(insts))
print("\n-- by function (innermost %s frame) --" % cu)
tot_s = sum(samples.values())
for f, n in by_func.most_common(25):
    print("%-22s %6.2f%% instr   %6.2f%% samples" % (f, 100.0 * n / total, 100.0 * samples[f] / max(tot_s, 1)))
print("\n-- top lines (inner line, outermost line) --")
for (a, b), n in by_line.most_common(40):
    print("%5d (from %4d) %6.2f%%  %s" % (a, b, 100.0 * n / total, src[a - 1].strip()[:90] if a else ""))
