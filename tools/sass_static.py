#!/usr/bin/env python
"""Static SASS accounting of the NVRTC-specialised bitstream kernel (no GPU needed): builds the
cubin for a pattern, disassembles it with the line table (`nvdisasm -gi`), attributes every SASS
instruction to the line of the KERNEL BODY it was inlined into, and sums per region of
scan_bits.cu (phase A tile loop, sweep 1, sweep 2, extraction, ...), split by issue pipe.
Loop bodies are straight-line code, so a region's static count is its count per iteration.
  python tools/sass_static.py [pattern] [mode] [extra -D defs via CGX_JIT_DEFS]
"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coregex_b200 as cg

pat = sys.argv[1] if len(sys.argv) > 1 else r"\d+\.\d+\.\d+\.\d+"
src = open(os.path.join(ROOT, "coregex_b200", "csrc", "scan_bits.cu")).read().splitlines()


def find(marker, start=0):
    for i in range(start, len(src)):
        if marker in src[i]:
            return i + 1
    raise KeyError(marker)


# regions of the kernel body by marker comments (1-based line numbers, [lo, hi))
marks = [("prologue", "CGX_DYN_SMEM(smem_raw)"),
         ("chunk head (flush wait)", "const int64_t cb = chunk_origin(cur);"),
         ("phase A: tile loop", "phase A: classify the chunk's tiles"),
         ("sweep 1 (right to left)", "phase B: lane-serial marker sweeps"),
         ("sweep 2 (left to right)", "sweep 2, left to right"),
         ("replay clear + hand-over + counts", "const bool bad = badbits != 0ull"),
         ("publish / mail", "if (P_MODE == M_FINDALL) {\n      const uint32_t rk"),
         ("epilogue", "if (P_MODE == M_FINDALL) {\n    flush(sb, true);")]
kstart = find("CGX_DYN_SMEM(smem_raw)")
bounds = []
text = "\n".join(src)
for name, m in marks:
    if "\n" in m:
        off = text.index(m)
        ln = text[:off].count("\n") + 1
    else:
        ln = find(m, kstart - 1)
    bounds.append((ln, name))
bounds.sort()


def region_of(line):
    name = "before kernel"
    for lo, nm in bounds:
        if line >= lo:
            name = nm
    return name


r = cg.Compile(pat)
cg._lib.cgx_debug_jit_compile.restype = C.c_long
cg._lib.cgx_debug_jit_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
buf = C.create_string_buffer(8 << 20)
n = cg._lib.cgx_debug_jit_compile(r._h, buf, len(buf))
assert n > 0, cg._lib.cgx_last_error()
tmp = tempfile.mkdtemp()
cubin = os.path.join(tmp, "jit.cubin")
open(cubin, "wb").write(buf.raw[:n])
sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout.strip().splitlines()[-1]

ALU = ("LOP3", "SHF", "ISETP", "IADD3", "IADD", "SEL", "PRMT", "LEA", "IABS", "IMNMX", "VIMNMX", "PLOP3", "SGXT", "BMSK", "ICMP", "FSEL", "MOV", "CS2R", "UMOV")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD")
XU = ("POPC", "BREV", "FLO", "MUFU")
LSU = ("LDS", "STS", "LDG", "STG", "LD.", "ST.", "LDL", "STL", "ATOM", "RED", "LDC", "SHFL", "ATOMS", "ATOMG", "REDUX", "MATCH", "VOTE", "SYNCS", "UBLKCP")


def pipe(op):
    b = op.split(".")[0]
    if b in ("IMAD", "IDP", "IDP4A"):
        return "fma"
    if b in XU:
        return "xu"
    if b in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "LDC", "LDCU", "SHFL", "REDUX", "MATCH", "VOTE", "SYNCS", "UBLKCP", "LDSM", "VOTEU"):
        return "lsu/other"
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "NANOSLEEP", "BAR", "YIELD", "NOP", "BREAK", "BMOV", "JMP", "BRX", "DEPBAR", "MEMBAR", "FENCE", "ERRBAR", "CCTL", "ELECT", "UTMACCTL", "ACQBULK"):
        return "ctrl"
    if b.startswith("U") and b not in ("UBLKCP",):
        return "uniform"
    return "alu"


counts = collections.defaultdict(lambda: collections.Counter())
in_kernel = False
cur_line = None
pend_outer = None
for ln in sass.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        in_kernel = ".text.cgx_flat_jit" in ln
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f1, l1, f2, l2 = m.group(1), int(m.group(2)), m.group(3), m.group(4)
        # nvdisasm prints the chain innermost first, one line per level; the last line of a chain
        # has no "inlined at" part or names the kernel body
        if f2:
            cand = (f2, int(l2))
        else:
            cand = (f1, l1)
        if "scan_bits.cu" in cand[0]:
            cur_line = cand[1]
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur_line is not None:
        op = m.group(1)
        reg = region_of(cur_line) if cur_line >= kstart else ("fn@%d" % cur_line)
        counts[reg][pipe(op)] += 1
        counts[reg]["total"] += 1

# functions outside the kernel body (noinline / outlined): group by containing function name
fn_lines = [(i + 1, l) for i, l in enumerate(src) if re.match(r"__device__ .*\b(\w+)\(", l)]


def fn_of(line):
    nm = "?"
    for lo, l in fn_lines:
        if line >= lo:
            nm = re.search(r"(\w+)\(", l.split("__device__")[1]).group(1)
    return nm


agg = collections.defaultdict(collections.Counter)
for reg, c in counts.items():
    key = reg
    if reg.startswith("fn@"):
        key = "fn " + fn_of(int(reg[3:]))
    agg[key].update(c)
print("pattern %s   %s" % (pat, res))
print("%-40s %6s %5s %5s %4s %5s %5s %5s" % ("region", "total", "alu", "fma", "xu", "lsu", "ctrl", "unif"))
tot = collections.Counter()
for reg in sorted(agg, key=lambda k: (k.startswith("fn "), k)):
    c = agg[reg]
    print("%-40s %6d %5d %5d %4d %5d %5d %5d" % (reg, c["total"], c["alu"], c["fma"], c["xu"], c["lsu/other"], c["ctrl"], c["uniform"]))
    tot.update(c)
print("%-40s %6d %5d %5d %4d %5d %5d %5d" % ("whole kernel", tot["total"], tot["alu"], tot["fma"], tot["xu"], tot["lsu/other"], tot["ctrl"], tot["uniform"]))
