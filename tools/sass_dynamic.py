#!/usr/bin/env python
"""Joins an ncu source-page export (SASS rows in address order, `ncu -i rep --page source --csv`)
with the line table of the same cubin rebuilt here (`nvdisasm -gi`): executed warp instructions,
issue-stall samples and pipe mix per region of scan_bits.cu's kernel body.
  python tools/sass_dynamic.py <src.csv> <cubin> <scan_bits.cu as profiled> <input bytes>
"""
import collections
import csv
import re
import subprocess
import sys

csvp, cubin, srcp, nbytes = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = open(srcp).read().splitlines()
text = "\n".join(src)


def find(marker):
    off = text.index(marker)
    return text[:off].count("\n") + 1


kstart = find("CGX_DYN_SMEM(smem_raw)")
marks = [("prologue", "CGX_DYN_SMEM(smem_raw)"),
         ("chunk head: flush (wait+extract)", "const int64_t cb = chunk_origin(cur);"),
         ("phase A: tile loop", "phase A: classify the chunk's tiles"),
         ("sweep 1 (right to left)", "phase B: lane-serial marker sweeps"),
         ("sweep 2 (left to right)", "sweep 2, left to right"),
         ("replay clear + merge + counts", "const bool bad = badbits != 0ull"),
         ("publish / mail", "if (P_MODE == M_FINDALL) {\n      const uint32_t rk"),
         ("epilogue", "if (P_MODE == M_FINDALL) {\n    flush(sb, true);")]
bounds = sorted((find(m), n) for n, m in marks)


def region_of(line):
    name = "?"
    for lo, nm in bounds:
        if line >= lo:
            name = nm
    return name


fn_lines = [(i + 1, re.search(r"(\w+)\(", l.split("__device__")[1]).group(1)) for i, l in enumerate(src)
            if re.match(r"__device__ .*\b(\w+)\(", l)]


def fn_of(line):
    nm = "?"
    for lo, n in fn_lines:
        if line >= lo:
            nm = n
    return nm


sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
insts = []  # (opcode, kernel-body line, innermost line)
in_kernel, cur, inner = False, None, None
for ln in sass.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        in_kernel = ".text.cgx_flat_jit" in ln
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if inner is None:
            inner = int(m.group(2))
        cand = (m.group(3), int(m.group(4))) if m.group(3) else (m.group(1), int(m.group(2)))
        if "scan_bits" in cand[0]:
            cur = cand[1]
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m:
        insts.append((m.group(2), cur, inner))
        inner = None

rows = list(csv.reader(open(csvp)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
if abs(len(data) - len(insts)) > 1:
    # the cubin rebuilt here is not instruction-for-instruction the profiled one (another NVRTC
    # build of nearly the same source): align the two opcode sequences and carry the line table over
    import difflib
    ops_prof = [re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", r[ci["Source"]]).group(1) for r in data]
    sm = difflib.SequenceMatcher(None, ops_prof, [i[0] for i in insts], autojunk=False)
    mapped, last = [], insts[0]
    j_of = {}
    for a0, b0, sz in sm.get_matching_blocks():
        for k in range(sz):
            j_of[a0 + k] = b0 + k
    for i, op in enumerate(ops_prof):
        if i in j_of:
            last = insts[j_of[i]]
        mapped.append((op, last[1], last[2]))
    sys.stderr.write("aligned %d profiled instructions with %d rebuilt ones (%d matched)\n" % (len(data), len(insts), len(j_of)))
    insts = mapped


def pipe(op):
    b = op.split(".")[0]
    if b in ("IMAD", "IDP", "IDP4A"):
        return "fma"
    if b in ("POPC", "BREV", "FLO", "MUFU"):
        return "xu"
    if b in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "LDC", "LDCU", "SHFL", "REDUX",
             "MATCH", "VOTE", "SYNCS", "UBLKCP", "VOTEU"):
        return "lsu"
    if b in ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "NANOSLEEP", "BAR", "YIELD", "NOP", "BREAK", "BMOV"):
        return "ctrl"
    if b.startswith("U") or b in ("R2UR", "ELECT"):
        return "unif"
    return "alu"


agg = collections.defaultdict(collections.Counter)
for (op, line, inner), row in zip(insts, data):
    ex = float(row[ci["Instructions Executed"]] or 0)
    smp = float(row[ci["# Samples"]] or 0)
    key = region_of(line) if line and line >= kstart else "fn " + fn_of(line or 0)
    a = agg[key]
    a["exec"] += ex
    a[pipe(op)] += ex
    a["samples"] += smp
    if op.startswith("IMAD.MOV") or op.startswith("MOV"):
        a["mov"] += ex
tile = 2048.0
tot = collections.Counter()
print("%-36s %9s %7s %6s %6s %5s %5s %5s %5s %5s %7s" % ("region", "per 2 KB", "alu", "fma", "xu", "lsu", "ctrl", "unif", "mov", "", "samples"))
ntiles = nbytes / tile
for k in sorted(agg, key=lambda k: -agg[k]["exec"]):
    a = agg[k]
    print("%-36s %9.1f %7.1f %6.1f %6.1f %5.1f %5.1f %5.1f %5.1f %5s %6.1f%%" % (
        k, a["exec"] / ntiles, a["alu"] / ntiles, a["fma"] / ntiles, a["xu"] / ntiles, a["lsu"] / ntiles, a["ctrl"] / ntiles,
        a["unif"] / ntiles, a["mov"] / ntiles, "", 0))
    tot.update(a)
print("%-36s %9.1f %7.1f %6.1f %6.1f %5.1f %5.1f %5.1f %5.1f" % ("total", tot["exec"] / ntiles, tot["alu"] / ntiles, tot["fma"] / ntiles,
                                                          tot["xu"] / ntiles, tot["lsu"] / ntiles, tot["ctrl"] / ntiles, tot["unif"] / ntiles, tot["mov"] / ntiles))
print("warp-instructions per byte: %.4f" % (tot["exec"] / nbytes))
allsmp = sum(a["samples"] for a in agg.values())
print("stall samples by region: " + ", ".join("%s %.1f%%" % (k, 100 * agg[k]["samples"] / allsmp) for k in sorted(agg, key=lambda k: -agg[k]["samples"])[:8]))
