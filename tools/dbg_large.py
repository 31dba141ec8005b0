#!/usr/bin/env python
"""debug aid: locate the match the large-set literal engine loses (test_large_prefix_free_sets...[250])"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import coregex_b200 as cg
from oracle_lib import Oracle
from test_gpu_teddy import _words
n = 250
words = _words(n)
pat = "|".join(words)
r, o = cg.Compile(pat), Oracle(pat)
rng = np.random.default_rng(n)
pieces = [w.encode() for w in words] + [w[:-1].encode() for w in words[:40]] + [(w[:2] + "Q" + w[3:]).encode() for w in words[:40]]
pieces += [b" ", b" ", b"\n", b"", b"zz", b","]
for size in (0, 5, 400, 30000):
    hay = b"".join(pieces[int(i)] + (b" " if rng.integers(0, 3) else b"") for i in rng.integers(0, len(pieces), size))
want = o.find_all(hay)
got = r.find_all_index_array(hay)
print("len", len(hay), "want", len(want), "got", len(got), "count", r.Count(hay))
ws, gs = set(map(tuple, want.tolist())), set(map(tuple, got.tolist()))
for s, e in sorted(ws - gs)[:5]:
    lit = hay[s:e]
    i = words.index(lit.decode())
    ls = hay.rfind(b"\n", 0, s) + 1
    le = hay.find(b"\n", s)
    print("MISSING", (s, e), lit, "id", i, "bucket", i % 16, "s%32768", s % 32768, "s%31744", s % 31744, "line", (ls, le), "line len", le - ls,
          "pos in line", s - ls, "ctx", hay[max(0, s - 20):e + 12])
for s, e in sorted(gs - ws)[:5]:
    print("EXTRA", (s, e), hay[s:e])
# is it the haystack position or the content? the line alone, and the haystack cut before / after
line = hay[ls:le + 1]
print("line alone:", np.array_equal(r.find_all_index_array(line), o.find_all(line)))
for cut in (ls, max(0, s - 4096) and hay.rfind(b"\n", 0, s - 4096) + 1):
    sub = hay[cut:]
    print("from", cut, np.array_equal(r.find_all_index_array(sub), o.find_all(sub)))
r.set_bitstream(1)
