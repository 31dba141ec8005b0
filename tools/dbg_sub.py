#!/usr/bin/env python
"""debug aid: FindAllSubmatchIndex of a nullable pattern, plain and pipelined, first differing row"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import coregex_b200 as cg
from oracle_lib import Oracle
log = cg.synth_host(cg.SYNTH_LOG, 22, 4096 * 500)
for pat, hay in [(r"(a*)(\d*)", log[:300000]), (r"(a*)(\d*)", log[:3000]), (r"(\d*)", log[:300000])]:
    r, o = cg.Compile(pat), Oracle(pat)
    want = o.find_all_submatch(hay)
    got = np.array(r.FindAllSubmatchIndex(hay), dtype=np.int64)
    pairs = r.find_all_index_array(hay)
    print(pat, len(hay), "piece", os.environ.get("CGX_PIPELINE_PIECE"), "want", want.shape, "got", got.shape, "pairs ok",
          np.array_equal(pairs, want[:, :2]))
    if got.shape != want.shape:
        k = 0
        while k < min(len(got), len(want)) and (got[k] == want[k]).all():
            k += 1
        print("  first diff row", k, "got", got[max(0, k - 1):k + 3].tolist(), "want", want[max(0, k - 1):k + 3].tolist())
    else:
        bad = np.nonzero((got != want).any(axis=1))[0]
        print("  rows differing", len(bad), bad[:5].tolist())
        for k in bad[:3]:
            print("   ", got[k].tolist(), want[k].tolist(), bytes(hay[want[k][0] - 3:want[k][1] + 3]))
