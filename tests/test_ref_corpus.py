"""Replay of the reference's own stdlib-comparison test (reference meta/stdlib_compat_test.go:
pattern table :30-72, corpus generateTestInput :144-199, both harvested by
tests/golden/harvest_stdlib_corpus.py).  The reference asserts FindAllIndex == Go stdlib for every
pattern on this corpus; Go is not installed here, so the expected side is recomputed with Python
`re` on bytes plus Go's empty-match rule (an empty match right after a non-empty one is dropped,
meta/findall.go:251-259).  CPU tier: the oracle against that; GPU tier: the product against the
oracle, through the C ABI."""
import json
import os
import re

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle, OracleError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPEC = json.load(open(os.path.join(GOLDEN, "ref_stdlib_patterns.json")))
PATS = [(p["name"], p["pattern"]) for p in SPEC["patterns"]]


def corpus(repeat=None):
    with open(os.path.join(GOLDEN, "ref_stdlib_corpus.txt"), "rb") as fh:
        block = fh.read()
    assert block.count(b"\n") == 41
    return block * (repeat or SPEC["repeat"])


def go_find_all(pat, hay):
    """Go's FindAllIndex from Python's finditer: same leftmost-first engine semantics on bytes; Go
    drops an empty match that starts where the previous non-empty match ended."""
    out, last = [], -1
    for m in re.finditer(pat.encode(), hay):
        s, e = m.start(), m.end()
        if s == e and s == last:
            continue
        out.append([s, e])
        if e > s:
            last = e
    return out


@pytest.mark.parametrize("name,pat", PATS, ids=[n for n, _ in PATS])
def test_oracle_equals_stdlib_on_reference_corpus(name, pat):
    hay = corpus(20)
    try:
        o = Oracle(pat)
    except OracleError as ex:
        pytest.skip("oracle: %s" % ex)
    want = go_find_all(pat, hay)
    got = o.find_all(hay).tolist()
    assert got == want, (name, got[:3], want[:3])
    assert o.count(hay) == len(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,pat", PATS, ids=[n for n, _ in PATS])
def test_gpu_equals_oracle_on_reference_corpus(name, pat):
    hay = corpus()
    try:
        r = cg.Compile(pat)
    except cg.UnsupportedError as ex:
        pytest.skip("outside the GPU engines: %s" % str(ex)[:80])
    want = Oracle(pat).find_all(np.frombuffer(hay, dtype=np.uint8))
    got = r.find_all_index_array(hay, cap=len(hay) + 8 if want.shape[0] > len(hay) // 64 else None)
    assert got.shape == want.shape and np.array_equal(got, want), (name, r.engine, got[:3], want[:3])
    assert r.Count(hay) == len(want)
    assert r.Match(hay) == (len(want) > 0)
