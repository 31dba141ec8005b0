"""PikeVM search kernel (coregex_b200/csrc/pike_search.cu) vs the oracle's PikeVM restatement
(oracle/pikevm.cpp, reference nfa/pikevm.go:1711-2100): patterns whose anchored DFA exceeds the
table kernels' 160 states, unanchored FindAll / Count / Match over multi-record haystacks."""
import random

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle

# every pattern here must select the PikeVM engine (checked on the CPU tier too)
PATS = [
    r"[ab]*a[ab]{8}",                       # the classic 2^n subset blow-up
    r"x[ab]*a[ab]{9}y",
    r"\d{1,3}(,\d{3}){2,6}\.\d\d\b.{0,10}USD",
    r"(?i)[a-f0-9]*c[a-f0-9]{7}\b",
    r"(?m)^[ab]*b[ab]{8}$",
    r"\b\w*e\w{8}\b",
    r"(ab|a)[ab]*b[ab]{7}(c|cd)",           # alternation priorities on both ends
]


@pytest.mark.parametrize("pat", PATS)
def test_patterns_select_the_pikevm_engine(pat):
    assert cg.Compile(pat).engine == "pikevm"


def _corpus(rng, n_lines, alphabet, specials):
    lines = []
    for _ in range(n_lines):
        parts = []
        for _ in range(rng.randrange(0, 6)):
            if rng.random() < 0.35:
                parts.append(rng.choice(specials))
            else:
                parts.append(bytes(rng.choice(alphabet) for _ in range(rng.randrange(1, 40))))
        lines.append(b" ".join(parts))
    return b"\n".join(lines) + (b"\n" if rng.random() < 0.5 else b"")


@pytest.mark.gpu
@pytest.mark.parametrize("pat", PATS)
def test_find_all_count_match_against_oracle(pat):
    rng = random.Random(hash(pat) & 0xFFFF)
    alphabet = b"abcdefABCDEF0123456789,.xy USD"
    specials = [b"abababaabababab", b"xabaaabababbay", b"xbbbabababababay", b"1,234,567.89 USD", b"12,345,678,901.00   USD", b"deadc0ffee12",
                b"ababbbababab", b"sentences", b"everywhere", b"abbbbbbbbbc", b"aabbbbbbbbbcd", b"DEADC0FFEE123"]
    r = cg.Compile(pat)
    o = Oracle(pat)
    for n_lines in (0, 1, 7, 60, 900):
        hay = _corpus(rng, n_lines, alphabet, specials)
        a = np.frombuffer(hay, dtype=np.uint8)
        want = o.find_all(a)
        got = r.find_all_index_array(hay)
        assert got.shape == want.shape and np.array_equal(got, want), (pat, n_lines, got[:4], want[:4])
        assert r.Count(hay) == len(want)
        assert r.Match(hay) == (len(want) > 0)
        # limit: the first n matches
        if len(want) > 3:
            assert np.array_equal(r.find_all_index_array(hay, n=3), want[:3])


@pytest.mark.gpu
def test_long_records_straddle_slices_and_no_trailing_delimiter():
    pat = r"[ab]*a[ab]{8}"
    r, o = cg.Compile(pat), Oracle(pat)
    rng = random.Random(9)
    # records of a few KB: most slices of 1 KB own no record start at all
    recs = [bytes(rng.choice(b"abc ") for _ in range(rng.randrange(1, 5000))) for _ in range(40)]
    hay = b"\n".join(recs)
    want = o.find_all(np.frombuffer(hay, dtype=np.uint8))
    got = r.find_all_index_array(hay)
    assert np.array_equal(got, want) and len(want) > 50
    # one single record without any delimiter
    one = bytes(rng.choice(b"ab ") for _ in range(20000))
    assert np.array_equal(r.find_all_index_array(one), o.find_all(np.frombuffer(one, dtype=np.uint8)))


@pytest.mark.gpu
def test_device_entry_base_offset_and_records():
    import torch
    from gpu_util import scan_device
    pat = r"x[ab]*a[ab]{9}y"
    r, o = cg.Compile(pat), Oracle(pat)
    hay = (b"xabbaabababbay zz xaaaaaaaaaaay\n" * 300)
    a = np.frombuffer(hay, dtype=np.uint8)
    want = o.find_all(a)
    t = torch.from_numpy(a.copy()).cuda()
    tot, flag, pairs = scan_device(r, t, base=1 << 36)
    assert tot == len(want) and flag == 1 and np.array_equal(pairs, want + (1 << 36))
    tot, flag, _ = scan_device(r, t, mode=cg.MODE_COUNT)
    assert tot == len(want)


# ---- patterns that can match the empty string (reference meta/findall.go:247-279) --------------------
NULLABLE = [r"a*", r"\d*", r"(?m)^", r"(?m)$", r"\b", r"x*y?", r"foo|bar|", r"(?m)^\s*", r"\w*@?"]


@pytest.mark.gpu
def test_empty_match_rule_reference_vector():
    # the reference's own example (meta/findall.go:249): "a*" on "ab" -> [[0 1] [2 2]]
    assert cg.Compile("a*").FindAllIndex(b"ab") == [[0, 1], [2, 2]]
    assert cg.Compile("a*").FindAllIndex(b"") == [[0, 0]]
    assert cg.Compile("a*").FindAllIndex(b"baaa\n") == [[0, 0], [1, 4], [5, 5]]
    assert cg.Compile(r"\d*").Count(b"12 3") == 2


@pytest.mark.gpu
@pytest.mark.parametrize("pat", NULLABLE)
def test_nullable_patterns_against_oracle(pat):
    rng = random.Random(hash(pat) & 0xFFFF)
    r, o = cg.Compile(pat), Oracle(pat)
    assert r.engine == "pikevm"
    for hay in [b"", b"\n", b"a", b"ab", b"aa\n", b"\n\n", b"xy foo 12\nbar@\n\n aaa", b" \t x\n\ty"]:
        want = o.find_all(np.frombuffer(hay, dtype=np.uint8))
        got = r.find_all_index_array(hay, cap=4 * len(hay) + 8)
        assert np.array_equal(got, want), (pat, hay, got.tolist(), want.tolist())
    for n_lines in (3, 40, 400):
        hay = _corpus(rng, n_lines, b"ab12 xy@\t", [b"foo", b"bar", b"aaa", b"123", b"x@y", b"  "])
        a = np.frombuffer(hay, dtype=np.uint8)
        want = o.find_all(a)
        got = r.find_all_index_array(hay, cap=2 * len(hay) + 8)
        assert got.shape == want.shape and np.array_equal(got, want), (pat, n_lines)
        assert r.Count(hay) == len(want) and r.Match(hay)


@pytest.mark.gpu
def test_patterns_without_a_record_delimiter_scan_the_haystack_as_one_record():
    rng = random.Random(3)
    for pat in [r"(?s)x.y", r"(?m)^POST\s+\S+", r"(?s)a.+?b"]:
        r, o = cg.Compile(pat), Oracle(pat)
        assert r.engine == "pikevm-serial"
        hay = _corpus(rng, 300, b"abxy \tPOST", [b"POST /a", b"x\ny", b"POST\n\n /b c", b"a\nb"])
        want = o.find_all(np.frombuffer(hay, dtype=np.uint8))
        got = r.find_all_index_array(hay)
        assert got.shape == want.shape and np.array_equal(got, want) and len(want) > 5, (pat, len(want))
        assert r.Count(hay) == len(want)
