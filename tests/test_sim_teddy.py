"""The multi-literal flavour of the bitstream kernel (csrc/scan_teddy.cu) on the CPU SIMT emulator
against the oracle's restatement of reference prefilter/teddy.go under the FindAll loop
(meta/findall.go:176-290, meta/find_indices.go:925).  Under -m gpu the same cases run on the device
through the C ABI (tests/test_gpu_teddy.py holds the larger ones)."""
import random

import numpy as np
import pytest

import coregex_b200 as cg
import sim_lib
from oracle_lib import Oracle

LIT16 = [b"error", b"warning", b"fatal", b"critical", b"timeout", b"refused", b"denied", b"panic",
         b"overflow", b"invalid", b"missing", b"corrupt", b"expired", b"blocked", b"aborted", b"unknown"]
LIT64 = [("k%02dz%s" % (i, "q" * (i % 4))).encode() for i in range(64)]


BACKEND = "sim"


@pytest.fixture(autouse=True, params=["sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    global BACKEND
    BACKEND = request.param
    yield
    BACKEND = "sim"


def _scan(pat, hay, grid, mode=0):
    if BACKEND == "sim":
        return sim_lib.scan_teddy(pat, hay, grid=grid, mode=mode)
    import torch
    from gpu_util import scan_device
    r = cg.Compile(pat)
    assert "teddy" in r.engine
    a = np.frombuffer(bytes(hay), dtype=np.uint8) if not isinstance(hay, np.ndarray) else hay
    t = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda")
    t[a.size:] = ord("e")  # bytes after the haystack must never be interpreted
    if a.size:
        t[: a.size] = torch.from_numpy(a.copy()).cuda()
    r.set_bitstream(3)  # scan_teddy.cu (opt-in on the device: the scan_dfa.cu engine measured faster)
    tot, flag, pairs = scan_device(r, t[: a.size], mode=mode, cap=a.size + 16)
    if mode == 0:
        r.set_bitstream(1)  # the scan_dfa.cu engine: an independent implementation of the same loop
        tot2, _, pairs2 = scan_device(r, t[: a.size], mode=0, cap=a.size + 16)
        assert tot2 == tot and np.array_equal(pairs, pairs2)
    return tot, flag, pairs


def check(pat, hay, grid=2):
    if isinstance(hay, (bytes, bytearray)):
        hay = np.frombuffer(bytes(hay), dtype=np.uint8)
    o = Oracle(pat)
    assert o.strategy == "UseTeddy", (pat, o.strategy)
    want = o.find_all(hay)
    tot, flag, pairs = _scan(pat, hay, grid)
    assert tot == len(want), (pat, tot, len(want))
    assert np.array_equal(pairs, want), (pat, pairs[:5], want[:5])
    assert flag == (1 if len(want) else 0)
    return want


def test_known_answers():
    assert check("cat|dog", b"a cat and dog").tolist() == [[2, 5], [10, 13]]   # meta/findall_extra_test.go:336
    assert check("foo|bar", b"hello foo world").tolist() == [[6, 9]]          # prefilter/teddy_test.go:95
    assert len(check("foo|bar", b"hello world")) == 0
    assert check("|".join("p%02d" % i for i in range(50)), b"test p42 here").tolist() == [[5, 8]]
    assert len(check("foo|bar", b"")) == 0


@pytest.mark.parametrize("blocks", [1, 3, 9, 40])
def test_slim16_on_synthetic_text(blocks):
    hay = cg.synth_host(cg.SYNTH_TEXT, 31 + blocks, 4096 * blocks, literals=LIT16)
    assert len(check(b"|".join(LIT16).decode(), hay, grid=1 + blocks % 2)) > 0


@pytest.mark.parametrize("blocks", [1, 9, 30])
def test_fat64_on_synthetic_text(blocks):
    hay = cg.synth_host(cg.SYNTH_TEXT, 77 + blocks, 4096 * blocks, literals=LIT64)
    assert len(check(b"|".join(LIT64).decode(), hay)) > 0


def test_overlapping_and_prefix_literals():
    # a literal inside another one, common prefixes, a literal that ends where another starts:
    # the chain (continue at the match end) decides, lane and chunk borders must not matter
    rng = random.Random(3)
    pats = ["timeout|out|meo", "err|error|rorre", "abc|bcd|cde|dea", "aaa|aab|baa", "foo|foobar|bar|oba",
            "the|then|hen|enx", "xyz|yzx|zxy", "timeout|outer|meout",
            "aba|abab|bab|ababa|baba|abb|bba|aab|baa|aaa|bbb", "abc|abcd|abcde|bcd|bcde|cde|xab|xabc|cdea",
            "foo|foobar|barfoo|bar|oba|oof|rfo"]
    pats = [p for p in pats if Oracle(p).strategy == "UseTeddy"]  # (prefix factoring sends some sets elsewhere)
    assert len(pats) >= 5
    for pat in pats:
        lits = [l.encode() for l in pat.split("|")]
        for trial in range(4):
            parts = []
            size = 0
            target = rng.choice([200, 3000, 9000, 30000])
            while size < target:
                p = rng.choice(lits) if rng.random() < 0.5 else bytes(rng.choice(b"abcdefortimuxyz \n") for _ in range(rng.randrange(1, 9)))
                parts.append(p)
                size += len(p)
            check(pat, b"".join(parts), grid=rng.choice([1, 2]))


def test_dense_literal_soup_has_no_safe_point():
    # every position starts a literal: the word before a lane's region is covered end to end
    check("aaa|aab|aba", b"a" * 20000 + b"b" + b"a" * 3000)
    check("abab|baba", b"ab" * 9000)
    check("abcdefghijklmnopqrstuvwxyzabcdef|zab", b"abcdefghijklmnopqrstuvwxyz" * 800)  # a 32-byte literal


@pytest.mark.parametrize("n", [0, 1, 2, 3, 15, 16, 17, 63, 64, 65, 2047, 2048, 2049, 8063, 8064, 8065, 8127, 8128, 8129,
                               16127, 16128, 16129, 16191, 16192])
def test_sizes_around_tile_and_chunk_edges(n):
    rng = random.Random(n)
    hay = bytearray(rng.choice(b"abc dog\n") for _ in range(n))
    for at in range(0, n - 3, 61):
        hay[at:at + 3] = b"cat" if (at // 61) % 2 else b"dog"
    check("cat|dog|bird", bytes(hay[:n]), grid=1)


def test_match_straddles_every_word_offset():
    for off in range(0, 200, 7):
        hay = b"x" * (8064 - 100 + off) + b"warning" + b" " * 300 + b"error"
        check(b"|".join(LIT16).decode(), hay)


def test_end_of_haystack():
    # fewer than 16 bytes left from the search start: the reference switches to its scalar loop and
    # tries literals in pattern order (prefilter/teddy.go:401, :447-458); lanes near the end take the
    # exact replay.  (Sets in which one literal is a prefix of another — where that order would show —
    # are factored by the parser and never reach this engine.)
    for pat in ["abc|bcd|cde|dea", "timeout|out|meo", "cat|dog|bird", "xyz|yzx|zxy"]:
        lits = [l.encode() for l in pat.split("|")]
        for pre in (0, 5, 40, 8050, 8100, 16140):
            for tail in (lits[0], b"x" + lits[1], lits[0] + lits[1], lits[2] + b" " + lits[0] + b" " + lits[1], lits[0][:-1],
                         lits[1] + lits[0][:2], b"abcdea", b"timeoutout"):
                check(pat, b"y" * pre + tail)


def test_count_and_ismatch_modes():
    hay = cg.synth_host(cg.SYNTH_TEXT, 5, 4096 * 6, literals=LIT16)
    pat = b"|".join(LIT16).decode()
    want = Oracle(pat).find_all(np.frombuffer(hay, dtype=np.uint8))
    tot, flag, _ = _scan(pat, hay, 2, mode=1)
    assert tot == len(want) and flag == 1
    tot, flag, _ = _scan(pat, hay, 2, mode=2)
    assert flag == 1
    tot, flag, _ = _scan(pat, b"nothing to see here " * 500, 2, mode=2)
    assert flag == 0
