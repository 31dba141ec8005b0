"""-m gpu: parity of the multi-literal (Slim/Fat Teddy) engine against the CPU oracle, including
the reference's verify-order regimes (bucket-major vs literal order in the last 16 bytes)."""
import os

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LIT16 = [b"error", b"warning", b"fatal", b"critical", b"timeout", b"refused", b"denied", b"panic",
         b"overflow", b"invalid", b"missing", b"corrupt", b"expired", b"blocked", b"aborted", b"unknown"]
LIT64 = [("k%02dz%s" % (i, "q" * (i % 4))).encode() for i in range(64)]


def check(pat, hay):
    r = cg.Compile(pat)
    o = Oracle(pat)
    assert o.strategy == "UseTeddy" and r.strategy == "UseTeddy" and "teddy" in r.engine
    want = o.find_all(hay)
    got = r.find_all_index_array(hay)
    assert got.shape == want.shape and np.array_equal(got, want), (pat, len(hay))
    assert r.Count(hay) == len(want) and r.Match(hay) == (len(want) > 0)


def test_known_answers():
    r = cg.Compile("cat|dog")
    assert r.FindAllIndex(b"a cat and dog") == [[2, 5], [10, 13]]       # meta/findall_extra_test.go:336
    r = cg.Compile("foo|bar")
    assert r.FindAllIndex(b"hello foo world") == [[6, 9]]              # prefilter/teddy_test.go:95
    assert r.FindAllIndex(b"hello world") is None
    r = cg.Compile("|".join("p%02d" % i for i in range(50)))           # meta/fat_teddy_fallback_test.go
    assert r.engine == "fat-teddy" and r.FindAllIndex(b"test p42 here") == [[5, 8]]


# reference meta/ahocorasick_test.go:235-274 (70 literals: above Fat Teddy's 64) and :201-231.  The
# Aho-Corasick automaton itself is an external module (github.com/coregx/ahocorasick, not in the
# reference tree): what the reference pins is the result — here the set runs on the generic engines.
AC70 = ("alpha|bravo|charlie|delta|echo|foxtrot|golf|hotel|india|juliet|kilo|lima|mike|november|oscar|papa|quebec|romeo|"
        "sierra|tango|uniform|victor|whiskey|xray|yankee|zulu|anise|basil|cilantro|dill|endive|fennel|ginger|hops|ivory|"
        "jasmine|kelp|lavender|mint|nutmeg|oregano|parsley|quassia|rosemary|sage|thyme|urtica|verbena|wasabi|xylose|"
        "yarrow|zinnia|acacia|bamboo|cactus|dahlia|ebony|fern|grass|holly|iris|juniper|kudzu|lotus|moss|nettle|oak|plum|"
        "reed|sorrel")


def test_reference_large_literal_sets():
    r = cg.Compile(AC70)
    hay = b"this is alpha and omega, with bravo and tango at the end"
    assert r.Match(hay) and r.Count(hay) == 3
    assert r.FindAllIndex(hay)[0] == [8, 13]
    r = cg.Compile("mon|tue|wed|thu|fri|sat|sun|day|week|month")
    # (:216 also lists "monday" -> 1, but that test skips itself unless the strategy is UseAhoCorasick,
    # which ten literals never select; leftmost-first FindAll yields "mon" and "day")
    for h, want in [(b"monday", 2), (b"mon tue wed", 3), (b"year", 0), (b"day week month", 3)]:
        assert r.Count(h) == want == len(Oracle("mon|tue|wed|thu|fri|sat|sun|day|week|month").find_all(np.frombuffer(h, dtype=np.uint8)))
    rng = np.random.default_rng(5)
    words = AC70.encode().split(b"|") + [b"omega", b"alp", b"oakplum", b"fernn", b"  ", b"\n"]
    hay = b" ".join(words[i] for i in rng.integers(0, len(words), 3000))
    for pat, h in [(AC70, hay), ("|".join("w%03dx" % i for i in range(100)),
                                 b" ".join(b"w%03dx" % i for i in rng.integers(0, 140, 2000)))]:
        r = cg.Compile(pat)
        want = Oracle(pat).find_all(np.frombuffer(h, dtype=np.uint8))
        got = r.find_all_index_array(h)
        assert got.shape == want.shape and np.array_equal(got, want), (r.engine, len(h))


def _words(n, seed=1, length=6):
    import random
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        w = "".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(length + len(out) % 3))
        if not any(x.startswith(w) or w.startswith(x) for x in out):
            out.append(w)
    return out


@pytest.mark.parametrize("n", [70, 130, 250])
def test_large_prefix_free_sets_run_on_the_literal_engine(n):
    """68..255 complete literals are the reference's Aho-Corasick strategy (meta/strategy.go:1165, the
    automaton an external module).  A prefix-free set has at most one literal per start position, so
    the multi-literal engine (16 buckets, exact byte fingerprints, byte-for-byte verification) returns
    the leftmost-first list — compared with the oracle on text made of the words, their truncations
    and one-letter corruptions."""
    words = _words(n)
    pat = "|".join(words)
    r, o = cg.Compile(pat), Oracle(pat)
    assert o.strategy == "UseAhoCorasick" and r.strategy == "UseAhoCorasick" and r.engine == "teddy-large", (r.strategy, r.engine)
    rng = np.random.default_rng(n)
    pieces = [w.encode() for w in words] + [w[:-1].encode() for w in words[:40]] + [(w[:2] + "Q" + w[3:]).encode() for w in words[:40]]
    pieces += [b" ", b" ", b"\n", b"", b"zz", b","]
    for size in (0, 5, 400, 30000):
        hay = b"".join(pieces[int(i)] + (b" " if rng.integers(0, 3) else b"") for i in rng.integers(0, len(pieces), size))
        want = o.find_all(hay)
        got = r.find_all_index_array(hay)
        assert got.shape == want.shape and np.array_equal(got, want), (n, size)
        assert r.Count(hay) == len(want) and r.Match(hay) == (len(want) > 0)
    big = cg.synth_host(cg.SYNTH_TEXT, 900 + n, 4096 * 600, literals=[w.encode() for w in words[:64]])
    assert np.array_equal(r.find_all_index_array(big), o.find_all(big))


def test_large_sets_with_prefix_related_literals_keep_the_generic_engines():
    words = _words(70) + ["zzzzzzq", "zzzzzzqx"]
    r = cg.Compile("|".join(words))
    assert r.engine != "teddy-large"
    hay = b"a zzzzzzqx b " + " ".join(words[:30]).encode()
    assert np.array_equal(r.find_all_index_array(hay), Oracle("|".join(words)).find_all(hay))


def test_fixture_corpus():
    corpus = open(os.path.join(ROOT, "tests", "golden", "stdlib_corpus.txt"), "rb").read()
    check("error|warning|fatal|critical", corpus)
    check("apple|banana|cherry|grape|lemon|mango|melon|olive|peach|plum|kiwi|lime", corpus)
    check("(?i)error", corpus)
    check("GET|POST|PUT", corpus)


@pytest.mark.parametrize("blocks", [1, 9, 64, 700])
def test_slim16_on_synthetic_text(blocks):
    hay = cg.synth_host(cg.SYNTH_TEXT, 31 + blocks, 4096 * blocks, literals=LIT16)
    check(b"|".join(LIT16).decode(), hay)


@pytest.mark.parametrize("blocks", [1, 9, 300])
def test_fat64_on_synthetic_text(blocks):
    hay = cg.synth_host(cg.SYNTH_TEXT, 77 + blocks, 4096 * blocks, literals=LIT64)
    r = cg.Compile(b"|".join(LIT64).decode())
    assert r.engine == "fat-teddy"
    check(b"|".join(LIT64).decode(), hay)


def test_prefix_overlapping_literals_and_tail_regime():
    """One literal being a prefix of another makes the verify ORDER observable; the reference
    switches order when fewer than 16 bytes remain from the call position (App. B hazard 1)."""
    rng = np.random.default_rng(9)
    sets = [["aba", "abab", "bab", "ababa", "baba", "abb", "bba", "aab", "baa", "aaa", "bbb"],
            ["abc", "abcd", "abcde", "bcd", "bcde", "cde", "xab", "xabc", "cdea"],
            ["foo", "foobar", "barfoo", "bar", "oba", "oof", "rfo"]]
    for lits in sets:
        pat = "|".join(lits)
        r, o = cg.Compile(pat), Oracle(pat)
        if o.strategy != "UseTeddy":
            continue
        alphabet = np.frombuffer("".join(sorted(set("".join(lits)))).encode() + b" \n", dtype=np.uint8)
        for it in range(120):
            n = int(rng.integers(0, 120))
            h = bytes(alphabet[rng.integers(0, len(alphabet), n)])
            assert np.array_equal(r.find_all_index_array(h), o.find_all(h)), (pat, h)
        for n in [15, 16, 17, 31, 32, 33, 4095, 4096, 4100, 31744, 31760, 40000]:
            h = bytes(alphabet[rng.integers(0, len(alphabet) - 1, n)])
            assert np.array_equal(r.find_all_index_array(h), o.find_all(h)), (pat, n)


def test_dense_and_boundaries():
    check("foo|bar", b"foobar" * 20000)
    check("foo|bar", b"foo\n" * 30000)
    for off in range(31744 - 8, 31744 + 3):
        check("error|warning|fatal|critical", b"x" * off + b"critical error\nfatal")
    check("error|warning|fatal|critical", b"z" * 100000 + b"warning")
    # a record that runs past its owner's window (31744 + 1024 bytes) with a literal starting at the
    # last positions of that window: the second / third fingerprint byte lies outside the window
    for off in range(32768 - 9, 32768 + 3):
        check("error|warning|fatal|critical", b"x" * off + b"critical error fatal\nwarning")
        check("error|warning|fatal|critical", b"y" * 31744 + b"x" * off + b"errorfatalerror\n")


def test_c3_sized_properties_512mb():
    """BASELINE config 3 shape (16 literals) at 512 MB: ordered, non-overlapping, every match is a
    literal; oracle agreement on sampled 256 KB windows."""
    import torch
    from gpu_util import dev_corpus, scan_device
    n = 512 << 20
    t = dev_corpus(cg.SYNTH_TEXT, 0xC0FFEE + 3, n, literals=LIT16)
    pat = b"|".join(LIT16).decode()
    r = cg.Compile(pat)
    total, _, pairs = scan_device(r, t, cap=n // 64)
    assert total == len(pairs) and total > n // 400
    s, e = pairs[:, 0], pairs[:, 1]
    assert np.all(s[1:] >= e[:-1])
    lens = set(len(x) for x in LIT16)
    assert set(np.unique(e - s).tolist()) <= lens
    o = Oracle(pat)
    rng = np.random.default_rng(2)
    win = 64 * 4096
    for b in rng.integers(0, n // win, 24):
        lo = int(b) * win
        hay = cg.synth_host(cg.SYNTH_TEXT, 0xC0FFEE + 3, win, first_block=lo // 4096, literals=LIT16)
        want = o.find_all(hay) + lo
        i0, i1 = np.searchsorted(s, lo), np.searchsorted(s, lo + win)
        assert np.array_equal(pairs[i0:i1], want)
