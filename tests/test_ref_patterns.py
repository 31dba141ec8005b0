"""Sweep over EVERY pattern the reference's own test files mention (570, harvested by
tests/golden/harvest_test_patterns.py into tests/golden/ref_test_patterns.json).

CPU tier — the oracle against stdlib semantics: Go is not installed, so the expected side is Python
`re` on bytes plus Go's empty-match rule, for the patterns whose syntax and semantics coincide in the
two dialects on ASCII input (the rest is skipped with the reason counted); the product's compiler
must accept / reject exactly the patterns the oracle does, with the same error text.
GPU tier — the product against the oracle, through the C ABI: FindAllIndex, Count and Match on an
ASCII haystack and on one with multi-byte and malformed UTF-8."""
import json
import os
import re
import signal
import warnings

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle, OracleError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "ref_test_patterns.json"), encoding="utf-8") as _fh:
    PATS = json.load(_fh)

PIECES = [b"a", b"b", b"c", b"ab", b"abc", b"foo", b"bar", b"foobar", b"hello", b"world", b"test", b"123", b"4", b"56", b" ", b" ",
          b"\n", b".", b",", b"@", b"-", b"_", b"x", b"y", b"z", b"A", b"B", b"Hello", b"ERROR", b"error", b"user@example.com",
          b"192.168.1.1", b"2024-01-15", b"http://a.b/c", b"=", b"\t", b"aaa", b"bbb", b"the", b"fox"]
UTF8_PIECES = ["é".encode(), "привет".encode(), "мир".encode(), "世界".encode(), "😀".encode(), "ß".encode(), b"\xff", b"\xc3",
               b"\xe0\x80", b"\xed\xa0\x80", "ПРИВЕТ".encode(), "αβγ".encode()]


def haystacks():
    with open(os.path.join(GOLDEN, "ref_stdlib_corpus.txt"), "rb") as fh:
        corpus = fh.read()
    rng = np.random.default_rng(3)
    ascii_hay = corpus[:1500] + b"".join(PIECES[int(i)] for i in rng.integers(0, len(PIECES), 900)) + b"end"
    both = PIECES + UTF8_PIECES
    utf8_hay = b"".join(both[int(i)] for i in rng.integers(0, len(both), 1200)) + b"end"
    return ascii_hay, utf8_hay


class _Timeout(Exception):
    pass


def _go_find_all(rx, hay):
    """Go's FindAllIndex from Python's finditer; None when the two loops are not comparable for this
    pattern: after an empty match at p Go continues at p + 1, Python retries a NON-empty match at p
    (`a*?`, `(|a)*`: an empty-preferring pattern that could also consume)."""
    out, last, prev_empty_at = [], -1, -1
    for m in rx.finditer(hay):
        s, e = m.start(), m.end()
        if e > s and s == prev_empty_at:
            return None
        prev_empty_at = s if s == e else -1
        if s == e and s == last:
            continue
        out.append([s, e])
        if e > s:
            last = e
    return out


def test_oracle_equals_stdlib_semantics_on_every_reference_test_pattern():
    ascii_hay, _ = haystacks()
    stats = {"equal": 0, "dialect": 0, "rejected": 0, "loop": 0, "slow": 0}
    bad = []

    def on_alarm(*_):
        raise _Timeout()

    old = signal.signal(signal.SIGALRM, on_alarm)
    try:
        for p in PATS:
            try:
                o = Oracle(p)
            except OracleError:
                stats["rejected"] += 1
                continue
            # constructs that exist in one dialect only or mean something else in Python
            if any(ord(ch) > 127 for ch in p) or "[:" in p or "{," in p or r"\C" in p:
                stats["dialect"] += 1
                continue
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("error")
                    rx = re.compile(p.encode())
            except Exception:
                stats["dialect"] += 1
                continue
            signal.alarm(5)
            try:
                want = _go_find_all(rx, ascii_hay)
            except _Timeout:
                stats["slow"] += 1
                continue
            finally:
                signal.alarm(0)
            if want is None:
                stats["loop"] += 1
                continue
            if o.find_all(ascii_hay).tolist() == want:
                stats["equal"] += 1
            else:
                bad.append((p, o.strategy))
    finally:
        signal.signal(signal.SIGALRM, old)
    assert not bad, bad[:10]
    assert stats["equal"] >= 500, stats


def test_oracle_captures_equal_stdlib_semantics_on_every_reference_test_pattern():
    """The same sweep for FindAllSubmatchIndex: every pattern with capture groups (125 of the 570)
    against Python `re` group spans.  One reference rule is applied to the expected side: the groups
    of the empty match AT len(haystack) are unset (nfa/pikevm.go:2201-2206)."""
    ascii_hay, _ = haystacks()
    equal, bad = 0, []

    def on_alarm(*_):
        raise _Timeout()

    old = signal.signal(signal.SIGALRM, on_alarm)
    try:
        for p in PATS:
            try:
                o = Oracle(p)
            except OracleError:
                continue
            if o.num_captures < 2 or any(ord(ch) > 127 for ch in p) or "[:" in p or "{," in p or r"\C" in p:
                continue
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("error")
                    rx = re.compile(p.encode())
            except Exception:
                continue
            signal.alarm(5)
            try:
                want, last, prev_empty_at, comparable = [], -1, -1, True
                for m in rx.finditer(ascii_hay):
                    s, e = m.start(), m.end()
                    if e > s and s == prev_empty_at:
                        comparable = False   # see _go_find_all
                        break
                    prev_empty_at = s if s == e else -1
                    if s == e and s == last:
                        continue
                    row = [s, e]
                    for g in range(1, rx.groups + 1):
                        row += [m.start(g), m.end(g)]
                    want.append(row)
                    if e > s:
                        last = e
            except _Timeout:
                continue
            finally:
                signal.alarm(0)
            if not comparable:
                continue
            got = o.find_all_submatch(ascii_hay).tolist()
            if want and got and want[-1][0] == len(ascii_hay) == got[-1][0] and got[-1][2:] == [-1] * (len(got[-1]) - 2):
                # the shortcut fired (it depends on where the last search started)
                want[-1] = [want[-1][0], want[-1][1]] + [-1] * (len(want[-1]) - 2)
            if got == want:
                equal += 1
            else:
                bad.append(p)
    finally:
        signal.signal(signal.SIGALRM, old)
    assert not bad, bad[:10]
    assert equal >= 110, equal


def test_product_and_oracle_accept_the_same_patterns():
    for p in PATS:
        try:
            Oracle(p)
            oerr = None
        except OracleError as ex:
            oerr = str(ex)
        try:
            cg.Compile(p)
            perr = None
        except cg.UnsupportedError as ex:
            pytest.fail("no GPU engine for a pattern of the reference's tests: %r: %s" % (p, ex))
        except cg.Error as ex:
            perr = str(ex)
        assert (oerr is None) == (perr is None) and (oerr is None or perr in oerr or oerr in perr), (p, oerr, perr)


@pytest.mark.gpu
def test_gpu_equals_oracle_on_every_reference_test_pattern():
    ascii_hay, utf8_hay = haystacks()
    bad, n, nbits = [], 0, 0
    for p in PATS:
        try:
            o = Oracle(p)
        except OracleError:
            continue
        r = cg.Compile(p)
        if "bitstream" in r.engine:
            # NVRTC specialises the bitstream kernel per pattern and mode (about a second each): every
            # tenth such pattern takes that path, the others the generic build of the same kernel
            nbits += 1
            if nbits % 10:
                r.set_bitstream(2)
        for hay in (ascii_hay, utf8_hay):
            want = o.find_all(hay)
            got = r.find_all_index_array(hay)
            if got.shape != want.shape or not np.array_equal(got, want) or r.Count(hay) != len(want) \
                    or r.Match(hay) != (len(want) > 0):
                bad.append((p, r.engine, o.strategy, len(got), len(want)))
        n += 1
    assert n >= 550
    assert not bad, (len(bad), bad[:12])
