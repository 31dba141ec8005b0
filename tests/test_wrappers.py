"""N4 wrappers (coregex_b200/wrappers.py): ReplaceAll*/Expand/Split/text forms are host plumbing over
the batch match lists.  CPU tier: the pure functions are fed the ORACLE's match lists and must give
the answers of the reference's own tables (tests/golden/ref_wrapper_vectors.json, harvested by
tests/golden/harvest_wrapper_vectors.py from replace_test.go and stdlib_compat_test.go).  GPU tier:
the same tables through `Regex` on the device."""
import json
import os

import numpy as np
import pytest

import coregex_b200 as cg
from coregex_b200 import wrappers as W
from oracle_lib import Oracle

VEC = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wrapper_vectors.json"),
                     encoding="utf-8"))


class OracleRegex(W.RegexWrappers):
    """the wrappers over the CPU oracle's lists (test infrastructure: what Regex does with the device)"""

    def __init__(self, pat):
        self._pat, self._o = pat, Oracle(pat)

    def FindAllIndex(self, b, n=-1):
        if n == 0:
            return None
        m = self._o.find_all(W._b(b), n).tolist()
        return m or None

    def FindAllSubmatchIndex(self, b, n=-1):
        if n == 0:
            return None
        m = self._o.find_all_submatch(W._b(b), n).tolist()
        return m or None

    def Match(self, b):
        return self._o.is_match(W._b(b))

    def Count(self, b, n=-1):
        return len(self.FindAllIndex(b, n) or [])

    def NumSubexp(self):
        return self._o.num_captures - 1

    def String(self):
        return self._pat


def _mk(kind, pat):
    return cg.Compile(pat) if kind == "gpu" else OracleRegex(pat)


KINDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("kind", KINDS)
def test_reference_replace_literal_tables(kind):
    for v in VEC["replace_literal"]:
        r = _mk(kind, v["pattern"])
        assert r.ReplaceAllLiteralString(v["input"], v["repl"]) == v["want"], v
        assert r.ReplaceAllLiteral(v["input"].encode(), v["repl"].encode()) == v["want"].encode(), v


@pytest.mark.parametrize("kind", KINDS)
def test_reference_replace_and_expand_tables(kind):
    for v in VEC["replace"]:
        r = _mk(kind, v["pattern"])
        assert r.ReplaceAllString(v["input"], v["repl"]) == v["want"], v
    for v in VEC["expand"]:
        r = _mk(kind, v["pattern"])
        m = r.FindSubmatchIndex(v["input"].encode())
        assert bytes(r.Expand(bytearray(), v["template"], v["input"], m)).decode() == v["want"], v


@pytest.mark.parametrize("kind", KINDS)
def test_reference_split_tables(kind):
    for v in VEC["split"]:
        assert _mk(kind, v["pattern"]).Split(v["input"], v["n"]) == v["want"], v


@pytest.mark.parametrize("kind", KINDS)
def test_func_and_text_forms(kind):
    r = _mk(kind, r"\d+")
    assert r.ReplaceAllStringFunc("1 2 3", lambda s: str(int(s) * 2)) == "2 4 6"          # replace_test.go:115-157
    assert r.ReplaceAllFunc(b"abc", lambda s: b"X") == b"abc"
    assert r.ReplaceAllFunc(b"foo123bar456", lambda s: b"<" + s + b">") == b"foo<123>bar<456>"   # stdlib_compat_test.go:1060
    assert _mk(kind, "[a-c]").ReplaceAllStringFunc("defabcdef", lambda s: "x" + s + "y") == "defxayxbyxcydef"
    assert r.FindAllString("1 2 3", -1) == ["1", "2", "3"] and r.FindAllString("abc", -1) is None   # regex.go:418
    assert r.FindString("age: 42") == "42" and r.FindString("none") == "" and r.FindStringIndex("age: 42") == [5, 7]
    assert r.Find(b"none") is None and r.FindIndex(b"x") is None
    assert list(r.AllStringIndex("1 22")) == [(0, 1), (2, 4)] and list(r.AllString("1 22")) == ["1", "22"]
    assert r.AppendAllIndex([(9, 9)], b"1 2") == [(9, 9), (0, 1), (2, 3)]
    assert r.MatchString("a1") and not r.MatchString("a") and r.CountString("1 2 3") == 3
    e = _mk(kind, r"(\w+)@(\w+)\.(\w+)")
    assert e.FindStringSubmatch("user@example.com") == ["user@example.com", "user", "example", "com"]   # regex.go:638
    assert e.FindSubmatch(b"nothing") is None
    assert e.FindAllStringSubmatch("a@b.c x@y.z", -1) == [["a@b.c", "a", "b", "c"], ["x@y.z", "x", "y", "z"]]  # :1400
    assert e.FindAllSubmatch(b"a@b.c", -1) == [[b"a@b.c", b"a", b"b", b"c"]]
    o = _mk(kind, r"(a)|(b)")
    assert o.FindSubmatch(b"b") == [b"b", None, b"b"] and o.FindStringSubmatch("b") == ["b", "", "b"]
    import io
    assert r.MatchReader(io.StringIO("abc 7")) and not r.MatchReader(io.StringIO("abc"))           # regex.go:1619
    assert r.FindReaderIndex(io.StringIO("ab 77")) == [3, 5] and e.FindReaderSubmatchIndex(io.StringIO("a@b.c")) == [0, 5, 0, 1, 2, 3, 4, 5]
    u = _mk(kind, "мир")
    assert u.FindStringIndex("привет мир") == [13, 19] and u.ReplaceAllString("привет мир", "world") == "привет world"
    assert u.Split("aмирb", -1) == ["a", "b"]


def test_quote_meta_and_names():
    assert cg.QuoteMeta("hello.world") == "hello\\.world" and cg.QuoteMeta("abc") == "abc"     # regex.go:229
    assert cg.QuoteMeta(r"\.+*?()|[]{}^$") == "".join("\\" + c for c in r"\.+*?()|[]{}^$")
    r = cg.Compile(r"(?P<year>\d+)-(?P<month>\d+)")
    assert r.SubexpNames() == ["", "year", "month"]                                           # regex.go:570-574
    assert (r.SubexpIndex("year"), r.SubexpIndex("month"), r.SubexpIndex("day"), r.SubexpIndex("")) == (1, 2, -1, -1)
    assert cg.Compile(r"(?P<bob>a+)(?P<bob>b+)|(c)").SubexpIndex("bob") == 1                   # :582-585
    assert cg.Compile(r"(a)(?:b)(?P<n>c)").SubexpNames() == ["", "", "n"]
    with pytest.raises(cg.Error) as ei:
        cg.MustCompilePOSIX("a(")
    assert str(ei.value).startswith("regexp: CompilePOSIX(`a(`): error parsing regexp: missing closing )")   # regex.go:162
    assert cg.CompilePOSIX("a+|a+b").String() == "a+|a+b"
    c = r.Copy()
    assert c.String() == r.String() and c is not r and r.MarshalText() == rb"(?P<year>\d+)-(?P<month>\d+)"
