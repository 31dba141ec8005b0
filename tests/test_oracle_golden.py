"""Pins the CPU oracle against the reference's own known-answer vectors (SURVEY.md App. A, D;
each case cites the reference test that asserts it) and against Python `re` on bytes, which has
the same leftmost-first semantics as Go's regexp for these patterns (the reference's own tests
use Go stdlib as their oracle, meta/stdlib_compat_test.go:18-141)."""
import json
import os
import random
import re

import numpy as np
import pytest

from oracle_lib import Oracle, OracleError, dump_ast, lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def fa(p, h):
    return Oracle(p).find_all(h).tolist()


# (pattern, haystack, expected FindAllIndex, reference source)
KAT = [
    (r"\w+", b"the fox", [[0, 3], [4, 7]], "meta/findall_extra_test.go:318-323"),
    (r"\d+", b"a1b22c333", [[1, 2], [3, 5], [6, 9]], "meta/findall_extra_test.go:324-329"),
    (r"ab", b"xababx", [[1, 3], [3, 5]], "meta/findall_extra_test.go:330-335"),
    (r"cat|dog", b"a cat and dog", [[2, 5], [10, 13]], "meta/findall_extra_test.go:336-341"),
    (r"\d+", b"no digits", [], "meta/findall_extra_test.go:342-347"),
    (r"[1-9][0-9]*|0", b"port 8080 pid 1234", [[5, 9], [14, 18]], "meta/digit_prefilter_integration_test.go:53"),
    (r"[1-9][0-9]*", b"leading zeros: 007", [[17, 18]], "meta/digit_prefilter_integration_test.go:120"),
    (r"\d+\.\d+\.\d+", b"version 1.2.3 here", [[8, 13]], "meta/strategy_selection_test.go:61,481"),
    (r"\d+\.\d+\.\d+\.35", b"192.168.1.35 and 10.0.0.35 end", [[0, 12], [17, 26]], "meta/issue124_test.go:358"),
    (r"\w+@\w+\.org", b"a@b.org c@d.org", [[0, 7], [8, 15]], "meta/issue124_test.go:361"),
    (r"\w+@\w+\.\w+", b"Contact: user@example.com for info", [[9, 25]], "regex_test.go:233-238"),
    (r"\d+\.\d+\.\d+\.\d+", b"x10.0.0.1 y 1.2.3 z 8.8.8.8", [[1, 9], [20, 27]], "SURVEY App. A"),
    (r"\d+\.\d+\.\d+\.\d+", b"1.2.3.4.5", [[0, 7]], "SURVEY App. A"),
    (r"\d+\.\d+\.\d+\.\d+", b"1.2.3.45 ", [[0, 8]], "SURVEY App. A"),
    (r"\d+\.\d+\.\d+\.\d+", b"1.2.3.", [], "SURVEY App. A"),
]


@pytest.mark.parametrize("pat,hay,want,src", KAT)
def test_known_answers(pat, hay, want, src):
    assert fa(pat, hay) == want, src


def test_limit_n():
    # meta/digit_prefilter_integration_test.go:179: `[1-9][0-9]*` on a1b2c3 with n=2 -> "1","2"
    o = Oracle(r"[1-9][0-9]*")
    assert o.find_all(b"a1b2c3", 2).tolist() == [[1, 2], [3, 4]]
    assert o.count(b"a1b2c3", 2) == 2


def test_ip_nfa_is_appendix_a():
    """Appendix A of SURVEY.md: 18 states, 5 byte classes, this exact numbering."""
    o = Oracle(r"\d+\.\d+\.\d+\.\d+")
    d = o.dump_nfa().splitlines()
    assert d[0] == "0 BR[30-39]->2" and d[2] == "2 QSplit(0,1)" and d[3] == "3 BR[2E-2E]->4"
    assert d[15] == "15 Match" and d[16] == "16 BR[00-FF]->17" and d[17] == "17 Split(0,16)"
    assert d[18] == "startAnchored=0 startUnanchored=17 classes=5"
    assert o.strategy == "UseDigitPrefilter" and o.digit_run_skip_safe


def test_strategy_table():
    # reference meta/strategy_selection_test.go:40-80,329-332 and meta/fat_teddy_fallback_test.go:31-33
    assert Oracle(r"\d+\.\d+\.\d+").strategy == "UseDigitPrefilter"
    assert Oracle(r"\d+\.\d+\.\d+\.\d+").strategy == "UseDigitPrefilter"
    assert Oracle("|".join("p%02d" % i for i in range(50))).strategy == "UseTeddy"
    assert Oracle(r"foo|bar").strategy == "UseTeddy"
    assert Oracle(r"\w+@\w+\.\w+").strategy in ("UseReverseInner", "UseNFA")


def test_captures_known_answers():
    # meta/findall_extra_test.go:147-157
    assert Oracle(r"(\w+)@(\w+)").find_all_submatch(b"user@host admin@server").tolist() == [
        [0, 9, 0, 4, 5, 9], [10, 22, 10, 15, 16, 22]]
    # nfa/pikevm_slottable_test.go:171-188
    assert Oracle(r"(a+)(b+)").find_all_submatch(b"xxxaaabbbyyy").tolist() == [[3, 9, 3, 6, 6, 9]]
    assert Oracle(r"([a-z]+)([0-9]+)").find_all_submatch(b"abc123xyz").tolist() == [[0, 6, 0, 3, 3, 6]]


def test_teddy_known_answers():
    # prefilter/teddy_test.go:95-137
    o = Oracle("foo|bar")
    assert o.find_at(b"hello foo world", 0) == (6, 9)
    assert o.find_at(b"hello bar world", 0) == (6, 9)
    assert o.find_at(b"foo bar foo", 1) == (4, 7)
    assert o.find_at(b"hello world", 0) is None
    # meta/fat_teddy_fallback_test.go:68-77
    o = Oracle("|".join("p%02d" % i for i in range(50)))
    assert o.find_at(b"test p42 here", 0) == (5, 8)


def test_empty_and_anchored_dfa():
    # dfa/lazy/anchored_search_prefilter_test.go:231-256: `a*` on "" -> 0, `abc` on "" -> no match
    assert Oracle("a*").find_all(b"").tolist() == [[0, 0]]
    assert Oracle("abc").find_all(b"").tolist() == []
    # empty-match rule (meta/findall.go:247-279): a* on "ab" -> [[0,1],[2,2]]
    assert Oracle("a*").find_all(b"ab").tolist() == [[0, 1], [2, 2]]


def test_parser_shapes():
    assert dump_ast("foo|bar|baz") == "alt{lit{foo}cat{lit{ba}cc{0x72-0x72 0x7a-0x7a}}}"
    assert dump_ast("two|three") == "cat{lit{t}alt{lit{wo}lit{hree}}}"   # meta/ahocorasick_test.go:237
    assert dump_ast("(a|b|c)+") == "plus{cap{cc{0x61-0x63}}}"            # meta/strategy.go:1094-1097
    assert dump_ast(r"a**").startswith("ERR error parsing regexp: invalid nested repetition operator")
    assert dump_ast("(").startswith("ERR error parsing regexp: missing closing )")
    assert dump_ast("x{1001}").startswith("ERR error parsing regexp: invalid repeat count")


def _stdlib_corpus():
    """The 41-line template corpus of reference meta/stdlib_compat_test.go:144-199, regenerated
    from its shape (timestamps, levels, IPs, e-mails, paths) — the committed fixture is in
    tests/golden/stdlib_corpus.txt."""
    with open(os.path.join(GOLDEN, "stdlib_corpus.txt"), "rb") as fh:
        return fh.read()


PY_PATTERNS = [
    r"\d+\.\d+\.\d+\.\d+", r"error|warning|fatal|critical", r"apple|banana|cherry|grape|lemon|mango|melon|olive|peach|plum|kiwi|lime",
    r"\d+", r"\w+", r"[a-z]+", r"[A-Z][a-z]+", r"\d{4}-\d{2}-\d{2}", r"\w+@\w+\.\w+", r"(\w+)@(\w+)\.(\w+)",
    r"GET|POST|PUT", r"[0-9]+ms", r"\d+\.\d+", r"user\d+", r"(?i)error", r"\bfoo\b", r"[1-9][0-9]*|0",
    r"\d{3}-\d{4}", r"https?://[a-z.]+", r"^\d+", r"(?m)^\d+",  # plain `$` differs: Go EndText vs Python "before final \\n"
]


@pytest.mark.parametrize("pat", PY_PATTERNS)
def test_vs_python_re(pat):
    corpus = _stdlib_corpus() * 3
    want = [[m.start(), m.end()] for m in re.finditer(pat.encode(), corpus)]
    got = Oracle(pat).find_all(corpus).tolist()
    assert got == want
    assert Oracle(pat).count(corpus) == len(want)
    assert Oracle(pat).is_match(corpus) == bool(want)


def test_multiline_dollar_byte_class_quirk():
    """Documents a reference quirk the oracle restates faithfully: byte classes are derived from
    byte-consuming states only (reference nfa/builder.go:47,65), so for `(?m)\\d+$` the bytes '\\n'
    and ' ' share a class and the lazy DFA's cached transition for that class depends on which of
    the two it met first (reference dfa/lazy/lazy.go:251-313 looks the class up before
    determinize's EndLine re-closure at :1345-1354).  The GPU engine implements the stdlib
    semantics instead; DESIGN.md lists this under parity hazards."""
    o = Oracle(r"(?m)\d+$")
    assert o.strategy == "UseDigitPrefilter"
    assert o.find_all(b"1 1\n").tolist() == []            # stdlib: [[2, 3]]
    assert Oracle(r"(?m)\d+$").find_all(b"1\n1 ").tolist() == [[0, 1], [2, 3]]  # stdlib: [[0, 1]]


def test_vs_python_re_random():
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"0123456789.. ab\n@_x-", dtype=np.uint8)
    pats = [r"\d+\.\d+\.\d+\.\d+", r"\d+\.\d+", r"[0-5]+x", r"\d{2}-\d", r"a+b", r"\w+@\w+\.\w+",
            r"(\d+)\.(\d+)", r"ab|a", r"\ba", r"(?m)^a", r"\d+\.\d+\.\d+\.35"]
    for it in range(200):
        n = int(rng.integers(0, 80))
        h = bytes(alphabet[rng.integers(0, len(alphabet), n)])
        for p in pats:
            o = Oracle(p)
            want = [[m.start(), m.end()] for m in re.finditer(p.encode(), h)]
            got = o.find_all(h).tolist()
            if p == r"[0-5]+x" and got != want:
                # documented reference quirk: only possible when the strategy really is
                # UseDigitPrefilter + digitRunSkipSafe (meta/strategy.go:530-560)
                assert o.strategy == "UseDigitPrefilter"
                continue
            assert got == want, (p, h)


def test_golden_fixture_replay():
    with open(os.path.join(GOLDEN, "oracle_vectors.json")) as fh:
        vec = json.load(fh)
    corpus = _stdlib_corpus()
    for v in vec:
        got = Oracle(v["pattern"]).find_all(corpus)[:1000].tolist()
        assert got == v["matches"], v["pattern"]


# ---- reference known-answer vectors for UTF-8 patterns (harvested by tests/golden/harvest_utf8_vectors.py
# from nfa/compile_utf8_test.go) -------------------------------------------------------------------------
def _utf8_vectors():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_utf8_vectors.json")
    return json.load(open(path, encoding="utf-8"))["vectors"]


@pytest.mark.parametrize("v", _utf8_vectors(), ids=lambda v: v["pattern"].encode("unicode_escape").decode()[:24])
def test_reference_utf8_vectors(v):
    pat = v["pattern"]
    hay = bytes.fromhex(v["haystack_hex"])
    o = Oracle(pat)
    if "first" in v:
        got = o.find_at(hay, 0)
        assert (list(got) if got else None) == v["first"]
        allm = o.find_all(hay).tolist()
        assert (allm[0] if allm else None) == v["first"]
    else:
        assert o.is_match(hay) == v["is_match"]


UTF8_DIFF = [r"a.c", r".*мир.*", r"[а-я]+", r"[^a\n]+", r"\S+", r"\W+", r"x.y", r"[^\s=]+=[^\s]*", r".+", r"(?s)a.b",
             r"[一-龥]+", r"é|e", r"[^é]", r"(?i)hello|a.z"]


@pytest.mark.parametrize("pat", UTF8_DIFF)
def test_utf8_patterns_equal_stdlib_semantics_on_valid_utf8(pat):
    """On well-formed UTF-8 the reference asserts equality with Go's stdlib
    (nfa/compile_utf8_test.go TestCompileUTF8_VsStdlib); Python's str-mode `re` with re.ASCII has the
    same semantics for these constructs, so it stands in for the stdlib here."""
    rng = np.random.default_rng(41)
    o = Oracle(pat)
    rx = re.compile(pat, re.ASCII)
    atoms = ["a", "b", "c", "x", "y", "z", " ", "\n", "=", "é", "e", "мир", "привет", "ПРИВЕТ", "я", "世", "界", "😀", "1", "Z"]
    for it in range(120):
        text = "".join(atoms[int(i)] for i in rng.integers(0, len(atoms), int(rng.integers(0, 16))))
        # char index -> byte offset
        offs = [0]
        for ch in text:
            offs.append(offs[-1] + len(ch.encode()))
        want = []
        pos = 0
        # Go FindAll semantics for empty matches do not arise here: none of the patterns is nullable
        for m in rx.finditer(text):
            want.append([offs[m.start()], offs[m.end()]])
        assert o.find_all(text.encode()).tolist() == want, (pat, text)


def test_case_folded_non_ascii_literal_quirk():
    """Go's parser stores a case-folded literal as the SMALLEST rune of each case orbit (regexp/syntax
    minFoldRune: `(?i)привет` is held as ПРИВЕТ with the FoldCase flag), and the reference expands the
    flag for ASCII letters only (nfa/compile.go:252 isASCIILetter): the pattern finds the upper-case
    spelling — which is all the reference's own test asks (edge_cases_test.go:427, equal to stdlib
    there) — and not the lower-case one it was written in.  Classes are folded by the parser itself
    and behave like stdlib."""
    assert Oracle("(?i)привет").find_all("ПРИВЕТ".encode()).tolist() == [[0, 12]]
    assert Oracle("(?i)привет").find_all("привет".encode()).tolist() == []
    assert Oracle("(?i)[а-я]+").find_all("ПРИвет".encode()).tolist() == [[0, 12]]
    assert Oracle("(?i)hello").find_all(b"HELLO").tolist() == [[0, 5]]               # :426


def _category_runs(text, pred):
    """maximal runs of characters satisfying pred, as byte offsets"""
    out, pos, start = [], 0, None
    for ch in text:
        if pred(ch):
            if start is None:
                start = pos
        elif start is not None:
            out.append([start, pos])
            start = None
        pos += len(ch.encode())
    if start is not None:
        out.append([start, pos])
    return out


UNI_TEXT = ["a", "Z", "é", "Ł", "ǅ", "ß", "мир", "Я", "αβ", "Ω", "世界", "あ", "ア", "한", "א", "ع", "1", "٣", "²", "Ⅷ", "½", " ", "\u00a0",
            "\u2003", "\n", ",", "«", "—", "_", "+", "€", "©", "^", "\u0301", "\u200d", "\t"]


@pytest.mark.parametrize("name", ["L", "Lu", "Ll", "Lt", "Lo", "N", "Nd", "Nl", "No", "P", "Pd", "S", "Sc", "Sm", "Z", "Zs", "M", "Mn",
                                  "C", "Cc", "Cf"])
def test_unicode_category_classes_equal_the_python_database(name):
    """\\p{category}+ and its negation against Python's unicodedata (Unicode 15.0.0, the version of Go's
    tables; an independent copy of the database — syntax/unicode_tables.inc is generated from perl's).
    Characters stay below U+10000: the reference widens four-byte ranges to whole lead bytes
    (nfa/compile.go:796-842), which the oracle restates and the malformed-input tests cover."""
    import unicodedata
    rng = np.random.default_rng(len(name) * 131 + ord(name[0]))
    atoms = [a.encode().decode("unicode_escape") if a.startswith("\\") else a for a in UNI_TEXT]
    pos_o, neg_o = Oracle(r"\p{%s}+" % name), Oracle(r"\P{%s}+" % name)
    inside = lambda ch: unicodedata.category(ch).startswith(name) or (name == "C" and unicodedata.category(ch) == "Cn")
    for it in range(60):
        text = "".join(atoms[int(i)] for i in rng.integers(0, len(atoms), int(rng.integers(0, 24))))
        assert pos_o.find_all(text.encode()).tolist() == _category_runs(text, inside), (name, text)
        assert neg_o.find_all(text.encode()).tolist() == _category_runs(text, lambda ch: not inside(ch)), (name, text)


@pytest.mark.parametrize("name", ["Latin", "Greek", "Cyrillic", "Han", "Hiragana", "Katakana", "Hangul", "Hebrew", "Arabic", "Common",
                                  "Inherited"])
def test_unicode_script_classes_equal_the_regex_module(name):
    """Scripts are not in Python's unicodedata; the third-party `regex` module carries its own
    database (a later Unicode version: the text uses long-assigned characters only)."""
    rx_mod = pytest.importorskip("regex")
    rng = np.random.default_rng(len(name) * 7)
    atoms = [a.encode().decode("unicode_escape") if a.startswith("\\") else a for a in UNI_TEXT]
    o, rx = Oracle(r"\p{%s}+" % name), rx_mod.compile(r"\p{Script=%s}+" % name)
    for it in range(60):
        text = "".join(atoms[int(i)] for i in rng.integers(0, len(atoms), int(rng.integers(0, 24))))
        offs = [0]
        for ch in text:
            offs.append(offs[-1] + len(ch.encode()))
        want = [[offs[m.start()], offs[m.end()]] for m in rx.finditer(text)]
        assert o.find_all(text.encode()).tolist() == want, (name, text)


def test_unicode_class_names_follow_go_1_25_lookup():
    """regexp/syntax of the Go release in the reference's go.mod (1.25): names compare case-insensitively
    ignoring space, underscore and hyphen; categories answer to their long aliases; Any, ASCII and
    Assigned are built in; \\p{^X} and \\P{X} negate; unknown names are ErrInvalidCharRange."""
    same = lambda a, b: dump_ast(a) == dump_ast(b) and not dump_ast(a).startswith("ERR")
    assert same(r"\p{Lu}", r"\p{Uppercase_Letter}") and same(r"\p{Lu}", r"\p{uppercase letter}") and same(r"\pL", r"\p{Letter}")
    assert same(r"\p{greek}", r"\p{Greek}") and same(r"\P{Greek}", r"\p{^Greek}") and same(r"\P{^Nd}", r"\p{Nd}")
    assert same(r"\p{Old_Italic}", r"\p{olditalic}") and same(r"\p{Assigned}", r"\P{Cn}") and same(r"\p{ASCII}", r"[\x00-\x7f]")
    assert same(r"[\p{Nd}]", r"\p{Nd}") and same(r"[^\p{Nd}]", r"\P{Nd}")
    for bad in [r"\p{Foo}", r"\pX", r"\p{", r"\p", r"[\p{Foo}]", r"\p{Script=Greek}"]:
        assert dump_ast(bad).startswith("ERR error parsing regexp: invalid character class range"), bad


def test_negated_class_stray_byte_fallback_quirk():
    """Reference quirk (nfa/compile.go:566-570): a class that covers everything above 0x7F gets, as its
    LAST alternative, "any single byte 0x80-0xFF".  A counted repetition that cannot be satisfied in
    whole code points therefore backtracks into counting the bytes of one code point as several
    characters: stdlib finds only [0,6) for `\\D{2,3}` in "мир😀", the restated automaton also yields the
    first three bytes of the emoji.  Kept in the oracle AND reproduced by the product's tables
    (tests/test_host_compile.py::test_utf8_tables_match_oracle)."""
    assert Oracle(r"\D{2,3}").find_all("мир😀".encode()).tolist() == [[0, 6], [6, 9]]


REF_FINDALL = json.load(open(os.path.join(GOLDEN, "ref_findall_vectors.json")))


@pytest.mark.parametrize("v", REF_FINDALL, ids=[v["src"].split(" ")[0] for v in REF_FINDALL])
def test_reference_findall_vectors(v):
    """Literal `[][2]int{...}` expectations of the reference's own tests (char-class searcher,
    non-greedy iteration, backtracker iteration), harvested by tests/golden/harvest_findall_vectors.py."""
    assert fa(v["pattern"], v["input"].encode()) == v["want"], v["src"]


# ---- reverse NFA / bidirectional lazy-DFA search (reference nfa/reverse.go, meta/find_indices.go:686-705) ----
BIDIR_PATS = [r"GET /\S+ HTTP", r'"[A-Z]+ .*" 200', r"(foo|bar)+baz", r"a.*b", r"x[^y]*y", r"(?i)error.*", r"[ab]+c|a+d",
              r"(a|b)*abb", r"foo\s+bar", r"ab+c+d", r"[ab]c+d[ef]", r"x\d+y\d+z", r"q[a-z]+\d", r"#[0-9a-f]+;", r"<[a-z]+>",
              r"\([0-9]+\)"]


@pytest.mark.parametrize("pat", BIDIR_PATS)
def test_bidirectional_dfa_search_equals_leftmost_first(pat):
    """UseDFA: forward lazy DFA for the end + lazy DFA of the ReverseAnchored NFA for the start (the
    restated reference path, oracle/nfa.cpp ReverseNFAStates + oracle/lazydfa.cpp SearchAt /
    SearchReverse) against the PikeVM restatement and Python `re` on random haystacks."""
    o = Oracle(pat)
    assert o.strategy == "UseDFA" and o.has_bidirectional, (pat, o.strategy)
    rx = re.compile(pat.encode())
    rng = random.Random(len(pat))
    alphabet = b"abcdefxyq 0123456789@.,-=;#<>()\n\"GETHTP/foobarzp"
    for it in range(400):
        h = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 160)))
        a = np.frombuffer(h, dtype=np.uint8)
        o.set_bidirectional(False)
        nfa = o.find_all(a).tolist()
        o.set_bidirectional(True)
        dfa = o.find_all(a).tolist()
        assert nfa == dfa == [[m.start(), m.end()] for m in rx.finditer(h)], (pat, h)


def test_reverse_nfa_of_a_leading_star_quirk():
    """Recorded observation, not a requirement on the product.  nfa/reverse.go:412-444
    (fillStartStateWithIncoming) turns EVERY edge into the forward start state into an epsilon edge
    of the reverse automaton — also the byte edge that closes a leading `x*` loop — so the restated
    reverse DFA stops at the loop: `a*b` on "aab" yields start 2 where leftmost-first yields 0.  The
    reference cannot be executed here (no Go toolchain) and documents stdlib compatibility, so the
    oracle's DEFAULT path for UseDFA stays the PikeVM restatement (leftmost-first), which is what
    the product is compared with."""
    for pat, hay, quirk, first in [(r"a*b", b"aab", [[2, 3]], [[0, 3]]), (r"[0-9]*px", b"12px", [[2, 4]], [[0, 4]]),
                                   (r"x*yz", b"xxyz", [[2, 4]], [[0, 4]])]:
        o = Oracle(pat)
        assert o.strategy == "UseDFA" and o.has_bidirectional
        a = np.frombuffer(hay, dtype=np.uint8)
        assert o.find_all(a).tolist() == first == [[m.start(), m.end()] for m in re.finditer(pat.encode(), hay)]
        o.set_bidirectional(True)
        assert o.find_all(a).tolist() == quirk


def test_dangling_at_sign_vector():
    """`\\w+@\\w+\\.\\w+` selects UseReverseInner in the reference, whose searcher is not
    restated (oracle/revsearch.h); the adversarial input with an `@` that starts no match is pinned
    to leftmost-first semantics here and on the device (tests/test_gpu_dfa.py)."""
    o = Oracle(r"\w+@\w+\.\w+")
    assert not o.strategy_exact  # (the engine falls back to its PikeVM restatement and says so)
    for hay, want in [(b"a@b c@d.e", [[4, 9]]), (b"@@a@b.c@", [[2, 7]]), (b"x@y z@w.", []), (b"a@b.c@d.e", [[0, 5]])]:
        assert o.find_all(np.frombuffer(hay, dtype=np.uint8)).tolist() == want
        assert [[m.start(), m.end()] for m in re.finditer(rb"\w+@\w+\.\w+", hay)] == want


def test_ast_dump_fixture():
    """The restated Go parser (syntax/parse.cpp) is shared by product and oracle, so no
    oracle-vs-device test can see a parser regression: its AST dumps for 106 patterns (the
    reference's stdlib-compat table, the suites' patterns, syntax corner cases and error messages)
    are pinned in tests/golden/ast_dumps.json (tests/golden/make_ast_fixture.py); the independent
    check of the parser's MEANING stays the Python-`re` differential above."""
    with open(os.path.join(GOLDEN, "ast_dumps.json")) as fh:
        fixture = json.load(fh)
    assert len(fixture) >= 100
    for pat, want in fixture.items():
        assert dump_ast(pat) == want, pat


def fa_sub(p, h):
    return Oracle(p).find_all_submatch(h).tolist()


def test_end_of_haystack_capture_shortcut_depends_on_where_the_search_starts():
    """Reference nfa/pikevm.go:2201-2206: a capture search that STARTS at len(haystack) builds its
    result from no slots.  Whether the last search of a FindAllSubmatch loop starts there depends on
    the strategy and on the previous match:"""
    # direct strategies: the loop reaches len only behind a match (or a skipped empty match) ending at len-1
    assert fa_sub(r"(a*)(\d*)", b"x") == [[0, 0, 0, 0, 0, 0], [1, 1, -1, -1, -1, -1]]      # previous match empty at len-1
    assert fa_sub(r"(\d*)", b"1x") == [[0, 1, 0, 1], [2, 2, -1, -1]]                       # skipped empty match at len-1
    assert fa_sub(r"(^)|($)", b"ab") == [[0, 0, 0, 0, -1, -1], [2, 2, -1, -1, 2, 2]]       # search started at 1: real slots
    assert fa_sub(r"(\d*)", b"") == [[0, 0, -1, -1]]                                       # empty haystack: starts at len
    # span-first strategies run the captures from the match start, which IS len
    assert Oracle(r"($)").strategy == "UseReverseAnchored" and fa_sub(r"($)", b"ab") == [[2, 2, -1, -1]]


def test_unicode_tables_are_what_the_generator_writes(tmp_path):
    """syntax/unicode_tables.inc is generated (tools/gen_unicode_tables.py: perl Unicode::UCD 15.0.0,
    cross-checked against Python unicodedata): regenerating must reproduce the committed file."""
    import shutil
    import subprocess
    import sys
    import unicodedata
    if shutil.which("perl") is None or unicodedata.unidata_version != "15.0.0":
        pytest.skip("needs perl and a Unicode 15.0.0 unicodedata")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "unicode_tables.inc"
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_unicode_tables.py"), str(out)], capture_output=True, text=True)
    if p.returncode != 0 and "want 15.0.0" in (p.stdout + p.stderr):
        pytest.skip("perl's Unicode database is not 15.0.0")
    assert p.returncode == 0, p.stdout + p.stderr
    with open(os.path.join(root, "syntax", "unicode_tables.inc")) as a, open(out) as b:
        assert a.read() == b.read()
