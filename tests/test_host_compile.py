"""CPU-tier checks of the product's host compiler: the C-ABI library loads and exports every
declared symbol, compile errors carry the Go-formatted text, the reference-strategy
classification agrees with the oracle's restatement, and the compiled DFA tables + filter choice
reproduce the oracle's matches when replayed by tests/table_model.py (no GPU involved)."""
import ctypes
import os
import re

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle, OracleError
from table_model import FlatModel, TableModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "coregex_b200.h")).read()
    names = sorted(set(re.findall(r"\b(cgx_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 14
    lib = ctypes.CDLL(os.path.join(ROOT, "coregex_b200", "lib", "libcoregex_b200.so"))
    for n in names:
        assert hasattr(lib, n), n


def test_compile_errors_are_go_formatted():
    cases = {
        "(": "error parsing regexp: missing closing ): `(`",
        "a)": "error parsing regexp: unexpected ): `a)`",
        "a**": "error parsing regexp: invalid nested repetition operator: `**`",
        "[a": "error parsing regexp: missing closing ]: `[a`",
        "x{1001}": "error parsing regexp: invalid repeat count: `{1001}`",
        "\\8": "error parsing regexp: invalid escape sequence: `\\8`",
    }
    for pat, msg in cases.items():
        with pytest.raises(cg.Error) as ei:
            cg.Compile(pat)
        assert str(ei.value) == msg
    with pytest.raises(cg.Error) as ei:
        cg.MustCompile("(")
    assert str(ei.value).startswith("regexp: Compile(`(`): error parsing regexp")


def test_unicode_property_classes_compile_to_the_table_engines():
    # \p{..} resolves through the generated Unicode 15.0.0 tables (syntax/unicode_tables.inc); the
    # UTF-8 automaton of a large class shares prefixes and suffixes (host/prog.cpp emitShared), which
    # keeps even \pL (650 ranges) inside the 160-state table kernels
    for pat in [r"\pL", r"\pL+", r"\pN+", r"\p{Greek}+", r"\p{Han}+", r"[\p{Lu}\d]+", r"\p{Lu}\p{Ll}+"]:
        assert cg.Compile(pat).engine.startswith("dfa"), pat
    for pat, msg in [(r"\p{Foo}", "invalid character class range: `\\p{Foo}`"), (r"\pX", "invalid character class range: `\\pX`"),
                     (r"\p{Greek", "invalid character class range: `\\p{Greek`")]:
        with pytest.raises(cg.Error) as ei:
            cg.Compile(pat)
        assert msg in str(ei.value)


def test_record_engine_needs_an_ascii_delimiter():
    # found by the sweep over the reference's test patterns: `\P{Han}+` can contain every ASCII byte, its
    # record delimiter is 0xC0, and the record engine's phase A marks delimiters with a SWAR test that
    # is exact below 0x80 only — such patterns stay on the candidate + anchored-walk engine
    for pat in [r"\P{Han}+", r"\P{Greek}+"]:
        r = cg.Compile(pat)
        assert r.delimiter[0] >= 0x80 and r.engine != "line-dfa", (pat, r.engine, r.delimiter)
    assert cg.Compile(r"[^\x00-\x7f]+").engine == "line-dfa"


def test_unsupported_patterns_fail_loudly():
    # matches that may contain every byte value leave no record delimiter: one record, one lane
    for pat in [r"(?s)x.y", r"(?s).+", r"(?m)^POST\s+\S+"]:
        r = cg.Compile(pat)
        assert r.engine == "pikevm-serial" and r.delimiter is None


def test_nullable_and_large_automata_go_to_the_pikevm_engine():
    # patterns that can match the empty string need the sequential empty-match rules of the
    # reference's FindAll loop (meta/findall.go:247-279); automata over 160 DFA states do not fit
    # the table kernels: both run on the PikeVM search kernel instead of being refused
    for pat in ["a*", r"\d*", "(?m)^", r"\b", "foo|bar|", r"[ab]*a[ab]{8}"]:
        assert cg.Compile(pat).engine == "pikevm"


def test_literal_sets_stay_on_the_multi_literal_engines():
    # 16 literals: Slim Teddy; 64 literals: Fat Teddy (reference prefilter/teddy_fat.go:348) — the
    # DFA of 64 literals does not fit 160 states, which must not push the set to the PikeVM engine
    lit16 = ["error", "warning", "fatal", "critical", "timeout", "refused", "denied", "panic", "overflow", "invalid",
             "missing", "corrupt", "expired", "blocked", "aborted", "unknown"]
    lit64 = ["k%02dz%s" % (i, "q" * (i % 4)) for i in range(64)]
    assert cg.Compile("|".join(lit16)).engine == "teddy"
    assert cg.Compile("|".join(lit64)).engine == "fat-teddy"
    # 68..255 complete, prefix-free literals (the reference's Aho-Corasick strategy): same engine, 16 buckets
    big = ["w%03dx%s" % (i, "abcdefghij"[i % 10] * (i % 3)) for i in range(0, 990, 9)]
    r = cg.Compile("|".join(big))
    assert (r.strategy, r.engine) in (("UseAhoCorasick", "teddy-large"), ("UseNFA", r.engine)), (r.strategy, r.engine)


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = cg.Compile(r"\d+")
    with pytest.raises(cg.NoDeviceError):
        r.Match(b"123")
    with pytest.raises(cg.NoDeviceError):
        r.FindAllIndex(b"123")


STRAT_PATTERNS = [
    r"\d+\.\d+\.\d+\.\d+", r"[1-9][0-9]*|0", "cat|dog", r"\w+@\w+\.\w+", r"(\w+)@(\w+)\.(\w+)", r"\d+",
    "ab", "error|warning|fatal|critical", r"\d{3}-\d{4}", r"[0-5]+x", r"foo\d+", r"\bfoo\b", "^abc",
    "abc$", "(?m)^abc$", r"\d+\.\d+\.\d+\.35", "|".join("p%02d" % i for i in range(50)),
    "|".join("w%03dx" % i for i in range(64)), "|".join("w%03dx" % i for i in range(70)),
    r"(?i)error", r"(?m)^(GET|POST|PUT)", r"[a-zA-Z]+\d+", r"\w+[0-9]+", r"user\d+", r"https?://[a-z.]+",
    r"[a-f0-9]{32,}", r"(?i)(error|fail|panic)", r"\d+\.\d+", "foo|foobar|barfoo", r"a+b",
]


@pytest.mark.parametrize("pat", STRAT_PATTERNS)
def test_reference_strategy_agrees_with_oracle(pat):
    o = Oracle(pat)
    try:
        r = cg.Compile(pat)
    except cg.UnsupportedError:
        pytest.skip("outside GPU scope")
    want = o.strategy
    if want == "UseNFA" and not o.strategy_exact:
        pytest.skip("oracle fell back")
    assert r.strategy == want


def test_ip_engine_choice():
    r = cg.Compile(r"\d+\.\d+\.\d+\.\d+")
    assert (r.strategy, r.engine) == ("UseDigitPrefilter", "dfa-runstart+bitstream")
    m = TableModel(r)
    assert m.nstates <= 12 and m.filter_kind == 0 and m.skip_safe and m.ranges == [(0x30, 0x39)]
    assert m.find_all(b"x10.0.0.1 y 1.2.3 z 8.8.8.8") == [[1, 9], [20, 27]]
    assert m.find_all(b"1.2.3.4.5.6.7") == [[0, 7]]


DFA_PATTERNS = [
    r"\d+\.\d+\.\d+\.\d+", r"[1-9][0-9]*|0", r"\d{3}-\d{4}", r"\w+@\w+\.\w+", r"(\w+)@(\w+)\.(\w+)", r"\d+",
    "ab", r"foo\d+", r"\bfoo\b", r"(?m)^\d+", r"(?m)\d+$", r"\d+\.\d+", r"user\d+", r"[a-zA-Z]+\d+",
    r"(?i)error", r"[0-9]+ms", r"\d{4}-\d{2}-\d{2}", r"a+b", r"ab|a", r"\ba", r"\d+\.\d+\.\d+\.35", r"[a-f0-9]{8,}",
    r"(?m)^(GET|POST|PUT)", r"\Babc\B", r"\d+\b",
]


@pytest.mark.parametrize("pat", DFA_PATTERNS)
def test_tables_match_python_re_on_corpus(pat):
    """Eager DFA + filter tables == stdlib leftmost-first semantics (Python re on bytes)."""
    corpus = open(os.path.join(ROOT, "tests", "golden", "stdlib_corpus.txt"), "rb").read()[:6000]
    r = cg.Compile(pat)
    if "teddy" in r.engine:
        pytest.skip("literal engine (covered by the teddy tests)")
    m = TableModel(r)
    want = [[x.start(), x.end()] for x in re.finditer(pat.encode(), corpus)]
    assert m.find_all(corpus) == want


def test_tables_match_oracle_random():
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"0123456789.. ab\n@_x-", dtype=np.uint8)
    pats = [r"\d+\.\d+\.\d+\.\d+", r"\d+\.\d+", r"\d{2}-\d", r"a+b", r"\w+@\w+\.\w+", r"ab|a",
            r"[1-9][0-9]*|0", r"\d+\.\d+\.\d+\.35", r"\d+x"]
    models = {p: TableModel(cg.Compile(p)) for p in pats}
    oracles = {p: Oracle(p) for p in pats}
    for it in range(300):
        n = int(rng.integers(0, 90))
        h = bytes(alphabet[rng.integers(0, len(alphabet), n)])
        for p in pats:
            assert models[p].find_all(h) == oracles[p].find_all(h).tolist(), (p, h)


def test_skip_safe_quirk_is_reproduced_only_when_reference_has_it():
    # `[0-5]+x`: the reference picks UseReverseSuffix (suffix literal "x"), so stdlib semantics apply
    r = cg.Compile(r"[0-5]+x")
    assert r.strategy == "UseReverseSuffix"
    assert TableModel(r).find_all(b"65x") == [[1, 3]]
    # `[0-5]+\.\d+`: UseDigitPrefilter + digitRunSkipSafe -> the run skip hides the match at 1,
    # exactly as the oracle (reference meta/find_indices.go:1079-1084) does
    r = cg.Compile(r"[0-5]+\.\d+")
    o = Oracle(r"[0-5]+\.\d+")
    assert r.strategy == o.strategy == "UseDigitPrefilter" and o.digit_run_skip_safe
    assert TableModel(r).find_all(b"65.1 ") == o.find_all(b"65.1 ").tolist() == []


def test_synth_host_is_deterministic_and_line_aligned():
    a = cg.synth_host(cg.SYNTH_LOG, 123, 4096 * 4)
    b = cg.synth_host(cg.SYNTH_LOG, 123, 4096 * 2, first_block=2)
    assert bytes(a[8192:]) == bytes(b)
    for k in range(1, 5):
        assert a[4096 * k - 1] == 10
    lits = [b"error", b"warning", b"fatal", b"critical"]
    t = cg.synth_host(cg.SYNTH_TEXT, 5, 4096 * 8, literals=lits)
    assert t[-1] == 10 and sum(bytes(t).count(x) for x in lits) > 20
    e = cg.synth_host(cg.SYNTH_EMAIL, 5, 80 * 100)
    lines = bytes(e).split(b"\n")[:-1]
    assert len(lines) == 100 and all(len(x) == 79 and x.count(b"@") == 1 for x in lines)
    assert all(re.search(rb"[a-z]+@[a-z]+\.[a-z]+", x) for x in lines)


FLAT_PATTERNS = [r"\d+\.\d+\.\d+\.\d+", r"\d+\.\d+", r"\w+@\w+\.\w+", r"\d{3}-\d{4}", r"[a-zA-Z]+\d+", r"a+b",
                 r"\d+x", r"[0-5]+\.\d+", r"(\w+)@(\w+)\.(\w+)", r"\d+\.\d+\.\d+\.35", r"ab?c*d",
                 r"(?i)ab\d{2,4}z"]


@pytest.mark.parametrize("pat", FLAT_PATTERNS)
def test_flat_start_filter_is_exact_superset(pat):
    """The bit-parallel filter must contain every position the anchored DFA matches from (it may
    contain more).  On flat patterns it is in fact exact."""
    r = cg.Compile(pat)
    if "teddy" in r.engine:
        pytest.skip("literal engine")
    assert r.engine.endswith(("+flat", "+bitstream"))
    fm, tm = FlatModel(r), TableModel(r)
    rng = np.random.default_rng(17)
    alphabet = np.frombuffer(b"0123456789..  abABxz\n@_-d35", dtype=np.uint8)
    for it in range(150):
        n = int(rng.integers(0, 70))
        h = bytes(alphabet[rng.integers(0, len(alphabet), n)])
        S = fm.start_set(h)
        for p in range(n):
            ok = tm.walk(h, p) >= 0
            assert ok == bool((S >> p) & 1), (pat, h, p)


UTF8_PATTERNS = [r"a.c", r"foo.*bar", r"[^a\n]+", r"\S+", r"[α-ω]+", r"x.y.z", r"[^\s\"]+=.", r"[\x{100}-\x{17F}]+x",
                 r"[^\x00-\x{7FF}\n]+", r"é+", r"(?i)straße|x.z", r"[\x{10000}-\x{10200}]", r"[\x{90000}-\x{10FFFF}]x", r"a[^b\n]c",
                 r"[\x{7F0}-\x{810}]+", r"[\x{D700}-\x{E010}]", r"[^\d\n]{2,3}", r"[^x\n]{3}y",
                 # Unicode property classes: large tables take the shared-prefix/suffix emission (host/prog.cpp emitShared)
                 r"\pL+", r"\pL", r"\p{Lu}\p{Ll}+", r"\p{Greek}+", r"[\p{Lu}\d]+x", r"\pN+", r"\p{Han}", r"\pS", r"(?i)\p{Lu}+",
                 r"\p{Latin}+", r"(?i)[а-в]+", r"(?i)я",
                 # only high bytes are safe record delimiters: not the record engine (its delimiter test is 7-bit SWAR)
                 r"\P{Han}+", r"\P{Greek}+x", r"[\x00-\x7f]+x"]


@pytest.mark.parametrize("pat", UTF8_PATTERNS)
def test_utf8_tables_match_oracle(pat):
    """`.`, negated and non-ASCII classes: the product's own UTF-8 byte automaton (host/prog.cpp) must
    define the same leftmost-first matches as the restated reference automaton (oracle/nfa.cpp,
    nfa/compile.go:440-1223), on well-formed AND malformed UTF-8."""
    from table_model import LineModel
    rng = np.random.default_rng(23)
    r = cg.Compile(pat)
    m = LineModel(r) if r.engine == "line-dfa" else TableModel(r)
    o = Oracle(pat)
    pieces = [b"a", b"b", b"c", b"x", b"y", b"z", b"foo", b"bar", b" ", b"\n", b"=", b"\"", "é".encode(), "α".encode(),
              "ω".encode(), "β".encode(), "Ł".encode(), "ſ".encode(), "ß".encode(), "€".encode(), "ߵ".encode(),
              "ࠅ".encode(), "퟿".encode(), "".encode(), "𐀀".encode(), "𐀂".encode(), "😀".encode(),
              "\U0010ffff".encode(), b"\x80", b"\xbf", b"\xc0", b"\xc3", b"\xe0", b"\xe0\x80", b"\xed\xa0\x80",
              b"\xf0\x90", b"\xf4\x90\x80\x80", b"\xf5", b"\xff", b"stra", b"STRASSE", b"Stra\xc3\x9fe"]
    for it in range(250):
        k = int(rng.integers(0, 14))
        h = b"".join(pieces[int(i)] for i in rng.integers(0, len(pieces), k))
        assert m.find_all(h) == o.find_all(h).tolist(), (pat, h)


LINE_PATTERNS = [r".*error", r"\S+@\S+", r".+", r"[^,\n]+,[^,\n]+", r".*\d{3}.*", r"(?m)^.*error.*$", r".*?b", r"\S+\.\S+",
                 r"[^\n]*a[^\n]*b", r".*\bfoo\b", r"(?i).*warn(ing)?", r".+?,", r"[^\s]+=[^\s]*;?"]


@pytest.mark.parametrize("pat", LINE_PATTERNS)
def test_record_engine_tables_match_oracle(pat):
    """Patterns without a useful first-byte filter: unanchored forward DFA + reverse DFA tables
    (host/dfa.cpp, the reference's UseDFA shape) must give the oracle's leftmost-first spans."""
    from table_model import LineModel
    rng = np.random.default_rng(31)
    r = cg.Compile(pat)
    assert r.engine == "line-dfa", r.engine
    m, o = LineModel(r), Oracle(pat)
    pieces = [b"a", b"b", b"x", b" ", b" ", b"\n", b",", b"@", b".", b"=", b";", b"error", b"foo", b"food", b"warn",
              b"WARNING", b"12", b"123", b"4567", "é".encode(), "мир".encode(), b"\xff", b"\xe0", b"user@host.tld"]
    for it in range(300):
        h = b"".join(pieces[int(i)] for i in rng.integers(0, len(pieces), int(rng.integers(0, 18))))
        assert m.find_all(h) == o.find_all(h).tolist(), (pat, h)


def test_record_delimiter_choice():
    """'\\n' whenever no match can contain it; otherwise a byte the pattern cannot consume."""
    assert cg.Compile(r"\d+\.\d+").delimiter == b"\n"
    assert cg.Compile(r"foo.*bar").delimiter == b"\n"
    assert cg.Compile(r"\s+").delimiter == b"e"
    assert cg.Compile(r"[^a]+").delimiter == b"a"
    assert cg.Compile(r"[a-z]+\s+[a-z]+").delimiter == b","


NON_LF_PATTERNS = [r"\s+", r"[^a]+", r"[a-z]+\s+[a-z]+", r"\W+", r"[^e]{3}", r"\D+"]


@pytest.mark.parametrize("pat", NON_LF_PATTERNS)
def test_non_newline_delimiter_tables_match_oracle(pat):
    from table_model import LineModel
    rng = np.random.default_rng(43)
    r = cg.Compile(pat)
    assert r.delimiter != b"\n"
    m = LineModel(r) if r.engine == "line-dfa" else TableModel(r)
    o = Oracle(pat)
    pieces = [b"a", b"e", b"t", b"x", b" ", b"  ", b"\n", b"\t", b",", b"12", b"3", b"word", b"ea", "é".encode(), b"\xff"]
    for it in range(300):
        h = b"".join(pieces[int(i)] for i in rng.integers(0, len(pieces), int(rng.integers(0, 20))))
        assert m.find_all(h) == o.find_all(h).tolist(), (pat, h)


def test_bitstream_kernel_specialises_with_nvrtc():
    """jit.cu: the bitstream kernel source, embedded in the library, compiles with NVRTC for
    sm_100a with the generated per-pattern header (no device needed for this step)."""
    import ctypes as C
    L = cg._lib
    L.cgx_debug_jit_compile.restype = C.c_long
    L.cgx_debug_jit_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    for pat in (r"\d+\.\d+\.\d+\.\d+", r"[a-z]+=\d+[x-z]?"):
        r = cg.Compile(pat)
        assert r.engine.endswith("+bitstream")
        n = L.cgx_debug_jit_compile(r._h, None, 0)
        assert n > 10000, L.cgx_last_error()
    r = cg.Compile(r"GET|POST")
    assert L.cgx_debug_jit_compile(r._h, None, 0) == -1
