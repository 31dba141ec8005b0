"""CompileWithConfig / Config.Validate / Longest — the boundary rows the reference tests in
meta/config_test.go and regex_test.go:982-1040 (vectors below are the reference's own)."""
import pytest

import coregex_b200 as cg


def test_default_config_values():
    # reference meta/config_test.go:9-37 (TestDefaultConfigValues)
    c = cg.DefaultConfig()
    assert (c.EnableDFA, c.EnablePrefilter, c.MaxDFAStates, c.DeterminizationLimit, c.MinLiteralLen, c.MaxLiterals,
            c.MaxRecursionDepth, c.EnableASCIIOptimization) == (1, 1, 10000, 1000, 1, 256, 100, 1)
    assert c.Validate() is None  # :40-45


# (field, value, valid) — reference meta/config_test.go:54-59, 92-96, 120-125, 149-153, 177-181
BOUNDS = [("MaxDFAStates", 0, False), ("MaxDFAStates", 1, True), ("MaxDFAStates", 10000, True),
          ("MaxDFAStates", 1_000_000, True), ("MaxDFAStates", 1_000_001, False), ("MaxDFAStates", 10_000_000, False),
          ("DeterminizationLimit", 5, False), ("DeterminizationLimit", 10, True), ("DeterminizationLimit", 1000, True),
          ("DeterminizationLimit", 100_000, True), ("DeterminizationLimit", 100_001, False),
          ("MinLiteralLen", 0, False), ("MinLiteralLen", 1, True), ("MinLiteralLen", 2, True), ("MinLiteralLen", 64, True),
          ("MinLiteralLen", 65, False), ("MinLiteralLen", -1, False),
          ("MaxLiterals", 0, False), ("MaxLiterals", 1, True), ("MaxLiterals", 256, True), ("MaxLiterals", 1000, True),
          ("MaxLiterals", 1001, False),
          ("MaxRecursionDepth", 5, False), ("MaxRecursionDepth", 10, True), ("MaxRecursionDepth", 100, True),
          ("MaxRecursionDepth", 1000, True), ("MaxRecursionDepth", 1001, False)]
MESSAGES = {"MaxDFAStates": "must be between 1 and 1,000,000", "DeterminizationLimit": "must be between 10 and 100,000",
            "MinLiteralLen": "must be between 1 and 64", "MaxLiterals": "must be between 1 and 1,000",
            "MaxRecursionDepth": "must be between 10 and 1,000"}


@pytest.mark.parametrize("field,value,valid", BOUNDS)
def test_validate_bounds(field, value, valid):
    c = cg.DefaultConfig()
    setattr(c, field, value)
    err = c.Validate()
    if valid:
        assert err is None
        assert cg.CompileWithConfig("hello", c).String() == "hello"
    else:
        # ConfigError.Error(), reference meta/config.go:179-181
        assert err == "regexp: invalid config: %s: %s" % (field, MESSAGES[field])
        with pytest.raises(cg.ConfigError) as ei:
            cg.CompileWithConfig("hello", c)
        assert str(ei.value) == err


def test_disabled_sections_are_not_validated():
    # reference meta/config.go:133,148: DFA / prefilter limits are only checked when that part is on
    # (meta/engine_test.go TestEngineCompileWithConfig "DFA disabled": zero limits are accepted)
    c = cg.Config(EnableDFA=0, EnablePrefilter=0, MaxRecursionDepth=100)
    assert c.Validate() is None
    assert cg.CompileWithConfig("hello", c).strategy == "UseNFA"


def test_config_steers_strategy_selection():
    ip = r"\d+\.\d+\.\d+\.\d+"
    assert cg.Compile(ip).strategy == "UseDigitPrefilter"
    c = cg.DefaultConfig()
    c.EnableDFA = 0                       # meta/strategy.go:1447
    assert cg.CompileWithConfig(ip, c).strategy == "UseNFA"
    c = cg.DefaultConfig()
    c.EnablePrefilter = 0                 # meta/strategy.go:515 (digit prefilter), meta/compile.go:466 (no literals)
    r = cg.CompileWithConfig(ip, c)
    assert r.strategy == "UseDFA"
    lits = "error|warning|fatal"
    assert cg.Compile(lits).strategy == "UseTeddy"
    assert cg.CompileWithConfig(lits, c).strategy != "UseTeddy"


# reference regex_test.go:985-1022 (TestLongest): pattern, input, leftmost-first, leftmost-longest
LONGEST = [(r"(a|ab)", b"ab", b"a", b"ab"), (r"(#|#!)", b"#!a", b"#", b"#!"),
           (r"(cat|catalog)", b"catalog", b"cat", b"catalog"), (r"a+", b"aaaa", b"aaaa", b"aaaa")]


@pytest.mark.gpu
@pytest.mark.parametrize("pat,inp,first,longest", LONGEST)
def test_longest(pat, inp, first, longest):
    r = cg.Compile(pat)
    m = r.FindAllIndex(inp)
    assert inp[m[0][0]:m[0][1]] == first
    r.Longest()
    m = r.FindAllIndex(inp)
    assert inp[m[0][0]:m[0][1]] == longest
    # a second regex compiled from the same pattern keeps leftmost-first (regex_stdlib_compat_test.go:180)
    m2 = cg.Compile(pat).FindAllIndex(inp)
    assert inp[m2[0][0]:m2[0][1]] == first


@pytest.mark.gpu
def test_longest_find_all_and_engine_reference_vector():
    # reference meta/engine_test.go:283-306 (TestEngineSetLongest)
    r = cg.Compile("a+")
    r.Longest()
    assert r.FindAllIndex(b"aaa") == [[0, 3]]
    r = cg.Compile(r"se|set|x\w*|settle")
    assert r.FindAllIndex(b"settle sets xy") == [[0, 2], [7, 9], [12, 14]]
    r.Longest()
    assert r.FindAllIndex(b"settle sets xy") == [[0, 6], [7, 10], [12, 14]]
