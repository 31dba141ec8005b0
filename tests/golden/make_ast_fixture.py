#!/usr/bin/env python
"""Writes tests/golden/ast_dumps.json: the AST dump of every pattern below as the restated Go parser
(syntax/parse.cpp, shared by product and oracle) produces it TODAY.  The fixture pins the parser:
product and oracle share it, so an oracle-vs-device test cannot see a parser regression — this
replay (tests/test_oracle_golden.py::test_ast_dump_fixture) and the Python-`re` differentials can.
Patterns: the reference's stdlib-compat table (ref_stdlib_patterns.json, harvested from
meta/stdlib_compat_test.go), the patterns of the GPU/CPU suites, and syntax corner cases."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import dump_ast  # noqa: E402

EXTRA = [
    r"\d+\.\d+\.\d+\.\d+", r"\w+@\w+\.\w+", r"(\w+)@(\w+)\.(\w+)", r"[a-z]+/\d+", r"\d+", r"\w+", r"a*", r"\d*", r"(?m)^", r"\b",
    r"foo|bar|", r"[ab]*a[ab]{8}", r"(?s)x.y", r"(?s).+", r"(?m)^POST\s+\S+", r".* 404 .*", r"\S+@\S+", r"[^,\n]+,[^,\n]+",
    r"GET /\S+ HTTP", r'"[A-Z]+ .*" 200', r"(foo|bar)+baz", r"a.*b", r"x[^y]*y", r"(?i)error.*", r"\bfoo\b", r"[ab]+c|a+d",
    r"(a|b)*abb", r"(?i)straße|x.z", r"[α-ω]+", r"[^\x00-\x{7FF}\n]+", r"é+", r"[^\d\n]{2,3}", r"a{2,}", r"a{,3}", r"a{3}?",
    r"(?P<user>\w+)@(?P<host>[a-z]+)", r"(a|ab)(c|bcd)(d*)x", r"((a)(b))+c", r"(\d+)-(\d+)?x", r"([a-z]+?)(\d+)", r"(?U)a+b",
    r"\Aabc\z", r"^$", r"(?:a|b)c", r"[[:alpha:]]+", r"[^[:space:]]", r"\p{Zs}", r"\P{Zs}", r"\p{^Zs}+", r"[\p{Zl}\p{Zp}x]", r"\pZ", r"\p{Lt}", r"(?i)\p{Lt}", r"\p{Braille}",
    r"\p{space separator}", r"[^\p{Nl}]", r"\p{Any}", r"\p{ASCII}", r"\p{Foo}", r"\p{Zs", r"(?i)я", r"(?i)[а-в]", r"(?i)ǆ", r"\x41\x{1F600}", r"\Q.*\E", r"a|b|c|d",
    r"abc|abd", r"two|three", r"(a|b|c)+", r"x{1,3}y{0}z", r"[a-c-e]", r"[]a]", r"[^]a]", r"\C", r"(?i)k", r"(?i)[k-l]s",
    r"a**", r"(", r")", r"x{1001}", r"[z-a]", r"\8", r"(?P<n>a)(?P<n>b)", r"a{2,1}", r"*a", r"(?z)", r"\pX", "[a",
]


def main():
    pats = list(EXTRA)
    ref = json.load(open(os.path.join(HERE, "ref_stdlib_patterns.json")))
    for e in (ref if isinstance(ref, list) else ref.get("patterns", [])):
        p = e["pattern"] if isinstance(e, dict) else e
        if p not in pats:
            pats.append(p)
    out = {p: dump_ast(p) for p in pats}
    json.dump(out, open(os.path.join(HERE, "ast_dumps.json"), "w"), indent=0, sort_keys=True, ensure_ascii=True)
    print(len(out), "patterns")


if __name__ == "__main__":
    main()
