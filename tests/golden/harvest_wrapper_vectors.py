#!/usr/bin/env python
"""Harvests the known-answer tables of the reference's Replace/Split/Expand tests into
tests/golden/ref_wrapper_vectors.json:

  replace_test.go          TestReplaceAllLiteral, TestReplaceAll, TestSplit, TestExpandEdgeCases
  stdlib_compat_test.go    replaceTests (minus the rows its own hasReplaceDifference skips and the
                           rows with `$`), replaceLiteralTests, splitTests

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/harvest_wrapper_vectors.py
Only table rows (pattern, input, replacement / count, expected) are extracted — no reference code."""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_wrapper_vectors.json")
TOK = r'(`[^`]*`|"(?:[^"\\]|\\.)*")'


def go_string(tok):
    if tok.startswith("`"):
        return tok[1:-1]
    return json.loads(re.sub(r"\\x([0-9a-fA-F]{2})", r"\\u00\1", tok))


def block(src, start_marker):
    """source text from start_marker to the closing brace of its top-level block (gofmt puts it at
    the start of a line; braces inside string literals make counting unreliable)"""
    i = src.index(start_marker)
    return src[i:src.index("\n}\n", i)]


def rows(text):
    for line in text.splitlines():
        line = line.strip()
        if line.startswith("{") and not line.startswith("//"):
            yield line


def string_list(tok):
    tok = tok.strip()
    if tok == "nil":
        return None
    return [go_string(t) for t in re.findall(TOK, tok)]


def main():
    out = {"replace_literal": [], "replace": [], "split": [], "expand": []}
    rt = open(os.path.join(REF, "replace_test.go")).read()
    for line in rows(block(rt, "func TestReplaceAllLiteral(")):
        f = re.findall(TOK, line)
        if len(f) == 4:
            p, i, r, w = map(go_string, f)
            out["replace_literal"].append({"pattern": p, "input": i, "repl": r, "want": w, "src": "replace_test.go TestReplaceAllLiteral"})
    for line in rows(block(rt, "func TestReplaceAll(")):
        f = re.findall(TOK, line)
        if len(f) == 4:
            p, i, r, w = map(go_string, f)
            out["replace"].append({"pattern": p, "input": i, "repl": r, "want": w, "src": "replace_test.go TestReplaceAll"})
    for line in rows(block(rt, "func TestSplit(")):
        m = re.match(r"\{%s,\s*%s,\s*(-?\d+),\s*(nil|\[\]string\{.*\})\}," % (TOK, TOK), line)
        if m:
            out["split"].append({"pattern": go_string(m.group(1)), "input": go_string(m.group(2)), "n": int(m.group(3)),
                                 "want": string_list(m.group(4)), "src": "replace_test.go TestSplit"})
    eb = block(rt, "func TestExpandEdgeCases(")
    for line in rows(eb[eb.index("tests :="):]):
        f = re.findall(TOK, line)
        if len(f) == 2:
            out["expand"].append({"pattern": r"(\d+)", "input": "test 123 end", "template": go_string(f[0]), "want": go_string(f[1]),
                                  "src": "replace_test.go TestExpandEdgeCases"})
    sc = open(os.path.join(REF, "stdlib_compat_test.go")).read()
    diffs = set(go_string(t) for t in re.findall(TOK + r":\s*true", block(sc, "var replacePatternsWithDiffs")))
    for line in rows(block(sc, "var replaceTests")):
        f = re.findall(TOK, line.split("//")[0])
        if len(f) == 4:
            p, r, i, w = map(go_string, f)
            if "$" in r or (i == "" and p in diffs):   # the rows the reference's own test skips
                continue
            out["replace_literal"].append({"pattern": p, "input": i, "repl": r, "want": w, "src": "stdlib_compat_test.go replaceTests"})
    for line in rows(block(sc, "var replaceLiteralTests")):
        f = re.findall(TOK, line)
        if len(f) == 4:
            p, r, i, w = map(go_string, f)
            out["replace_literal"].append({"pattern": p, "input": i, "repl": r, "want": w, "src": "stdlib_compat_test.go replaceLiteralTests"})
    for line in rows(block(sc, "var splitTests")):
        m = re.match(r"\{%s,\s*%s,\s*(-?\d+),\s*(nil|\[\]string\{.*\})\}," % (TOK, TOK), line)
        if m:
            out["split"].append({"input": go_string(m.group(1)), "pattern": go_string(m.group(2)), "n": int(m.group(3)),
                                 "want": string_list(m.group(4)), "src": "stdlib_compat_test.go splitTests"})
    json.dump(out, open(OUT, "w"), indent=1, ensure_ascii=False)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
