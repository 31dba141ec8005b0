#!/usr/bin/env python
"""Harvests the FindAll known-answer vectors that the reference's own tests assert with literal
`[][2]int{...}` expectations into tests/golden/ref_findall_vectors.json.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/harvest_findall_vectors.py
Only (pattern, input, expected) triples are extracted — no reference code is copied.  Table rows
of the shape {"name", "input", [][2]int{..}|nil} take their pattern from the searcher the test
builds (stated per table below); rows with pattern/haystack/want fields carry their own."""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_findall_vectors.json")

# tables whose pattern is fixed by the searcher the test constructs: (file, test function, pattern)
FIXED = [
    ("nfa/charclass_searcher_test.go", "TestCharClassSearcher_FindAllIndices", r"[a-zA-Z0-9_]+"),
    ("nfa/charclass_searcher_extra_test.go", "TestCharClassSearcher_FindAllIndices_Comprehensive", r"[0-9]+"),
]
# tables with explicit pattern / haystack / want fields
KEYED = [("meta/issue124_test.go", "TestIssue124_FindIndicesAt_NonGreedyIteration")]
# single assertions: (file, line of the expectation, pattern, input) read off the test body
SINGLE = [
    ("nfa/backtrack_search_test.go", "TestBacktracker_SearchAtWithState_Iteration", r"[a-z]+", "abc 123 def 456 ghi"),
]


def go_string(tok):
    tok = tok.strip()
    if tok.startswith("`"):
        return tok[1:-1]
    return bytes(tok[1:-1], "utf-8").decode("unicode_escape")


def pairs(tok):
    tok = tok.strip()
    if tok == "nil":
        return []
    return [[int(a), int(b)] for a, b in re.findall(r"\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", tok)]


def func_body(text, name):
    m = re.search(r"^func %s\(.*?^}" % re.escape(name), text, re.S | re.M)
    assert m, name
    return m.group(0), text[: m.start()].count("\n") + 1


def main():
    out = []
    for path, fn, pat in FIXED:
        body, line0 = func_body(open(os.path.join(REF, path)).read(), fn)
        for k, ln in enumerate(body.splitlines()):
            m = re.match(r'\s*\{("(?:[^"\\]|\\.)*")\s*,\s*("(?:[^"\\]|\\.)*"|`[^`]*`)\s*,\s*(nil|\[\]\[2\]int\{.*\})\s*\},?\s*$', ln)
            if m:
                out.append({"pattern": pat, "input": go_string(m.group(2)), "want": pairs(m.group(3)),
                            "src": "%s:%d (%s)" % (path, line0 + k, go_string(m.group(1)))})
    for path, fn in KEYED:
        body, line0 = func_body(open(os.path.join(REF, path)).read(), fn)
        for m in re.finditer(r'name:\s*("[^"]*")\s*,\s*pattern:\s*(`[^`]*`|"(?:[^"\\]|\\.)*")\s*,\s*'
                             r'haystack:\s*(`[^`]*`|"(?:[^"\\]|\\.)*")\s*,\s*want:\s*(nil|\[\]\[2\]int\{.*?\}\})', body, re.S):
            out.append({"pattern": go_string(m.group(2)), "input": go_string(m.group(3)), "want": pairs(m.group(4)),
                        "src": "%s:%d (%s)" % (path, line0 + body[: m.start()].count("\n"), go_string(m.group(1)))})
    for path, fn, pat, inp in SINGLE:
        body, line0 = func_body(open(os.path.join(REF, path)).read(), fn)
        m = re.search(r"expected := (\[\]\[2\]int\{.*\})", body)
        out.append({"pattern": pat, "input": inp, "want": pairs(m.group(1)),
                    "src": "%s:%d" % (path, line0 + body[: m.start()].count("\n"))})
    json.dump(out, open(OUT, "w"), indent=1)
    print("%d vectors -> %s" % (len(out), OUT))


if __name__ == "__main__":
    main()
