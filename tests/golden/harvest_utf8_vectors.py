#!/usr/bin/env python
"""Harvests the known-answer vectors of the reference's own UTF-8 tests
(/root/reference/nfa/compile_utf8_test.go: struct literals with pattern / haystack / wantPos) into
tests/golden/ref_utf8_vectors.json, so that the oracle can be pinned against them on machines where
/root/reference does not exist.  Test DATA only — no reference source is copied.

Run here (the reference is mounted read-only):  python tests/golden/harvest_utf8_vectors.py
"""
import json
import os
import re

SRC = "/root/reference/nfa/compile_utf8_test.go"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_utf8_vectors.json")


def unquote(tok):
    """Go interpreted / raw string literal -> bytes."""
    if tok[0] == "`":
        return tok[1:-1].encode()
    body, out, i = tok[1:-1], bytearray(), 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out += c.encode()
            i += 1
            continue
        e = body[i + 1]
        if e == "x":
            out.append(int(body[i + 2:i + 4], 16)); i += 4
        elif e == "u":
            out += chr(int(body[i + 2:i + 6], 16)).encode(); i += 6
        elif e == "U":
            out += chr(int(body[i + 2:i + 10], 16)).encode(); i += 10
        else:
            out += {"n": b"\n", "t": b"\t", "r": b"\r", "\\": b"\\", '"': b'"', "'": b"'", "0": b"\0"}[e]; i += 2
    return bytes(out)


STR = r'("(?:[^"\\]|\\.)*"|`[^`]*`)'
text = open(SRC, encoding="utf-8").read()
vectors = []
# keyed struct literals: pattern: "...", haystack: "...", wantPos: []int{a, b} | nil
for m in re.finditer(r"pattern:\s*" + STR + r",\s*haystack:\s*" + STR + r",(?:\s*//[^\n]*)?\s*wantPos:\s*(nil|\[\]int\{(\d+),\s*(\d+)\})", text):
    pat, hay = unquote(m.group(1)), unquote(m.group(2))
    want = None if m.group(3) == "nil" else [int(m.group(4)), int(m.group(5))]
    vectors.append({"pattern": pat.decode("utf-8", "surrogateescape"), "haystack_hex": hay.hex(), "first": want})
# positional literals of TestCompileUTF8_BoundaryRunes: {"name", "pattern", "haystack", true|false}
for m in re.finditer(r"\{" + STR + r",\s*" + STR + r",\s*" + STR + r",\s*(true|false)\}", text):
    pat, hay = unquote(m.group(2)), unquote(m.group(3))
    vectors.append({"pattern": pat.decode("utf-8", "surrogateescape"), "haystack_hex": hay.hex(),
                    "is_match": m.group(4) == "true"})
json.dump({"source": "nfa/compile_utf8_test.go", "vectors": vectors}, open(OUT, "w"), ensure_ascii=False, indent=1)
print(len(vectors), "vectors ->", OUT)
