#!/usr/bin/env python
"""Harvests every pattern the reference's own test files mention in a `pattern:`-like field or as
the first back-quoted column of a table row into tests/golden/ref_test_patterns.json — the list the
GPU parity sweep (tests/test_ref_patterns.py) replays against the oracle.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/harvest_test_patterns.py
Only the pattern strings are extracted."""
import glob
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_test_patterns.json")


def main():
    pats = set()
    for f in glob.glob(os.path.join(REF, "**", "*_test.go"), recursive=True):
        src = open(f, errors="replace").read()
        for m in re.finditer(r'(?:pattern|pat|Pattern|re|regex)\s*[:=]\s*(`[^`\n]*`|"(?:[^"\\\n]|\\.)*")', src):
            t = m.group(1)
            try:
                pats.add(t[1:-1] if t[0] == "`" else json.loads(re.sub(r"\\x([0-9a-fA-F]{2})", r"\\u00\1", t)))
            except ValueError:
                continue
        for m in re.finditer(r"^\s*\{\s*(`[^`\n]*`)\s*,", src, re.M):
            pats.add(m.group(1)[1:-1])
    pats = sorted(p for p in pats if len(p) <= 300)
    json.dump(pats, open(OUT, "w"), indent=0, ensure_ascii=False)
    print(len(pats), "patterns")


if __name__ == "__main__":
    main()
