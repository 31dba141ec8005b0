#!/usr/bin/env python
"""Harvests the reference's own test corpus — the `lines` table of generateTestInput
(reference meta/stdlib_compat_test.go:145-187, SURVEY.md §8d "replay the reference's own 41-line
template corpus") — into tests/golden/ref_stdlib_corpus.txt (one line per table row, in order;
the test repeats the block 100 times, the replaying test does the same).

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/harvest_stdlib_corpus.py
Only the string literals of the table are extracted — no reference code is copied."""
import os
import re

REF = "/root/reference/meta/stdlib_compat_test.go"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stdlib_corpus.txt")


def main():
    src = open(REF).read()
    body = src[src.index("func generateTestInput()"):]
    body = body[body.index("lines := []string{") + len("lines := []string{"):]
    body = body[:body.index("\n\t}\n")]
    lines = []
    for row in body.splitlines():
        row = row.strip().rstrip(",")
        if not row:
            continue
        if row.startswith("`"):
            lines.append(row[1:-1])
        elif row.startswith('"'):
            lines.append(bytes(row[1:-1], "utf-8").decode("unicode_escape"))
        elif row.startswith("fmt.Sprintf("):
            # fmt.Sprintf("long token: %s end", strings.Repeat("abcdef0123456789", 3))
            m = re.match(r'fmt\.Sprintf\("([^"]*)", strings\.Repeat\("([^"]*)", (\d+)\)\)', row)
            assert m, row
            lines.append(m.group(1).replace("%s", m.group(2) * int(m.group(3))))
        else:
            raise SystemExit("unrecognised table row: " + row)
    assert len(lines) == 41, len(lines)
    with open(OUT, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("%d lines -> %s" % (len(lines), OUT))
    # the pattern table of the same test (TestStdlibCompatibility, :30-72): {"name", `pattern`} rows
    head = src[src.index("func TestStdlibCompat"):src.index("func generateTestInput()")]
    pats = re.findall(r'\{"([a-z_0-9]+)", `([^`]*)`\}', head)
    assert len(pats) >= 30, len(pats)
    import json
    with open(os.path.join(os.path.dirname(OUT), "ref_stdlib_patterns.json"), "w") as fh:
        json.dump({"source": "reference meta/stdlib_compat_test.go (pattern table of the stdlib comparison test)",
                   "repeat": 100, "patterns": [{"name": n, "pattern": p} for n, p in pats]}, fh, indent=1)
    print("%d patterns" % len(pats))


if __name__ == "__main__":
    main()
