"""Generates tests/golden/stdlib_corpus.txt and tests/golden/oracle_vectors.json.

The corpus is our own text with the same KINDS of lines as the reference's differential-test
corpus (reference meta/stdlib_compat_test.go:144-199: request lines, access-log lines with IPs,
level-prefixed messages, e-mails, URLs, versions, word lists, hex tokens).  The expected
FindAllIndex results (first 1000 matches, as the reference's test caps them) come from Python
`re` on bytes — an independent leftmost-first engine, which is the role Go's stdlib regexp plays
for the reference's own tests.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))

LINES = [
    "HTTP/1.0 404 Not Found",
    "GET /v1/items HTTP/1.1",
    "POST /v1/session HTTP/1.1",
    "PUT /v1/settings HTTP/1.0",
    '203.0.113.7 - - [02/Feb/2025:09:14:03 +0000] "GET /home.html HTTP/1.1" 200 4312',
    '10.20.30.40 - root [02/Feb/2025:11:45:59 +0000] "POST /signin?pwd=hunter2 HTTP/1.1" 302 17',
    '198.51.100.250 - - [02/Feb/2025:23:00:00 +0000] "GET /v1/report HTTP/1.1" 500 0',
    "error: upstream timed out after 3000ms",
    "warning: cache hit ratio below 0.75",
    "fatal: out of file descriptors",
    "critical: certificate expires in 2 days",
    "[Error] handler crashed at frame 17 (ERROR code 12)",
    "alice@example.net forwarded the memo to bob_smith@corp.example.org",
    "write to help+desk@support-site.co for assistance",
    "see https://docs.example.net/guide?page=2#intro or http://mirror.example.org/pub",
    "release 4.5.6 replaces 4.5.5; kernel 6.1.0 needs firmware 20.04",
    "apple banana cherry grape lemon mango melon olive peach plum kiwi lime",
    "sid=9f8e7d6c5b4a token=00112233445566778899aabbccddeeff",
    "failed login for user42 from 192.0.2.33 at 07:08:09",
    "wrote summary.txt, rotated server.log, updated README.md",
    "abc123 def456 test789 hello world123",
    "word word2 word34 word567 word8901",
    "foo bar foobar barfoo foo_bar foo",
    "call 555-0199 or 555-1234 before 2025-03-01",
    "latency 12ms 340ms 7ms p99=1200ms",
    "1.2.3 1.2.3. 1..2.3.4 1.2.3.4.5.6.7.8 999.999.999.999x",
    "x",
    "",
    "trailing digits 12345",
    "0 00 007 7 70 700",
]

PATTERNS = [
    r"\d+\.\d+\.\d+\.\d+", r"error|warning|fatal|critical",
    r"apple|banana|cherry|grape|lemon|mango|melon|olive|peach|plum|kiwi|lime",
    r"\d+", r"\w+", r"[a-z]+", r"[A-Z][a-z]+", r"\d{4}-\d{2}-\d{2}", r"\w+@\w+\.\w+",
    r"GET|POST|PUT", r"[0-9]+ms", r"\d+\.\d+", r"user\d+", r"(?i)error", r"\bfoo\b",
    r"[1-9][0-9]*|0", r"\d{3}-\d{4}", r"https?://[a-z.]+", r"(?m)^\d+",
    r"\d+\.\d+\.\d+", r"[a-f0-9]{32,}", r"(?m)^(GET|POST|PUT)", r"word\d+", r"[a-zA-Z]+\d+",
]


def main():
    corpus = ("\n".join(LINES) + "\n").encode() * 40
    with open(os.path.join(HERE, "stdlib_corpus.txt"), "wb") as fh:
        fh.write(corpus)
    vec = []
    for p in PATTERNS:
        ms = [[m.start(), m.end()] for m in re.finditer(p.encode(), corpus)][:1000]
        vec.append({"pattern": p, "matches": ms})
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as fh:
        json.dump(vec, fh, separators=(",", ":"))
    print("corpus bytes:", len(corpus), "patterns:", len(PATTERNS))


if __name__ == "__main__":
    main()
