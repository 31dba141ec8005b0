"""-m gpu: FindAllSubmatchIndex (scan kernel + one Pike lane per match) against the CPU oracle's
restatement of the reference PikeVM capture search."""
import os

import numpy as np
import pytest

import coregex_b200 as cg
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(pat, hay):
    r, o = cg.Compile(pat), Oracle(pat)
    want = o.find_all_submatch(hay)
    got = r.FindAllSubmatchIndex(hay)
    if len(want) == 0:
        assert got is None
        return
    assert np.array_equal(np.array(got, dtype=np.int64), want), (pat, len(hay))


def test_reference_known_answers():
    # reference meta/findall_extra_test.go:147-157
    assert cg.Compile(r"(\w+)@(\w+)").FindAllSubmatchIndex(b"user@host admin@server") == [
        [0, 9, 0, 4, 5, 9], [10, 22, 10, 15, 16, 22]]
    # reference nfa/pikevm_slottable_test.go:171-188
    assert cg.Compile(r"(a+)(b+)").FindAllSubmatchIndex(b"xxxaaabbbyyy") == [[3, 9, 3, 6, 6, 9]]
    assert cg.Compile(r"([a-z]+)([0-9]+)").FindAllSubmatchIndex(b"abc123xyz") == [[0, 6, 0, 3, 3, 6]]
    r = cg.Compile(r"\w+@\w+\.\w+")          # no groups: stride 2 (reference meta/findall.go:109-112)
    assert r.NumSubexp() == 0
    assert r.FindAllSubmatchIndex(b"Contact: user@example.com for info") == [[9, 25]]
    assert cg.Compile(r"(\d+)").FindAllSubmatchIndex(b"none") is None
    assert cg.Compile(r"(\d+)").FindAllSubmatchIndex(b"1 2 3", 2) == [[0, 1, 0, 1], [2, 3, 2, 3]]


PATS = [r"(\w+)@(\w+)\.(\w+)", r"(\d+)\.(\d+)\.(\d+)\.(\d+)", r"(\w+)@(\w+)", r"(a)?b", r"(foo)|(bar)",
        r"((a)(b))+c", r"(\d+)-(\d+)?x", r"([a-z]+?)(\d+)", r"(?P<user>\w+)@(?P<host>[a-z]+)", r"(a|ab)(c|bcd)(d*)x"]


@pytest.mark.parametrize("pat", PATS)
def test_random_haystacks(pat):
    rng = np.random.default_rng(13)
    alphabet = np.frombuffer(b"0123456789.. abcdx\n@_-foobar", dtype=np.uint8)
    r, o = cg.Compile(pat), Oracle(pat)
    for it in range(60):
        n = int(rng.integers(0, 400))
        h = bytes(alphabet[rng.integers(0, len(alphabet), n)])
        want = o.find_all_submatch(h)
        got = r.FindAllSubmatchIndex(h)
        if len(want) == 0:
            assert got is None, (pat, h)
        else:
            assert np.array_equal(np.array(got, dtype=np.int64), want), (pat, h)


# flat deterministic patterns: group offsets = item boundaries of the forced greedy walk (flat_caps.cu)
FLAT_PATS = [r"(\w+)@(\w+)\.(\w+)", r"(\d+)\.(\d+)\.(\d+)\.(\d+)", r"((\w+)@(\w+))\.(\w+)", r"([a-z]+)=(\d{1,3})",
             r"x(\d{2})(\d*)-", r"(?P<k>[a-z]+)(=)(?P<v>\d+)", r"(\d+)(\.)(\d+)", r"([a-f]+)([0-9]+)(x?)\.",
             r"(\d+\.\d+)\.(\d+\.(\d+))"]
# groups under a quantifier keep the Pike captures kernel
PIKE_PATS = [r"(\d)+\.", r"(a)?b+@", r"([a-z])*=\d+"]


def _engine(r):
    return cg._lib.cgx_debug_captures_engine(r._h)


@pytest.mark.parametrize("pat", FLAT_PATS + PIKE_PATS)
def test_flat_pattern_groups(pat):
    rng = np.random.default_rng(29)
    alphabet = np.frombuffer(b"0123456789..== abcdefx\n@_-", dtype=np.uint8)
    r, o = cg.Compile(pat), Oracle(pat)
    assert _engine(r) == (1 if pat in FLAT_PATS else 2), (pat, r.engine)
    specials = [b"user@example.com", b"10.20.30.40", b"key=123 other=4567", b"x12345- x12-", b"abc123x. ff00.", b"1.2.3.4.5.6.7.8"]
    for it in range(80):
        n = int(rng.integers(0, 600))
        h = bytearray(alphabet[rng.integers(0, len(alphabet), n)])
        for _ in range(int(rng.integers(0, 6))):
            sp = specials[int(rng.integers(0, len(specials)))]
            at = int(rng.integers(0, max(1, len(h))))
            h[at:at] = sp
        h = bytes(h)
        want = o.find_all_submatch(h)
        got = r.FindAllSubmatchIndex(h)
        if len(want) == 0:
            assert got is None, (pat, h)
        else:
            assert np.array_equal(np.array(got, dtype=np.int64), want), (pat, h)


# groups around `.`, `\S`, negated and non-ASCII classes: UTF-8 byte automata, more instructions than the
# 64-instruction form of the captures kernel holds — the 512-instruction form (engine code 3), its
# thread lists sized by the proved widest generation (host/pike_pack.cpp MaxLiveThreads)
LARGE_PATS = [r"(\S+)@(\S+)", r"(.*)=(.*)", r"([^,\n]+),([^,\n]+)", r"(\w{1,40})-(\w{1,40})", r"([а-я]+) ([а-я]+)",
              r"(GET|POST) (/\S*)", r'"([A-Z]+) (.*)" (\d+)', r"(?i)(error|warn): (.+)", r"(\S+?)=(\S*?);", r"(é|e)+(.)"]


@pytest.mark.parametrize("pat", LARGE_PATS)
def test_groups_around_utf8_classes(pat):
    rng = np.random.default_rng(31)
    pieces = [b"a", b"b", b"x", b" ", b" ", b"\n", b",", b"@", b"=", b";", b"-", b"e", b"error", b"WARN: ", b"12", b"GET /",
              b"POST /a", b'"GET ', b'" 200', "é".encode(), "мир".encode(), "привет".encode(), "世界".encode(),
              "😀".encode(), b"\xff", b"\xe0\x80", b"\xc3", b"user@host", b"k=v;"]
    r, o = cg.Compile(pat), Oracle(pat)
    assert _engine(r) in (2, 3), (pat, r.engine, _engine(r))
    for it in range(50):
        h = b"".join(pieces[int(i)] for i in rng.integers(0, len(pieces), int(rng.integers(0, 90))))
        want = o.find_all_submatch(h)
        got = r.FindAllSubmatchIndex(h)
        if len(want) == 0:
            assert got is None, (pat, h)
        else:
            assert np.array_equal(np.array(got, dtype=np.int64), want), (pat, h)
    h = b"".join(pieces[int(i)] for i in rng.integers(0, len(pieces), 60000))
    assert np.array_equal(np.array(r.FindAllSubmatchIndex(h), dtype=np.int64), o.find_all_submatch(h)), pat


def test_large_form_is_selected_by_instruction_count_or_width():
    assert _engine(cg.Compile(r"(\S+)@(\S+)")) == 3 and _engine(cg.Compile(r"(\w{1,40})-(\w{1,40})")) == 3
    assert _engine(cg.Compile(r"(GET|POST) (/\S*)")) == 2      # 40 consuming instructions, 9 alive at once


@pytest.mark.parametrize("lines", [1, 100, 4000, 100000])
def test_c4_email_lines(lines):
    hay = cg.synth_host(cg.SYNTH_EMAIL, 0xC0FFEE + 4, 80 * lines)
    check(r"(\w+)@(\w+)\.(\w+)", hay)
    check(r"\w+@\w+\.\w+", hay)


def test_ip_groups_on_log_corpus():
    hay = cg.synth_host(cg.SYNTH_LOG, 5, 4096 * 200)
    check(r"(\d+)\.(\d+)\.(\d+)\.(\d+)", hay)


def test_c4_sized_device_entry():
    """BASELINE config 4 shape: 10M x 80-byte lines, captures form; properties + sampled oracle."""
    import torch
    from gpu_util import dev_corpus
    nlines = 10_000_000
    n = 80 * nlines
    t = dev_corpus(cg.SYNTH_EMAIL, 0xC0FFEE + 4, n)
    r = cg.Compile(r"(\w+)@(\w+)\.(\w+)")
    cap = nlines + 1024
    out = torch.empty((cap, 8), dtype=torch.int64, device="cuda")
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    r.scan_submatch_device(t.data_ptr(), n, out.data_ptr(), cap, res.data_ptr())
    torch.cuda.synchronize()
    total = int(res[0].item())
    assert total == nlines                       # exactly one e-mail per line
    m = out[:total].cpu().numpy()
    assert np.all(m[:, 0] // 80 == np.arange(nlines))            # one per line, in order
    assert np.all(m[:, 2] == m[:, 0]) and np.all(m[:, 7] == m[:, 1])
    assert np.all(m[:, 3] + 1 == m[:, 4]) and np.all(m[:, 5] + 1 == m[:, 6])
    o = Oracle(r"(\w+)@(\w+)\.(\w+)")
    rng = np.random.default_rng(4)
    for b in rng.integers(0, nlines // 1000, 16):
        lo = int(b) * 1000
        hay = cg.synth_host(cg.SYNTH_EMAIL, 0xC0FFEE + 4, 80 * 1000, first_block=lo)
        want = o.find_all_submatch(hay)
        want = np.where(want >= 0, want + lo * 80, want)
        assert np.array_equal(m[lo:lo + 1000], want)


# ---- FindAllSubmatchIndex of a large HOST buffer: pieces cut at record delimiters flow through
# H2D | scan + captures | D2H (capi.cu host_scan_pipelined, rows of 2*(groups+1) int64).  The piece size
# is read once per process, so the pipelined run happens in a child.
_PIPE_SUB_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import coregex_b200 as cg
from oracle_lib import Oracle
ok = True
email = cg.synth_host(cg.SYNTH_EMAIL, 21, 80 * 30000)
log = cg.synth_host(cg.SYNTH_LOG, 22, 4096 * 500)
for pat, hay in [(r"(\w+)@(\w+)\.(\w+)", email),            # flat captures pass (flat_caps.cu)
                 (r"(\d+)\.(\d+)\.(\d+)\.(\d+)", log),
                 (r"(\w+)@((\w+)\.)+(\w+)", email),           # group under a quantifier: Pike captures pass
                 (r"(GET|POST) (/\S*)", log),
                 (r"(a*)(\d*)", log[:300000]),                # nullable: PikeVM search kernel + captures
                 (r"\w+@\w+", email)]:                        # no groups: rows are the pairs
    r, o = cg.Compile(pat), Oracle(pat)
    want = o.find_all_submatch(hay)
    got = r.FindAllSubmatchIndex(hay)
    good = (got is None and len(want) == 0) or np.array_equal(np.array(got, dtype=np.int64), want)
    part = r.FindAllSubmatchIndex(hay, 777)
    good = good and np.array_equal(np.array(part, dtype=np.int64), want[:777])
    small = np.full((100, want.shape[1]), -7, dtype=np.int64)   # caller buffer smaller than the result
    cnt = r.find_all_into(hay, small, submatch=True)
    good = good and cnt == len(want) and np.array_equal(small, want[:100])
    print(pat, len(hay), len(want), r.engine, good)
    ok = ok and good
sys.exit(0 if ok else 1)
"""


def test_pipelined_host_submatch_matches_oracle():
    import subprocess
    import sys
    env = dict(os.environ, CGX_PIPELINE_PIECE=str(160 * 1024))
    p = subprocess.run([sys.executable, "-c", _PIPE_SUB_CHILD, ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr


def test_groups_of_the_empty_match_at_the_end_of_the_haystack_are_unset():
    """Reference nfa/pikevm.go:2201-2206: a capture search that starts AT len(haystack) answers from
    matchesEmptyAt and builds the result from no slots — `(a*)(\\d*)` on "x" is [0 0 0 0 0 0] and
    [1 1 -1 -1 -1 -1] (stdlib: [1 1 1 1 1 1]).  Restated in the oracle, reproduced by the captures kernel,
    for the LAST shard only."""
    import torch
    assert cg.Compile(r"(a*)(\d*)").FindAllSubmatchIndex(b"x") == [[0, 0, 0, 0, 0, 0], [1, 1, -1, -1, -1, -1]]
    assert cg.Compile(r"(\d*)").FindAllSubmatchIndex(b"") == [[0, 0, -1, -1]]
    for pat in [r"(a*)(\d*)", r"(\d*)", r"(x?)(y?)"]:
        check(pat, cg.synth_host(cg.SYNTH_LOG, 22, 4096 * 20)[:70001])
        check(pat, b"ab 12\n\nx9\n")
    # two shards: the first shard's end is not the end of the haystack
    pat, hay = r"(a*)(\d*)", b"ab 12\nxa7\n"
    r, want = cg.Compile(pat), Oracle(pat).find_all_submatch(b"ab 12\nxa7\n")
    parts = []
    for piece, base, after in ((hay[:6], 0, 4), (hay[6:], 6, 0)):
        t = torch.frombuffer(bytearray(piece + b"\0" * 16), dtype=torch.uint8).cuda()
        res = torch.zeros(2, dtype=torch.int64, device="cuda")
        out = torch.empty((64, 6), dtype=torch.int64, device="cuda")
        r.scan_submatch_device(t.data_ptr(), len(piece), out.data_ptr(), 64, res.data_ptr(), base_offset=base, bytes_after=after)
        torch.cuda.synchronize()
        parts.append(out[: int(res[0].item())].cpu().numpy())
    assert np.array_equal(np.concatenate(parts), want)
