"""Bitstream kernel (coregex_b200/csrc/scan_bits.cu) on the CPU SIMT emulator vs the oracle.

The kernel SOURCE is compiled with g++ against tests/sim/simt_cpu.h (one fiber per CUDA thread),
so these tests exercise the real warp-level code — marker passes, ownership by sync bytes, the
alternation check, the serial replay, bitmap extraction and the epoch look-back — in
a container without a GPU.  The -m gpu tests repeat the same comparisons on the device.
"""
import random

import numpy as np
import pytest

import coregex_b200 as cg
import sim_lib
from oracle_lib import Oracle

IP = r"\d+\.\d+\.\d+\.\d+"
# geometry of the kernel (scan_bits.cu): 64-byte words, K = 8 words per lane, 2 KB tiles, chunks of
# 32 K words that overlap by one word; STRIDE = one lane's region
K = 8
STRIDE, TILE, CHUNK = 64 * K, 2048, (32 * K - 1) * 64
BACKEND = "sim"


@pytest.fixture(autouse=True, params=["sim", "simjit", "simjit_k8", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """Every case runs on the emulator — the interpreting build and the per-pattern specialised
    build (what jit.cu compiles for the device) — and, under -m gpu, on the device through the C ABI."""
    global BACKEND
    BACKEND = request.param
    yield
    BACKEND = "sim"


def scan(pat, hay, mode=0, cap=None, grid=2, base=0):
    """(total, flag, pairs) from the selected backend."""
    if BACKEND in ("sim", "simjit", "simjit_k8"):
        # simjit: the per-pattern, per-mode specialised build (what jit.cu compiles for the device);
        # simjit_k8: the same with 8 words per lane and a 4-deep window ring (other chunk geometry);
        # every emulated scan runs twice on the same scratch (stale look-back words of the first launch)
        k4 = BACKEND == "simjit_k8"
        return sim_lib.scan(pat, hay, mode=mode, cap=cap, grid=grid, base=base, jit=BACKEND != "sim",
                            defs="-DCGX_K=8 -DCGX_NB=4 -DCGX_WARPS=8 -DCGX_CAP=300" if k4 else "", tag="k8" if k4 else "", launches=2)
    import torch
    from gpu_util import scan_device
    r = cg.Compile(pat)
    if not r.engine.endswith("+bitstream"):
        raise sim_lib.NotEligible(pat)
    a = np.frombuffer(bytes(hay), dtype=np.uint8) if not isinstance(hay, np.ndarray) else hay
    t = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda")
    t[a.size:] = ord("1")  # bytes after the haystack must never be interpreted
    if a.size:
        t[: a.size] = torch.from_numpy(a.copy()).cuda()
    return scan_device(r, t[: a.size], mode=mode, cap=(a.size + 16 if cap is None else cap), base=base)


def check(pat, hay, grid=2, **kw):
    if isinstance(hay, (bytes, bytearray)):
        hay = np.frombuffer(bytes(hay), dtype=np.uint8)
    want = Oracle(pat).find_all(hay)
    tot, _, pairs = scan(pat, hay, grid=grid, **kw)
    assert tot == len(want), (pat, tot, len(want))
    assert np.array_equal(pairs, want), (pat, pairs[:5], want[:5])
    return want


def test_log_corpus_multi_cta():
    hay = cg.synth_host(cg.SYNTH_LOG, 7, 4096 * 48)
    w = check(IP, hay, grid=3)
    assert len(w) > 1500


@pytest.mark.parametrize("n", [0, 1, 7, 15, 16, 17, 63, 64, 65, 255, 256, 257, STRIDE - 1, STRIDE, STRIDE + 1,
                               TILE - 1, TILE, TILE + 1, 2 * STRIDE - 1, 2 * STRIDE, 2 * STRIDE + 1,
                               STRIDE + TILE - 1, STRIDE + TILE, STRIDE + TILE + 1, 8128 - 1, 8128, 8128 + 1,
                               8192, CHUNK - 1, CHUNK, CHUNK + 1, CHUNK + 63,
                               CHUNK + 64, CHUNK + 65, 2 * CHUNK, 2 * CHUNK + 64])
def test_sizes_around_tile_and_chunk_edges(n):
    unit = b"a 1.2.3.4 b 10.20.30.40.50 c9.9.9.9\n"
    hay = (unit * (n // len(unit) + 1))[:n]
    check(IP, hay)
    # the input ends inside / right after a match
    hay2 = bytearray(hay)
    tail = b"7.7.7.77"
    if n >= len(tail) + 1:
        hay2[n - len(tail) - 1] = ord(" ")
        hay2[n - len(tail):] = tail
        check(IP, hay2)


def test_match_straddles_every_tile_offset():
    ip = b"123.45.67.89"
    for off in range(STRIDE - 16, STRIDE + 70):
        hay = bytearray(b"x" * (3 * STRIDE))
        hay[off:off + len(ip)] = ip
        hay[off + TILE:off + TILE + len(ip)] = ip
        w = check(IP, hay, grid=1)
        assert len(w) == 2


def test_overlapping_candidates_take_the_serial_path():
    hay = (b"v 1.2.3.4.5.6.7.8 w 1.2.3.4.5 z\n" * 200)
    check(IP, hay)
    hay = b"1.2.3.4.5.6.7.8.9.10.11.12" * 400  # no sync byte at all
    check(IP, hay)


def test_long_spans_without_sync_bytes():
    rng = random.Random(5)
    parts = []
    for _ in range(40):
        parts.append(b" " * rng.randrange(1, 50))
        parts.append(bytes(rng.choice(b"0123456789.") for _ in range(rng.randrange(1, 5000))))
        parts.append(b" 1.2.3.4 ")
    check(IP, b"".join(parts), grid=2)
    check(IP, b"9" * 5000 + b".1.1.1 " + b"8" * 3000)
    check(IP, b"1.1.1." + b"9" * 6000)


def test_open_segments_take_the_bit_parallel_replay():
    # stretches without a sync byte that run past a lane's neighbour word (64 .. 500 bytes): the owning
    # lane replays them bit-parallel from global memory (replay_bits), longer ones byte by byte
    rng = random.Random(11)
    for pat in (r"[a-z]+/\d+", r"\w+@\w+\.\w+", IP):
        parts = []
        for _ in range(300):
            n = rng.choice([3, 20, 70, 130, 200, 260, 400, 450, 700])
            if pat == IP:
                body = b"".join(rng.choice([b"1", b"22", b"333"]) + rng.choice([b".", b".", b""]) for _ in range(n // 3))
                parts.append(body + rng.choice([b" ", b" x ", b"\n"]))
            elif "@" in pat:
                parts.append(b"w" * n + b"@" + b"h" * rng.randrange(1, 90) + b"." + b"c" * rng.randrange(1, 40) + rng.choice([b" ", b"\n", b". "]))
            else:
                parts.append(bytes(rng.choice(b"abcxyz") for _ in range(n)) + b"/" + b"7" * rng.randrange(1, 80) + rng.choice([b" ", b"/x ", b"\n"]))
        check(pat, b"".join(parts), grid=2)


def test_class_words_of_all_ones_use_exact_carries():
    # a 64-byte piece made of class bytes only: the fast carry chain (one ballot) is not valid there
    rng = random.Random(21)
    parts = []
    for _ in range(120):
        parts.append(b"x" * rng.randrange(1, 30) + b" ")
        parts.append(b"9" * rng.randrange(60, 400) + b"." + b"1" * rng.randrange(1, 200) + b".22.3 ")
        parts.append(b"w" * rng.randrange(64, 300) + b"@" + b"h" * rng.randrange(1, 150) + b".org ")
    hay = b"".join(parts)
    check(IP, hay, grid=2)
    check(r"\w+@\w+\.\w+", hay, grid=2)
    check(r"\d+", hay, grid=2)


def test_alternation_criterion_exhaustive():
    """word_misordered() in scan_bits.cu, restated: per word, with `inn` = starts below - ends below,
    D = E - S - inn must satisfy D ^ (D << 1 | inn) == S ^ E and inn in {0,1}; plus equal totals.
    Checked against the definition (spans sorted by start never overlap, ends distinct) for every
    assignment of ends to every set of starts on three 3-bit words."""
    import itertools
    B, NW = 3, 3
    W, M = B * NW, (1 << B) - 1

    def truth(pairs):
        last = -1
        for s_, e_ in sorted(pairs):
            if s_ < last:
                return False
            last = e_
        return len({e_ for _, e_ in pairs}) == len(pairs)

    def kernel(S, E):
        if bin(S).count("1") != bin(E).count("1"):
            return False
        cs = ce = 0
        for w in range(NW):
            Sw, Ew = (S >> (B * w)) & M, (E >> (B * w)) & M
            inn = cs - ce
            if inn not in (0, 1):
                return False
            D = (Ew - Sw - inn) & M
            if D ^ (((D << 1) | inn) & M) != Sw ^ Ew:
                return False
            cs += bin(Sw).count("1")
            ce += bin(Ew).count("1")
        return True

    if BACKEND != "sim":
        pytest.skip("backend-independent")
    for S in range(1, 1 << W):
        starts = [p_ for p_ in range(W) if S >> p_ & 1]
        for ends in itertools.product(*[range(s_ + 1, W) for s_ in starts]):
            E = 0
            for e_ in ends:
                E |= 1 << e_
            assert kernel(S, E) == truth(list(zip(starts, ends))), (starts, ends)


def test_random_digit_dot_soup():
    rng = random.Random(11)
    for trial in range(30):
        alphabet = rng.choice([b"0123456789. ", b"01.", b"0. \n", b"12345.x"])
        n = rng.choice([50, 500, 2100, 4100, 9000, 17000])
        hay = bytes(rng.choice(alphabet) for _ in range(n))
        check(IP, hay, grid=rng.choice([1, 2]))


PATTERNS = [
    r"\w+@\w+\.\w+",
    r"[a-z]+=\d+",
    r"ab+c",
    r"a+ba",           # a match can end in the middle of a run of the first class
    r"\d{1,3}\.\d{1,3}",
    r"[A-Z][a-z]+",
    r"x\d*y?z",
    r"\d+",
    r"[a-c]+[x-z]?",
    r"GET|POST",       # not flat: must be refused by the bitstream engine
    r"fo+\d+b",
    r"\d{4}-\d{2}-\d{2}",
]


@pytest.mark.parametrize("pat", PATTERNS)
def test_pattern_zoo(pat):
    rng = random.Random(hash(pat) & 0xFFFF)
    alphabet = b"abcxyz0123456789@.=- ABZGETPOSfor\n"
    words = [b"user@host.com", b"key=123", b"abbbc", b"aabaab", b"1.22.333", b"Hello", b"x12yz", b"xz", b"foo42bar", b"fooo7b",
             b"2024-01-31", b"aaabaaaba", b"abcabcx", b"GET", b"POST"]
    for trial in range(6):
        parts = []
        size = 0
        target = rng.choice([300, 2500, 5000, 20000])
        while size < target:
            if rng.random() < 0.4:
                p = rng.choice(words)
            else:
                p = bytes(rng.choice(alphabet) for _ in range(rng.randrange(1, 12)))
            parts.append(p)
            size += len(p)
        hay = b"".join(parts)
        try:
            check(pat, hay, grid=rng.choice([1, 2]))
        except sim_lib.NotEligible:
            # not flat, or neighbouring starts would share their end (`\d \d?`): candidate/DFA kernel
            assert pat in (r"GET|POST", r"\d{1,3}\.\d{1,3}"), pat
            return


def test_dense_matches():
    hay = b"1 22 333 4 55 6 7 8 9 0 " * 2500  # every second byte starts a match
    w = check(r"\d+", hay, grid=2)
    assert len(w) > 20000
    check(IP, b"1.1.1.1 " * 6000, grid=2)


def test_count_ismatch_and_cap():
    hay = cg.synth_host(cg.SYNTH_LOG, 3, 4096 * 16)
    want = Oracle(IP).find_all(hay)
    tot, flag, _ = scan(IP, hay, mode=cg.MODE_COUNT)
    assert tot == len(want)
    tot, flag, _ = scan(IP, hay, mode=cg.MODE_ISMATCH)
    assert flag == 1
    tot, flag, _ = scan(IP, b"no address here 1.2.3 \n" * 3000, mode=cg.MODE_ISMATCH)
    assert flag == 0
    tot, flag, _ = scan(IP, b"x" * 40000 + b" 1.2.3.4", mode=cg.MODE_ISMATCH)
    assert flag == 1
    # capped output: the first `cap` matches, total still counts all
    tot, _, pairs = scan(IP, hay, cap=100)
    assert tot == len(want) and np.array_equal(pairs, want[:100])


def test_base_offset_is_added():
    hay = b"a 1.2.3.4 b\n" * 500
    want = Oracle(IP).find_all(np.frombuffer(hay, dtype=np.uint8))
    tot, _, pairs = scan(IP, hay, base=1 << 40)
    assert np.array_equal(pairs, want + (1 << 40))


@pytest.mark.gpu
def test_device_runs_the_nvrtc_specialised_kernel():
    import ctypes as C
    r = cg.Compile(IP)
    assert r.FindAllIndex(b"a 1.2.3.4 b") == [[2, 9]]
    st = cg._lib.cgx_debug_jit_state(r._h)
    assert st == 1, (st, cg._lib.cgx_last_error())
