"""-m gpu: the byte-search family (csrc/memchr.cu through coregex_b200/simd.py, which mirrors the
reference's `simd` package) against the reference's own test vectors and Python's bytes.find /
index semantics (the role bytes.IndexByte / bytes.Index play in simd/memchr_test.go, memmem_test.go)."""
import random

import numpy as np
import pytest

from coregex_b200 import simd

pytestmark = pytest.mark.gpu
FOX = b"the quick brown fox jumps over the lazy dog"


def test_memchr_reference_vectors():
    # reference simd/memchr_test.go:14-43
    for hay, needle, want in [(b"", ord("a"), -1), (b"a", ord("a"), 0), (b"a", ord("b"), -1), (b"hello", ord("h"), 0),
                              (b"hello", ord("l"), 2), (b"hello", ord("o"), 4), (b"hello", ord("x"), -1),
                              (b"hello world", ord("o"), 4), (bytes([0, 1, 2, 3]), 0, 0), (bytes([1, 2, 3, 4]), 0, -1),
                              (bytes([1, 2, 255, 4]), 255, 2), (bytes([5, 5, 5, 5]), 5, 0), (FOX, ord("q"), 4),
                              (FOX, ord("z"), 37), (FOX, ord("t"), 0), (FOX, ord("g"), 42)]:
        assert simd.Memchr(hay, needle) == want == hay.find(bytes([needle]))


@pytest.mark.parametrize("size", [1, 2, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1023, 1024, 1025, 4095, 4096,
                                  4097, 16383, 16384, 32767, 32768, 32769, 65535, 65536, 65537, 1 << 20])
def test_memchr_sizes(size):
    # reference simd/memchr_test.go:62-110: needle at the first, middle, last position, absent
    hay = bytearray(b"a" * size)
    assert simd.Memchr(bytes(hay), ord("X")) == -1
    for pos in sorted({0, size // 2, size - 1}):
        h = bytearray(hay)
        h[pos] = ord("X")
        assert simd.Memchr(bytes(h), ord("X")) == pos
        assert simd.Memchr2(bytes(h), ord("Y"), ord("X")) == pos
        assert simd.Memchr3(bytes(h), ord("Y"), ord("Z"), ord("X")) == pos


def test_memchr23_first_of_any():
    rng = random.Random(2)
    for _ in range(40):
        n = rng.randrange(0, 70000)
        h = bytes(rng.choice(b"abcdefgh") for _ in range(n))
        a, b, c = (rng.choice(b"abcdefghxyz") for _ in range(3))
        want2 = min([p for p in (h.find(bytes([a])), h.find(bytes([b]))) if p >= 0], default=-1)
        want3 = min([p for p in (h.find(bytes([a])), h.find(bytes([b])), h.find(bytes([c]))) if p >= 0], default=-1)
        assert simd.Memchr2(h, a, b) == want2
        assert simd.Memchr3(h, a, b, c) == want3


def test_digit_and_word_classes():
    # reference simd/memchr_digit_test.go, memchr_class_test.go
    assert simd.MemchrDigit(b"") == -1 and simd.MemchrDigit(b"abc") == -1
    assert simd.MemchrDigit(b"abc5def") == 3 and simd.MemchrDigit(b"0") == 0
    assert simd.MemchrDigit(b"x" * 100000 + b"7") == 100000
    assert simd.MemchrDigitAt(b"1a2b3", 0) == 0 and simd.MemchrDigitAt(b"1a2b3", 1) == 2
    assert simd.MemchrDigitAt(b"1a2b3", 3) == 4 and simd.MemchrDigitAt(b"1a2b3", 5) == -1
    h = b"9" + b"x" * 70000 + b"8" + b"y" * 10
    for at, want in [(0, 0), (1, 70001), (33000, 70001), (70001, 70001), (70002, -1)]:
        assert simd.MemchrDigitAt(h, at) == want
    assert simd.MemchrWord(b"   hello") == 3 and simd.MemchrWord(b" .,;") == -1 and simd.MemchrWord(b"_") == 0
    assert simd.MemchrNotWord(b"hello world") == 5 and simd.MemchrNotWord(b"abc_123") == -1
    table = [False] * 256
    for b in b"aeiou":
        table[b] = True
    assert simd.MemchrInTable(b"rhythm and blues", table) == 7
    assert simd.MemchrNotInTable(b"aeixou", table) == 3
    assert simd.MemchrInTable(b"", table) == -1


def test_memchr_pair():
    # reference simd/memchr_test.go:684-711 (the doc example of memchr_amd64.go:196-200 — "@", "c", 9
    # -> 7 — contradicts memchrPairGeneric, simd/memchr_generic_impl.go:253-266, and is not a test)
    for hay, b1, b2, off, want in [(b"hello", "h", "e", 1, 0), (b"hello", "l", "o", 2, 2), (b"hello", "h", "o", 1, -1),
                                   (b"hello", "x", "e", 1, -1), (b"hello", "h", "x", 1, -1), (b"", "a", "b", 1, -1),
                                   (b"hello", "h", "h", 0, 0), (b"hello", "h", "e", 0, -1), (b"hi", "h", "i", 5, -1),
                                   (b"hello", "h", "e", -1, -1), (b"test@example.com", "@", ".", 8, 4),
                                   (b"http://localhost", ":", "/", 1, 4), (b'{"key": "value"}', '"', ":", 5, 1),
                                   (b"abababab", "a", "b", 1, 0), (b"xxabxxab", "a", "b", 1, 2),
                                   (b"a0123456789b", "a", "b", 11, 0), (b"xxxxxab", "a", "b", 1, 5), (b"ab", "a", "b", 1, 0)]:
        assert simd.MemchrPair(hay, ord(b1), ord(b2), off) == want, (hay, b1, b2, off)
    assert simd.MemchrPair(b"contact@test.com for info", ord("@"), ord("c"), 6) == 7
    assert simd.MemchrPair(b"abcabc", ord("a"), ord("c"), 2) == 0
    assert simd.MemchrPair(b"abxabc", ord("a"), ord("c"), 2) == 3
    assert simd.MemchrPair(b"abc", ord("a"), ord("c"), -1) == -1
    assert simd.MemchrPair(b"abc", ord("a"), ord("c"), 3) == -1
    assert simd.MemchrPair(b"abc", ord("a"), ord("b"), 0) == -1
    assert simd.MemchrPair(b"xxaxx", ord("a"), ord("a"), 0) == 2
    rng = random.Random(4)
    for _ in range(30):
        n = rng.randrange(2, 50000)
        h = bytes(rng.choice(b"abc") for _ in range(n))
        off = rng.randrange(1, 40)
        want = next((p for p in range(max(0, n - off)) if h[p] == ord("a") and h[p + off] == ord("c")), -1)
        assert simd.MemchrPair(h, ord("a"), ord("c"), off) == want


def test_memmem():
    # reference simd/memmem_test.go:17-63, :82-150
    for hay, needle, want in [(b"hello", b"", 0), (b"", b"", 0), (b"", b"a", -1), (b"hi", b"hello", -1), (b"hello", b"ll", 2),
                              (b"hello!", b"!", 5), (b"hello world", b"wo", 6), (b"aaaaaabaaaa", b"aab", 4),
                              (b"hello world", b"xyz", -1), (b"hello", b"hello", 0)]:
        assert simd.Memmem(hay, needle) == want == hay.find(needle)
    for m in (2, 4, 8, 16, 32, 64, 128, 256):
        hay = bytearray(b"a" * 100000)
        needle = b"a" * (m - 1) + b"X"
        hay[-m:] = needle
        assert simd.Memmem(bytes(hay), needle) == 100000 - m
        hay = bytearray(b"b" * 100000)
        needle = b"X" + b"a" * (m - 1)
        hay[:m] = needle
        assert simd.Memmem(bytes(hay), needle) == 0
    rng = random.Random(6)
    for _ in range(40):
        n = rng.randrange(0, 80000)
        h = bytes(rng.choice(b"ab") for _ in range(n))
        needle = bytes(rng.choice(b"ab") for _ in range(rng.randrange(1, 20)))
        assert simd.Memmem(h, needle) == h.find(needle)


def test_first_hit_far_into_a_large_buffer():
    import torch
    n = 1 << 30
    t = torch.full((n + 64,), ord("x"), dtype=torch.uint8, device="cuda")
    for pos in (n - 1, n // 2 + 12345, 3 * (1 << 28) + 1):
        t[pos] = ord("7")
        assert simd.memchr_table_device(t.data_ptr(), n, simd._DIGIT) == pos
        assert simd.memmem_device(t.data_ptr(), n, b"x7") == pos - 1
        t[pos] = ord("x")
    assert simd.memchr_table_device(t.data_ptr(), n, simd._DIGIT) == -1
