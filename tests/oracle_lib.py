"""ctypes binding to the CPU parity oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (coregex_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def build(force=False):
    """Compile the oracle if missing (g++ only; seconds)."""
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, u8p = C.c_void_p, C.c_int64, C.c_void_p
        L.orc_compile.restype = vp
        L.orc_compile.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_free.argtypes = [vp]
        L.orc_strategy.argtypes = [vp]
        L.orc_strategy_name.restype = C.c_char_p
        L.orc_strategy_name.argtypes = [vp]
        L.orc_strategy_exact.argtypes = [vp]
        L.orc_has_bidirectional.argtypes = [vp]
        L.orc_set_bidirectional.argtypes = [vp, C.c_int]
        L.orc_num_captures.argtypes = [vp]
        L.orc_digit_run_skip_safe.argtypes = [vp]
        L.orc_is_match.argtypes = [vp, u8p, i64]
        L.orc_find_all.restype = i64
        L.orc_find_all.argtypes = [vp, u8p, i64, i64, C.c_void_p, i64]
        L.orc_count.restype = i64
        L.orc_count.argtypes = [vp, u8p, i64, i64]
        L.orc_find_all_submatch.restype = i64
        L.orc_find_all_submatch.argtypes = [vp, u8p, i64, i64, C.c_void_p, i64]
        L.orc_find_at.argtypes = [vp, u8p, i64, i64, C.POINTER(i64), C.POINTER(i64)]
        L.orc_dump_ast.restype = C.c_char_p
        L.orc_dump_ast.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_dump_nfa.restype = C.c_char_p
        L.orc_dump_nfa.argtypes = [vp]
        L.orc_dump_prefixes.restype = C.c_char_p
        L.orc_dump_prefixes.argtypes = [vp]
        L.orc_memchr_digit_at.restype = i64
        L.orc_memchr_digit_at.argtypes = [u8p, i64, i64]
        L.orc_scan_mt.restype = i64
        L.orc_scan_mt.argtypes = [C.c_char_p, C.c_size_t, u8p, i64, C.c_int, C.c_int,
                                  C.POINTER(C.c_double)]
        L.orc_teddy_find.argtypes = [u8p, C.c_void_p, C.c_int, u8p, i64, i64,
                                     C.POINTER(i64), C.POINTER(i64)]
        _lib = L
    return _lib


def _buf(data):
    """bytes / bytearray / numpy uint8 -> (pointer, length, keepalive)."""
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(data), dtype=np.uint8)
    return a.ctypes.data, a.size, a


class OracleError(Exception):
    pass


class Oracle:
    """The reference engine restated on the CPU (see oracle/meta.h)."""

    def __init__(self, pattern):
        if isinstance(pattern, str):
            pattern = pattern.encode()
        self.pattern = pattern
        err = C.create_string_buffer(512)
        self._h = lib().orc_compile(pattern, len(pattern), err, 512)
        if not self._h:
            raise OracleError(err.value.decode())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_free(h)

    @property
    def strategy(self):
        return lib().orc_strategy_name(self._h).decode()

    @property
    def strategy_exact(self):
        return bool(lib().orc_strategy_exact(self._h))

    @property
    def has_bidirectional(self):
        """UseDFA pattern for which the reference builds a reverse DFA (meta/compile.go:176-219)."""
        return bool(lib().orc_has_bidirectional(self._h))

    def set_bidirectional(self, on):
        """Run UseDFA searches through the restated forward + reverse lazy-DFA pair
        (meta/find_indices.go:686-705 over nfa/reverse.go) instead of the PikeVM restatement."""
        lib().orc_set_bidirectional(self._h, int(bool(on)))

    @property
    def num_captures(self):
        return lib().orc_num_captures(self._h)

    @property
    def digit_run_skip_safe(self):
        return bool(lib().orc_digit_run_skip_safe(self._h))

    def is_match(self, data):
        p, n, keep = _buf(data)
        return bool(lib().orc_is_match(self._h, p, n))

    def find_all(self, data, limit=-1):
        """-> int64 array of shape (count, 2)."""
        p, n, keep = _buf(data)
        cap = max(16, n // 2 + 2)
        while True:
            out = np.empty((cap, 2), dtype=np.int64)
            c = lib().orc_find_all(self._h, p, n, limit, out.ctypes.data, cap)
            if c <= cap:
                return out[:c].copy()
            cap = c

    def count(self, data, limit=-1):
        p, n, keep = _buf(data)
        return lib().orc_count(self._h, p, n, limit)

    def find_all_submatch(self, data, limit=-1):
        p, n, keep = _buf(data)
        stride = 2 * self.num_captures
        cap = max(16, n + 2)
        out = np.empty((cap, stride), dtype=np.int64)
        c = lib().orc_find_all_submatch(self._h, p, n, limit, out.ctypes.data, cap)
        return out[:c].copy()

    def find_at(self, data, at):
        p, n, keep = _buf(data)
        s, e = C.c_int64(), C.c_int64()
        if lib().orc_find_at(self._h, p, n, at, C.byref(s), C.byref(e)):
            return s.value, e.value
        return None

    def dump_nfa(self):
        return lib().orc_dump_nfa(self._h).decode()

    def prefixes(self):
        out = []
        for line in lib().orc_dump_prefixes(self._h).decode().splitlines():
            flag, _, hx = line.partition(" ")
            out.append((bytes.fromhex(hx), flag == "C"))
        return out


def dump_ast(pattern):
    if isinstance(pattern, str):
        pattern = pattern.encode()
    return lib().orc_dump_ast(pattern, len(pattern)).decode()


def scan_mt(pattern, data, threads, mode="findall"):
    """Sharded multi-thread scan (CPU baseline). -> (total_matches, seconds)."""
    if isinstance(pattern, str):
        pattern = pattern.encode()
    p, n, keep = _buf(data)
    sec = C.c_double()
    c = lib().orc_scan_mt(pattern, len(pattern), p, n, threads, 1 if mode == "count" else 0,
                          C.byref(sec))
    return c, sec.value
