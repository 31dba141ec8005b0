"""ctypes binding to the CPU SIMT build of the bitstream kernel (tests/sim/).

TEST INFRASTRUCTURE ONLY: the kernel source compiled for a fiber-based warp emulator so that its
warp-level logic can be checked against the oracle in a container without a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "tests", "sim")
_SO = os.path.join(_DIR, "_build", "libcgxsim.so")
_lib = None
last_diag = {}


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR], stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        L.cgxsim_scan.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                  C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.c_uint, C.c_int, C.c_int]
        _lib = L
    return _lib


class NotEligible(Exception):
    pass


_jit_libs = {}


def jit_lib(pattern, tiles=0, defs="", tag=""):  # `tiles` = search mode the kernel is specialised for
    """The emulated kernel built the way jit.cu builds it for the device: -DCGX_JIT plus the
    generated cgx_jit_prog.h of this pattern (straight-line passes)."""
    if isinstance(pattern, str):
        pattern = pattern.encode()
    if (pattern, tiles, tag) in _jit_libs:
        return _jit_libs[(pattern, tiles, tag)]
    import hashlib
    L = lib()
    buf = C.create_string_buffer(1 << 16)
    L.cgxsim_jit_header.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    r = L.cgxsim_jit_header(pattern, len(pattern), buf, len(buf))
    if r == -2:
        raise NotEligible(pattern)
    assert r > 0, r
    d = os.path.join(_DIR, "_build", "jit_" + hashlib.sha1(buf.value).hexdigest()[:12])
    os.makedirs(d, exist_ok=True)
    hp = os.path.join(d, "cgx_jit_prog.h")
    if not os.path.exists(hp) or open(hp, "rb").read() != buf.value:
        open(hp, "wb").write(buf.value)
    subprocess.check_call(["make", "-C", _DIR, "jit", "JITDIR=" + d, "MODE=%d" % tiles, "DEFS=" + defs, "TAG=" + tag],
                          stdout=subprocess.DEVNULL)
    J = C.CDLL(os.path.join(d, "libcgxsim_jit%d%s.so" % (tiles, tag)))
    J.cgxsim_scan.argtypes = L.cgxsim_scan.argtypes
    _jit_libs[(pattern, tiles, tag)] = J
    return J


def scan(pattern, hay, mode=0, cap=None, grid=2, base=0, pad=ord("1"), jit=False, tiles=None, defs="", tag="",
         launches=1):
    """Runs the emulated kernel.  Returns (total, flag, pairs[ndarray k x 2])."""
    if isinstance(pattern, str):
        pattern = pattern.encode()
    a = np.frombuffer(bytes(hay), dtype=np.uint8) if not isinstance(hay, np.ndarray) else np.ascontiguousarray(hay)
    n = a.size
    if cap is None:
        cap = n + 16
    out = np.full((max(cap, 1), 2), -7, dtype=np.int64)
    res = (C.c_uint64 * 4)()
    r = (jit_lib(pattern, mode, defs, tag) if jit else lib()).cgxsim_scan(pattern, len(pattern), a.ctypes.data if n else None, n, base, mode,
                          out.ctypes.data, cap, res, grid, pad, launches)
    if r == -2:
        raise NotEligible(pattern)
    if r == -1:
        raise RuntimeError("compile failed: %r" % pattern)
    if r != 0:
        raise RuntimeError("emulated launch left dirty scratch behind (code %d): %r" % (r, pattern))
    total = int(res[0])
    global last_diag
    last_diag = {"serial": int(res[2]), "redo": int(res[3])}
    return total, int(res[1]), out[: min(total, cap)]


_teddy = None


def teddy_lib():
    global _teddy
    if _teddy is None:
        subprocess.check_call(["make", "-C", _DIR, "teddy"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_DIR, "_build", "libcgxsim_teddy.so"))
        L.cgxsim_scan_teddy.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                        C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.c_uint, C.c_int, C.c_int]
        _teddy = L
    return _teddy


def scan_teddy(pattern, hay, mode=0, cap=None, grid=2, base=0, after=0, pad=ord("e"), launches=2):
    """The multi-literal flavour of the kernel (scan_teddy.cu) on the emulator: (total, flag, pairs)."""
    if isinstance(pattern, str):
        pattern = pattern.encode()
    a = np.frombuffer(bytes(hay), dtype=np.uint8) if not isinstance(hay, np.ndarray) else np.ascontiguousarray(hay)
    n = a.size
    if cap is None:
        cap = n + 16
    out = np.full((max(cap, 1), 2), -7, dtype=np.int64)
    res = (C.c_uint64 * 4)()
    r = teddy_lib().cgxsim_scan_teddy(pattern, len(pattern), a.ctypes.data if n else None, n, base, after, mode,
                                      out.ctypes.data, cap, res, grid, pad, launches)
    if r == -2:
        raise NotEligible(pattern)
    if r == -1:
        raise RuntimeError("compile failed: %r" % pattern)
    if r != 0:
        raise RuntimeError("emulated launch left dirty scratch behind (code %d): %r" % (r, pattern))
    total = int(res[0])
    return total, int(res[1]), out[: min(total, cap)]
