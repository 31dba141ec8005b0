"""Host-side mirror of the reference's `simd` package (reference simd/memchr_amd64.go,
memchr_digit_amd64.go, memchr_class_amd64.go, memmem.go) over the device entry points of
csrc/memchr.cu: same names, same arguments, same return values (index of the first hit or -1).
The `*_device` functions take a device pointer and never copy; the plain ones copy the haystack to
the GPU first (tests, small inputs).  No CPU fallback: without a device they raise."""
import ctypes as C

import numpy as np

from . import _check, _lib

_vp = C.c_void_p
_lib.cgx_memchr_table_device.argtypes = [_vp, C.c_size_t, C.c_char_p, _vp, _vp]
_lib.cgx_memchr_table_at_device.argtypes = [_vp, C.c_size_t, C.c_size_t, C.c_char_p, _vp, _vp]
_lib.cgx_memchr_pair_device.argtypes = [_vp, C.c_size_t, C.c_uint8, C.c_uint8, C.c_int64, _vp, _vp]
_lib.cgx_memmem_device.argtypes = [_vp, C.c_size_t, C.c_char_p, C.c_size_t, _vp, _vp]


def _table(members):
    t = bytearray(256)
    for b in members:
        t[b] = 1
    return bytes(t)


_DIGIT = _table(range(ord("0"), ord("9") + 1))
_WORD = _table(list(range(48, 58)) + list(range(65, 91)) + list(range(97, 123)) + [95])
_NOT_WORD = bytes(1 - b for b in _WORD)


def _result(torch):
    return torch.empty(2, dtype=torch.int64, device="cuda")


def memchr_table_device(d_ptr, n, table256, stream=0):
    import torch
    res = _result(torch)
    _check(_lib.cgx_memchr_table_device(d_ptr, n, bytes(table256), res.data_ptr(), stream))
    return int(res[0].item())


def memchr_pair_device(d_ptr, n, byte1, byte2, offset, stream=0):
    import torch
    res = _result(torch)
    _check(_lib.cgx_memchr_pair_device(d_ptr, n, byte1, byte2, offset, res.data_ptr(), stream))
    return int(res[0].item())


def memmem_device(d_ptr, n, needle, stream=0):
    import torch
    res = _result(torch)
    _check(_lib.cgx_memmem_device(d_ptr, n, bytes(needle), len(needle), res.data_ptr(), stream))
    return int(res[0].item())


def _to_device(haystack):
    import torch
    a = np.frombuffer(bytes(haystack), dtype=np.uint8) if not isinstance(haystack, np.ndarray) else haystack
    t = torch.empty(a.size + 64, dtype=torch.uint8, device="cuda")
    if a.size:
        t[: a.size] = torch.from_numpy(a.copy()).cuda()
    return t, a.size


def Memchr(haystack, needle):  # reference simd/memchr_amd64.go:67
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _table([needle]))


def Memchr2(haystack, needle1, needle2):  # :114
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _table([needle1, needle2]))


def Memchr3(haystack, needle1, needle2, needle3):  # :159
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _table([needle1, needle2, needle3]))


def MemchrPair(haystack, byte1, byte2, offset):  # :202
    t, n = _to_device(haystack)
    return memchr_pair_device(t.data_ptr(), n, byte1, byte2, offset)


def MemchrDigit(haystack):  # reference simd/memchr_digit_amd64.go:17
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _DIGIT)


def MemchrDigitAt(haystack, at):  # :34 — absolute index of the first digit at or after `at`
    if at < 0 or at >= len(haystack):
        return -1
    import torch
    t, n = _to_device(haystack)
    res = _result(torch)
    _check(_lib.cgx_memchr_table_at_device(t.data_ptr(), n, at, _DIGIT, res.data_ptr(), 0))
    return int(res[0].item())


def MemchrWord(haystack):  # reference simd/memchr_class_amd64.go:35
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _WORD)


def MemchrNotWord(haystack):  # :58
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, _NOT_WORD)


def MemchrInTable(haystack, table):  # :76 — table: 256 truthy/falsy entries
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, bytes(1 if x else 0 for x in table))


def MemchrNotInTable(haystack, table):  # :90
    t, n = _to_device(haystack)
    return memchr_table_device(t.data_ptr(), n, bytes(0 if x else 1 for x in table))


def Memmem(haystack, needle):  # reference simd/memmem.go:53
    t, n = _to_device(haystack)
    return memmem_device(t.data_ptr(), n, needle)
