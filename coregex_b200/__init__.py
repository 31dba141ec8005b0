"""coregex_b200 — Python binding of the B200 bulk-scan engine (libcoregex_b200.so).

The method names and semantics follow the reference's public Go API for the bulk-scan path
(reference regex.go): Compile :110, MustCompile :129, (*Regex).Match :282, FindAllIndex :695,
Count :1349, FindAllSubmatchIndex :1423, NumSubexp :552, String :444 — so the parity tests read
like the reference's own tests.  Everything runs on the GPU through the C ABI declared in
include/coregex_b200.h; there is no CPU fallback: importing works anywhere, searching without a
CUDA device raises NoDeviceError, and a missing shared library raises at import time.
"""
import ctypes as C
import os

import numpy as np

from .wrappers import RegexWrappers, quote_meta as QuoteMeta  # noqa: E402  (reference regex.go:233)

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libcoregex_b200.so")

CGX_OK = 0
CGX_ERR_SYNTAX = -1
CGX_ERR_UNSUPPORTED = -2
CGX_ERR_NO_DEVICE = -3
CGX_ERR_CUDA = -4
CGX_ERR_ARGS = -5
CGX_ERR_NOMEM = -6
CGX_ERR_CONFIG = -7

MODE_FINDALL, MODE_COUNT, MODE_ISMATCH = 0, 1, 2


class Error(Exception):
    """Compile error; str() is the Go-formatted message (reference meta/compile.go:775-784)."""


class UnsupportedError(Error):
    """Valid pattern outside the GPU engines' current scope."""


class ConfigError(Error):
    """Invalid Config; str() is the reference's message (meta/config.go:179-181)."""


class Config(C.Structure):
    """meta.Config (reference meta/config.go:31-113), field for field; DefaultConfig() fills the
    reference's defaults."""
    _fields_ = [("EnableDFA", C.c_int), ("EnablePrefilter", C.c_int), ("MaxDFAStates", C.c_uint32),
                ("DeterminizationLimit", C.c_int), ("MinLiteralLen", C.c_int), ("MaxLiterals", C.c_int),
                ("MaxRecursionDepth", C.c_int), ("EnableASCIIOptimization", C.c_int)]

    def Validate(self):
        """None, or the reference's error text (Config.Validate, meta/config.go:132-170)."""
        err = C.create_string_buffer(256)
        rc = _lib.cgx_config_validate(C.byref(self), err, 256)
        return None if rc == CGX_OK else err.value.decode()


def DefaultConfig():
    """reference meta/config.go:101-112"""
    c = Config()
    _lib.cgx_default_config(C.byref(c))
    return c


class NoDeviceError(RuntimeError):
    pass


class CudaError(RuntimeError):
    pass


def _load():
    global _SO
    _SO = os.environ.get("COREGEX_B200_LIB", _SO)  # e.g. the -DCGX_TIMING build of tools/phase_timing.py
    if not os.path.exists(_SO):
        raise ImportError(
            "coregex_b200: %s is missing — run `python -m coregex_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % _SO)
    L = C.CDLL(_SO)
    vp, u8p, i64, sz = C.c_void_p, C.c_void_p, C.c_int64, C.c_size_t
    L.cgx_compile.argtypes = [C.c_char_p, sz, C.POINTER(vp), C.c_char_p, sz]
    L.cgx_free.argtypes = [vp]
    L.cgx_compile_cfg.argtypes = [C.c_char_p, sz, C.c_void_p, C.POINTER(vp), C.c_char_p, sz]
    L.cgx_default_config.argtypes = [C.c_void_p]
    L.cgx_config_validate.argtypes = [C.c_void_p, C.c_char_p, sz]
    L.cgx_set_longest.argtypes = [vp, C.c_int]
    L.cgx_strategy.restype = C.c_char_p
    L.cgx_strategy.argtypes = [vp]
    L.cgx_engine.restype = C.c_char_p
    L.cgx_engine.argtypes = [vp]
    L.cgx_num_captures.argtypes = [vp]
    L.cgx_subexp_name.restype = C.c_char_p
    L.cgx_subexp_name.argtypes = [vp, C.c_int]
    L.cgx_delimiter.argtypes = [vp]
    L.cgx_debug_set_bitstream.argtypes = [vp, C.c_int]
    L.cgx_last_error.restype = C.c_char_p
    L.cgx_is_match.argtypes = [vp, u8p, sz, C.POINTER(C.c_int)]
    L.cgx_find_all_index.argtypes = [vp, u8p, sz, i64, C.c_void_p, sz, C.POINTER(sz)]
    L.cgx_count.argtypes = [vp, u8p, sz, i64, C.POINTER(sz)]
    L.cgx_find_all_submatch_index.argtypes = [vp, u8p, sz, i64, C.c_void_p, sz, C.POINTER(sz)]
    L.cgx_scan_device.argtypes = [vp, u8p, sz, i64, C.c_int, C.c_void_p, sz, C.c_void_p, C.c_void_p]
    L.cgx_scan_shard_device.argtypes = [vp, u8p, sz, i64, i64, C.c_int, C.c_void_p, sz, C.c_void_p, C.c_void_p]
    L.cgx_scan_submatch_device.argtypes = [vp, u8p, sz, i64, C.c_void_p, sz, C.c_void_p, C.c_void_p]
    L.cgx_scan_submatch_shard_device.argtypes = [vp, u8p, sz, i64, i64, C.c_void_p, sz, C.c_void_p, C.c_void_p]
    L.cgx_scan_records_device.argtypes = [vp, u8p, sz, C.c_void_p, sz, i64, C.c_void_p, sz, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
    L.cgx_wire_bytes.restype = sz
    L.cgx_wire_bytes.argtypes = [sz, C.c_int]
    L.cgx_wire_segments.argtypes = [sz]
    L.cgx_pack_offsets_device.argtypes = [C.c_void_p, sz, i64, sz, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cgx_unpack_offsets_device.argtypes = [C.c_void_p, sz, i64, sz, C.c_void_p, C.c_void_p]
    L.cgx_launch_count.restype = C.c_uint64
    L.cgx_launch_count.argtypes = [vp]
    L.cgx_synth_device.argtypes = [C.c_int, C.c_uint64, C.c_uint64, u8p, sz, u8p, C.c_void_p, C.c_int, C.c_void_p]
    L.cgx_synth_host.argtypes = [C.c_int, C.c_uint64, C.c_uint64, u8p, sz, u8p, C.c_void_p, C.c_int]
    return L


_lib = _load()


def _check(rc):
    if rc == CGX_OK:
        return
    msg = (_lib.cgx_last_error() or b"").decode(errors="replace")
    if rc == CGX_ERR_NO_DEVICE:
        raise NoDeviceError(msg or "no CUDA device (coregex_b200 has no CPU fallback)")
    if rc == CGX_ERR_UNSUPPORTED:
        raise UnsupportedError(msg)
    if rc == CGX_ERR_ARGS:
        raise ValueError(msg or "bad arguments")
    raise CudaError("status %d: %s" % (rc, msg))


def _host_buf(b):
    """bytes-like / numpy uint8 / torch CPU uint8 tensor -> (ptr, len, keepalive)."""
    if isinstance(b, str):
        b = b.encode()
    if hasattr(b, "data_ptr") and hasattr(b, "numel"):  # torch tensor (pinned or pageable, CPU)
        return b.data_ptr(), b.numel(), b
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    a = np.frombuffer(bytes(b), dtype=np.uint8)
    return a.ctypes.data, a.size, a


class Regex(RegexWrappers):
    """A compiled pattern.  Mirrors reference regex.go `type Regex`: the searches below run on the
    device; the text / replace / split forms (wrappers.py) are host plumbing over their results."""

    def __init__(self, pattern, config=None):
        self._config = config
        if isinstance(pattern, str):
            pattern = pattern.encode()
        self._pattern = pattern
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        rc = _lib.cgx_compile_cfg(pattern, len(pattern), C.byref(config) if config is not None else None,
                                  C.byref(h), err, 1024)
        if rc == CGX_ERR_CONFIG:
            raise ConfigError(err.value.decode(errors="replace"))
        if rc == CGX_ERR_SYNTAX:
            raise Error(err.value.decode(errors="replace"))
        if rc == CGX_ERR_UNSUPPORTED:
            raise UnsupportedError(err.value.decode(errors="replace"))
        _check(rc)
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.cgx_free(h)

    def Longest(self):
        """reference regex.go:464: leftmost-longest matching for all later searches."""
        _check(_lib.cgx_set_longest(self._h, 1))
        self._longest = True

    def _subexp_name(self, i):
        return _lib.cgx_subexp_name(self._h, i).decode()

    # -- introspection --------------------------------------------------------------------------
    def String(self):
        return self._pattern.decode(errors="replace")

    def NumSubexp(self):
        return _lib.cgx_num_captures(self._h) - 1

    @property
    def strategy(self):
        """Strategy the reference would select (meta.Engine.Strategy)."""
        return _lib.cgx_strategy(self._h).decode()

    @property
    def engine(self):
        return _lib.cgx_engine(self._h).decode()

    @property
    def delimiter(self):
        """Record delimiter byte of this pattern (no match can contain it); b"\\n" when possible.
        None: matches may contain every byte value (the haystack is scanned as one record)."""
        d = _lib.cgx_delimiter(self._h)
        return None if d < 0 else bytes([d])

    def set_bitstream(self, on):
        """Tests / A-B runs: keep a flat deterministic pattern on the candidate+DFA kernel
        (on=False) instead of the bitstream kernel.  Returns the previous setting."""
        return bool(_lib.cgx_debug_set_bitstream(self._h, int(on)))

    @property
    def launches(self):
        return _lib.cgx_launch_count(self._h)

    # -- host-buffer API (drop-in for the Go methods) ---------------------------------------------
    def Match(self, b):
        p, n, keep = _host_buf(b)
        m = C.c_int(0)
        _check(_lib.cgx_is_match(self._h, p, n, C.byref(m)))
        return bool(m.value)

    def Count(self, b, n=-1):
        p, ln, keep = _host_buf(b)
        c = C.c_size_t(0)
        _check(_lib.cgx_count(self._h, p, ln, n, C.byref(c)))
        return c.value

    def find_all_index_array(self, b, n=-1, cap=None):
        """FindAllIndex returning an int64 array of shape (count, 2) (no Python list building)."""
        p, ln, keep = _host_buf(b)
        if n == 0:
            return np.empty((0, 2), dtype=np.int64)
        if cap is None:
            cap = max(64, ln // 64)
        while True:
            out = np.empty((cap, 2), dtype=np.int64)
            c = C.c_size_t(0)
            _check(_lib.cgx_find_all_index(self._h, p, ln, n, out.ctypes.data, cap, C.byref(c)))
            if c.value <= cap:
                return out[: c.value]
            cap = c.value

    def FindAllIndex(self, b, n=-1):
        """[[start, end], ...] or None when there is no match / n == 0 (reference regex.go:695-723)."""
        a = self.find_all_index_array(b, n)
        if a.shape[0] == 0:
            return None
        return a.tolist()

    FindAllStringIndex = FindAllIndex

    def FindAllSubmatchIndex(self, b, n=-1):
        p, ln, keep = _host_buf(b)
        if n == 0:
            return None
        stride = 2 * _lib.cgx_num_captures(self._h)
        cap = max(64, ln // 32)
        while True:
            out = np.empty((cap, stride), dtype=np.int64)
            c = C.c_size_t(0)
            _check(_lib.cgx_find_all_submatch_index(self._h, p, ln, n, out.ctypes.data, cap, C.byref(c)))
            if c.value <= cap:
                break
            cap = c.value
        if c.value == 0:
            return None
        return out[: c.value].tolist()

    def find_all_into(self, b, out, n=-1, submatch=False):
        """FindAllIndex / FindAllSubmatchIndex of a host buffer into a caller-owned int64 buffer
        (numpy array or CPU torch tensor, pinned for full PCIe speed) of shape (cap, 2) or
        (cap, 2 * (NumSubexp() + 1)); nothing is allocated per call.  Returns the number of matches,
        which may exceed cap (the first cap rows are written) — the buffered-append forms of the
        reference (regex.go AppendAllIndex family) in one call."""
        p, ln, keep = _host_buf(b)
        if n == 0:
            return 0
        stride = 2 * _lib.cgx_num_captures(self._h) if submatch else 2
        optr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        cap = (out.numel() if hasattr(out, "numel") else out.size) // stride
        c = C.c_size_t(0)
        fn = _lib.cgx_find_all_submatch_index if submatch else _lib.cgx_find_all_index
        _check(fn(self._h, p, ln, n, optr, cap, C.byref(c)))
        return c.value

    # -- device-resident API -----------------------------------------------------------------------
    def scan_device(self, d_ptr, length, mode=MODE_FINDALL, out_ptr=0, cap_pairs=0, result_ptr=0,
                    base_offset=0, stream=0, bytes_after=0):
        """Enqueue a scan of device memory [d_ptr, d_ptr+length).  Pointers are raw ints.
        base_offset / bytes_after place the buffer inside a larger logical haystack (a shard)."""
        _check(_lib.cgx_scan_shard_device(self._h, d_ptr, length, base_offset, bytes_after, mode, out_ptr,
                                          cap_pairs, result_ptr, stream))

    def scan_records_device(self, d_ptr, length, rec_off_ptr, nrec, out_ptr, cap_pairs, rec_prefix_ptr,
                            result_ptr, base_offset=0, stream=0):
        """Batch of independent delimiter-terminated records in one scan (include/coregex_b200.h
        cgx_scan_records_device): pairs in global order + per-record pair-index prefix."""
        _check(_lib.cgx_scan_records_device(self._h, d_ptr, length, rec_off_ptr, nrec, base_offset, out_ptr,
                                            cap_pairs, rec_prefix_ptr, result_ptr, stream))

    def scan_submatch_device(self, d_ptr, length, out_ptr, cap_matches, result_ptr, base_offset=0,
                             stream=0, bytes_after=0):
        _check(_lib.cgx_scan_submatch_shard_device(self._h, d_ptr, length, base_offset, bytes_after, out_ptr,
                                                   cap_matches, result_ptr, stream))


def Compile(pattern):
    """reference regex.go:110"""
    return Regex(pattern)


def CompileWithConfig(pattern, config):
    """reference regex.go:198"""
    return Regex(pattern, config)


def MustCompile(pattern):
    """reference regex.go:129 — panics (raises) with the reference's message format"""
    try:
        return Regex(pattern)
    except Error as e:
        raise Error("regexp: Compile(`%s`): %s" % (pattern if isinstance(pattern, str) else pattern.decode(), e))


def CompilePOSIX(pattern):
    """reference regex.go:146: Compile + Longest()"""
    re = Regex(pattern)
    re.Longest()
    return re


def MustCompilePOSIX(pattern):
    """reference regex.go:159"""
    try:
        return CompilePOSIX(pattern)
    except Error as e:
        raise Error("regexp: CompilePOSIX(`%s`): %s" % (pattern if isinstance(pattern, str) else pattern.decode(), e))


def MatchReader(pattern, reader):
    """reference regex.go:1672"""
    return Regex(pattern).MatchReader(reader)


def Match(pattern, b):
    """reference regex.go:170: compile + Match in one call"""
    return Regex(pattern).Match(b)


def MatchString(pattern, s):
    """reference regex.go:181"""
    return Regex(pattern).MatchString(s)


# ---- synthetic corpora (bench / tests) ----------------------------------------------------------
SYNTH_LOG, SYNTH_TEXT, SYNTH_EMAIL = 0, 1, 2
SYNTH_BLOCK = {0: 4096, 1: 4096, 2: 80}


def _pack_literals(literals):
    if not literals:
        return None, None, 0
    blob = b"".join(literals)
    offs = np.zeros(len(literals) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(x) for x in literals])
    return np.frombuffer(blob, dtype=np.uint8).copy(), offs, len(literals)


def synth_host(kind, seed, nbytes, first_block=0, literals=None):
    """Deterministic synthetic corpus slice as a numpy uint8 array (host twin of synth_device)."""
    out = np.empty(nbytes, dtype=np.uint8)
    blob, offs, nlit = _pack_literals(literals)
    _check(_lib.cgx_synth_host(kind, seed, first_block, out.ctypes.data, nbytes,
                               blob.ctypes.data if blob is not None else None,
                               offs.ctypes.data if offs is not None else None, nlit))
    return out


def synth_device(kind, seed, d_ptr, nbytes, first_block=0, d_lit_ptr=0, d_off_ptr=0, nlit=0, stream=0):
    _check(_lib.cgx_synth_device(kind, seed, first_block, d_ptr, nbytes, d_lit_ptr, d_off_ptr, nlit, stream))


# ---- compact offset wire format (multi-GPU offset gather) ------------------------------------------
def wire_segments(shard_len):
    return _lib.cgx_wire_segments(shard_len)


def wire_bytes(count, shard_len):
    return _lib.cgx_wire_bytes(count, _lib.cgx_wire_segments(shard_len))


def pack_offsets_device(pairs_ptr, count, shard_base, shard_len, wire_ptr, bad_ptr, stream=0):
    _check(_lib.cgx_pack_offsets_device(pairs_ptr, count, shard_base, shard_len, wire_ptr, bad_ptr, stream))


def unpack_offsets_device(wire_ptr, count, shard_base, shard_len, out_ptr, stream=0):
    _check(_lib.cgx_unpack_offsets_device(wire_ptr, count, shard_base, shard_len, out_ptr, stream))
