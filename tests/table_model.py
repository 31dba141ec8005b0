"""CPU model of what scan_dfa.cu computes, driven by the product's own compiled tables.

This is NOT the oracle and not a fallback: it replays the candidate-filter + anchored-walk +
chain algorithm of the kernel in numpy/Python over the tables exported by the debug entry points,
so the host compiler (parser -> program -> eager DFA -> filter choice) can be checked against the
oracle in the CPU-only test tier.  The kernel's parallel decomposition (chunk ownership, batches,
look-back) is only exercised by the -m gpu tests.
"""
import ctypes as C

import numpy as np

import coregex_b200 as cg

_L = cg._lib
_L.cgx_debug_dfa_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
_L.cgx_debug_dfa_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def _kind(b):
    if b == 10:
        return 3
    if b == 13:
        return 4
    c = chr(b)
    return 1 if (c.isalnum() and b < 128) or c == "_" else 0


class TableModel:
    def __init__(self, regex):
        ns, fk, ss, kl, nr = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        start = np.zeros(5, dtype=np.uint16)
        rng = np.zeros(8, dtype=np.uint8)
        ok = _L.cgx_debug_dfa_info(regex._h, C.byref(ns), start.ctypes.data, C.byref(fk), C.byref(ss),
                                   C.byref(kl), rng.ctypes.data, C.byref(nr))
        assert ok == 1, "not a DFA-engine pattern"
        self.nstates, self.filter_kind, self.skip_safe = ns.value, fk.value, bool(ss.value)
        self.kind_lut, self.start = bool(kl.value), start
        self.ranges = [(int(rng[2 * k]), int(rng[2 * k + 1])) for k in range(nr.value)]
        self.trans = np.zeros(self.nstates * 256, dtype=np.uint16)
        self.eoi = np.zeros(self.nstates, dtype=np.uint8)
        self.lut = np.zeros(256, dtype=np.uint8)
        _L.cgx_debug_dfa_copy(regex._h, self.trans.ctypes.data, self.eoi.ctypes.data, self.lut.ctypes.data)

    def in_set(self, b):
        if self.filter_kind == 2:
            return bool(self.lut[b])
        return any(lo <= b <= hi for lo, hi in self.ranges)

    def walk(self, h, p0):
        n = len(h)
        s = int(self.start[0])
        if self.kind_lut:
            s = int(self.start[2 if p0 == 0 else _kind(h[p0 - 1])])
        last, p = -1, p0
        while s:
            if p >= n:
                if self.eoi[s]:
                    last = n
                break
            e = int(self.trans[s * 256 + h[p]])
            if e & 0x8000:
                last = p
            s = e & 0x7FFF
            p += 1
        return last

    def find_all(self, h):
        """reference findAllIndicesLoop + digit-prefilter style candidate loop over the tables"""
        h = bytes(h)
        n, pos, out = len(h), 0, []
        while pos < n:
            d = pos
            while d < n and not self.in_set(h[d]):
                d += 1
            if d >= n:
                break
            e = self.walk(h, d)
            if e >= 0:
                out.append([d, e])
                pos = e if e > d else d + 1
            else:
                pos = d + 1
                if self.filter_kind == 0:
                    while pos < n and self.in_set(h[pos]):
                        pos += 1
        return out


_L.cgx_debug_flat.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]


class FlatModel:
    """Exact (whole-haystack, no tile edges) evaluation of the kernel's bit-parallel start filter
    with Python big integers: bit p of the result <=> the flat program can match starting at p."""

    def __init__(self, regex):
        ops = np.zeros(48, dtype=np.uint8)
        nr = np.zeros(4, dtype=np.uint8)
        rg = np.zeros(32, dtype=np.uint8)
        nc = C.c_int()
        self.nops = _L.cgx_debug_flat(regex._h, ops.ctypes.data, C.byref(nc), nr.ctypes.data, rg.ctypes.data)
        self.ops = [(int(ops[2 * i]), int(ops[2 * i + 1])) for i in range(self.nops)]
        self.classes = [[(int(rg[(c * 4 + r) * 2]), int(rg[(c * 4 + r) * 2 + 1])) for r in range(nr[c])]
                        for c in range(nc.value)]

    def start_set(self, h):
        n = len(h)
        full = (1 << (n + 1)) - 1          # positions 0..n (n = end of input)
        cm = []
        for ranges in self.classes:
            m = 0
            for p, b in enumerate(h):
                if any(lo <= b <= hi for lo, hi in ranges):
                    m |= 1 << p
            cm.append(m)
        M = full
        for kind, cls in reversed(self.ops):
            Cm = cm[cls]
            step = (M >> 1) & Cm           # p in class and p+1 in M
            if kind == 0:
                M = step
            elif kind == 3:
                M = M | step
            else:
                # propagate leftwards through runs of the class
                plus = step
                while True:
                    nxt = plus | ((plus >> 1) & Cm)
                    if nxt == plus:
                        break
                    plus = nxt
                M = plus if kind == 1 else (M | plus)
        return M


_L.cgx_debug_line_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
_L.cgx_debug_line_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


class LineModel:
    """CPU replay of the record engine (one lane per record in scan_dfa.cu): unanchored forward DFA
    for the leftmost-first end, reverse DFA (last flag wins) for the start, resume at the end."""

    def __init__(self, regex):
        un, rn = C.c_int(), C.c_int()
        us, rs = np.zeros(5, dtype=np.uint16), np.zeros(5, dtype=np.uint16)
        ok = _L.cgx_debug_line_info(regex._h, C.byref(un), C.byref(rn), us.ctypes.data, rs.ctypes.data)
        assert ok == 1, "not a record-engine pattern"
        self.un, self.rn, self.us, self.rs = un.value, rn.value, us, rs
        self.delim = regex.delimiter[0]
        self.ut = np.zeros(self.un * 256, dtype=np.uint16)
        self.rt = np.zeros(self.rn * 256, dtype=np.uint16)
        self.ueoi = np.zeros(self.un, dtype=np.uint8)
        self.reoi = np.zeros(self.rn, dtype=np.uint8)
        _L.cgx_debug_line_copy(regex._h, self.ut.ctypes.data, self.rt.ctypes.data, self.ueoi.ctypes.data,
                               self.reoi.ctypes.data)

    def _reverse(self, h, e, lo):
        n = len(h)
        st = int(self.rs[2 if e >= n else _kind(h[e])])
        last, q = lo, e
        while st:
            if q == lo:
                if q == 0:
                    if self.reoi[st]:
                        last = q
                elif int(self.rt[st * 256 + h[q - 1]]) & 0x8000:
                    last = q
                break
            t = int(self.rt[st * 256 + h[q - 1]])
            if t & 0x8000:
                last = q
            st = t & 0x7FFF
            q -= 1
        return last

    def find_all(self, h, delim=None):
        h = bytes(h)
        delim = self.delim if delim is None else delim
        n, out = len(h), []
        line = 0
        while line < n:
            pos = line
            line_end = None
            while True:
                st = int(self.us[2 if pos == 0 else _kind(h[pos - 1])])
                last, i = -1, pos
                while st:
                    if i >= n:
                        if self.ueoi[st]:
                            last = i
                        line_end = n
                        break
                    b = h[i]
                    t = int(self.ut[st * 256 + b])
                    if t & 0x8000:
                        last = i
                    st = t & 0x7FFF
                    if b == delim:
                        line_end = i + 1
                        break
                    i += 1
                if last < 0:
                    break
                out.append([self._reverse(h, last, pos), last])
                pos = last
            if line_end is None:  # the scan died before the delimiter: find the record's end
                j = h.find(bytes([delim]), pos)
                line_end = n if j < 0 else j + 1
            line = line_end
        return out
