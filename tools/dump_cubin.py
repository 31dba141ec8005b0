#!/usr/bin/env python
"""Builds the NVRTC-specialised bitstream kernel for a pattern here (no GPU needed) and writes the
cubin: python tools/dump_cubin.py <out.cubin> [pattern]   (CGX_JIT_DEFS adds -D options)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coregex_b200 as cg

pat = sys.argv[2] if len(sys.argv) > 2 else r"\d+\.\d+\.\d+\.\d+"
r = cg.Compile(pat)
cg._lib.cgx_debug_jit_compile.restype = C.c_long
cg._lib.cgx_debug_jit_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
buf = C.create_string_buffer(8 << 20)
n = cg._lib.cgx_debug_jit_compile(r._h, buf, len(buf))
assert n > 0
open(sys.argv[1], "wb").write(buf.raw[:n])
print(sys.argv[1], n)
