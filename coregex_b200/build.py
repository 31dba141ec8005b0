"""Builds libcoregex_b200.so in-tree (coregex_b200/lib/) with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the gpurun snapshot.  Usage: python -m coregex_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libcoregex_b200.so")

SOURCES = [
    "capi.cu",
    "scan_dfa.cu",
    "scan_flat.cu",
    "scan_teddy.cu",
    "pikevm_kernel.cu",
    "host/prog.cpp",
    "host/dfa.cpp",
    "host/analysis.cpp",
    "host/engine.cpp",
    "host/pike_pack.cpp",
]
EXTRA = [os.path.join(ROOT, "syntax", "parse.cpp")]
HEADERS_DIRS = [CSRC, os.path.join(CSRC, "host"), os.path.join(ROOT, "syntax"), os.path.join(ROOT, "include")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--cudart", "static",
    "-shared",
]


def _sources():
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return srcs + EXTRA


def _fingerprint():
    h = hashlib.sha256()
    files = list(_sources())
    for d in HEADERS_DIRS:
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                files.append(os.path.join(d, f))
    for f in sorted(set(files)):
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, ".stamp")
    fp = _fingerprint()
    if not force and os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read() == fp:
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIBDIR, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libcoregex_b200.so (see %s)" % log)
    if verbose:
        print(res.stderr)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
