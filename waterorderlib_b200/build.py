"""Builds libwol.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m waterorderlib_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwol.so")
STAMP = LIB + ".srchash"  # hash of the sources and flags the library was built from
SOURCES = ["wol_capi.cu", "wol_cells.cu", "wol_q3b.cu", "wol_q3b_tpc.cu", "wol_q3b_brick.cu", "wol_q3b_brick_ws.cu", "wol_q3b_brick32.cu", "wol_q3b_tpc32.cu", "wol_aux.cu", "wol_slab.cu", "wol_pairs.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "--threads", "4", "-ldl"]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _source_hash(extra):
    """sha256 over the flags and every file the library is built from (an mtime says nothing about a .so that
    travelled with a snapshot of the tree)."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "wol_capi.h")]
    for d in deps:
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _stale(extra):
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != _source_hash(extra)


def build_lib(force=False, verbose=True):
    extra = os.environ.get("WOL_NVCC_EXTRA", "").split()  # development switches, e.g. -DWOL_WS_PROF
    if not force and not _stale(extra):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + _sources()
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as fh:
        fh.write(_source_hash(extra) + "\n")
    return LIB


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv)
