"""Builds libwol.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m waterorderlib_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwol.so")
SOURCES = ["wol_capi.cu", "wol_cells.cu", "wol_q3b.cu", "wol_q3b_tpc.cu", "wol_q3b_brick.cu", "wol_q3b_brick_ws.cu", "wol_q3b_tpc32.cu", "wol_aux.cu", "wol_slab.cu", "wol_pairs.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "--threads", "4", "-ldl"]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "wol_capi.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=True):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("WOL_NVCC_EXTRA", "").split()  # development switches, e.g. -DWOL_WS_PROF
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", LIB] + _sources()
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv)
