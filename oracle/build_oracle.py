"""TEST INFRASTRUCTURE ONLY -- builds the oracle's native pieces.

  oracle/libwol_oracle.so      the C restatement (oracle/wol_oracle.c), OpenMP, no FMA contraction
  oracle/_ref/libgfortran.so.3 the stub runtime the reference's prebuilt f2py module needs
  oracle/_ref/{waterlib,sortlib}.cpython-37m-x86_64-linux-gnu.so   staged only while /root/reference is visible
        (build container): the reference's own prebuilt BINARY of fortran/waterlib.f90 (gfortran is not
        in this image, so it cannot be recompiled), so that the reference's compiled arithmetic can also
        run on a GPU box where /root/reference does not exist.  No reference source text is staged.
        oracle/_ref/ is git-ignored (never enters history) but travels with the gpurun snapshot.

Run: python oracle/build_oracle.py
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"
STAGED = (
    (os.path.join("fortran", "waterlib.cpython-37m-x86_64-linux-gnu.so"), "waterlib.cpython-37m-x86_64-linux-gnu.so"),
    (os.path.join("fortran", "sortlib.cpython-37m-x86_64-linux-gnu.so"), "sortlib.cpython-37m-x86_64-linux-gnu.so"),
)


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(verbose=True):
    os.makedirs(REF, exist_ok=True)
    src = os.path.join(HERE, "wol_oracle.c")
    out = os.path.join(HERE, "libwol_oracle.so")
    if _newer(src, out):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
               "-Wall", "-o", out, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    stub_src = os.path.join(HERE, "gfortran_stub.c")
    stub = os.path.join(REF, "libgfortran.so.3")
    if _newer(stub_src, stub):
        cmd = ["gcc", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-o", stub, stub_src,
               "-Wl,--version-script=" + os.path.join(HERE, "gfortran_stub.map"),
               "-Wl,-soname,libgfortran.so.3"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if os.path.isdir(REFERENCE_ROOT):
        for rel, name in STAGED:
            s = os.path.join(REFERENCE_ROOT, rel)
            d = os.path.join(REF, name)
            if os.path.exists(s) and _newer(s, d):
                shutil.copyfile(s, d)
                if verbose:
                    print("staged", s, "->", d)
    return out


if __name__ == "__main__":
    build()
    sys.exit(0)
