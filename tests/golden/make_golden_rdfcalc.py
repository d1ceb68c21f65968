"""Golden fixture for the rdfCalc driver, generated from the LIVE reference (build container only).

    python tests/golden/make_golden_rdfcalc.py

Runs the reference's own, unmodified ``rdfCalc`` body (AST-extracted from structureLibs/orderParam_lib.py:575-727)
over a small synthetic trajectory, with
  * ``wl``         = the reference's compiled Fortran through ctypes (oracle/ref_fortran.py),
  * ``TrajObject`` = a duck-typed stand-in (frames with .xyz / .box.values, index LISTS -- the reference's
                     ``solInds==[]`` tests only work on lists under NumPy 2),
  * ``argrelmin``  = scipy.signal.argrelmin, as the reference imports it,
  * ``simps``      = the composite Simpson rule of SciPy < 1.11 with even='avg' (what ``scipy.integrate.simps`` was when
                     the reference was written; the name no longer exists in the installed SciPy), restated below.
Stores the inputs, the return values and the two tables the function writes.
"""
import ast
import os
import sys
import tempfile

import numpy as np
from scipy.signal import argrelmin

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_fortran  # noqa: E402
from waterorderlib_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def legacy_simps(y, x):
    """scipy.integrate.simps(y, x) of SciPy < 1.11 (even='avg'), from its documentation: Simpson over consecutive pairs
    of intervals (unequal spacing formula); for an even sample count the average of "first N-2 intervals + trapezoid
    on the last" and "trapezoid on the first + last N-2 intervals"."""
    from scipy.integrate import simpson  # identical for odd sample counts

    y, x = np.asarray(y, float), np.asarray(x, float)
    n = len(y)
    if n % 2 == 1:
        return simpson(y, x=x)
    last = 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    first = 0.5 * (x[1] - x[0]) * (y[1] + y[0])
    a = (simpson(y[:-1], x=x[:-1]) if n > 2 else 0.0) + last
    b = (simpson(y[1:], x=x[1:]) if n > 2 else 0.0) + first
    return 0.5 * (a + b)


class _Box:
    def __init__(self, v):
        self.values = np.concatenate([np.asarray(v, float), [90.0, 90.0, 90.0]])


class _Frame:
    def __init__(self, xyz, box):
        self.xyz, self.box = xyz, _Box(box)


def make_inputs(n_frames=10, n_sol=8):
    rng = np.random.default_rng(77)
    frames, boxes = [], []
    for f in range(n_frames):
        wat, box = synth.water_box(4, sigma=0.5, seed=300 + f)
        sol = (0.5 * box + rng.uniform(-3.0, 3.0, size=(n_sol, 3))).astype(np.float32).astype(np.float64)
        frames.append(np.concatenate([sol, wat]))
        boxes.append(box)
    return np.stack(frames), np.stack(boxes), n_sol


def main():
    xyz, boxes, n_sol = make_inputs()
    n_atoms = xyz.shape[1]
    sol_inds, wat_inds = list(range(n_sol)), list(range(n_sol, n_atoms))

    class StubTrajObject:
        def __init__(self, topFile, trajFile, stride, solResName, watResName):
            self.top, self.traj = topFile, trajFile

        def getWatInds(self):
            return wat_inds, [], 1

        def getSolInds(self):
            return self.sol, [], [], [], [], []

    src = "/root/reference/structureLibs/orderParam_lib.py"
    with open(src) as fh:
        tree = ast.parse(fh.read())
    node = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "rdfCalc"][0]
    ns = {"np": np, "wl": ref_fortran.RefWaterlib(), "TrajObject": StubTrajObject, "simps": legacy_simps, "argrelmin": argrelmin}
    exec(compile(ast.Module(body=[node], type_ignores=[]), src, "exec"), ns)
    traj = [_Frame(xyz[f], boxes[f]) for f in range(xyz.shape[0])]
    out = {"xyz": xyz.astype(np.float32), "boxes": boxes, "n_sol": n_sol}
    assert np.array_equal(out["xyz"].astype(np.float64), xyz)
    cwd = os.getcwd()
    for tag, sol in (("sol", sol_inds), ("nosol", [])):
        StubTrajObject.sol = sol
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                ret = ns["rdfCalc"](None, traj, binwidth=0.1, totbins=110)
                out["rdf_txt_" + tag] = np.loadtxt("rdf.txt")
                out["coord_txt_" + tag] = np.loadtxt("coord.txt")
            finally:
                os.chdir(cwd)
        out["ret_" + tag] = np.asarray(ret, dtype=np.float64)
        print(tag, ret)
    np.savez_compressed(os.path.join(OUT, "rdfcalc_n512.npz"), **out)


if __name__ == "__main__":
    main()
