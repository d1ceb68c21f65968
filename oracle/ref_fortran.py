"""TEST INFRASTRUCTURE ONLY -- the *real* reference, executed (not restated).

This module runs the reference's own compiled Fortran and its own, unmodified Python function
bodies so that parity is pinned on the arithmetic users of WaterOrderLib actually ran:

* ``RefWaterlib`` binds the raw Fortran entry points exported by the reference's prebuilt f2py
  module ``fortran/waterlib.cpython-37m-x86_64-linux-gnu.so`` (built from ``fortran/waterlib.f90``)
  through ``ctypes`` and gives them the f2py call surface (lower-case names, hidden dimension
  arguments, ``intent(out)`` arrays returned, ``logical`` -> int32).  The module needs
  ``libgfortran.so.3`` which this image lacks; ``oracle/gfortran_stub.c`` supplies the handful of
  symbols (none are reached on the hot path).
* ``load_reference_functions`` AST-extracts named ``def`` bodies from the reference's
  ``structureLibs/water_properties.py`` and ``exec``s them with ``np`` and a ``RefWaterlib`` as
  ``wl`` in scope.  The reference module itself is never imported (its import preamble chdir's,
  may ``rm *.so`` and needs removed SciPy names) and its text is never copied into tracked files.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may use this module.  Nothing under ``waterorderlib_b200/`` imports it.

Where the reference lives: ``/root/reference`` in the build container; on a GPU box only the
prebuilt Fortran binary staged into the git-ignored ``oracle/_ref/`` (by ``oracle/build_oracle.py``
while ``/root/reference`` is visible) exists -- the Python bodies are then supplied by the
restatement in ``oracle/ref_driver.py`` (pinned against the live bodies by
tests/test_oracle_vs_reference.py).
"""
import ast
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_SO_NAME = "waterlib.cpython-37m-x86_64-linux-gnu.so"

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int32)


def _find(rel_in_reference, staged_name):
    """Prefer the read-only reference tree, fall back to the staged copy in oracle/_ref/."""
    for cand in (os.path.join("/root/reference", rel_in_reference), os.path.join(_REF_DIR, staged_name)):
        if os.path.exists(cand):
            return cand
    return None


def reference_available():
    """True when the reference's compiled Fortran can be loaded (build container or staged binary)."""
    return (
        os.path.exists(os.path.join(_REF_DIR, "libgfortran.so.3"))
        and _find(os.path.join("fortran", _SO_NAME), _SO_NAME) is not None
    )


def reference_python_available():
    """True only where the reference's Python source is visible (the build container)."""
    return os.path.exists(os.path.join("/root/reference", "structureLibs", "water_properties.py"))


def _f64(a):
    """float64, column-major -- what f2py hands to the Fortran side."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def _box(boxl):
    b = np.ascontiguousarray(np.asarray(boxl, dtype=np.float64).reshape(-1))
    if b.size != 3:
        raise ValueError("boxl must hold 3 values")
    return b


class RefWaterlib:
    """f2py-compatible face of the reference's compiled waterlib (signatures recovered from the
    module's docstrings).  Every argument is passed by reference, arrays are
    column-major, integer/logical are 4 bytes."""

    def __init__(self):
        stub = os.path.join(_REF_DIR, "libgfortran.so.3")
        so = _find(os.path.join("fortran", _SO_NAME), _SO_NAME)
        if not os.path.exists(stub) or so is None:
            raise RuntimeError(
                "reference oracle unavailable: run `python oracle/build_oracle.py` where /root/reference exists"
            )
        self._stub = ctypes.CDLL(stub, mode=ctypes.RTLD_GLOBAL)
        self._lib = ctypes.CDLL(so)
        self._lib.cosangle3_.restype = ctypes.c_double
        self._lib.angbetween_.restype = ctypes.c_double

    # fortran/waterlib.f90:830-862
    def allnearneighbors(self, pos, boxl, lowcut, highcut):
        pos = _f64(pos)
        n = pos.shape[0]
        out = np.zeros((n, n), dtype=np.int32, order="F")
        self._lib.allnearneighbors_(
            _dp(pos), _dp(_box(boxl)), ctypes.byref(ctypes.c_double(lowcut)),
            ctypes.byref(ctypes.c_double(highcut)), _ip(out), ctypes.byref(ctypes.c_int32(n)))
        return out

    # fortran/waterlib.f90:710-743
    def nearneighbors(self, subpos, pos, boxl, lowcut, highcut):
        subpos = _f64(subpos)
        pos = _f64(pos)
        m, n = subpos.shape[0], pos.shape[0]
        out = np.zeros((m, n), dtype=np.int32, order="F")
        self._lib.nearneighbors_(
            _dp(subpos), _dp(pos), _dp(_box(boxl)), ctypes.byref(ctypes.c_double(lowcut)),
            ctypes.byref(ctypes.c_double(highcut)), _ip(out),
            ctypes.byref(ctypes.c_int32(m)), ctypes.byref(ctypes.c_int32(n)))
        return out

    # fortran/waterlib.f90:32-47
    def reimage(self, pos, refpos, boxl):
        pos = _f64(pos)
        n, dim = pos.shape
        ref = np.ascontiguousarray(np.asarray(refpos, dtype=np.float64).reshape(-1))
        out = np.zeros((n, dim), dtype=np.float64, order="F")
        self._lib.reimage_(_dp(pos), _dp(ref), _dp(_box(boxl)), _dp(out),
                           ctypes.byref(ctypes.c_int32(n)), ctypes.byref(ctypes.c_int32(dim)))
        return out

    # fortran/waterlib.f90:867-895 (diagonal is never written by the Fortran; zeros here)
    def tetracosang(self, refpos, neighpos, boxl):
        neigh = _f64(neighpos)
        k = neigh.shape[0]
        ref = np.ascontiguousarray(np.asarray(refpos, dtype=np.float64).reshape(-1))
        out = np.zeros((k, k), dtype=np.float64, order="F")
        self._lib.tetracosang_(_dp(ref), _dp(neigh), _dp(_box(boxl)), _dp(out),
                               ctypes.byref(ctypes.c_int32(k)))
        return out

    # fortran/waterlib.f90:900-918
    def lsidists(self, refpos, neighpos, boxl):
        neigh = _f64(neighpos)
        k = neigh.shape[0]
        ref = np.ascontiguousarray(np.asarray(refpos, dtype=np.float64).reshape(-1))
        out = np.zeros(k, dtype=np.float64)
        self._lib.lsidists_(_dp(ref), _dp(neigh), _dp(_box(boxl)), _dp(out),
                            ctypes.byref(ctypes.c_int32(k)))
        return out

    # fortran/waterlib.f90:1156-1210
    def generalhbonds(self, acceptorpos, donorpos, donorhpos, boxl, distcut, angcut):
        acc, don, donh = _f64(acceptorpos), _f64(donorpos), _f64(donorhpos)
        na, nd, nh = acc.shape[0], don.shape[0], donh.shape[0]
        if nd != nh:
            raise ValueError("donor heavy atoms and hydrogens differ in number (Fortran would `stop`)")
        out = np.zeros((na, nd), dtype=np.int32, order="F")
        self._lib.generalhbonds_(
            _dp(acc), _dp(don), _dp(donh), _dp(_box(boxl)),
            ctypes.byref(ctypes.c_double(distcut)), ctypes.byref(ctypes.c_double(angcut)), _ip(out),
            ctypes.byref(ctypes.c_int32(na)), ctypes.byref(ctypes.c_int32(nd)),
            ctypes.byref(ctypes.c_int32(nh)))
        return out

    # fortran/waterlib.f90:1550-1593
    def histrr3b(self, pos, boxl, distwidth, dnum, angwidth, anum):
        pos = _f64(pos)
        n = pos.shape[0]
        out = np.zeros((dnum, dnum, anum), dtype=np.float64, order="F")
        self._lib.histrr3b_(
            _dp(pos), _dp(_box(boxl)), ctypes.byref(ctypes.c_double(distwidth)),
            ctypes.byref(ctypes.c_int32(dnum)), ctypes.byref(ctypes.c_double(angwidth)),
            ctypes.byref(ctypes.c_int32(anum)), _dp(out), ctypes.byref(ctypes.c_int32(n)))
        return out

    # fortran/waterlib.f90:1286-1341 -> (densvals (nx,ny,nz), densnorms (nx,ny,nz,3)), C-ordered copies
    def willarddensityfield(self, pos, gridx, gridy, gridz, boxl, smoothlen):
        pos = _f64(pos)
        gx, gy, gz = (np.ascontiguousarray(np.asarray(g, dtype=np.float64)) for g in (gridx, gridy, gridz))
        nx, ny, nz = gx.size, gy.size, gz.size
        dens = np.zeros((nx, ny, nz), dtype=np.float64, order="F")
        norms = np.zeros((nx, ny, nz, 3), dtype=np.float64, order="F")
        self._lib.willarddensityfield_(
            _dp(pos), _dp(gx), _dp(gy), _dp(gz), _dp(_box(boxl)), ctypes.byref(ctypes.c_double(smoothlen)),
            _dp(dens), _dp(norms), ctypes.byref(ctypes.c_int32(pos.shape[0])), ctypes.byref(ctypes.c_int32(nx)),
            ctypes.byref(ctypes.c_int32(ny)), ctypes.byref(ctypes.c_int32(nz)))
        return np.ascontiguousarray(dens), np.ascontiguousarray(norms)

    # fortran/waterlib.f90:1219-1268 -> densvals (nx,ny,nz), C-ordered copy
    def densityfield(self, pos, gridx, gridy, gridz, boxl):
        pos = _f64(pos)
        gx, gy, gz = (np.ascontiguousarray(np.asarray(g, dtype=np.float64).reshape(-1)) for g in (gridx, gridy, gridz))
        dens = np.zeros((gx.size, gy.size, gz.size), dtype=np.float64, order="F")
        self._lib.densityfield_(_dp(pos), _dp(gx), _dp(gy), _dp(gz), _dp(_box(boxl)), _dp(dens),
                                ctypes.byref(ctypes.c_int32(pos.shape[0])), ctypes.byref(ctypes.c_int32(gx.size)),
                                ctypes.byref(ctypes.c_int32(gy.size)), ctypes.byref(ctypes.c_int32(gz.size)))
        return np.ascontiguousarray(dens)

    # fortran/waterlib.f90:1351-1398
    def willarddensitypoints(self, pos, denspts, boxl, smoothlen):
        pos, pts = _f64(pos), _f64(denspts)
        npts = pts.shape[0]
        dens = np.zeros(npts, dtype=np.float64)
        norms = np.zeros((npts, 3), dtype=np.float64, order="F")
        self._lib.willarddensitypoints_(
            _dp(pos), _dp(pts), _dp(_box(boxl)), ctypes.byref(ctypes.c_double(smoothlen)), _dp(dens), _dp(norms),
            ctypes.byref(ctypes.c_int32(pos.shape[0])), ctypes.byref(ctypes.c_int32(npts)))
        return dens, np.ascontiguousarray(norms)

    # fortran/waterlib.f90:193-231
    def radialdist(self, pos1, pos2, binwidth, totbins, bulkdens, boxl):
        p1, p2 = _f64(pos1), _f64(pos2)
        rdf = np.zeros(totbins, dtype=np.float64)
        self._lib.radialdist_(_dp(p1), _dp(p2), ctypes.byref(ctypes.c_double(binwidth)), ctypes.byref(ctypes.c_int32(totbins)),
                              ctypes.byref(ctypes.c_double(bulkdens)), _dp(_box(boxl)), _dp(rdf),
                              ctypes.byref(ctypes.c_int32(p1.shape[0])), ctypes.byref(ctypes.c_int32(p2.shape[0])))
        return rdf

    # fortran/waterlib.f90:316-353
    # fortran/waterlib.f90:237-314 (its matmul goes through the stub runtime's _gfortran_matmul_r8)
    def radialdistplane(self, pos1, pos2, binwidth, totbins, bulkdens, boxl):
        p1, p2 = _f64(pos1), _f64(pos2)
        if p1.shape != (3, 3):
            raise ValueError("pos1 must be (3, 3)")
        rdf = np.zeros((totbins, totbins), dtype=np.float64, order="F")
        self._lib.radialdistplane_(_dp(p1), _dp(p2), ctypes.byref(ctypes.c_double(binwidth)), ctypes.byref(ctypes.c_int32(totbins)),
                                   ctypes.byref(ctypes.c_double(bulkdens)), _dp(_box(boxl)), _dp(rdf),
                                   ctypes.byref(ctypes.c_int32(p2.shape[0])))
        return rdf

    def radialdistsame(self, pos, binwidth, totbins, bulkdens, boxl):
        p = _f64(pos)
        rdf = np.zeros(totbins, dtype=np.float64)
        self._lib.radialdistsame_(_dp(p), ctypes.byref(ctypes.c_double(binwidth)), ctypes.byref(ctypes.c_int32(totbins)),
                                  ctypes.byref(ctypes.c_double(bulkdens)), _dp(_box(boxl)), _dp(rdf),
                                  ctypes.byref(ctypes.c_int32(p.shape[0])))
        return rdf

    # fortran/waterlib.f90:358-389
    def pairdistancehistogram(self, pos1, pos2, binwidth, totbins, boxl):
        p1, p2 = _f64(pos1), _f64(pos2)
        hist = np.zeros(totbins, dtype=np.float64)
        self._lib.pairdistancehistogram_(_dp(p1), _dp(p2), ctypes.byref(ctypes.c_double(binwidth)),
                                         ctypes.byref(ctypes.c_int32(totbins)), _dp(_box(boxl)), _dp(hist),
                                         ctypes.byref(ctypes.c_int32(p1.shape[0])), ctypes.byref(ctypes.c_int32(p2.shape[0])),
                                         ctypes.byref(ctypes.c_int32(3)))
        return hist

    # fortran/waterlib.f90:683-703
    def cosangle3(self, p1, p2, p3):
        a = [np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(3)) for p in (p1, p2, p3)]
        return float(self._lib.cosangle3_(_dp(a[0]), _dp(a[1]), _dp(a[2])))

    # fortran/waterlib.f90:954-965
    def angbetween(self, v1, v2):
        a = [np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3)) for v in (v1, v2)]
        return float(self._lib.angbetween_(_dp(a[0]), _dp(a[1])))

    # fortran/waterlib.f90:1414-1469 (numwater / watclose are accumulated into uninitialised
    # outputs by the Fortran; zeroed here as f2py users would see on fresh pages)
    def interfacewater(self, pos, gridpos, gridnorm, cutoff, boxl):
        pos, gp, gn = _f64(pos), _f64(gridpos), _f64(gridnorm)
        n, g = pos.shape[0], gp.shape[0]
        watclose = np.zeros(n, dtype=np.int32)
        surfclose = np.zeros(g, dtype=np.int32)
        numwater = ctypes.c_int32(0)
        allwatdists = np.zeros(n, dtype=np.float64)
        self._lib.interfacewater_(
            _dp(pos), _dp(gp), _dp(gn), ctypes.byref(ctypes.c_double(cutoff)), _dp(_box(boxl)),
            _ip(watclose), _ip(surfclose), ctypes.byref(numwater), _dp(allwatdists),
            ctypes.byref(ctypes.c_int32(n)), ctypes.byref(ctypes.c_int32(g)))
        return watclose, surfclose, int(numwater.value), allwatdists


    # fortran/waterlib.f90:973-1011
    def watorient(self, opos, hpos, refvec, boxl):
        opos, hpos = _f64(opos), _f64(hpos)
        no, nh = opos.shape[0], hpos.shape[0]
        if nh != 2 * no:
            raise ValueError("Number of hydrogens must be two times number of oxygens.")  # the Fortran STOPs here
        ref = np.ascontiguousarray(np.asarray(refvec, dtype=np.float64).reshape(-1))
        angdip, angplane = np.zeros(no), np.zeros(no)
        self._lib.watorient_(_dp(opos), _dp(hpos), _dp(ref), _dp(_box(boxl)), _dp(angdip), _dp(angplane),
                             ctypes.byref(ctypes.c_int32(no)), ctypes.byref(ctypes.c_int32(nh)))
        return angdip, angplane

    # fortran/waterlib.f90:1047-1099
    def binongrid(self, opos, xbins, ybins, zbins):
        opos = _f64(opos)
        xb, yb, zb = (np.ascontiguousarray(np.asarray(g, dtype=np.float64).reshape(-1)) for g in (xbins, ybins, zbins))
        if (yb[1] - yb[0]) != (xb[1] - xb[0]) or (zb[1] - zb[0]) != (xb[1] - xb[0]):
            raise ValueError("Must break volume into CUBES. Currently, bin-widths do not match.")  # the Fortran STOPs here
        out = np.zeros((xb.size - 1, yb.size - 1, zb.size - 1), dtype=np.int32, order="F")
        self._lib.binongrid_(_dp(opos), _dp(xb), _dp(yb), _dp(zb), _ip(out), ctypes.byref(ctypes.c_int32(opos.shape[0])),
                             ctypes.byref(ctypes.c_int32(xb.size)), ctypes.byref(ctypes.c_int32(yb.size)),
                             ctypes.byref(ctypes.c_int32(zb.size)))
        return out


_WP_FUNCS = ("getCosAngs", "getOrderParamq", "tetrahedralMetrics", "getLSI", "HBondsGeneral", "getOrderParamPsi")


def load_reference_functions(names=_WP_FUNCS, wl=None):
    """Return {name: function} built from the reference's own source text
    (structureLibs/water_properties.py:210-391, :681-719), executed with a RefWaterlib as `wl`."""
    src_path = os.path.join("/root/reference", "structureLibs", "water_properties.py")
    if not os.path.exists(src_path):
        raise RuntimeError("reference water_properties.py not visible (only in the build container)")
    with open(src_path) as fh:
        tree = ast.parse(fh.read())
    wl = wl if wl is not None else RefWaterlib()
    ns = {"np": np, "wl": wl}
    wanted = set(names)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, src_path, "exec"), ns)
    missing = wanted - set(ns)
    if missing:
        raise RuntimeError("reference functions not found: %s" % sorted(missing))
    return {k: ns[k] for k in names}


_SORT_SO = "sortlib.cpython-37m-x86_64-linux-gnu.so"


def sortlib_available():
    return os.path.exists(os.path.join(_REF_DIR, "libgfortran.so.3")) and _find(os.path.join("fortran", _SORT_SO), _SORT_SO) is not None


class RefSortlib:
    """f2py-compatible face of the reference's compiled sortlib: depthfirstsort (fortran/sortlib.f90:26-72), reached
    through the wrapper f2py generated for its assumed-shape argument (f2pywrapdepthfirstsort_)."""

    def __init__(self):
        stub = os.path.join(_REF_DIR, "libgfortran.so.3")
        so = _find(os.path.join("fortran", _SORT_SO), _SORT_SO)
        if not os.path.exists(stub) or so is None:
            raise RuntimeError("reference sortlib unavailable: run `python oracle/build_oracle.py` where /root/reference exists")
        self._stub = ctypes.CDLL(stub, mode=ctypes.RTLD_GLOBAL)
        self._lib = ctypes.CDLL(so)

    def depthfirstsort(self, vertex, array, visited, m, n=None):
        """final_visited = depthfirstsort(vertex, array, visited, m, [n]) -- visited is updated in place like f2py's
        intent(inout) would; vertex is 1-based."""
        a = np.asfortranarray(np.asarray(array, dtype=np.int32))
        n = a.shape[0] if n is None else int(n)
        vis = np.ascontiguousarray(visited, dtype=np.int32)
        final = np.zeros(n, dtype=np.int32)  # f2py hands the Fortran a fresh intent(out) array
        self._lib.f2pywrapdepthfirstsort_(ctypes.byref(ctypes.c_int32(int(vertex))), _ip(a), _ip(vis), ctypes.byref(ctypes.c_int32(int(m))),
                                          ctypes.byref(ctypes.c_int32(n)), _ip(final), ctypes.byref(ctypes.c_int32(vis.shape[0])))
        return final


def load_reference_driver_functions(names, sortlib):
    """AST-extract def bodies from the reference's structureLibs/orderParam_lib.py (getClusters, ...) and exec them with
    np and the given sortlib in scope.  Build container only."""
    path = os.path.join("/root/reference", "structureLibs", "orderParam_lib.py")
    tree = ast.parse(open(path).read())
    ns = {"np": np, "sortlib": sortlib}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return {k: ns[k] for k in names}
