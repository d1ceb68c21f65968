"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/wol_capi.h declares, and its
host-only entry points (grid plan, angle-bin table, argument validation) behave.  No kernel is launched."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import port
from waterorderlib_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from waterorderlib_b200 import build
    build.build_lib(verbose=False)
    return _capi.lib()


def test_exports_every_declared_symbol(L):
    header = open(os.path.join(ROOT, "include", "wol_capi.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(wol_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "libwol.so does not export " + name
    assert declared == set(_capi.SIGNATURES), "ctypes binding and header disagree: %s" % (declared ^ set(_capi.SIGNATURES))
    assert L.wol_abi_version() == 2
    assert ctypes.sizeof(_capi.Q3bArgs) % 8 == 0


def test_plan_grid(L):
    box = np.array([[49.655, 49.655, 49.655], [49.0, 50.0, 24.83]])
    nc = (ctypes.c_int32 * 3)()
    edge = ctypes.c_double()
    _capi.check(L.wol_plan_grid(box.ctypes.data_as(ctypes.c_void_p), 2, 3.413, ctypes.byref(nc), ctypes.byref(edge), None), "plan")
    assert tuple(nc) == (14, 14, 7)
    assert edge.value >= 3.413 * (1 + 1e-9) and abs(edge.value - 49.0 / 14) < 1e-12
    tiny = np.array([[5.0, 5.0, 5.0]])
    _capi.check(L.wol_plan_grid(tiny.ctypes.data_as(ctypes.c_void_p), 1, 3.413, ctypes.byref(nc), ctypes.byref(edge), None), "plan")
    assert tuple(nc) == (1, 1, 1)
    bad = np.array([[10.0, -1.0, 10.0]])
    rc = L.wol_plan_grid(bad.ctypes.data_as(ctypes.c_void_p), 1, 3.4, ctypes.byref(nc), ctypes.byref(edge), None)
    assert rc == -2 and b"non-periodic" in L.wol_last_error()
    with pytest.raises(ValueError):
        _capi.check(rc, "plan")


def table_position(tab, nbins, c):
    """What the device does with the table (wol_q3b.cu: angle_position), vectorised."""
    thr = tab[: nbins + 1]
    pos = (c[:, None] <= thr[None, :]).sum(axis=1) - 1  # thresholds decrease, so this is the largest k
    return np.where(c == -1.0, int(tab[nbins + 1]), pos)


@pytest.mark.parametrize("lo,hi,nbins", [(0.0, 180.0, 500), (0.0, 180.0, 180), (30.0, 150.0, 97), (0.0, 1.0, 10)])
def test_angle_table_reproduces_reference_binning(L, lo, hi, nbins):
    tab = np.zeros(nbins + 1 + _capi.WOL_TABLE_EXTRA)
    _capi.check(L.wol_angle_table(lo, hi, nbins, 100.0, 120.0, tab.ctypes.data_as(ctypes.c_void_p)), "table")
    assert tab[nbins + 5] == 1.0
    rng = np.random.default_rng(1)
    c = np.concatenate([rng.random(200000) * 2.0 - 1.0, np.cos(np.deg2rad(np.linspace(0, 180, 2001))),
                        [1.0, -1.0, 0.0, np.nextafter(-1.0, 0.0), np.nextafter(1.0, 0.0)]])
    # every threshold and its neighbours a few ulps either side
    thr = tab[: nbins + 1]
    thr = thr[(thr > -1.0) & (thr <= 1.0)]
    near = [thr]
    up, dn = thr.copy(), thr.copy()
    for _ in range(3):
        up = np.nextafter(up, 2.0)
        dn = np.nextafter(dn, -2.0)
        near += [up[up <= 1.0], dn[dn >= -1.0]]
    c = np.concatenate([c] + near)
    ang = port.angles_from_cos(c)
    want = port.histogram(ang, nbins, lo, hi)
    pos = table_position(tab, nbins, c)
    got = np.bincount(pos[(pos >= 0) & (pos < nbins)], minlength=nbins)
    assert np.array_equal(got, want)
    # out-of-range agreement element-wise
    outside = (ang < lo) | (ang > hi)
    assert np.array_equal((pos < 0) | (pos >= nbins), outside)
    # tetrahedral window thresholds
    in_win = (ang >= 100.0) & (ang <= 120.0)
    assert np.array_equal((c <= tab[nbins + 3]) & (c >= tab[nbins + 4]) & (c != -1.0), in_win)


def test_minus_180_and_zero_positions(L):
    tab = np.zeros(500 + 1 + _capi.WOL_TABLE_EXTRA)
    _capi.check(L.wol_angle_table(0.0, 180.0, 500, 100.0, 120.0, tab.ctypes.data_as(ctypes.c_void_p)), "table")
    assert tab[501] == -1.0  # c == -1 -> -180 degrees -> dropped by np.histogram(range=[0,180])
    assert tab[502] == 0.0   # coincident positions -> 0 degrees -> bin 0
    assert tab[0] == 1.0 and tab[500] == -2.0
    assert np.all(np.diff(tab[:500]) < 0)


def test_argument_validation_without_gpu(L):
    a = _capi.Q3bArgs()
    assert L.wol_q3b_frames(ctypes.byref(a), None) == -1 and b"struct_size" in L.wol_last_error()
    a.struct_size = ctypes.sizeof(_capi.Q3bArgs)
    assert L.wol_q3b_frames(ctypes.byref(a), None) == -1
    nc = (ctypes.c_int32 * 3)(4, 4, 4)
    assert L.wol_workspace_bytes(2, 1000, 1000, ctypes.byref(nc)) > 2 * 1000 * 32
    assert L.wol_cell_build(None, 0, None, 1, 10, ctypes.byref(nc), 0, None, 0, None) == -1


def test_all_atoms_convention_is_validated_without_gpu(L):
    """centres == NULL means "every atom of the cell list is a centre" (walked in cell order) and needs n_centres == n_pos;
    mode 1 of wol_pair_hist pairs a set with itself.  Both are rejected before anything is launched."""
    nc = (ctypes.c_int32 * 3)(8, 8, 8)
    p = ctypes.c_void_p(4096)  # never dereferenced: the calls fail in their argument checks
    rc = L.wol_lsi(None, 0, p, 1, 100, 99, ctypes.byref(nc), 8.0, 0.0, 3.7, p, 1 << 20, p, p, None)
    assert rc == -1 and b"every atom is a centre" in L.wol_last_error()
    rc = L.wol_neighbors_csr(None, 0, p, 1, 100, 99, ctypes.byref(nc), 8.0, 0.0, 3.5, p, 1 << 20, p, p, None, 0, None)
    assert rc == -1 and b"every atom is a centre" in L.wol_last_error()
    rc = L.wol_angles_fill(None, 0, p, 1, 100, 99, ctypes.byref(nc), 8.0, 0.0, 3.4, p, 1 << 20, p, p, None)
    assert rc == -1 and b"every atom is a centre" in L.wol_last_error()
    rc = L.wol_pair_hist(1, p, 0, 100, p, 99, ctypes.byref(nc), 8.0, 0.1, 50, p, 1 << 20, p, None)
    assert rc == -1 and b"pairs a set with itself" in L.wol_last_error()
