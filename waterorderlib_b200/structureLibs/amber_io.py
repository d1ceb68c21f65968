"""Minimal AMBER readers so that ``TrajObject(topFile="x.parm7", trajFile="x.nc")`` works without parmed / pytraj
(the reference loads both through those packages, structureLibs/TrajObject.py:30,33):

  * ``read_parm7``         the topology fields the hot path needs from a parm7 / prmtop text file: atom names, residue
                           labels and pointers, bonds (for getHBInds' ``atom.bond_partners``)
  * ``NetCDFTrajectory``   AMBER NetCDF (convention "AMBER", NetCDF-3) trajectories through ``scipy.io.netcdf_file``
                           with memory mapping: ``coordinates`` (frame, atom, 3) float32 and ``cell_lengths`` (frame, 3)

Trajectory I/O is outside the hot path: nothing here touches the GPU.  Frames are handed on
as float32 -- exactly the values the file stores (pytraj upcasts the same numbers to float64); the kernels take float32
storage and compute in fp64.
"""
import numpy as np

from .TrajObject import ArrayTrajectory, Frame, Topology


def _sections(path):
    """{flag: (format, [lines])} of a parm7 file."""
    out, flag, fmt, lines = {}, None, None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith("%FLAG"):
                if flag is not None:
                    out[flag] = (fmt, lines)
                flag, fmt, lines = line.split()[1], None, []
            elif line.startswith("%FORMAT"):
                fmt = line.strip()[8:-1]
            elif line.startswith("%"):
                continue
            elif flag is not None:
                lines.append(line.rstrip("\n"))
    if flag is not None:
        out[flag] = (fmt, lines)
    return out


def _fixed(lines, width):
    vals = []
    for ln in lines:
        vals.extend(ln[i:i + width] for i in range(0, len(ln), width))
    return vals


def read_parm7(path):
    """parm7 / prmtop -> Topology (names, residue names, residue ids, bonds)."""
    sec = _sections(path)
    if "POINTERS" not in sec or "ATOM_NAME" not in sec:
        raise ValueError("%s does not look like an AMBER parm7 topology" % path)
    pointers = [int(v) for v in _fixed(sec["POINTERS"][1], 8) if v.strip()]
    natom, nres = pointers[0], pointers[11]
    names = [v.strip() for v in _fixed(sec["ATOM_NAME"][1], 4)][:natom]
    labels = [v.strip() for v in _fixed(sec["RESIDUE_LABEL"][1], 4)][:nres]
    first = [int(v) for v in _fixed(sec["RESIDUE_POINTER"][1], 8) if v.strip()][:nres]  # 1-based first atom of each residue
    bounds = np.array(first + [natom + 1], dtype=np.int64) - 1
    resids = np.repeat(np.arange(nres), np.diff(bounds))
    resnames = np.asarray(labels, dtype=str)[resids]
    bonds = []
    for flag in ("BONDS_INC_HYDROGEN", "BONDS_WITHOUT_HYDROGEN"):
        if flag in sec:
            v = [int(x) for x in _fixed(sec[flag][1], 8) if x.strip()]
            trip = np.asarray(v, dtype=np.int64).reshape(-1, 3)
            bonds.append(trip[:, :2] // 3)  # coordinate-array offsets -> atom indices; third entry is the bond type
    bonds = np.concatenate(bonds) if bonds else np.zeros((0, 2), dtype=np.int64)
    return Topology(names, resnames, resids, bonds)


class _NativeFrames:
    """Array-like view of the file's coordinates that returns native-endian arrays (NetCDF-3 stores big-endian)."""

    def __init__(self, var, stride):
        self._v, self._stride = var, stride
        n = var.shape[0]
        self.shape = ((n + stride - 1) // stride,) + tuple(var.shape[1:])
        self.ndim = 3
        self.dtype = np.dtype(np.float32)

    def __len__(self):
        return self.shape[0]

    def raw(self, begin, end):
        """Frames [begin, end) exactly as the file stores them (big-endian float32, a view of the memory map): the
        batched drivers upload these bytes and swap them on the device."""
        return self._v[begin * self._stride:end * self._stride:self._stride]

    def __getitem__(self, key):
        if isinstance(key, tuple):
            head, rest = key[0], key[1:]
        else:
            head, rest = key, ()
        if isinstance(head, slice):
            start, stop, step = head.indices(self.shape[0])
            raw = self._v[start * self._stride:stop * self._stride:step * self._stride]
        else:
            raw = self._v[int(head) * self._stride]
        out = np.ascontiguousarray(raw, dtype=np.float32)
        return out[(slice(None),) + rest] if (rest and isinstance(head, slice)) else (out[rest] if rest else out)


class NetCDFTrajectory(ArrayTrajectory):
    """AMBER NetCDF trajectory with the slice of pytraj's TrajectoryIterator interface the drivers use."""

    def __init__(self, path, top=None, stride=1):
        from scipy.io import netcdf_file
        self._nc = netcdf_file(path, "r", mmap=True)
        v = self._nc.variables
        if "coordinates" not in v:
            raise ValueError("%s has no 'coordinates' variable (not an AMBER NetCDF trajectory)" % path)
        self.xyz = _NativeFrames(v["coordinates"], stride)
        n = len(self.xyz)
        if "cell_lengths" in v:
            self.boxes = np.ascontiguousarray(v["cell_lengths"][::stride], dtype=np.float64)
            ang = np.ascontiguousarray(v["cell_angles"][::stride], dtype=np.float64) if "cell_angles" in v else np.full((n, 3), 90.0)
            if np.any(np.abs(ang - 90.0) > 1e-6):
                raise ValueError("non-orthorhombic cells are not supported (the reference ignores cell angles, orderParam_lib.py:1315)")
        else:
            raise ValueError("%s has no periodic box ('cell_lengths')" % path)
        self.top = top

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return Frame(self.xyz[int(i)], self.boxes[int(i)])
        if isinstance(i, slice):  # a sub-trajectory (copied out of the map), as pytraj's traj[a:b]
            return ArrayTrajectory(self.xyz[i], self.boxes[i], top=self.top)
        raise TypeError("frame indices must be integers or slices")

    def close(self):
        """Release the memory map (frames already handed out are copies and stay valid)."""
        import warnings
        self.xyz = None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # scipy warns that its own variable objects still reference the map
            self._nc.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def write_netcdf(path, xyz, boxes):
    """Write an AMBER-convention NetCDF-3 trajectory (used by the tests and for exporting synthetic boxes)."""
    from scipy.io import netcdf_file
    xyz = np.asarray(xyz, dtype=np.float32)
    boxes = np.asarray(boxes, dtype=np.float64).reshape(xyz.shape[0], 3)
    f = netcdf_file(path, "w", version=2)
    f.Conventions, f.ConventionVersion, f.program = "AMBER", "1.0", "waterorderlib_b200"
    f.createDimension("frame", None)
    f.createDimension("spatial", 3)
    f.createDimension("atom", xyz.shape[1])
    f.createDimension("cell_spatial", 3)
    f.createDimension("cell_angular", 3)
    c = f.createVariable("coordinates", "f", ("frame", "atom", "spatial"))
    c.units = "angstrom"
    cl = f.createVariable("cell_lengths", "d", ("frame", "cell_spatial"))
    ca = f.createVariable("cell_angles", "d", ("frame", "cell_angular"))
    for t in range(xyz.shape[0]):
        c[t] = xyz[t]
        cl[t] = boxes[t]
        ca[t] = (90.0, 90.0, 90.0)
    f.close()


def write_parm7(path, top):
    """Write the topology fields read_parm7 understands (enough for round trips and tests)."""
    n = top.n_atoms
    _, first = np.unique(top.resids, return_index=True)
    first = np.sort(first)
    labels = [top.resnames[i] for i in first]
    is_h = np.char.startswith(top.names, "H")
    with_h = [b for b in top.bonds if is_h[b[0]] or is_h[b[1]]]
    without = [b for b in top.bonds if not (is_h[b[0]] or is_h[b[1]])]
    pointers = [0] * 31
    pointers[0], pointers[2], pointers[3], pointers[11] = n, len(with_h), len(without), len(first)

    def ints(vals):
        return ["".join("%8d" % v for v in vals[i:i + 10]) for i in range(0, max(len(vals), 1), 10)]

    def strs(vals):
        return ["".join("%-4s" % v[:4] for v in vals[i:i + 20]) for i in range(0, max(len(vals), 1), 20)]

    with open(path, "w") as fh:
        fh.write("%VERSION  VERSION_STAMP = V0001.000  DATE = 01/01/01  00:00:00\n")
        for flag, fmt, lines in (("POINTERS", "(10I8)", ints(pointers)), ("ATOM_NAME", "(20a4)", strs(list(top.names))),
                                 ("RESIDUE_LABEL", "(20a4)", strs(labels)), ("RESIDUE_POINTER", "(10I8)", ints([int(i) + 1 for i in first])),
                                 ("BONDS_INC_HYDROGEN", "(10I8)", ints([v for b in with_h for v in (3 * int(b[0]), 3 * int(b[1]), 1)])),
                                 ("BONDS_WITHOUT_HYDROGEN", "(10I8)", ints([v for b in without for v in (3 * int(b[0]), 3 * int(b[1]), 1)]))):
            fh.write("%%FLAG %s\n%%FORMAT%s\n" % (flag, fmt))
            for ln in lines:
                fh.write(ln + "\n")
