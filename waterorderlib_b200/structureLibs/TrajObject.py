"""Drop-in for the reference's ``structureLibs/TrajObject.py`` (:15-103): topology + trajectory access and
the Amber-mask index selections the frame drivers use.

The reference builds on parmed + pytraj (TrajObject.py:30,33), which read AMBER files; neither is needed on
the hot path, which only consumes ``frame.xyz`` (natom,3), ``frame.box.values[:3]`` and index arrays
(orderParam_lib.py:1314-1316).  This module provides the same class and method names over

  * in-memory objects (``Topology`` / ``ArrayTrajectory``), the form synthetic benchmarks and tests use,
  * ``.npz`` files written by ``Topology.save`` / ``ArrayTrajectory.save``,
  * AMBER ``.parm7`` / ``.prmtop`` topologies and ``.nc`` NetCDF trajectories through the built-in readers of
    ``amber_io`` (orthorhombic boxes),
  * anything else is handed to parmed / pytraj exactly as the reference does (ImportError if absent).

Trajectory I/O is outside the hot path: nothing here touches the GPU.
"""
import re

import numpy as np


class Atom:
    """What getHBInds reads from a parmed atom (orderParam_lib.py:75-86): name, idx, bond_partners."""
    __slots__ = ("name", "idx", "resname", "resid", "bond_partners")

    def __init__(self, name, idx, resname, resid):
        self.name, self.idx, self.resname, self.resid = name, idx, resname, resid
        self.bond_partners = []


class Topology:
    """Atom names, residue names, residue ids and bonds; ``select`` understands the Amber-mask subset the
    reference uses: ``:RES``, ``@NAME``, ``@PREFIX=`` (wildcard), ``!``, ``&``, ``|`` and parentheses."""

    def __init__(self, names, resnames, resids=None, bonds=None):
        self.names = np.asarray(names, dtype=str)
        self.resnames = np.asarray(resnames, dtype=str)
        n = self.names.shape[0]
        if self.resnames.shape[0] != n:
            raise ValueError("names and resnames differ in length")
        self.resids = np.arange(n) if resids is None else np.asarray(resids, dtype=np.int64)
        self.bonds = np.zeros((0, 2), dtype=np.int64) if bonds is None else np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
        self._atoms = None
        self._selections = {}

    @property
    def n_atoms(self):
        return int(self.names.shape[0])

    @property
    def atoms(self):
        if self._atoms is None:
            atoms = [Atom(str(nm), i, str(rn), int(ri)) for i, (nm, rn, ri) in enumerate(zip(self.names, self.resnames, self.resids))]
            for a, b in self.bonds:
                atoms[a].bond_partners.append(atoms[b])
                atoms[b].bond_partners.append(atoms[a])
            self._atoms = atoms
        return self._atoms

    # ---- Amber masks -------------------------------------------------------------------------------
    _TOKEN = re.compile(r"\s*([()!&|]|[:@][^()!&|\s]+)")

    def select(self, mask):
        """Indices (ascending int array) of the atoms matching an Amber mask, like pytraj's top.select."""
        cached = self._selections.get(mask)
        if cached is not None:     # string matching over 10^6 atoms costs ~0.1 s; the drivers ask for the same masks again
            return cached.copy()
        tokens = []
        pos = 0
        mask = mask.strip()
        while pos < len(mask):
            m = self._TOKEN.match(mask, pos)
            if not m:
                raise ValueError("cannot parse mask %r at %r" % (mask, mask[pos:]))
            tokens.append(m.group(1))
            pos = m.end()
        self._tok, self._i = tokens, 0
        sel = self._parse_or()
        if self._i != len(tokens):
            raise ValueError("unbalanced mask %r" % mask)
        out = np.nonzero(sel)[0]
        self._selections[mask] = out
        return out.copy()

    def _peek(self):
        return self._tok[self._i] if self._i < len(self._tok) else None

    def _parse_or(self):
        v = self._parse_and()
        while self._peek() == "|":
            self._i += 1
            v = v | self._parse_and()
        return v

    def _parse_and(self):
        v = self._parse_not()
        while self._peek() == "&":
            self._i += 1
            v = v & self._parse_not()
        return v

    def _parse_not(self):
        if self._peek() == "!":
            self._i += 1
            return ~self._parse_not()
        return self._parse_atom()

    def _parse_atom(self):
        t = self._peek()
        if t is None:
            raise ValueError("mask ends unexpectedly")
        self._i += 1
        if t == "(":
            v = self._parse_or()
            if self._peek() != ")":
                raise ValueError("missing ')' in mask")
            self._i += 1
            return v
        if t[0] == ":":
            return self._match(self.resnames, t[1:])
        if t[0] == "@":
            return self._match(self.names, t[1:])
        raise ValueError("unexpected token %r in mask" % t)

    @staticmethod
    def _match(values, spec):
        out = np.zeros(values.shape[0], dtype=bool)
        for item in spec.split(","):
            if item.endswith("="):  # Amber wildcard: @H= matches every name starting with H
                out |= np.char.startswith(values, item[:-1])
            else:
                out |= values == item
        return out

    def n_residues(self, mask=None):
        idx = np.arange(self.n_atoms) if mask is None else self.select(mask)
        return int(np.unique(self.resids[idx]).size)

    def save(self, path):
        np.savez_compressed(path, names=self.names, resnames=self.resnames, resids=self.resids, bonds=self.bonds)

    @classmethod
    def load(cls, path):
        d = np.load(path)
        return cls(d["names"], d["resnames"], d["resids"], d["bonds"])

    @classmethod
    def water_box(cls, n_waters, solute_names=(), solute_resname="SOL", solute_bonds=()):
        """Topology of `solute_names` atoms (one residue) followed by n_waters 3-site waters (O, H1, H2), the
        contiguous O,H,H layout the reference assumes for water (TrajObject.py:45-52)."""
        ns = len(solute_names)
        names = list(solute_names) + ["O", "H1", "H2"] * n_waters
        resnames = [solute_resname] * ns + ["WAT"] * (3 * n_waters)
        resids = [0] * ns + list(np.repeat(np.arange(n_waters) + (1 if ns else 0), 3))
        o = ns + 3 * np.arange(n_waters)
        bonds = np.concatenate([np.asarray(solute_bonds, dtype=np.int64).reshape(-1, 2),
                                np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)])
        return cls(names, resnames, resids, bonds)


class _Box:
    """frame.box: ``.values`` = (a, b, c, alpha, beta, gamma) like pytraj's Box."""
    __slots__ = ("values",)

    def __init__(self, v):
        v = np.asarray(v, dtype=np.float64).reshape(-1)
        self.values = v if v.size == 6 else np.concatenate([v[:3], [90.0, 90.0, 90.0]])


class Frame:
    """One trajectory frame: ``.xyz`` (natom,3) in Angstrom and ``.box.values``."""
    __slots__ = ("xyz", "box")

    def __init__(self, xyz, box):
        self.xyz = xyz
        self.box = _Box(box)


class ArrayTrajectory:
    """In-memory trajectory with the slice of pytraj's TrajectoryIterator interface the drivers use:
    len(), iteration, integer indexing, ``.top.select(mask)``; plus the whole-array views (``xyz``, ``boxes``)
    the batched GPU path reads directly.  ``xyz`` may also be a torch tensor (page-locked host memory or CUDA): the
    batched drivers (tetOrderCalc, threeBodyCalc) then skip the staging copy numpy frames need."""

    def __init__(self, xyz, box, top=None, stride=1):
        if not (hasattr(xyz, "is_cuda") and hasattr(xyz, "ndim")):   # torch tensors (pinned host or CUDA) are kept as they are:
            xyz = np.asarray(xyz)                                    # the batched drivers move them without a staging copy
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise ValueError("xyz must have shape (frames, atoms, 3)")
        box = np.asarray(box, dtype=np.float64)
        if box.ndim == 1:
            box = np.broadcast_to(box, (xyz.shape[0], box.shape[0]))
        self.xyz = xyz[::stride]
        self.boxes = np.ascontiguousarray(box[::stride])
        self.top = top

    def __len__(self):
        return int(self.xyz.shape[0])

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return Frame(self.xyz[i], self.boxes[i])
        if isinstance(i, slice):  # a sub-trajectory, as pytraj's traj[a:b] (rdfCalc's chunks, orderParam_lib.py:617)
            return ArrayTrajectory(self.xyz[i], self.boxes[i], top=self.top)
        raise TypeError("frame indices must be integers or slices")

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def save(self, path):
        np.savez_compressed(path, xyz=self.xyz, box=self.boxes)

    @classmethod
    def load(cls, path, top=None, stride=1):
        d = np.load(path)
        return cls(d["xyz"], d["box"], top=top, stride=stride)


class TrajObject:
    """Same constructor and methods as the reference class (TrajObject.py:15-103)."""

    def __init__(self, topFile, trajFile=None, stride=1, solResName='(!:WAT)', watResName='(:WAT)'):
        self.topFile = topFile
        self.trajFile = trajFile
        self.stride = stride
        self.solResName = solResName
        self.watResName = watResName
        self.top = self._load_top(topFile)
        if trajFile is not None:
            self.traj = self._load_traj(trajFile, stride)

    @staticmethod
    def _load_top(topFile):
        if isinstance(topFile, Topology):
            return topFile
        if isinstance(topFile, str) and topFile.endswith(".npz"):
            return Topology.load(topFile)
        if isinstance(topFile, str) and topFile.lower().endswith((".parm7", ".prmtop", ".top")):
            from .amber_io import read_parm7  # built-in reader for the fields the hot path needs
            return read_parm7(topFile)
        import parmed as pmd  # anything else: as the reference does (TrajObject.py:30)
        return pmd.load_file(topFile)

    def _load_traj(self, trajFile, stride):
        if isinstance(trajFile, ArrayTrajectory):
            t = ArrayTrajectory(trajFile.xyz, trajFile.boxes, top=self.top, stride=stride)
            return t
        if isinstance(trajFile, (tuple, list)) and len(trajFile) == 2 and not isinstance(trajFile[0], str):
            return ArrayTrajectory(trajFile[0], trajFile[1], top=self.top, stride=stride)
        if isinstance(trajFile, str) and trajFile.endswith(".npz"):
            return ArrayTrajectory.load(trajFile, top=self.top, stride=stride)
        if isinstance(trajFile, str) and trajFile.lower().endswith((".nc", ".ncdf", ".netcdf")) and isinstance(self.top, Topology):
            from .amber_io import NetCDFTrajectory  # built-in AMBER NetCDF reader (scipy.io.netcdf_file, memory-mapped)
            return NetCDFTrajectory(trajFile, top=self.top, stride=stride)
        import pytraj as pt  # TrajObject.py:33
        return pt.iterload(trajFile, pt.load_parmed(self.top, traj=False), stride=stride)

    def _select(self, mask):
        return self.traj.top.select(mask) if hasattr(self, "traj") else self.top.select(mask)

    def getWatInds(self):
        """(watInds, watHInds, lenWat): water oxygens, water hydrogens, atoms per water (TrajObject.py:35-52)."""
        nWatAtoms = len(self._select(self.watResName))
        watInds = self._select(self.watResName + '&(!@H=)&(!@EP=)')
        watHInds = self._select(self.watResName + '&(@H=)')
        lenWat = int(nWatAtoms / len(watInds)) if len(watInds) != 0 else 0
        return watInds, watHInds, lenWat

    def getHeavyInds(self):
        return self._select('(!@H=)&(!@EP=)')

    def getPhobicInds(self):
        return self._select('(@C=)|(@S=)')

    def getPhilicInds(self):
        return self._select('(@O=)|(@N=)')

    def getSolInds(self):
        """(solInds, solHInds, solCInds, solNInds, solOInds, solSInds) (TrajObject.py:85-103)."""
        s = self.solResName
        return (self._select(s + '&(!@H=)'), self._select(s + '&(@H=)'), self._select(s + '&(@C=)'),
                self._select(s + '&(@N=)'), self._select(s + '&(@O=)'), self._select(s + '&(@S=)'))
