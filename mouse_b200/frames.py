"""Plain-array view of a frame: what the kernels consume.

``FrameArrays.from_config`` flattens a reference-style ``Config`` (any object with the
``atoms() / bonds() / box()`` accessors of ``lammpack_types.py:139-203`` whose atoms carry
``num, pos, mol_id, res_type, type`` and whose bonds carry ``num, type, atom1, atom2``);
``from_melt`` takes the synthetic generator's arrays.  String fields are turned into integer
codes; molecule / residue codes follow the reference's ``hash8`` (``ordering_functions.py:10-11``)
so that even its hash collisions are reproduced when a ``Config`` is given.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def hash8(s):
    """``abs(hash(s)) % 10**8`` - the reference's molecule / residue code (``ordering_functions.py:10-11``)."""
    return abs(hash(s)) % (10 ** 8)


@dataclass
class FrameArrays:
    positions: np.ndarray            # (N_a, 3) f64
    box: np.ndarray                  # (3,) f64
    atom_num: np.ndarray             # (N_a,) i64   reference atom.num
    atom_type: list                  # per-atom type string
    atom_res: list                   # per-atom res_type string
    atom_mol: list                   # per-atom mol_id string
    bond_atoms: np.ndarray           # (N_b, 2) i32 indices into the atom arrays (atom1, atom2)
    bond_num: np.ndarray             # (N_b,) i64   reference bond.num
    bond_type: list                  # per-bond type string
    _cache: dict = field(default_factory=dict, repr=False)

    @property
    def n_atoms(self):
        return len(self.atom_num)

    @property
    def n_bonds(self):
        return len(self.bond_num)

    # ---------------------------------------------------------------- constructors
    @classmethod
    def from_config(cls, config) -> "FrameArrays":
        atoms = config.atoms()
        index_of = {}
        pos = np.empty((len(atoms), 3), np.float64)
        num = np.empty(len(atoms), np.int64)
        for k, a in enumerate(atoms):
            pos[k, 0], pos[k, 1], pos[k, 2] = a.pos.x, a.pos.y, a.pos.z
            num[k] = a.num
            index_of[a.num] = k
        bonds = config.bonds()
        ba = np.empty((len(bonds), 2), np.int32)
        bnum = np.empty(len(bonds), np.int64)
        for k, b in enumerate(bonds):
            ba[k, 0] = index_of[b.atom1.num]
            ba[k, 1] = index_of[b.atom2.num]
            bnum[k] = b.num
        box = config.box()
        return cls(positions=pos, box=np.array([box.x, box.y, box.z], np.float64), atom_num=num,
                   atom_type=[a.type for a in atoms], atom_res=[a.res_type for a in atoms],
                   atom_mol=[a.mol_id for a in atoms], bond_atoms=ba, bond_num=bnum,
                   bond_type=[b.type for b in bonds])

    @classmethod
    def from_melt(cls, melt, positions=None) -> "FrameArrays":
        pos = melt.positions if positions is None else positions
        return cls(positions=np.ascontiguousarray(pos, np.float64), box=np.asarray(melt.box, np.float64),
                   atom_num=np.arange(1, melt.n_atoms + 1, dtype=np.int64),
                   atom_type=[melt.atom_type_names[t] for t in melt.atom_type],
                   atom_res=[melt.res_type_names[t] for t in melt.res],
                   atom_mol=melt.mol_id_strings(), bond_atoms=np.ascontiguousarray(melt.bond_atoms, np.int32),
                   bond_num=np.arange(1, melt.n_bonds + 1, dtype=np.int64),
                   bond_type=[melt.bond_type_names[t] for t in melt.bond_type])

    # ---------------------------------------------------------------- codes and masks
    def _hashes(self, strings):
        table = {}
        out = np.empty(len(strings), np.int64)
        for k, s in enumerate(strings):
            h = table.get(s)
            if h is None:
                h = table[s] = hash8(s)
            out[k] = h
        return out

    def atom_mol_hash(self):
        if "molh" not in self._cache:
            self._cache["molh"] = self._hashes(self.atom_mol)
        return self._cache["molh"]

    def atom_res_hash(self):
        if "resh" not in self._cache:
            self._cache["resh"] = self._hashes(self.atom_res)
        return self._cache["resh"]

    def bond_mol_hash(self):
        """hash8(bond.atom1.mol_id) per bond (``ordering_functions.py:31,39``)."""
        return self.atom_mol_hash()[self.bond_atoms[:, 0]]

    def atom_mol_code(self):
        """Dense, collision-free molecule codes (0 = the empty mol_id, then 1.. in order of first appearance): what
        the kernels compare for ``sameMolecule=False``.  Unlike ``hash8`` (kept for the reference's (9,N)/(6,N)
        array layout, whose collisions and per-process salt are part of that layout) these are identical on every
        rank and in every run."""
        if "molc" not in self._cache:
            table = {"": 0}
            out = np.empty(len(self.atom_mol), np.int32)
            for k, s in enumerate(self.atom_mol):
                c = table.get(s)
                if c is None:
                    c = table[s] = len(table)
                out[k] = c
            self._cache["molc"] = out
        return self._cache["molc"]

    def bond_mol_code(self):
        """``atom_mol_code`` of atom1 of every bond (the molecule the reference attributes the bond to, ``:31,39``)."""
        return self.atom_mol_code()[self.bond_atoms[:, 0]] if self.n_bonds else np.zeros(0, np.int32)

    def bond_res_hash(self):
        return self.atom_res_hash()[self.bond_atoms[:, 0]]

    @staticmethod
    def _passes(values, allowed):
        """Reference filter rule (``ordering_functions.py:25-28``): None or empty list passes all."""
        if allowed is None or len(allowed) == 0:
            return None
        allowed = set(allowed)
        return np.fromiter((v in allowed for v in values), bool, len(values))

    def bond_mask(self, bond_types=None, res_types=None, mol_ids=None, bond_nums=None):
        """Bonds passing the four filters; residue / molecule filters need BOTH atoms to match."""
        m = np.ones(self.n_bonds, bool)
        t = self._passes(self.bond_type, bond_types)
        if t is not None:
            m &= t
        r = self._passes(self.atom_res, res_types)
        if r is not None:
            m &= r[self.bond_atoms[:, 0]] & r[self.bond_atoms[:, 1]]
        q = self._passes(self.atom_mol, mol_ids)
        if q is not None:
            m &= q[self.bond_atoms[:, 0]] & q[self.bond_atoms[:, 1]]
        b = self._passes(self.bond_num.tolist(), bond_nums)
        if b is not None:
            m &= b
        return m

    # ---------------------------------------------------------------- selections on arrays
    def subset(self, atom_keep) -> "FrameArrays":
        """Atoms with ``atom_keep`` true plus the bonds whose BOTH atoms survive - what the reference's
        ``selectByAtomFields`` / ``select_rectangular`` build as a new ``Config`` (``lammpack_misc.py:96-144``).
        Atom and bond numbers are kept (the reference re-inserts the same objects)."""
        keep = np.asarray(atom_keep, bool)
        new_index = np.cumsum(keep) - 1
        bk = keep[self.bond_atoms[:, 0]] & keep[self.bond_atoms[:, 1]] if self.n_bonds else np.zeros(0, bool)
        ai = np.flatnonzero(keep)
        bi = np.flatnonzero(bk)
        return FrameArrays(positions=self.positions[ai], box=self.box.copy(), atom_num=self.atom_num[ai],
                           atom_type=[self.atom_type[k] for k in ai], atom_res=[self.atom_res[k] for k in ai],
                           atom_mol=[self.atom_mol[k] for k in ai],
                           bond_atoms=new_index[self.bond_atoms[bi]].astype(np.int32).reshape(len(bi), 2),
                           bond_num=self.bond_num[bi], bond_type=[self.bond_type[k] for k in bi])

    def select_chains(self, chain_list) -> "FrameArrays":
        """``select_chains`` (``lammpack_misc.py:120-121``): atoms whose ``mol_id`` is in ``chain_list``."""
        allowed = set(chain_list)
        return self.subset(np.fromiter((m in allowed for m in self.atom_mol), bool, self.n_atoms))

    def select_rectangular(self, cutx=False, xrelmin=0., xrelmax=1., cuty=False, yrelmin=0., yrelmax=1.,
                           cutz=False, zrelmin=0., zrelmax=1.) -> "FrameArrays":
        """``select_rectangular`` (``lammpack_misc.py:124-144``): box-relative coordinate ``pos / box`` inside
        ``[min, max]`` (both ends included) on every axis that is cut."""
        rel = self.positions / self.box
        keep = np.ones(self.n_atoms, bool)
        for a, (cut, lo, hi) in enumerate(((cutx, xrelmin, xrelmax), (cuty, yrelmin, yrelmax), (cutz, zrelmin, zrelmax))):
            if cut:
                keep &= (rel[:, a] >= lo) & (rel[:, a] <= hi)
        return self.subset(keep)
