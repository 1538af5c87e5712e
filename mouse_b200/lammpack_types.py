"""Minimal object model with the reference's accessor names (``lammpack_types.py:20-203``,
``vector3d.py:16-66``) so that reference-style callers and the CLI keep working: ``Config`` /
``Atom`` / ``Bond`` / ``Vector3d`` and the LAMMPS-data reader.  The kernels never see these
objects - ``mouse_b200.frames.FrameArrays.from_config`` flattens them."""
from __future__ import annotations

import math


class Vector3d:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z

    def set(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    def length_sq(self):
        return self.x * self.x + self.y * self.y + self.z * self.z

    def length(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def scale(self, s):
        self.x *= s
        self.y *= s
        self.z *= s

    def __repr__(self):
        return "Vector3d(%f, %f, %f)" % (self.x, self.y, self.z)


class Atom:
    def __init__(self):
        self.pos = Vector3d()
        self.pbc = Vector3d(0, 0, 0)
        self.type = ""
        self.id = ""
        self.mol_id = ""
        self.res_type = ""
        self.res_num = 0
        self.num = 0
        self.charge = 0.
        self.mass = 0.0
        self.is_hetatm = False
        self.res_insert = ""
        self.bfactor = 0.0
        self.occupancy = 0.0

    def pdb_str(self):
        """One ATOM / HETATM record, field widths of ``lammpack_types.py:41-53`` (5-column serial; the record
        name loses its last blank from atom 100 000 on, which PyMol still reads)."""
        name = "HETATM" if self.is_hetatm else ("ATOM  " if self.num < 100000 else "ATOM ")
        t = self.type
        if len(t) == 1:
            t = " " + t + "  "
        elif len(t) == 2:
            t = " " + t + " "
        elif len(t) == 3:
            t = t + " " if t[0].isdigit() else " " + t
        return "%s%5s %4s %4s%1s%4d%1s   %8.3f%8.3f%8.3f%6.2f%6.2f         %2s%2s" % (
            name, self.num, t, self.res_type, self.mol_id[:1], self.res_num, self.res_insert,
            self.pos.x, self.pos.y, self.pos.z, self.occupancy, self.bfactor, "  ", "  ")


class Bond:
    def __init__(self):
        self.atom1 = None
        self.atom2 = None
        self.type = ""
        self.num = 0

    def pdb_str(self):
        """CONECT record (``lammpack_types.py:94-99``): 5-column serials, 7 from atom 100 000 on."""
        w = 7 if (self.atom1.num >= 100000 or self.atom2.num >= 100000) else 5
        return "CONECT" + str(self.atom1.num).rjust(w) + str(self.atom2.num).rjust(w)


class Config:
    """Container with the accessors the hot path's callers use (``lammpack_types.py:139-203``)."""

    def __init__(self):
        self._atoms, self._bonds = [], []
        self._box, self._box_center = Vector3d(), Vector3d()
        self._atoms_by_num, self._bonds_by_num = {}, {}
        self.title = ""
        self.timestep = 0

    def box(self):
        return self._box

    def set_box(self, box):
        self._box = box

    def box_center(self):
        return self._box_center

    def set_box_center(self, c):
        self._box_center = c

    def write_pdb(self, pdb, hide_pbc_bonds=False):
        """PDB file as the reference writes it (``lammpack_types.py:316-339``): TITLE, orthorhombic CRYST1, one
        MODEL with the atoms, TER / ENDMDL, then the CONECT records; ``hide_pbc_bonds`` drops bonds that span
        more than half a box edge (they would be drawn across the cell).  The reference orders atoms and
        bonds with a comparator that only ever answers "less" or "equal"; for files read in ascending order
        (every LAMMPS data file) that is the stored order, which is what is written here."""
        box = self.box()
        with open(pdb, "w") as f:
            f.write("TITLE     " + self.title + "\n")
            f.write("%6s%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f\n" % ("CRYST1", box.x, box.y, box.z, 90., 90., 90.))
            f.write("MODEL     %4i\n" % 1)
            for atom in self._atoms:
                f.write(atom.pdb_str() + "\n")
            f.write("TER\n")
            f.write("ENDMDL\n")
            for bond in self._bonds:
                if hide_pbc_bonds:
                    a1, a2 = bond.atom1.pos, bond.atom2.pos
                    if (box.x > 0 and abs(a2.x - a1.x) / box.x > 0.5) or (box.y > 0 and abs(a2.y - a1.y) / box.y > 0.5) \
                            or (box.z > 0 and abs(a2.z - a1.z) / box.z > 0.5):
                        continue
                f.write(bond.pdb_str() + "\n")

    def n_atom(self):
        return len(self._atoms)

    def n_bond(self):
        return len(self._bonds)

    def atoms(self):
        return self._atoms

    def bonds(self):
        return self._bonds

    def atom_by_num(self, i):
        return self._atoms_by_num[i]

    def bond_by_num(self, i):
        return self._bonds_by_num[i]

    def insert_atom(self, atom):
        self._atoms.append(atom)
        self._atoms_by_num[atom.num] = atom

    def insert_bond(self, bond):
        self._bonds.append(bond)
        self._bonds_by_num[bond.num] = bond

    def bond_vector_by_num(self, bond_num):
        """(vector, midpoint) with minimum image - ``lammpack_types.py:508-515, 587-599``."""
        from .ordering_functions import _bond_vector
        b = self._bonds_by_num[bond_num]
        v, c = _bond_vector((b.atom1.pos.x, b.atom1.pos.y, b.atom1.pos.z),
                            (b.atom2.pos.x, b.atom2.pos.y, b.atom2.pos.z),
                            (self._box.x, self._box.y, self._box.z))
        return Vector3d(*v), Vector3d(*c)

    def read_lmp_data(self, lmp_data, bonds=True, angles=True, dihedrals=True):
        """LAMMPS data file: header counts, ``xlo xhi`` box lines, ``Atoms`` then ``Bonds``
        (``lammpack_types.py:400-464``, ``lammpack_lmp_functions.py:3-33``): atom lines are
        ``id mol type [q] x y z ix iy iz`` with the coordinates taken from the END of the line;
        bond atoms are sorted ascending and bonds renumbered 1.. in file order."""
        natom = nbond = 0
        with open(lmp_data, "r") as f:
            while True:
                line = f.readline()
                if not line:
                    break
                w = line.split()
                if not w:
                    continue
                if w[0] == "Atoms":
                    f.readline()
                    for _ in range(natom):
                        t = f.readline().split()
                        a = Atom()
                        a.id, a.num, a.mol_id, a.type = t[0], int(t[0]), t[1], t[2]
                        a.res_num = int(a.mol_id)
                        if len(t) == 9 + 1:
                            a.charge = float(t[3])
                        a.pos = Vector3d(float(t[-6]), float(t[-5]), float(t[-4]))
                        a.pbc = Vector3d(int(t[-3]), int(t[-2]), int(t[-1]))
                        self.insert_atom(a)
                elif w[0] == "Bonds":
                    f.readline()
                    for _ in range(nbond):
                        t = f.readline().split()
                        if not bonds:
                            continue
                        b = Bond()
                        b.type = t[1]
                        lo, hi = sorted((int(t[2]), int(t[3])))
                        b.atom1, b.atom2 = self._atoms_by_num[lo], self._atoms_by_num[hi]
                        b.num = self.n_bond() + 1
                        self.insert_bond(b)
                elif len(w) >= 4 and w[2] in ("xlo", "ylo", "zlo") and w[3] == w[2][0] + "hi":
                    lo, hi = float(w[0]), float(w[1])
                    setattr(self._box, w[2][0], abs(hi - lo))
                    setattr(self._box_center, w[2][0], (hi + lo) / 2.)
                elif len(w) >= 2 and w[1] == "atoms":
                    natom = int(w[0])
                elif len(w) >= 2 and w[1] == "bonds":
                    nbond = int(w[0])
