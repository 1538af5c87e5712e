"""Synthetic bead-spring polymer melts (SURVEY.md §8(d) "Synthetic input generator").

Random-walk Kremer-Grest-like melt: ``n_chains`` chains of ``n_beads`` beads,
number density 0.85 sigma^-3, bond length 0.97 sigma, cubic (or scaled
orthorhombic) periodic box centred on the origin.  Everything is float64 and
fully determined by ``seed`` so that the reference, the oracle and the CUDA
path all see the same bits.

Only NumPy is used here: the generator feeds the reference harness
(tests/golden/make_golden.py), the oracle, the tests and bench.py alike.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DENSITY = 0.85
BOND_LENGTH = 0.97


@dataclass
class Melt:
    """Plain-array description of one frame (the tensors the C-ABI consumes)."""

    positions: np.ndarray          # (N_a, 3) float64, wrapped into [-L/2, L/2)
    images: np.ndarray             # (N_a, 3) int32 periodic image flags
    box: np.ndarray                # (3,) float64 edge lengths
    bond_atoms: np.ndarray         # (N_b, 2) int32, 0-based atom indices, atom1 < atom2
    bond_type: np.ndarray          # (N_b,) int32 index into bond_type_names
    bond_type_names: list = field(default_factory=lambda: ["1"])
    atom_type: np.ndarray = None   # (N_a,) int32 index into atom_type_names
    atom_type_names: list = field(default_factory=lambda: ["1"])
    mol: np.ndarray = None         # (N_a,) int32, 0-based chain index (mol_id string = str(mol+1))
    res: np.ndarray = None         # (N_a,) int32 index into res_type_names
    res_type_names: list = field(default_factory=lambda: [""])

    @property
    def n_atoms(self) -> int:
        return int(self.positions.shape[0])

    @property
    def n_bonds(self) -> int:
        return int(self.bond_atoms.shape[0])

    def mol_id_strings(self):
        return [str(int(m) + 1) for m in self.mol]


def make_melt(n_chains: int, n_beads: int, seed: int = 0, *, density: float = DENSITY,
              bond_length: float = BOND_LENGTH, aspect=(1.0, 1.0, 1.0),
              n_bond_types: int = 1, n_res_types: int = 1, n_atom_types: int = 1) -> Melt:
    """Random-walk melt.  ``aspect`` stretches the box (volume preserved).

    With ``n_bond_types`` > 1 bond k of every chain gets type ``(k % n_bond_types)``;
    with ``n_res_types`` > 1 chain c gets residue type ``c % n_res_types`` (names
    "R0", "R1", ...); with ``n_atom_types`` > 1 bead k gets type ``k % n_atom_types``.
    """
    rng = np.random.default_rng(seed)
    n = n_chains * n_beads
    edge = (n / density) ** (1.0 / 3.0)
    asp = np.asarray(aspect, dtype=np.float64)
    asp = asp / np.prod(asp) ** (1.0 / 3.0)
    box = edge * asp
    starts = (rng.random((n_chains, 1, 3)) - 0.5) * box
    steps = rng.normal(size=(n_chains, n_beads, 3))
    steps /= np.linalg.norm(steps, axis=2, keepdims=True)
    steps *= bond_length
    steps[:, 0, :] = 0.0
    unwrapped = (starts + np.cumsum(steps, axis=1)).reshape(n, 3)
    img = np.floor((unwrapped + box / 2.0) / box)
    pos = unwrapped - img * box
    # guard the half-open interval against round-off at the upper face
    over = pos >= box / 2.0
    pos = np.where(over, pos - box, pos)
    img = img + over
    first = (np.arange(n_chains, dtype=np.int64)[:, None] * n_beads
             + np.arange(n_beads - 1, dtype=np.int64)[None, :]).reshape(-1)
    bond_atoms = np.stack([first, first + 1], axis=1).astype(np.int32)
    k_in_chain = np.tile(np.arange(n_beads - 1, dtype=np.int32), n_chains)
    bead_in_chain = np.tile(np.arange(n_beads, dtype=np.int32), n_chains)
    mol = np.repeat(np.arange(n_chains, dtype=np.int32), n_beads)
    return Melt(
        positions=np.ascontiguousarray(pos, dtype=np.float64),
        images=img.astype(np.int32),
        box=np.ascontiguousarray(box, dtype=np.float64),
        bond_atoms=bond_atoms,
        bond_type=(k_in_chain % n_bond_types).astype(np.int32),
        bond_type_names=[str(t + 1) for t in range(n_bond_types)],
        atom_type=(bead_in_chain % n_atom_types).astype(np.int32),
        atom_type_names=[str(t + 1) for t in range(n_atom_types)],
        mol=mol,
        res=(mol % n_res_types).astype(np.int32),
        res_type_names=([""] if n_res_types == 1 else ["R%d" % t for t in range(n_res_types)]),
    )


def jitter(melt: Melt, frame: int, sigma: float = 0.05) -> np.ndarray:
    """Positions of trajectory frame ``frame``: base melt + Gaussian jitter (seed = frame)."""
    rng = np.random.default_rng(1_000_003 + frame)
    return melt.positions + sigma * rng.normal(size=melt.positions.shape)


def write_lammps_data(melt: Melt, path: str) -> None:
    """LAMMPS data file the reference's ``read_lmp_data`` (lammpack_types.py:400-464) parses:
    9-column atom lines ``id mol type x y z ix iy iz``; ``Atoms`` precedes ``Bonds``."""
    half = melt.box / 2.0
    with open(path, "w") as f:
        f.write("LAMMPS data file via mouse_b200.melt\n\n")
        f.write("%d atoms\n%d bonds\n\n" % (melt.n_atoms, melt.n_bonds))
        f.write("%d atom types\n%d bond types\n\n" % (len(melt.atom_type_names), len(melt.bond_type_names)))
        for lo, hi, name in zip(-half, half, "xyz"):
            f.write("%s %s %slo %shi\n" % (repr(float(lo)), repr(float(hi)), name, name))
        f.write("\nAtoms\n\n")
        for i in range(melt.n_atoms):
            p = melt.positions[i]
            im = melt.images[i]
            f.write("%d %d %s %s %s %s %d %d %d\n" % (
                i + 1, melt.mol[i] + 1, melt.atom_type_names[melt.atom_type[i]],
                repr(float(p[0])), repr(float(p[1])), repr(float(p[2])), im[0], im[1], im[2]))
        f.write("\nBonds\n\n")
        for b in range(melt.n_bonds):
            a1, a2 = melt.bond_atoms[b]
            f.write("%d %s %d %d\n" % (b + 1, melt.bond_type_names[melt.bond_type[b]], a1 + 1, a2 + 1))


def tilt_melt(melt: Melt, tilt) -> np.ndarray:
    """Positions of ``melt`` re-wrapped into the triclinic cell with the same edge lengths and tilt factors
    ``tilt`` = (xy, xz, yz) (LAMMPS convention a = (lx,0,0), b = (xy,ly,0), c = (xz,yz,lz)): the chains are taken as
    grown (unwrapped) and every atom is mapped to fractional coordinates in [0, 1), so bonds cross the tilted faces."""
    lx, ly, lz = (float(v) for v in melt.box)
    xy, xz, yz = (float(v) for v in tilt)
    p = melt.positions + melt.images * melt.box
    sz = p[:, 2] / lz
    sy = (p[:, 1] - yz * sz) / ly
    sx = (p[:, 0] - xy * sy - xz * sz) / lx
    s = np.stack([sx, sy, sz], axis=1)
    s -= np.floor(s)
    return np.ascontiguousarray(np.stack([s[:, 0] * lx + s[:, 1] * xy + s[:, 2] * xz, s[:, 1] * ly + s[:, 2] * yz,
                                          s[:, 2] * lz], axis=1))
