"""Shared helpers for the test-suite: golden fixture loading and selection masks."""
from __future__ import annotations

import glob
import os

import numpy as np

from mouse_b200.melt import Melt

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


P2_CASES = [n for n in golden_names() if not n.startswith(("rdf_", "ingest_", "next_"))]
RDF_CASES = golden_names("rdf_")
# the Theta(N^2) literal restatement takes ~5 s on the 9 900-bond case; keep it out of the quick lists
P2_SMALL_CASES = [n for n in P2_CASES if not n.startswith("config1")]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    melt = Melt(positions=g["positions"], images=g["images"], box=g["box"], bond_atoms=g["bond_atoms"],
                bond_type=g["bond_type"], bond_type_names=[str(s) for s in g["bond_type_names"]],
                atom_type=g["atom_type"], atom_type_names=[str(s) for s in g["atom_type_names"]],
                mol=g["mol"], res=g["res"], res_type_names=[str(s) for s in g["res_type_names"]])
    return melt, g


def p2_masks(melt, g):
    """(nbr_mask, ref_mask) over bonds per ``ordering_functions.py:25-28,198-199``: an empty /
    None list passes everything; the residue filter needs BOTH atoms to match."""
    bondtypes = [str(s) for s in g["case_bondtypes"]]
    rrt = [str(s) for s in g["case_reference_res_types"]]
    nbr = np.ones(melt.n_bonds, bool)
    if bondtypes:
        ok = np.array([n in bondtypes for n in melt.bond_type_names])
        nbr = ok[melt.bond_type]
    ref = nbr.copy()
    if rrt:
        ok = np.array([n in rrt for n in melt.res_type_names])
        ref &= ok[melt.res[melt.bond_atoms[:, 0]]] & ok[melt.res[melt.bond_atoms[:, 1]]]
    return nbr, ref


def reference_dense(g, n_bonds):
    """Expand the golden per-reference lists to dense per-bond arrays in ORIGINAL bond order
    (npBonds columns hold bond.num = original index + 1).  Returns (mean, count, ptr, idx)."""
    col_to_bond = g["ref_npBonds"][0].astype(np.int64) - 1
    mean = np.full(n_bonds, np.nan)
    count = np.zeros(n_bonds, np.int64)
    lists = [np.zeros(0, np.int64)] * n_bonds
    for k, col in enumerate(g["ref_ref_index"]):
        b = col_to_bond[col]
        mean[b] = g["ref_mean"][k]
        count[b] = g["ref_count"][k]
        lists[b] = col_to_bond[g["ref_nbr_idx"][g["ref_nbr_ptr"][k]:g["ref_nbr_ptr"][k + 1]]]
    ptr = np.zeros(n_bonds + 1, np.int64)
    ptr[1:] = np.cumsum([len(l) for l in lists])
    idx = np.concatenate(lists) if n_bonds else np.zeros(0, np.int64)
    return mean, count, ptr, idx


def config_from_melt(melt, positions=None):
    """Reference-style Config (mouse_b200's own object model) built like read_lmp_data would."""
    from mouse_b200.lammpack_types import Atom, Bond, Config, Vector3d
    pos = melt.positions if positions is None else positions
    cfg = Config()
    cfg.set_box(Vector3d(float(melt.box[0]), float(melt.box[1]), float(melt.box[2])))
    names = melt.mol_id_strings()
    for i in range(melt.n_atoms):
        a = Atom()
        a.id, a.num, a.mol_id = str(i + 1), i + 1, names[i]
        a.type = melt.atom_type_names[melt.atom_type[i]]
        a.res_type = melt.res_type_names[melt.res[i]]
        a.pos = Vector3d(float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2]))
        cfg.insert_atom(a)
    for b in range(melt.n_bonds):
        bond = Bond()
        lo, hi = sorted((int(melt.bond_atoms[b, 0]) + 1, int(melt.bond_atoms[b, 1]) + 1))
        bond.atom1, bond.atom2 = cfg.atom_by_num(lo), cfg.atom_by_num(hi)
        bond.type = melt.bond_type_names[melt.bond_type[b]]
        bond.num = cfg.n_bond() + 1
        cfg.insert_bond(bond)
    return cfg
