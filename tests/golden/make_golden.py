#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference.

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

Needs /root/reference (build container only).  Every case stores the plain-array inputs
(positions, box, bond list, codes) and what the reference computed from a ``Config`` built
from them: ``npBonds`` (``createNumpyBondsArrayFromConfig``), per-reference mean cos^2,
neighbour counts and neighbour index lists (spy on ``np.ma.average`` inside
``calculateAveCosSqForReference``), the scalar S / histogram returned by
``CalculateOrientationOrderParameter``, and RDF counts from ``calculateRdfForReference``.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mouse_b200.melt import Melt, make_melt, write_lammps_data  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def melt_arrays(m: Melt):
    return dict(positions=m.positions, images=m.images, box=m.box, bond_atoms=m.bond_atoms,
                bond_type=m.bond_type, bond_type_names=np.array(m.bond_type_names),
                atom_type=m.atom_type, atom_type_names=np.array(m.atom_type_names),
                mol=m.mol, res=m.res, res_type_names=np.array(m.res_type_names))


def p2_case(name, m: Melt, bondtypes, rcut, same_molecule, reference_res_types=None,
            histo=None, colour=False, rmax_arg=None):
    mods = rh.load()
    cfg = rh.config_from_melt(m)
    r = rh.p2_per_vector(cfg, bondtypes, rcut, same_molecule, reference_res_types)
    out = melt_arrays(m)
    out.update(case_bondtypes=np.array(bondtypes if bondtypes else [], dtype=str),
               case_rcut=np.float64(rcut), case_same_molecule=np.bool_(same_molecule),
               case_reference_res_types=np.array(reference_res_types if reference_res_types else [], dtype=str),
               ref_npBonds=r["npBonds"], ref_ref_index=r["ref_index"], ref_mean=r["mean"],
               ref_count=r["count"], ref_nbr_ptr=r["nbr_ptr"], ref_nbr_idx=r["nbr_idx"],
               ref_skipped=r["skipped"])
    kw = dict(mode="average", sameMolecule=same_molecule)
    if reference_res_types is not None:
        kw["referenceResTypes"] = reference_res_types
    rmax = rcut if rmax_arg is None else rmax_arg
    try:
        s = mods.of.CalculateOrientationOrderParameter(cfg, bondtypes, 0., rmax, **kw)
        out["ref_S"] = np.float64(s)
    except UnboundLocalError:
        out["ref_S"] = np.float64(np.nan)
        out["ref_S_unbound"] = np.bool_(True)
    if histo is not None:
        smin, smax, sbins = histo
        try:
            h = mods.of.CalculateOrientationOrderParameter(cfg, bondtypes, 0., rmax, mode="histo", smin=smin,
                                                           smax=smax, sbins=sbins, sameMolecule=same_molecule,
                                                           **({"referenceResTypes": reference_res_types}
                                                              if reference_res_types is not None else {}))
            out.update(histo_args=np.array([smin, smax, sbins], np.float64),
                       ref_histo_values=np.array(h["values"], np.float64),
                       ref_histo_norm=np.array(h["norm"], np.float64),
                       ref_histo_step=np.float64(h["step"]),
                       ref_histo_print=np.array(mods.lm.printHistogram(h, norm=False)))
        except IndexError:
            out["ref_histo_indexerror"] = np.bool_(True)
    if colour:
        cfg2 = rh.config_from_melt(m)
        mods.of.CalculateOrientationOrderParameter(cfg2, bondtypes, 0., rmax, storeAsAtomtypes=True,
                                                   sameMolecule=same_molecule)
        out["ref_atom_types_after"] = np.array([a.type for a in cfg2.atoms()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s bonds=%6d refs=%6d pairs=%8d S=%r" % (name, m.n_bonds, len(r["ref_index"]),
                                                       int(r["count"].sum()), out["ref_S"]))


def hand_melt(positions, bonds, box, mol=None):
    positions = np.asarray(positions, np.float64)
    n = len(positions)
    bonds = np.asarray(bonds, np.int32)
    mol = np.arange(n, dtype=np.int32) // 2 if mol is None else np.asarray(mol, np.int32)
    return Melt(positions=positions, images=np.zeros((n, 3), np.int32), box=np.asarray(box, np.float64),
                bond_atoms=bonds, bond_type=np.zeros(len(bonds), np.int32),
                atom_type=np.zeros(n, np.int32), mol=mol, res=np.zeros(n, np.int32))


def kat_cases():
    # all bonds parallel -> S == 1.0 exactly; histo raises IndexError
    g = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(4), indexing="ij"), -1).reshape(-1, 3) * 2.0 - 3.5
    pos = np.concatenate([g, g + np.array([0.9, 0., 0.])])
    bonds = np.stack([np.arange(64), np.arange(64) + 64], 1)
    p2_case("kat_parallel", hand_melt(pos, bonds, [8., 8., 8.], mol=np.concatenate([np.arange(64)] * 2)),
            ["1"], 2.5, True, histo=(-0.5, 1.0, 75))
    # perpendicular pair + one tilted bond: cos^2 == 0 neighbours are dropped from the count
    pos = [[0, 0, 0], [1, 0, 0], [0.5, 1.0, -0.5], [0.5, 1.0, 0.5], [0.5, -1.2, 0.1], [1.2, -0.7, 0.3],
           [3., 3., 3.], [3., 4., 3.]]
    p2_case("kat_perpendicular", hand_melt(pos, [[0, 1], [2, 3], [4, 5], [6, 7]], [10., 10., 10.]),
            ["1"], 2.0, True)
    # midpoints exactly rc apart (inclusive <=), and one ulp beyond
    rc = 1.5
    pos = [[-0.5, 0, 0], [0.5, 0, 0], [-0.5, rc, 0.25], [0.5, rc, -0.25],
           [-0.5, -np.nextafter(rc, 2.), 0.25], [0.5, -np.nextafter(rc, 2.), -0.25],
           [3.0, 0.3, 0.0], [3.0, -0.3, 0.8]]
    p2_case("kat_exact_cutoff", hand_melt(pos, [[0, 1], [2, 3], [4, 5], [6, 7]], [9., 9., 9.]),
            ["1"], rc, True)
    # bonds crossing the periodic boundary: vector minimum-imaged, midpoint not wrapped
    pos = [[4.7, 0, 0], [-4.6, 0.2, 0.1], [4.9, 0.5, 0.5], [-4.9, 0.9, 0.2], [0.1, 4.8, 0], [0.3, -4.7, 0.4],
           [-4.5, 0.1, 4.9], [-4.4, 0.6, -4.8]]
    p2_case("kat_boundary", hand_melt(pos, [[0, 1], [2, 3], [4, 5], [6, 7]], [10., 10., 10.]),
            ["1"], 3.0, True)
    # rmax == 0 -> half box diagonal -> all pairs (ordering_functions.py:185-186)
    m = make_melt(4, 12, seed=5)
    cfg_box = np.sqrt((m.box**2).sum()) / 2.
    p2_case("kat_rmax_default", m, ["1"], cfg_box, True, rmax_arg=0.)
    # zero-length bonds: reference bond skipped, neighbour poisons with NaN
    pos = [[0, 0, 0], [1, 0, 0], [0.5, 1.0, 0.2], [0.5, 1.0, 0.2], [0.2, -1.0, 0.1], [1.0, -0.7, 0.3]]
    p2_case("kat_zero_length", hand_melt(pos, [[0, 1], [2, 3], [4, 5]], [10., 10., 10.]), ["1"], 2.0, True)
    # nothing in range anywhere -> UnboundLocalError in the reference
    pos = [[0, 0, 0], [1, 0, 0], [4, 4, 4], [4, 5, 4]]
    p2_case("kat_no_neighbours", hand_melt(pos, [[0, 1], [2, 3]], [12., 12., 12.]), ["1"], 1.0, True)


def rdf_case(name, m: Melt, atomtypes, rmin, rmax, nbin, same_molecule):
    cfg = rh.config_from_melt(m)
    types, data = rh.rdf_counts(cfg, atomtypes, rmin, rmax, nbin, same_molecule)
    out = melt_arrays(m)
    out.update(case_atomtypes=np.array(atomtypes if atomtypes else [], dtype=str),
               case_rmin=np.float64(rmin), case_rmax=np.float64(rmax), case_nbin=np.int64(nbin),
               case_same_molecule=np.bool_(same_molecule), ref_types=np.array(types),
               ref_keys=np.array(sorted(data)), ref_counts=np.stack([data[k] for k in sorted(data)]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%-28s atoms=%6d keys=%s total=%d" % (name, m.n_atoms, sorted(data),
                                                int(sum(v.sum() for v in data.values()))))


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":
        raise SystemExit("run with PYTHONHASHSEED=0")
    m = make_melt(10, 30, seed=1)
    p2_case("melt_290_rc15_same", m, ["1"], 1.5, True, histo=(-0.5, 1.0, 75), colour=True)
    p2_case("melt_290_rc15_diffmol", m, ["1"], 1.5, False)
    m = make_melt(20, 50, seed=0)
    p2_case("melt_980_rc15_same", m, ["1"], 1.5, True)
    p2_case("melt_980_rc25_diffmol", m, ["1"], 2.5, False, histo=(-0.5, 1.0, 40))
    m = make_melt(24, 41, seed=3, aspect=(1.0, 1.4, 0.8), n_bond_types=2, n_res_types=3)
    p2_case("ortho_960_types_sel", m, ["1"], 1.8, False, reference_res_types=["R0", "R2"])
    p2_case("ortho_960_alltypes", m, [], 2.2, True)
    m = make_melt(100, 100, seed=0)
    p2_case("config1_9900_rc15", m, ["1"], 1.5, True)
    write_lammps_data(make_melt(10, 30, seed=1), os.path.join(OUT, "melt_290.data"))
    kat_cases()
    m = make_melt(20, 50, seed=7, n_atom_types=2)
    rdf_case("rdf_1000_2types_5s_1000", m, [], 0., 5., 1000, True)
    rdf_case("rdf_1000_2types_diffmol", m, [], 0., 4., 100, False)
    rdf_case("rdf_1000_typesel_rmin", m, ["2"], 0.5, 3.0, 50, True)
    # lattice: many distances fall exactly on bin edges
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3) * 1.0 - 4.0
    lat = Melt(positions=g.astype(np.float64), images=np.zeros((512, 3), np.int32), box=np.array([8., 8., 8.]),
               bond_atoms=np.zeros((0, 2), np.int32), bond_type=np.zeros(0, np.int32),
               atom_type=np.zeros(512, np.int32), mol=np.arange(512, dtype=np.int32), res=np.zeros(512, np.int32))
    rdf_case("rdf_lattice_512_edges", lat, [], 0., 4., 16, True)


if __name__ == "__main__":
    main()
