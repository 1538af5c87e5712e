#!/usr/bin/env python3
"""Command-line front end with the flags of the reference's ``calculate_local_ordering.py``
(``:9-35``): one S (or one histogram) per input frame file on stdout, progress on stderr.
The arithmetic runs in the sm_100a kernels behind ``mouse_b200.ordering_functions``."""
import argparse
import sys

from mouse_b200 import ingest
from mouse_b200.lammpack_misc import (determine_data_type, printHistogram, read_data_typeselect, select_chains,
                                      select_rectangular)
from mouse_b200.ordering_functions import CalculateOrientationOrderParameter


def build_parser():
    ap = argparse.ArgumentParser(description="Local orientational order parameter of backbone bonds (B200).")
    ap.add_argument("frames", metavar="LAMMPS_data", type=str, nargs="+", help="frame files (LAMMPS data)")
    ap.add_argument("--bondtypes", type=str, nargs="+", help="bond types taken into account")
    ap.add_argument("--max", type=float, nargs="?", help="cutoff between bond midpoints")
    ap.add_argument("--min", type=float, nargs="?", help="accepted and ignored, as in the reference")
    ap.add_argument("--chains", type=str, nargs="+", help="molecule ids to analyse")
    ap.add_argument("--mode", type=str, nargs="?", default="average", help="average | histo")
    ap.add_argument("--histo-min", type=float, nargs="?", default=-0.5)
    ap.add_argument("--histo-max", type=float, nargs="?", default=1.)
    ap.add_argument("--histo-bins", type=int, nargs="?", default=75)
    ap.add_argument("--subcell", type=float, nargs="+", help="xmin xmax ymin ymax zmin zmax, box-relative")
    ap.add_argument("--reference-residue", type=str, nargs="+", help="residue types of the reference bonds")
    ap.add_argument("--same-molecule", action="store_true", help="also count bonds of the same molecule")
    ap.add_argument("--out-pdb", type=str, nargs="*", help="output .pdb with atom types encoding the local ordering "
                                                          "value, e.g. CXX, where XX is the value in %%")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    for path in args.frames:
        sys.stderr.write("\r")
        store = bool(args.out_pdb)
        # LAMMPS data files go straight into arrays (no per-atom objects); the object model is only built when the
        # atoms have to be recoloured for --out-pdb
        as_arrays = not store and determine_data_type(path) == "lammps_data"
        frame = ingest.read_lmp_data_arrays(path) if as_arrays else read_data_typeselect(path)
        natom = (lambda fr: fr.n_atoms) if as_arrays else (lambda fr: fr.n_atom())
        sys.stderr.write("Read frame (natom): " + str(natom(frame)) + "\n")
        if args.chains:
            frame = frame.select_chains(args.chains) if as_arrays else select_chains(frame, args.chains)
        sys.stderr.write("After molecules selection: " + str(natom(frame)) + "\n")
        if args.subcell is not None:
            lim = [0., 1., 0., 1., 0., 1.]
            cut = [False, False, False]
            for k in range(min(len(args.subcell) // 2, 3)):
                lim[2 * k], lim[2 * k + 1], cut[k] = args.subcell[2 * k], args.subcell[2 * k + 1], True
            sel = (cut[0], lim[0], lim[1], cut[1], lim[2], lim[3], cut[2], lim[4], lim[5])
            frame = frame.select_rectangular(*sel) if as_arrays else select_rectangular(frame, *sel)
        sys.stderr.write("After region selection: " + str(natom(frame)) + "\n")
        s = CalculateOrientationOrderParameter(frame, args.bondtypes, args.min, args.max, mode=args.mode,
                                               storeAsAtomtypes=store, smin=args.histo_min, smax=args.histo_max,
                                               sbins=args.histo_bins, referenceResTypes=args.reference_residue,
                                               sameMolecule=args.same_molecule)
        if store:
            frame.write_pdb(args.out_pdb[0], hide_pbc_bonds=True)     # calculate_local_ordering.py:66,74
        if args.mode == "average":
            sys.stdout.write(str(s) + "\n")
        elif args.mode == "histo":
            sys.stdout.write(printHistogram(s, norm=False))


if __name__ == "__main__":
    main()
