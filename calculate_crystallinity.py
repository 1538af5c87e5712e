#!/usr/bin/env python3
"""Command-line front end with the flags of the reference's ``calculate_crystallinity.py`` (``:9-31``): P2 of the
selected bonds against one director (a given vector or the mean bond vector).  Same selection order as the
reference (chains, residues, rectangular subcell, ``:38-69``); ``--mode histo`` prints the histogram with
``printHistogram(s, norm=True)`` (``:81``); ``--mode average`` raises the reference's own ``KeyError: 0`` (its average
branch looks up a bond number that does not exist, ``ordering_functions.py:257-261``).  The bond vectors are built by
the sm_100a K1 kernel behind ``mouse_b200.ordering_functions.CalculateCrystallinityParameter``."""
import argparse
import sys

from mouse_b200.lammpack_misc import (printHistogram, read_data_typeselect, selectByAtomFields, select_chains,
                                      select_rectangular)
from mouse_b200.lammpack_types import Vector3d
from mouse_b200.ordering_functions import CalculateCrystallinityParameter


def build_parser():
    ap = argparse.ArgumentParser(description="Calculate average value of crystallinity parameter (3cos^2(theta)-1)/2 (B200).")
    ap.add_argument("frames", metavar="LAMMPS_data", type=str, nargs="+", help="lammps data file")
    ap.add_argument("--mode", type=str, nargs=1, default=["average"],
                    help="average: return average value (default); histo: return histogram")
    ap.add_argument("--bins", type=int, nargs=1, default=[75], help="number of bins of the histogram")
    ap.add_argument("--bondtypes", type=str, nargs="+", help="types of bonds")
    ap.add_argument("--chains", type=str, nargs="*", help="chain numbers for analysis")
    ap.add_argument("--subcell", type=float, nargs="*",
                    help="rectangular subcell boundaries, normalized xmin, xmax, [ymin, ymax, [zmin, zmax]]]")
    ap.add_argument("--reference", type=float, nargs="*", help="optional reference direction vector components")
    ap.add_argument("--mol-id", type=str, nargs=1, default="resSeq",
                    help=".pdb value to identify individual molecules, resSeq (default), or chainId")
    ap.add_argument("--residue", type=str, nargs="*", help="residue names for analysis")
    ap.add_argument("--out-pdb", type=str, nargs="*",
                    help="output .pdb with atom types encoding local crystallinity value, e.g. CXX, where XX is "
                         "crystallinity value in %%")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    read_options = {"pdb": {"assignMolecules": {"type": args.mol_id, "cluster": True}}}
    for path in args.frames:
        sys.stderr.write("\r")
        frame = read_data_typeselect(path, options=read_options)
        sys.stderr.write("Initial configuration: " + str(frame.n_atom()) + " atoms\n")
        if args.chains:
            frame = select_chains(frame, args.chains)
            sys.stderr.write("Molecule selection: " + str(frame.n_atom()) + " atoms\n")
        if args.residue:
            frame = selectByAtomFields(frame, {"res_type": args.residue})
            sys.stderr.write("Residue selection: " + str(frame.n_atom()) + " atoms\n")
        # the reference fills the limits pairwise until the list runs out (:52-62): a cut is only switched on
        # once BOTH of its limits were given
        lim = [0., 1., 0., 1., 0., 1.]
        cut = [False, False, False]
        sub = args.subcell or []
        for k in range(3):
            if len(sub) >= 2 * k + 2:
                lim[2 * k], lim[2 * k + 1], cut[k] = sub[2 * k], sub[2 * k + 1], True
            else:
                if len(sub) == 2 * k + 1:
                    lim[2 * k] = sub[2 * k]
                break
        ref = args.reference or []
        reference_vector = Vector3d(ref[0], ref[1], ref[2]) if len(ref) >= 3 else Vector3d(0., 0., 0.)
        if any(cut):
            frame = select_rectangular(frame, cut[0], lim[0], lim[1], cut[1], lim[2], lim[3], cut[2], lim[4], lim[5])
            sys.stderr.write("Rectangular selection: " + str(frame.n_atom()) + " atoms\n")
        store = bool(args.out_pdb)
        s = CalculateCrystallinityParameter(frame, args.bondtypes, reference_vector=reference_vector,
                                            storeAsAtomtypes=store, mode=args.mode[0], sbins=args.bins[0])
        if store:
            frame.write_pdb(args.out_pdb[0], hide_pbc_bonds=True)
        if args.mode[0] == "average":
            sys.stdout.write(str(s) + "\n")
        if args.mode[0] == "histo":
            sys.stdout.write(printHistogram(s, norm=True))


if __name__ == "__main__":
    main()
