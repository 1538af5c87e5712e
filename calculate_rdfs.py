#!/usr/bin/env python3
"""Command-line front end with the flags of the reference's ``calculate_rdfs.py`` (``:30-46``): all-pairs radial
distribution functions per atom-type pair of one snapshot, table on stdout (``printRdfs``, ``:9-28``).  The reference
script itself cannot run under Python 3 (``print >>`` at ``:50``); the pair histogram runs in the sm_100a kernels
behind ``mouse_b200.ordering_functions.CalculatePairCorrelationFunctions``."""
import argparse
import sys

from mouse_b200 import ingest
from mouse_b200.lammpack_misc import determine_data_type, printRdfs, read_data_typeselect, select_chains
from mouse_b200.ordering_functions import CalculatePairCorrelationFunctions


def build_parser():
    ap = argparse.ArgumentParser(description="Calculate radial distribution functions (B200).")
    ap.add_argument("frame", metavar="data", type=str, nargs=1, help="snapshot file")
    ap.add_argument("--atomtypes", type=str, nargs="+", help="types of atoms")
    ap.add_argument("--rmin", type=float, nargs="?", default=0., help="minimum distance")
    ap.add_argument("--rmax", type=float, nargs="?", default=0., help="maximum distance")
    ap.add_argument("--nbin", type=int, nargs="?", default=100, help="number of bins")
    ap.add_argument("--chains", type=str, nargs="+", help="chain numbers for analysis")
    ap.add_argument("--same-molecule", action="store_true", help="calculate for atoms in the same molecule")
    ap.add_argument("--no-norm", action="store_false", help="do not normalize the distributions")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    sys.stderr.write("\r")
    path = args.frame[0]
    as_arrays = determine_data_type(path) == "lammps_data"
    frame = ingest.read_lmp_data_arrays(path) if as_arrays else read_data_typeselect(path)
    natom, nbond = (frame.n_atoms, frame.n_bonds) if as_arrays else (frame.n_atom(), frame.n_bond())
    sys.stderr.write("Frame natom: " + str(natom) + " Frame nbonds: " + str(nbond) + "\n")
    if args.chains:          # (the reference reaches the same call through a TypeError when --chains is absent)
        frame = frame.select_chains(args.chains) if as_arrays else select_chains(frame, args.chains)
    rdfs = CalculatePairCorrelationFunctions(frame, args.atomtypes, args.rmin, args.rmax, args.nbin, args.same_molecule)
    printRdfs(rdfs, args.no_norm)


if __name__ == "__main__":
    main()
