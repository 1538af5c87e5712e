#!/usr/bin/env python3
"""Command-line front end with the flags of the reference's ``calculate_com_rdfs.py`` (``:9-19``): radial
distribution function of the molecules' centres of mass, table on stdout (``printRdfs``).  The pair histogram runs in
the sm_100a RDF sweep behind ``mouse_b200.ordering_functions.CalculateComRdfs``."""
import argparse
import sys

from mouse_b200.lammpack_misc import printRdfs, read_data_typeselect
from mouse_b200.ordering_functions import CalculateComRdfs


def build_parser():
    ap = argparse.ArgumentParser(description="Calculate radial distribution functions of molecular centres of mass (B200).")
    ap.add_argument("frame", metavar="data", type=str, nargs=1, help="snapshot file")
    ap.add_argument("--rmin", type=float, nargs="?", default=0., help="minimum distance")
    ap.add_argument("--rmax", type=float, nargs="?", default=0., help="maximum distance")
    ap.add_argument("--nbin", type=int, nargs="?", default=100, help="number of bins")
    ap.add_argument("--no-norm", action="store_false", help="do not normalize the distributions")
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    sys.stderr.write("\r")
    frame = read_data_typeselect(args.frame[0])
    sys.stderr.write("Frame natom: " + str(frame.n_atom()) + " Frame nbonds: " + str(frame.n_bond()) + "\n")
    rdfs = CalculateComRdfs(frame, args.rmin, args.rmax, args.nbin)
    printRdfs(rdfs, args.no_norm)


if __name__ == "__main__":
    main()
