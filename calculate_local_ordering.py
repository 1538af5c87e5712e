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


def load_frame(path, args):
    """Reads one frame file and applies the selections (``calculate_local_ordering.py:39-70`` of the reference);
    returns (frame, progress lines for stderr).  LAMMPS data files go straight into arrays (no per-atom objects);
    the object model is only built when the atoms have to be recoloured for --out-pdb."""
    store = bool(args.out_pdb)
    as_arrays = not store and determine_data_type(path) == "lammps_data"
    frame = ingest.read_lmp_data_arrays(path) if as_arrays else read_data_typeselect(path)
    natom = (lambda fr: fr.n_atoms) if as_arrays else (lambda fr: fr.n_atom())
    log = ["\r", "Read frame (natom): " + str(natom(frame)) + "\n"]
    if args.chains:
        frame = frame.select_chains(args.chains) if as_arrays else select_chains(frame, args.chains)
    log.append("After molecules selection: " + str(natom(frame)) + "\n")
    if args.subcell is not None:
        lim = [0., 1., 0., 1., 0., 1.]
        cut = [False, False, False]
        for k in range(min(len(args.subcell) // 2, 3)):
            lim[2 * k], lim[2 * k + 1], cut[k] = args.subcell[2 * k], args.subcell[2 * k + 1], True
        sel = (cut[0], lim[0], lim[1], cut[1], lim[2], lim[3], cut[2], lim[4], lim[5])
        frame = frame.select_rectangular(*sel) if as_arrays else select_rectangular(frame, *sel)
    log.append("After region selection: " + str(natom(frame)) + "\n")
    return frame, log


def process_frame(frame, args, write_pdb):
    """One frame through the kernels; returns the text the reference prints for it on stdout."""
    store = bool(args.out_pdb)
    s = CalculateOrientationOrderParameter(frame, args.bondtypes, args.min, args.max, mode=args.mode,
                                           storeAsAtomtypes=store, smin=args.histo_min, smax=args.histo_max,
                                           sbins=args.histo_bins, referenceResTypes=args.reference_residue,
                                           sameMolecule=args.same_molecule)
    if store and write_pdb:
        frame.write_pdb(args.out_pdb[0], hide_pbc_bonds=True)     # calculate_local_ordering.py:66,74
    if args.mode == "average":
        return str(s) + "\n"
    elif args.mode == "histo":
        return printHistogram(s, norm=False)
    return ""


def main(argv=None):
    """The reference loops over the frame files one after the other (``calculate_local_ordering.py:39-76``).  Here
    the files are block-sharded over the ranks when the script runs under ``torchrun`` (one process per GPU; put ``--``
    between the script path and its arguments: the launcher rejects ``--max`` as an abbreviation of its own options), every
    rank reads file k+1 on a host thread while file k is on its GPU, and rank 0 prints the per-frame results in the
    order of the command line - the reference's stdout, byte for byte, whatever the number of ranks."""
    import os
    import queue
    import threading
    args = build_parser().parse_args(argv)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    files = list(args.frames)
    lo, hi = (rank * len(files)) // world, ((rank + 1) * len(files)) // world
    q = queue.Queue(maxsize=2)

    def produce():
        try:
            for k in range(lo, hi):
                q.put((k,) + load_frame(files[k], args))
            q.put(None)
        except BaseException as e:
            q.put(e)

    threading.Thread(target=produce, daemon=True).start()
    mine = {}
    while True:
        item = q.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        k, frame, log = item
        for line in log:
            sys.stderr.write(line)
        # (the reference rewrites --out-pdb for every frame: the last frame's file is what remains)
        text = process_frame(frame, args, write_pdb=(world == 1 or k == len(files) - 1))
        if world == 1:
            sys.stdout.write(text)
        else:
            mine[k] = text
    if world > 1:
        everything = [None] * world
        dist.all_gather_object(everything, mine)
        if rank == 0:
            merged = {}
            for part in everything:
                merged.update(part)
            for k in range(len(files)):
                sys.stdout.write(merged[k])
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
