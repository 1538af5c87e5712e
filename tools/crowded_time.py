"""Crowded-cell / few-cell cases: wall time of one frame.  usage: crowded_time.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mouse_b200 import engine
from mouse_b200.melt import make_melt
for chains, beads, rc in ((2500, 101, 4.3), (2500, 101, 4.8), (2500, 101, 5.0), (2500, 101, 7.0), (200, 101, -1.0), (500, 101, -1.0), (1000, 101, -1.0)):
    melt = make_melt(chains, beads, seed=0)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    out = []
    for exact in (False, True):
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=True, exact=exact)
            torch.cuda.synchronize(); t1 = time.time()
        out.append(((t1 - t0) * 1e3, res))
    same = bool((out[0][1].count == out[1][1].count).all().item())
    print("bonds %d rc %.1f grid %s: default %.2f ms, all-FP64 %.2f ms, counts equal %s  %s" % (
        melt.n_bonds, rc, list(out[0][1].grid.ncell_axis), out[0][0], out[1][0], same, out[0][1].totals_dict()), flush=True)
