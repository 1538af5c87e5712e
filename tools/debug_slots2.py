import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mouse_b200 import engine, _lib
from mouse_b200.melt import make_melt
from oracle import c_oracle
melt = make_melt(100, 101, seed=2)
molb = melt.mol[melt.bond_atoms[:, 0]]
vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
ora = c_oracle.p2(vec, ctr, molb, melt.box, 2.5, True)
for rep in range(3):
    for slots in (0, 4, 8, 4, 12, 4, 16, 4):
        _lib.lib().mouse_set_tuning(b"fast_slots", slots)
        r = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5)
        c = r.count.cpu().numpy().astype(np.int64)
        bad = np.flatnonzero(c != ora["count"])
        s = r.sum.cpu().numpy()
        print("rep", rep, "slots", slots, "mismatches", len(bad), "band", r.totals_dict()["n_band"], bad[:5], (c - ora["count"])[bad[:5]], "nan sums", int(np.isnan(s).sum()))
