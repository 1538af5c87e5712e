#!/usr/bin/env python3
"""Checking aid: 10 frames of the 1M-bond melt through the frames-in-flight pipeline vs one frame at a time."""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import engine
from mouse_b200.melt import make_melt, jitter
from mouse_b200.trajectory import local_ordering_trajectory
melt = make_melt(10000, 101, seed=0)
molb = melt.mol[melt.bond_atoms[:, 0]]
frames = [jitter(melt, f, 0.03) for f in range(10)]
per, tot = local_ordering_trajectory(frames, melt.bond_atoms, molb, melt.box, 2.5, same_molecule=True)
ok = True
for f, s in per:
    r = engine.p2_frame(frames[f], melt.bond_atoms, molb, melt.box, 2.5, same_molecule=True)
    if s != r.order_parameter(): ok = False; print("MISMATCH", f, s, r.order_parameter())
print("pipeline vs serial identical:", ok, tot.n_frames, tot.n_pairs)
