"""A few P2 frames of BASELINE config 2 (ncu target).  usage: sweep_only.py [same|diff] [frames] [slots] [chains]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mouse_b200 import engine
from mouse_b200.melt import make_melt
same = (sys.argv[1] == "same") if len(sys.argv) > 1 else True
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
slots = int(sys.argv[3]) if len(sys.argv) > 3 else 0
melt = make_melt(int(sys.argv[4]) if len(sys.argv) > 4 else 10000, 101, seed=0)
molb = melt.mol[melt.bond_atoms[:, 0]]
for _ in range(frames):
    res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5, same_molecule=same, slots=slots)
torch.cuda.synchronize()
print(res.totals_dict(), res.order_parameter())
