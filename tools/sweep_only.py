import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mouse_b200 import engine
from mouse_b200.melt import make_melt
from mouse_b200 import _lib
for key in ('fast_slots', 'fast_variant'):
    if os.environ.get('MOUSE_' + key.upper()):
        _lib.lib().mouse_set_tuning(key.encode(), int(os.environ['MOUSE_' + key.upper()]))
melt = make_melt(10000, 101, seed=0)
molb = melt.mol[melt.bond_atoms[:, 0]]
same = (sys.argv[1] == "same") if len(sys.argv) > 1 else True
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5, same_molecule=same)
torch.cuda.synchronize()
print(res.totals_dict(), res.order_parameter())
