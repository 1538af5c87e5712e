import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mouse_b200 import engine, _lib
from mouse_b200.melt import make_melt
chains, beads, rc = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
melt = make_melt(chains, beads, seed=0)
molb = melt.mol[melt.bond_atoms[:, 0]]
ex = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=True, exact=True, private_workspace=True)
torch.cuda.synchronize()
print("exact", ex.totals_dict(), list(ex.grid.ncell_axis))
res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=True, private_workspace=True)
torch.cuda.synchronize()
out8 = (C.c_int64 * 8)()
_lib.lib().mouse_debug_counters(C.byref(res.grid), melt.n_bonds, engine._ptr(res.ws), out8, engine._stream())
print("fast ", res.totals_dict(), "counters", list(out8))
print("count mismatch", int((res.count != ex.count).sum().item()), "max |dmean|", float((res.mean - ex.mean).abs().max().item()))
