import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from mouse_b200 import engine, _lib
from tests.util import load_golden, p2_masks, reference_dense
melt, g = load_golden("config1_9900_rc15")
nbr, ref = p2_masks(melt, g)
molb = melt.mol[melt.bond_atoms[:, 0]]
res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 1.5, same_molecule=True, nbr_mask=nbr, ref_mask=ref)
n = melt.n_bonds
L = _lib.lib()
ws = engine.workspace(L.mouse_workspace_bytes(n, C.byref(res.grid)), res.sum.device)
total = int(res.count.sum())
row_ptr = torch.full((n + 1,), -7, dtype=torch.int64, device="cuda")
nb = torch.full((total,), -5, dtype=torch.int32, device="cuda")
rc = L.mouse_p2_pair_lists(engine._ptr(res.vec), engine._ptr(res.ctr), n, C.byref(res.grid), 1, engine._ptr(ws), engine._ptr(res.count), engine._ptr(row_ptr), engine._ptr(nb), total, engine._stream())
torch.cuda.synchronize()
print("rc", rc, L.mouse_last_error())
print("row_ptr", row_ptr[:5].tolist(), row_ptr[-3:].tolist())
print("nb head", nb[:40].tolist())
print("n written", int((nb != -5).sum()), "of", total)
mean, count, ptr, idx = reference_dense(g, n)
print("expected head", idx[:40].tolist())
