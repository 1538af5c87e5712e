import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mouse_b200 import engine
from mouse_b200.melt import make_melt
chains, beads = int(sys.argv[1]), int(sys.argv[2])
melt = make_melt(chains, beads, seed=0)
pos = torch.as_tensor(melt.positions).cuda(); t = torch.as_tensor(melt.atom_type).cuda(); mol = torch.as_tensor(melt.mol).cuda()
for same in (True, False):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h, e = engine.rdf_frame(pos, t, mol, 1, melt.box, 0., 5., int(os.environ.get("NBIN", "1000")), same_molecule=same)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rdf n=%d same=%s: %.2f ms, pairs %d -> %.3e pairs/s" % (melt.n_atoms, same, dt * 1e3, int(h.sum()), int(h.sum()) / dt))
