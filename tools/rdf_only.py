"""A few RDF frames of BASELINE config 4 (ncu target).  usage: rdf_only.py [chains beads frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mouse_b200 import engine
from mouse_b200.melt import make_melt
chains, beads, frames = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40000, 100, 3)
melt = make_melt(chains, beads, seed=0)
pos = torch.as_tensor(melt.positions).cuda(); t = torch.as_tensor(melt.atom_type).cuda(); mol = torch.as_tensor(melt.mol).cuda()
for _ in range(frames):
    h, e = engine.rdf_frame(pos, t, mol, 1, melt.box, 0., 5., 1000, same_molecule=True)
torch.cuda.synchronize()
print(int(h.sum()))
