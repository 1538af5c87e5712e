#!/usr/bin/env python3
"""Tuning aid: frames alternating over S streams (own workspace and outputs each) vs one stream."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import _lib, engine
from mouse_b200.melt import make_melt, jitter

chains, beads, rc = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
melt = make_melt(chains, beads, seed=0)
dev = torch.device("cuda")
n = melt.n_bonds
L = _lib.lib()
grid = _lib.grid_plan(melt.box, rc, n)
bonds = torch.as_tensor(melt.bond_atoms).to(dev)
mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
frames = [torch.as_tensor(jitter(melt, f) if f else melt.positions).to(dev) for f in range(4)]
p = engine._ptr
nbytes = L.mouse_workspace_bytes(n, C.byref(grid))

def lane():
    return dict(ws=torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=dev), vec=torch.empty((3, n), dtype=torch.float64, device=dev),
                ctr=torch.empty((3, n), dtype=torch.float64, device=dev), s=torch.empty(n, dtype=torch.float64, device=dev),
                c=torch.empty(n, dtype=torch.int32, device=dev), m=torch.empty(n, dtype=torch.float64, device=dev),
                t=torch.zeros(6, dtype=torch.int64, device=dev), st=torch.cuda.Stream(device=dev))

for key in ('fast_slots', 'fast_warps'):
    if os.environ.get('MOUSE_' + key.upper()):
        L.mouse_set_tuning(key.encode(), int(os.environ['MOUSE_' + key.upper()]))
for S in (1, 3, 4):
    lanes = [lane() for _ in range(S)]
    def run(K):
        for k in range(K):
            l = lanes[k % S]
            rcode = L.mouse_p2_frame(p(frames[k % 4]), p(bonds), p(mol), None, None, n, C.byref(grid), _lib.MOUSE_SAME_MOLECULE,
                                     p(l["ws"]), l["ws"].numel(), p(l["vec"]), p(l["ctr"]), p(l["s"]), p(l["c"]), p(l["m"]),
                                     p(l["t"]), C.c_void_p(l["st"].cuda_stream))
            _lib.check(rcode, "mouse_p2_frame")
    run(400)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 1000
    run(K)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print("streams=%d: %.4f ms per frame (n_pairs %d)" % (S, dt * 1e3, int(lanes[0]["t"][2])))
