#!/usr/bin/env python3
"""Scheduling experiment (staged C-ABI calls): does it pay to run the builds of G frames together before their G sweeps?
mode 0: every frame build -> sweep -> reduce on its lane (what the fused frame call does); mode G: the builds of G frames
are submitted first (own streams), every sweep of the group waits for ALL of them, then the G sweeps + reductions."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import _lib, engine
from mouse_b200.melt import make_melt, jitter

chains, beads, rc = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (10000, 101, 2.5)
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
box3 = (C.c_double * 3)(*[float(b) for b in melt.box])
flags = _lib.MOUSE_SAME_MOLECULE

def lane():
    return dict(ws=torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=dev), vec=torch.empty((3, n), dtype=torch.float64, device=dev),
                ctr=torch.empty((3, n), dtype=torch.float64, device=dev), s=torch.empty(n, dtype=torch.float64, device=dev),
                c=torch.empty(n, dtype=torch.int32, device=dev), m=torch.empty(n, dtype=torch.float64, device=dev),
                t=torch.zeros(6, dtype=torch.int64, device=dev), st=torch.cuda.Stream(device=dev), built=torch.cuda.Event())

def build(l, k):
    st = C.c_void_p(l["st"].cuda_stream)
    _lib.check(L.mouse_bond_build(p(frames[k % 4]), p(bonds), n, box3, p(l["vec"]), p(l["ctr"]), st), "bond_build")
    _lib.check(L.mouse_cell_build(p(l["vec"]), p(l["ctr"]), p(mol), None, None, n, C.byref(grid), p(l["ws"]), l["ws"].numel(), st), "cell_build")

def sweep(l):
    st = C.c_void_p(l["st"].cuda_stream)
    _lib.check(L.mouse_p2_sweep(p(l["vec"]), p(l["ctr"]), n, C.byref(grid), flags, p(l["ws"]), l["ws"].numel(), p(l["s"]), p(l["c"]), st), "sweep")
    _lib.check(L.mouse_p2_reduce(p(l["s"]), p(l["c"]), n, C.byref(grid), p(l["ws"]), l["ws"].numel(), p(l["m"]), p(l["t"]), None, 0.0, 0.0, 0, st), "reduce")

for G, S in ((0, 3), (0, 4), (2, 4), (2, 6), (3, 6), (4, 8)):
    lanes = [lane() for _ in range(S)]
    def run(K):
        k = 0
        while k < K:
            if G == 0:
                l = lanes[k % S]; build(l, k); sweep(l); k += 1
            else:
                grp = [lanes[(k + j) % S] for j in range(G)]
                for j, l in enumerate(grp):
                    build(l, k + j); l["built"].record(l["st"])
                for l in grp:
                    for o in grp:
                        if o is not l: l["st"].wait_event(o["built"])
                    sweep(l)
                k += G
    run(240); torch.cuda.synchronize()
    t0 = time.perf_counter(); K = 720; run(K); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print("group=%d lanes=%d: %.4f ms per frame (pairs %d)" % (G, S, dt * 1e3, int(lanes[0]["t"][2])), flush=True)
    del lanes
