import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mouse_b200 import engine, _lib
from mouse_b200.melt import make_melt
from oracle import c_oracle
melt = make_melt(100, 101, seed=2)
molb = melt.mol[melt.bond_atoms[:, 0]]
vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
ora = c_oracle.p2(vec, ctr, molb, melt.box, 2.5, True, want_lists=True)
g = _lib.grid_plan(melt.box, 2.5, melt.n_bonds)
n = list(g.ncell_axis); w = np.array(list(g.cell_w)); L = np.array(melt.box)
x = ctr - L * np.floor(ctr / L)
k = np.minimum((x / w).astype(int), np.array(n) - 1)
cid = (k[:, 2] * n[1] + k[:, 1]) * n[0] + k[:, 0]
order = np.lexsort((np.arange(len(cid)), cid))          # canonical: by cell, then bond index
start = np.searchsorted(cid[order], np.arange(n[0] * n[1] * n[2] + 1))
L4 = _lib.lib()
for slots in (4,):
    L4.mouse_set_tuning(b"fast_slots", slots)
    res = {}
    for variant in (2, 1):
        L4.mouse_set_tuning(b"fast_variant", variant)
        r = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5)
        res[variant] = r.count.cpu().numpy().astype(np.int64)
    bad = np.flatnonzero(res[1] != ora["count"])
    print("slots", slots, "mismatch refs", bad, "gpu", res[1][bad], "nofix", res[2][bad], "oracle", ora["count"][bad])
    for b in bad:
        kx, ky, kz = k[b]
        tlist = []
        for lane in range(27):
            dx, dy, dz = lane % 3 - 1, (lane // 3) % 3 - 1, lane // 9 - 1
            c = (((kz + dz) % n[2]) * n[1] + (ky + dy) % n[1]) * n[0] + (kx + dx) % n[0]
            tlist += list(order[start[c]:start[c + 1]])
        tlist = np.array(tlist)
        nb = ora["nbr_idx"][ora["nbr_ptr"][b]:ora["nbr_ptr"][b + 1]]
        tpos = np.array([np.flatnonzero(tlist == j)[0] for j in nb])
        NC = slots * 32
        print(" ref", b, "T", len(tlist), "self t", np.flatnonzero(tlist == b), "neighbours per pass", np.bincount(tpos // NC, minlength=6))
L4.mouse_set_tuning(b"fast_slots", 0); L4.mouse_set_tuning(b"fast_variant", 1)
