#!/usr/bin/env python3
"""One-shot runs of the other BASELINE.json configurations (not bench lines): timing + oracle spot checks."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import _lib, engine
from mouse_b200.melt import make_melt, jitter

def time_frames(melt, rc, same, nbr=None, ref=None, reps=20):
    dev = torch.device("cuda")
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = None
    for _ in range(3):
        res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=same, nbr_mask=nbr, ref_mask=ref, out=res)
    pos = torch.as_tensor(melt.positions).to(dev); bonds = torch.as_tensor(melt.bond_atoms).to(dev)
    mol = torch.as_tensor(molb.astype(np.int32)).to(dev)
    nb = None if nbr is None else torch.as_tensor(nbr).to(dev); rf = None if ref is None else torch.as_tensor(ref).to(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        res = engine.p2_frame(pos, bonds, mol, melt.box, rc, same_molecule=same, nbr_mask=nb, ref_mask=rf, out=res)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    t = res.totals_dict()
    return dt, t, res

which = sys.argv[1:] or ["3", "5"]
if "3" in which:
    melt = make_melt(2500, 101, seed=0)
    dt, t, res = time_frames(melt, 2.5, True, reps=200)
    print("config 3 (250k bonds, rc 2.5): %.3f ms/frame -> %.0f frames/s, %.3e pairs/s, S=%.6g, grid %s" % (
        dt * 1e3, 1 / dt, t["n_pairs"] / dt, res.order_parameter(), list(res.grid.ncell_axis)))
if "5" in which:
    melt = make_melt(80000, 101, seed=1, n_bond_types=2, n_res_types=3)
    nbr = (melt.bond_type == 0).astype(np.uint8)
    resb = melt.res[melt.bond_atoms[:, 0]]
    ref = (nbr.astype(bool) & (resb != 2)).astype(np.uint8)
    dt, t, res = time_frames(melt, 3.0, False, nbr, ref, reps=5)
    print("config 5 (8M bonds orthorhombic, rc 3.0, bond type + residue selections, other-molecule pairs only): "
          "%.2f ms/frame, %.3e pairs/s, refs %d, S=%.6g, grid %s, workspace %.2f GB" % (
              dt * 1e3, t["n_pairs"] / dt, t["n_refs"], res.order_parameter(), list(res.grid.ncell_axis),
              _lib.lib().mouse_workspace_bytes(melt.n_bonds, C.byref(res.grid)) / 1e9))
    from oracle import c_oracle
    vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    t0 = time.perf_counter()
    ora = c_oracle.p2(vec, ctr, molb, melt.box, 3.0, False, ref_mask=ref, nbr_mask=nbr)
    print("  C oracle (OpenMP) %.1f s; counts bit-exact: %s" % (time.perf_counter() - t0,
          bool(np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"]))))
    ok = ora["count"] > 0
    print("  max |mean - oracle| = %.3e" % float(np.nanmax(np.abs(res.mean.cpu().numpy()[ok] - ora["mean"][ok]))))
