#!/usr/bin/env python3
"""Per-stage CUDA-event timings of the P2 pipeline (tuning aid, not the benchmark)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import _lib, engine  # noqa: E402
from mouse_b200.melt import make_melt  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    chains, beads, rc = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    same = (sys.argv[4] == "same") if len(sys.argv) > 4 else True
    melt = make_melt(chains, beads, seed=0)
    dev = torch.device("cuda")
    pos = torch.as_tensor(melt.positions).to(dev)
    bonds = torch.as_tensor(melt.bond_atoms).to(dev)
    mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
    n = melt.n_bonds
    L = _lib.lib()
    grid = _lib.grid_plan(melt.box, rc, n)
    ws = engine.workspace(L.mouse_workspace_bytes(n, C.byref(grid)), dev)
    vec = torch.empty((3, n), dtype=torch.float64, device=dev)
    ctr = torch.empty_like(vec)
    ssum = torch.empty(n, dtype=torch.float64, device=dev)
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    mean = torch.empty(n, dtype=torch.float64, device=dev)
    tot = torch.zeros(6, dtype=torch.int64, device=dev)
    box3 = (C.c_double * 3)(*[float(b) for b in melt.box])
    p = engine._ptr
    st = engine._stream()
    flags = _lib.MOUSE_SAME_MOLECULE if same else 0
    stages = {
        "bond_build": lambda: L.mouse_bond_build(p(pos), p(bonds), n, box3, p(vec), p(ctr), st),
        "cell_build": lambda: L.mouse_cell_build(p(vec), p(ctr), p(mol), None, None, n, C.byref(grid), p(ws), ws.numel(), st),
        "p2_sweep": lambda: L.mouse_p2_sweep(p(vec), p(ctr), n, C.byref(grid), flags, p(ws), ws.numel(), p(ssum), p(cnt), st),
        "p2_reduce": lambda: L.mouse_p2_reduce(p(ssum), p(cnt), n, C.byref(grid), p(ws), ws.numel(), p(mean), p(tot), None, 0., 0., 0, st),
        "frame": lambda: L.mouse_p2_frame(p(pos), p(bonds), p(mol), None, None, n, C.byref(grid), flags, p(ws), ws.numel(),
                                          p(vec), p(ctr), p(ssum), p(cnt), p(mean), p(tot), st),
    }
    print("n_bonds=%d rc=%g grid=%s fast=%d same=%s" % (n, rc, list(grid.ncell_axis), grid.fast, same))
    for name, fn in stages.items():
        med, mn = timeit(fn)
        print("%-12s median %.3f ms  min %.3f ms" % (name, med, mn))
    pairs = int(tot.cpu()[2])
    for slots in (4, 6, 8):
        med, mn = timeit(lambda: L.mouse_p2_sweep(p(vec), p(ctr), n, C.byref(grid), flags | _lib.MOUSE_SLOTS(slots), p(ws),
                                                  ws.numel(), p(ssum), p(cnt), st))
        print("sweep slots=%-2d median %.3f ms  min %.3f ms  -> %.3e pairs/s" % (slots, med, mn, pairs / (mn * 1e-3)))
    med, mn = timeit(lambda: L.mouse_p2_sweep(p(vec), p(ctr), n, C.byref(grid), flags | _lib.MOUSE_FORCE_EXACT, p(ws),
                                              ws.numel(), p(ssum), p(cnt), st), reps=3, warm=1)
    print("sweep exact   median %.3f ms  min %.3f ms" % (med, mn))
    med, mn = timeit(stages["frame"])
    print("pairs=%d  frame %.3f ms -> %.3e pairs/s; n_band=%d" % (pairs, mn, pairs / (mn * 1e-3), int(tot.cpu()[5])))


if __name__ == "__main__":
    main()
