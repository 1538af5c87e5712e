#!/usr/bin/env python3
"""Times the K3 sweep of every tuning build under tools/variants/ on BASELINE config 2 (one subprocess per build,
MOUSE_B200_LIB selects the library).  Usage: sweep_variants.py [--one] [chains beads rc]"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(chains, beads, rc):
    import numpy as np
    import torch
    from mouse_b200 import _lib, engine
    from mouse_b200.melt import make_melt
    melt = make_melt(chains, beads, seed=0)
    dev = torch.device("cuda")
    pos = torch.as_tensor(melt.positions).to(dev)
    bonds = torch.as_tensor(melt.bond_atoms).to(dev)
    mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
    n = melt.n_bonds
    L = _lib.lib()
    grid = _lib.grid_plan(melt.box, rc, n)
    ws = torch.empty(L.mouse_workspace_bytes(n, C.byref(grid)) + 4096, dtype=torch.uint8, device=dev)
    vec = torch.empty((3, n), dtype=torch.float64, device=dev)
    ctr = torch.empty_like(vec)
    ssum = torch.empty(n, dtype=torch.float64, device=dev)
    cnt = torch.empty(n, dtype=torch.int32, device=dev)
    mean = torch.empty(n, dtype=torch.float64, device=dev)
    tot = torch.zeros(6, dtype=torch.int64, device=dev)
    p, st = engine._ptr, engine._stream()
    name = os.path.basename(os.path.dirname(os.environ.get("MOUSE_B200_LIB", "default/x")))
    combos = [(same, slots, 0) for same in (True, False) for slots in (4, 6, 8)]
    if os.environ.get("SWEEP_WARPS"):      # occupancy scan of the default slot count: SWEEP_WARPS=4,6,8,10
        combos = [(True, int(os.environ.get("SWEEP_SLOTS", "6")), int(w)) for w in os.environ["SWEEP_WARPS"].split(",")]
    if os.environ.get("SWEEP_QUICK"):
        combos = [(True, int(s), 0) for s in os.environ["SWEEP_QUICK"].split(",")]
    for same, slots, warps in combos:
        if True:
            flags = (_lib.MOUSE_SAME_MOLECULE if same else 0) | _lib.MOUSE_SLOTS(slots) | _lib.MOUSE_WARPS(warps)
            e0 = [torch.cuda.Event(enable_timing=True) for _ in range(24)]
            e1 = [torch.cuda.Event(enable_timing=True) for _ in range(24)]
            f0 = [torch.cuda.Event(enable_timing=True) for _ in range(24)]
            f1 = [torch.cuda.Event(enable_timing=True) for _ in range(24)]
            for e in e0 + e1:       # torch creates the CUDA event at the first record
                e.record()
            for k in range(24):
                f0[k].record()
                rcode = L.mouse_p2_frame_timed(p(pos), p(bonds), p(mol), None, None, n, C.byref(grid), flags, p(ws), ws.numel(),
                                               p(vec), p(ctr), p(ssum), p(cnt), p(mean), p(tot),
                                               C.c_void_p(e0[k].cuda_event), C.c_void_p(e1[k].cuda_event), st)
                assert rcode == 0, L.mouse_last_error()
                f1[k].record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in zip(e0[4:], e1[4:]))
            tf = sorted(a.elapsed_time(b) for a, b in zip(f0[4:], f1[4:]))
            t = tot.cpu()
            print("%-12s same=%d slots=%d warps=%d sweep median %.4f min %.4f ms | frame median %.4f ms | pairs=%d band=%d S=%.15g" % (
                name, same, slots, warps, ts[len(ts) // 2], ts[0], tf[len(tf) // 2], int(t[2]), int(t[5]),
                1.5 * float(t.view(torch.float64)[0]) / max(int(t[1]), 1) - 0.5), flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    chains, beads, rc = (int(args[0]), int(args[1]), float(args[2])) if len(args) >= 3 else (10000, 101, 2.5)
    if "--one" in sys.argv:
        return one(chains, beads, rc)
    libs = [None] + sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "*", "libmouse_b200.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["MOUSE_B200_LIB"] = lib
        else:
            env.pop("MOUSE_B200_LIB", None)
        subprocess.call([sys.executable, os.path.abspath(__file__), "--one", str(chains), str(beads), str(rc)], env=env)


if __name__ == "__main__":
    main()
