#!/usr/bin/env python3
"""cProfile of the device dump parser over a 24-frame, 252 500-atom dump (steady state: frames 2..)."""
import cProfile, os, pstats, sys, tempfile
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import ingest

n = 252500
rng = np.random.default_rng(0)
path = os.path.join(tempfile.gettempdir(), "prof.dump")
with open(path, "w") as f:
    xs = rng.random((n, 3))
    rows = np.column_stack([np.arange(1, n + 1), np.ones(n), xs, np.zeros((n, 3))])
    for k in range(24):
        f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n-5 5\n-5 5\n-5 5\nITEM: ATOMS id type xs ys zs ix iy iz\n" % (k, n))
        np.savetxt(f, rows, fmt="%d %d %.8g %.8g %.8g %d %d %d")
index = ingest.index_lmp_dump(path)
num = np.arange(1, n + 1)
it = ingest.iter_lmp_dump_frames(path, num, "scaled", index=index, device="cuda")
next(it); next(it)
pr = cProfile.Profile()
pr.enable()
cnt = sum(1 for _ in it)
torch.cuda.synchronize()
pr.disable()
print("frames profiled:", cnt)
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
os.remove(path)
