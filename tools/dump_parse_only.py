#!/usr/bin/env python3
"""Two 252 500-atom dump frames through the device parser (ncu target for the dump_parse.cu kernels)."""
import io, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import ingest

n = 252500
rng = np.random.default_rng(0)
body = io.BytesIO()
np.savetxt(body, np.column_stack([np.arange(1, n + 1), np.ones(n), rng.random((n, 3)), np.zeros((n, 3))]),
           fmt="%d %d %.8g %.8g %.8g %d %d %d")
path = os.path.join(tempfile.gettempdir(), "ncu.dump")
with open(path, "wb") as f:
    for k in range(3):
        f.write(("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n-5 5\n-5 5\n-5 5\n"
                 "ITEM: ATOMS id type xs ys zs ix iy iz\n" % (k, n)).encode())
        f.write(body.getvalue())
for fr in ingest.iter_lmp_dump_frames(path, np.arange(1, n + 1), "scaled", device="cuda"):
    pass
os.remove(path)
