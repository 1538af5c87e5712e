#!/usr/bin/env python3
"""Where the time of the device dump parser goes (one 252 500-atom frame, 12 MB of text): the threaded pread into the
pinned staging buffer by thread count, and the whole parse() call."""
import os, sys, tempfile, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import ingest

n = 252500
rng = np.random.default_rng(0)
path = os.path.join(tempfile.gettempdir(), "parts.dump")
with open(path, "w") as f:
    for k in range(8):
        xs = rng.random((n, 3))
        f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n-5 5\n-5 5\n-5 5\nITEM: ATOMS id type xs ys zs ix iy iz\n" % (k, n))
        np.savetxt(f, np.column_stack([np.arange(1, n + 1), np.ones(n), xs, np.zeros((n, 3))]), fmt="%d %d %.8g %.8g %.8g %d %d %d")
index = ingest.index_lmp_dump(path)
num = np.arange(1, n + 1)
with open(path, "rb") as f:
    for threads in (1, 2, 4, 8, 16):
        p = ingest.DeviceDumpParser(ingest._RowMap(num), "cuda", fd=f.fileno(), read_threads=threads)
        nbytes = int(np.diff(index).min()) - 200
        with torch.cuda.device(p.device):
            p._room(nbytes); p._stage(0, nbytes)
        for rep in range(3):
            t0 = time.perf_counter()
            for k in range(8):
                p._read(0, None, int(index[k]) + 150, nbytes)
            dt = (time.perf_counter() - t0) / 8
        print("read_threads=%2d: %.2f ms per frame (%.1f GB/s)" % (threads, 1e3 * dt, nbytes / dt / 1e9))
        p.close()
for threads in (4, 8):
    for rep in range(3):
        t0 = time.perf_counter()
        with open(path, "rb") as f:
            pass
        t0 = time.perf_counter()
        cnt = sum(1 for _ in ingest.iter_lmp_dump_frames(path, num, "scaled", index=index, device="cuda"))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / cnt
    print("iter_lmp_dump_frames(device): %.2f ms per frame" % (1e3 * dt))
    break
os.remove(path)
