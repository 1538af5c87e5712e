#!/usr/bin/env python3
"""Steady-state breakdown of DeviceDumpParser.parse per 252 500-atom frame (wall clock around its stages)."""
import os, sys, tempfile, time, io
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import ingest

n = 252500
rng = np.random.default_rng(0)
path = os.path.join(tempfile.gettempdir(), "brk.dump")
body = io.BytesIO()
np.savetxt(body, np.column_stack([np.arange(1, n + 1), np.ones(n), rng.random((n, 3)), np.zeros((n, 3))]), fmt="%d %d %.8g %.8g %.8g %d %d %d")
body = body.getvalue()
with open(path, "wb") as f:
    for k in range(40):
        f.write(("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n-5 5\n-5 5\n-5 5\nITEM: ATOMS id type xs ys zs ix iy iz\n" % (k, n)).encode())
        f.write(body)
index = ingest.index_lmp_dump(path)
num = np.arange(1, n + 1)
acc = {}
def wrap(cls, name):
    fn = getattr(cls, name)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    setattr(cls, name, w)
for name in ("parse", "_run", "_read", "_prefetch", "_room"):
    wrap(ingest.DeviceDumpParser, name)
orig_hdr = ingest._parse_dump_header
def hdr(*a, **k):
    t0 = time.perf_counter()
    try:
        return orig_hdr(*a, **k)
    finally:
        acc["header"] = acc.get("header", 0.0) + time.perf_counter() - t0
ingest._parse_dump_header = hdr
it = ingest.iter_lmp_dump_frames(path, num, "scaled", index=index, device="cuda")
for _ in range(4):
    next(it)
acc.clear()
t0 = time.perf_counter()
cnt = sum(1 for _ in it)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("steady state: %.3f ms per frame over %d frames" % (1e3 * dt / cnt, cnt))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-10s %.3f ms per frame" % (k, 1e3 * v / cnt))
os.remove(path)
