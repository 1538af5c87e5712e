#!/usr/bin/env python3
"""Ingestion rate of a config-3-sized LAMMPS dump (252 500 atoms per frame): parse time per frame with the indexed,
multi-threaded reader, peak RSS, and the end-to-end frames/s of local_ordering_from_dump (parse thread + pipeline)."""
import os, resource, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import ingest
from mouse_b200.frames import FrameArrays
from mouse_b200.melt import jitter, make_melt

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 12
melt = make_melt(2500, 101, seed=0)
top = FrameArrays.from_melt(melt)
half = melt.box / 2
path = os.path.join(tempfile.gettempdir(), "cfg3.dump")
t0 = time.perf_counter()
with open(path, "w") as f:
    for k in range(nframes):
        pos = jitter(melt, k, 0.02)
        xs = (pos + half) / melt.box
        f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (1000 * k, melt.n_atoms))
        for a in range(3):
            f.write("%.10g %.10g\n" % (-half[a], half[a]))
        f.write("ITEM: ATOMS id type xs ys zs ix iy iz\n")
        rows = np.column_stack([np.arange(1, melt.n_atoms + 1), np.ones(melt.n_atoms), xs, melt.images])
        np.savetxt(f, rows, fmt="%d %d %.8g %.8g %.8g %d %d %d")
print("wrote %d frames, %.1f MB in %.1f s" % (nframes, os.path.getsize(path) / 1e6, time.perf_counter() - t0))
t0 = time.perf_counter(); index = ingest.index_lmp_dump(path); print("index: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for threads in (1, 4, 8, 16):
    t0 = time.perf_counter()
    n = sum(1 for _ in ingest.iter_lmp_dump_frames(path, top.atom_num, "scaled", index=index, threads=threads))
    print("threads=%2d: %.1f ms per frame" % (threads, 1e3 * (time.perf_counter() - t0) / n))
print("peak RSS %.0f MB" % (resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e3))
try:
    import torch
    if torch.cuda.is_available():
        from mouse_b200.trajectory import local_ordering_from_dump
        for rep in range(2):      # (the first pass warms the device parser's buffers and the allocator)
            st = {}
            t0 = time.perf_counter()
            n = sum(1 for _ in ingest.iter_lmp_dump_frames(path, top.atom_num, "scaled", index=index, device="cuda", stats=st))
            torch.cuda.synchronize()
            print("device parser: %.2f ms per frame (text -> pinned -> GPU -> positions on the device) %s" % (
                1e3 * (time.perf_counter() - t0) / n, st))
        for parse_on in ("device", "host"):
            local_ordering_from_dump(top, path, ["1"], 2.5, "scaled", parse_on=parse_on)          # warm
            t0 = time.perf_counter()
            per_frame, tot = local_ordering_from_dump(top, path, ["1"], 2.5, "scaled", parse_on=parse_on)
            dt = time.perf_counter() - t0
            print("local_ordering_from_dump(parse_on=%s): %.2f ms per frame end to end (%d frames, %d pairs), peak RSS %.0f MB" % (
                parse_on, 1e3 * dt / tot.n_frames, tot.n_frames, tot.n_pairs, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e3))
except ImportError:
    pass
os.remove(path)
