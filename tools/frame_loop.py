#!/usr/bin/env python3
"""Back-to-back frame loop timings (tuning aid): same buffer / cycled buffers / with sweep events."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200 import _lib, engine
from mouse_b200.melt import make_melt, jitter
melt = make_melt(10000, 101, seed=0)
dev = torch.device("cuda")
n = melt.n_bonds
frames = [torch.from_numpy(np.ascontiguousarray(jitter(melt, f, 0.02))).to(dev) for f in range(4)]
bonds = torch.as_tensor(melt.bond_atoms).to(dev)
mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
L = _lib.lib(); grid = _lib.grid_plan(melt.box, 2.5, n)
ws = engine.workspace(L.mouse_workspace_bytes(n, C.byref(grid)), dev)
vec = torch.empty((3, n), dtype=torch.float64, device=dev); ctr = torch.empty_like(vec)
ssum = torch.empty(n, dtype=torch.float64, device=dev); cnt = torch.empty(n, dtype=torch.int32, device=dev)
mean = torch.empty(n, dtype=torch.float64, device=dev); tot = torch.zeros((64, 6), dtype=torch.int64, device=dev)
p = engine._ptr; st = engine._stream(); stream = torch.cuda.current_stream()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
for a, b in evs: a.record(); b.record()
def run(cycle, events, steps=50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for k in range(steps):
        pos = frames[k % 4] if cycle else frames[0]
        ev0 = C.c_void_p(evs[k][0].cuda_event) if events else None
        ev1 = C.c_void_p(evs[k][1].cuda_event) if events else None
        L.mouse_p2_frame_timed(p(pos), p(bonds), p(mol), None, None, n, C.byref(grid), 1, p(ws), ws.numel(), p(vec), p(ctr),
                               p(ssum), p(cnt), p(mean), p(tot[k]), ev0, ev1, st)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (t1 - t0) / steps * 1e3
for cycle in (0, 1):
    for events in (0, 1):
        run(cycle, events)
        gpu, cpu = run(cycle, events)
        print("cycle=%d events=%d: %.3f ms/frame on the GPU, %.3f ms/frame of host enqueue time" % (cycle, events, gpu, cpu))
