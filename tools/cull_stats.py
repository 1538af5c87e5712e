#!/usr/bin/env python3
"""CPU model of the half-shell sweep's work items (no GPU): per home cell the fetched candidates T (14 cells), the
survivors M of the bounding-box cull and the references nhome, for a synthetic melt.  Used to evaluate the tiling
parameters (landing zone, register slots, chunking) before touching the kernel.
usage: cull_stats.py [chains beads rc]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mouse_b200.melt import make_melt

chains, beads, rc = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (10000, 101, 2.5)
melt = make_melt(chains, beads, seed=0)
L = melt.box
p = melt.positions
a, b = melt.bond_atoms[:, 0], melt.bond_atoms[:, 1]
d = p[b] - p[a]; d -= L * np.round(d / L)
mid = p[a] + 0.5 * d
mid -= L * np.floor(mid / L + 0.5)           # wrapped to [-L/2, L/2)
nax = np.floor(L / (rc * (1 + 1e-6))).astype(int)
cw = L / nax
ci = np.minimum(((mid + L / 2) / cw).astype(int), nax - 1)
cid = (ci[:, 2] * nax[1] + ci[:, 1]) * nax[0] + ci[:, 0]
ncell = int(np.prod(nax))
cnt = np.bincount(cid, minlength=ncell)
# bounding boxes per cell
lo = np.full((ncell, 3), 1e30); hi = np.full((ncell, 3), -1e30)
for ax in range(3):
    np.minimum.at(lo[:, ax], cid, mid[:, ax]); np.maximum.at(hi[:, ax], cid, mid[:, ax])
offs = [(0, 0, 0), (1, 0, 0)] + [(dx, dy, dz) for (dy, dz) in ((1, 0), (-1, 1), (0, 1), (1, 1)) for dx in (-1, 0, 1)]
assert len(offs) == 14
T = np.zeros(ncell, int); M = np.zeros(ncell, int)
for (dx, dy, dz) in offs:
    # candidate j in cell c_j is a candidate of home h = c_j - o
    h = np.stack([(ci[:, 0] - dx) % nax[0], (ci[:, 1] - dy) % nax[1], (ci[:, 2] - dz) % nax[2]], 1)
    hid = (h[:, 2] * nax[1] + h[:, 1]) * nax[0] + h[:, 0]
    # position of the candidate in the home cell's frame (nearest image to the home cell)
    shift = np.stack([((ci[:, 0] - dx) - h[:, 0]) , ((ci[:, 1] - dy) - h[:, 1]), ((ci[:, 2] - dz) - h[:, 2])], 1)  # multiples of nax (wraps)
    q = mid - shift / nax * L * 1.0
    e = np.maximum(np.maximum(lo[hid] - q, q - hi[hid]), 0.0)
    d2 = (e ** 2).sum(1)
    ok = cnt[hid] > 0
    keep = (d2 <= rc * rc) & ok
    T += np.bincount(hid[ok], minlength=ncell)
    M += np.bincount(hid[keep], minlength=ncell)
ne = cnt > 0
print("cells %d non-empty %d  bonds/cell %.2f" % (ncell, ne.sum(), cnt[ne].mean()))
print("T mean %.1f ref-weighted %.1f | M mean %.1f ref-weighted %.1f" % (T[ne].mean(), (T * cnt).sum() / cnt.sum(), M[ne].mean(), (M * cnt).sum() / cnt.sum()))
for q in (50, 75, 90, 95, 99, 99.9):
    print("  pct %5.1f: nhome %3d  T %4d  M %4d" % (q, np.percentile(cnt[ne], q), np.percentile(T[ne], q), np.percentile(M[ne], q)))
for lz in (192, 224, 256, 288, 320, 384):
    print("  T > %d: %.1f%% of cells" % (lz, 100.0 * (T[ne] > lz).mean()))
for nc in (128, 160, 192, 224, 256):
    print("  M > %d: %.1f%% of cells" % (nc, 100.0 * (M[ne] > nc).mean()))
np.savez("/tmp/cull_stats.npz", cnt=cnt, T=T, M=M)

def model(LZ, NC, MAXHOME, F=40, U=52, gran=64):
    """loop instructions under the kernel's chunking policy (survivors per chunk assumed proportional)"""
    it = 0; units = 0; subp = 0
    c = cnt[ne]; t = T[ne]; m = M[ne]
    for nh, tt, mm in zip(c, t, m):
        nwin = -(-nh // MAXHOME)
        for w in range(nwin):
            nref = min(MAXHOME, nh - w * MAXHOME)
            nchunk = -(-tt // LZ)
            for ch in range(nchunk):
                tl = min(LZ, tt - ch * LZ)
                ml = int(round(mm * tl / tt)) if ch else mm - int(round(mm * (tt - min(LZ, tt)) / tt))
                k = 0
                while True:
                    nv = min(NC, ml - k)
                    npp = -(-max(nv, 0) // gran)
                    subp += 1
                    it += -(-nref // 2) * (1 if npp else 0)
                    units += -(-nref // 2) * npp
                    k += NC
                    if k >= ml: break
    return subp, it, units
for (LZ, NC, MH) in ((256, 192, 32), (320, 192, 32), (256, 256, 32), (384, 256, 32)):
    s_, i_, u_ = model(LZ, NC, MH)
    print("LZ %d NC %d MAXHOME %d: sub-passes %.1fk iterations %.0fk pair-units %.0fk -> loop instr %.1fM" % (LZ, NC, MH, s_ / 1e3, i_ / 1e3, u_ / 1e3, (i_ * 40 + u_ * 52) / 1e6))
