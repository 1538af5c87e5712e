#!/usr/bin/env python3
"""Golden vectors for the ingestion row, produced by the UNMODIFIED reference readers.

    PYTHONHASHSEED=0 python tests/golden/make_golden_ingest.py

Writes ``melt_290.dump`` (three frames derived from ``melt_290.data``: scaled, cut and un-cut coordinate columns,
image flags, shuffled atom order, a box that changes from frame to frame) and ``ingest_290.npz`` holding what
``Config.read_lmp_data`` and ``Config.return_extra_frames_from_lmp_dump`` make of them.  Needs /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    m = rh.load()
    data = os.path.join(OUT, "melt_290.data")
    cfg = m.lt.Config()
    cfg.read_lmp_data(data)
    atoms, bonds = cfg.atoms(), cfg.bonds()
    base = np.array([[a.pos.x, a.pos.y, a.pos.z] for a in atoms])
    img = np.array([[a.pbc.x, a.pbc.y, a.pbc.z] for a in atoms], dtype=int)
    nums = np.array([a.num for a in atoms])
    box0 = np.array([cfg.box().x, cfg.box().y, cfg.box().z])
    rng = np.random.default_rng(7)
    dump = os.path.join(OUT, "melt_290.dump")
    with open(dump, "w") as f:
        for k in range(3):
            scale = 1.0 + 0.01 * k
            L = box0 * scale
            lo, hi = -0.5 * L + 0.1 * k, 0.5 * L + 0.1 * k
            p = base * scale + 0.1 * k + rng.normal(0, 0.02, base.shape)
            p = lo + np.mod(p - lo, L)                      # wrapped into [lo, hi)
            order = rng.permutation(len(atoms))
            f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (1000 * k, len(atoms)))
            for a in range(3):
                f.write("%.16e %.16e\n" % (lo[a], hi[a]))
            f.write("ITEM: ATOMS id type xs ys zs xc yc zc xu yu zu ix iy iz\n")
            for i in order:
                xs = (p[i] - lo) / L
                xc = p[i] - lo
                xu = p[i] + img[i] * L
                f.write("%d 1 %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e %.16e %d %d %d\n" % (
                    nums[i], xs[0], xs[1], xs[2], xc[0], xc[1], xc[2], xu[0], xu[1], xu[2], img[i, 0], img[i, 1], img[i, 2]))
    out = dict(data_pos=base, data_box=box0, data_num=nums, data_mol=np.array([a.mol_id for a in atoms]),
               data_type=np.array([a.type for a in atoms]),
               bond_atoms=np.array([[b.atom1.num, b.atom2.num] for b in bonds]),
               bond_num=np.array([b.num for b in bonds]), bond_type=np.array([b.type for b in bonds]))
    for ct in ("scaled", "cut", "uncut"):
        frames = cfg.return_extra_frames_from_lmp_dump(dump, coords_type=ct)
        # the reference appends a frame only when the NEXT timestep header is read: the last frame is dropped
        out["dump_%s_pos" % ct] = np.array([[[a.pos.x, a.pos.y, a.pos.z] for a in fr.atoms()] for fr in frames])
        out["dump_%s_box" % ct] = np.array([[fr.box().x, fr.box().y, fr.box().z] for fr in frames])
        out["dump_%s_timestep" % ct] = np.array([fr.timestep for fr in frames])
    np.savez_compressed(os.path.join(OUT, "ingest_290.npz"), **out)
    print("wrote", dump, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
