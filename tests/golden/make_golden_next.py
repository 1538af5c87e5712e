#!/usr/bin/env python3
"""Golden vectors for the "next" rows CalculateCrystallinityParameter / CalculateComRdfs,
produced by the UNMODIFIED reference on the committed ``melt_290.data``.

    PYTHONHASHSEED=0 python tests/golden/make_golden_next.py        (needs /root/reference)
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

    def fresh():
        cfg = m.lt.Config()
        cfg.read_lmp_data(data)
        return cfg

    out = {}
    for tag, rv in (("mean", m.v3.Vector3d(0., 0., 0.)), ("z", m.v3.Vector3d(0., 0., 1.)), ("tilt", m.v3.Vector3d(0.3, -1.2, 0.5))):
        cfg = fresh()
        h = m.of.CalculateCrystallinityParameter(cfg, ['1'], reference_vector=rv, storeAsAtomtypes=True, mode='histo',
                                                 smin=-0.5, smax=1.0, sbins=40)
        out["cryst_%s_values" % tag] = np.array(h["values"])
        out["cryst_%s_norm" % tag] = np.array(h["norm"])
        out["cryst_%s_step" % tag] = np.float64(h["step"])
        out["cryst_%s_types" % tag] = np.array([a.type for a in cfg.atoms()])
    try:
        m.of.CalculateCrystallinityParameter(fresh(), ['1'], mode='average')
        out["cryst_average_error"] = np.array("none")
    except Exception as e:      # the reference crashes here (KeyError: 0)
        out["cryst_average_error"] = np.array(type(e).__name__ + ":" + str(e))
    for tag, masses in (("nomass", False), ("mass", True)):
        cfg = fresh()
        if masses:
            for k, a in enumerate(cfg.atoms()):
                a.mass = 1.0 + (k % 3) if k % 2 == 0 else 0.0
        r = m.of.CalculateComRdfs(cfg, 0., 4., 32)
        for t, row in r["data"].items():
            out["com_%s_data_%s" % (tag, t)] = np.asarray(row)
            out["com_%s_norm_%s" % (tag, t)] = np.asarray(r["norm"][t])
        out["com_%s_r" % tag] = np.array(r["r"])
        out["com_%s_v" % tag] = np.array(r["v"])
    np.savez_compressed(os.path.join(OUT, "next_290.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
