#!/usr/bin/env python3
"""Soak test (not part of the suite): many random P2 / RDF frames against the C oracle, plus repeated runs of one big
frame checking run-to-run bit-identical results."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_stress import _case
from mouse_b200 import engine
from mouse_b200.melt import make_melt
from oracle import c_oracle
bad = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    pos, bonds, mol, box, rc, nbr, ref = _case(100 + seed)
    vec, ctr = c_oracle.bond_build(pos, bonds, box)
    for same in (True, False):
        ora = c_oracle.p2(vec, ctr, mol, box, rc, same, ref_mask=ref, nbr_mask=nbr)
        res = engine.p2_frame(pos, bonds, mol, box, rc, same_molecule=same, nbr_mask=nbr, ref_mask=ref)
        ok = ora["count"] > 0
        if not np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"]) or \
           not np.allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=1e-14):
            bad += 1
            print("MISMATCH p2 seed", seed, same)
    molp = np.repeat(mol, 2).astype(np.int32)
    typ = np.zeros(len(pos), np.int32)
    h, _ = engine.rdf_frame(pos, typ, molp, 1, box, 0., rc, 64, same_molecule=False)
    exp, _ = c_oracle.rdf(pos, typ, molp, 1, box, 0., rc, 64, False)
    if not np.array_equal(h.cpu().numpy(), exp):
        bad += 1
        print("MISMATCH rdf seed", seed)
melt = make_melt(10000, 101, seed=0)
molb = melt.mol[melt.bond_atoms[:, 0]]
first = None
for rep in range(20):
    r = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5)
    s, c = r.sum.clone(), r.count.clone()
    if first is None:
        first = (s, c)
    elif not (torch.equal(first[0], s) and torch.equal(first[1], c)):
        bad += 1
        print("run-to-run difference at repetition", rep)
print("soak done, mismatches:", bad)
