"""Randomised parity sweep of the fast P2 / RDF sweeps against the C oracle: box aspect ratios, cutoffs from 3 cells
per axis to many, densities with empty and crowded cells, neighbour / reference selections, both molecule rules,
every candidate-slot variant.  Small systems, many seeds."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    box = rng.uniform(7.0, 16.0, 3)
    n = int(rng.integers(400, 6000))
    kind = seed % 3
    if kind == 0:      # uniform gas of bonds
        ctr = rng.uniform(-0.5, 0.5, (n, 3)) * box
    elif kind == 1:    # clustered: a few dense blobs + background
        centres = rng.uniform(-0.5, 0.5, (6, 3)) * box
        which = rng.integers(0, 6, n)
        ctr = centres[which] + rng.normal(0, 0.6, (n, 3))
        ctr[: n // 3] = rng.uniform(-0.5, 0.5, (n // 3, 3)) * box
    else:              # slab: many empty cells
        ctr = rng.uniform(-0.5, 0.5, (n, 3)) * box
        ctr[:, 2] *= 0.15
    ctr += rng.integers(-2, 3, (n, 3)) * box * (seed % 2)      # un-wrapped midpoints every other seed
    vec = rng.normal(size=(n, 3))
    vec *= rng.uniform(0.5, 1.2, (n, 1)) / np.linalg.norm(vec, axis=1)[:, None]
    pos = np.empty((2 * n, 3))
    pos[0::2], pos[1::2] = ctr - 0.5 * vec, ctr + 0.5 * vec
    bonds = np.stack([np.arange(0, 2 * n, 2), np.arange(1, 2 * n, 2)], axis=1).astype(np.int32)
    mol = rng.integers(0, max(2, n // 9), n).astype(np.int32)
    rc = float(rng.uniform(0.8, min(box) / 3.05))
    nbr = (rng.random(n) < 0.9).astype(np.uint8) if seed % 4 == 1 else None
    ref = ((rng.random(n) < 0.6) & (nbr.astype(bool) if nbr is not None else True)).astype(np.uint8) if seed % 4 in (1, 2) else None
    return pos, bonds, mol, box, rc, nbr, ref


@pytest.mark.parametrize("seed", range(18))
def test_random_p2_frames_match_the_c_oracle(seed):
    from mouse_b200 import engine
    from oracle import c_oracle
    engine.require_cuda()
    pos, bonds, mol, box, rc, nbr, ref = _case(seed)
    vec, ctr = c_oracle.bond_build(pos, bonds, box)
    for same in (True, False):
        ora = c_oracle.p2(vec, ctr, mol, box, rc, same, ref_mask=ref, nbr_mask=nbr)
        ok = ora["count"] > 0
        # slots = -1: the library's own choice (crowded frames leave the fast sweep, mouse_p2_uses_fast); the other
        # variants force the FP32-filter sweep: several chunks / sub-passes per home cell, oversized home cells
        for slots in (-1, 0, 4, 6, 8) if seed % 3 == 0 else (-1, 0):
            res = engine.p2_frame(pos, bonds, mol, box, rc, same_molecule=same, nbr_mask=nbr, ref_mask=ref,
                                  slots=max(slots, 0), force_fast=slots >= 0)
            assert res.grid.fast == 1
            assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"]), (seed, same, slots)
            np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=1e-14)
            assert res.totals_dict()["n_pairs"] == int(ora["count"].sum())


@pytest.mark.parametrize("seed", range(8))
def test_random_rdf_frames_match_the_c_oracle(seed):
    from mouse_b200 import engine
    from oracle import c_oracle
    engine.require_cuda()
    pos, _, mol, box, rc, _, _ = _case(seed)
    rng = np.random.default_rng(seed)
    ntypes = 1 + seed % 3
    typ = rng.integers(0, ntypes, len(pos)).astype(np.int32)
    if seed % 2:
        typ[rng.random(len(pos)) < 0.1] = -1          # atoms outside the type selection
    molp = np.repeat(mol, 2).astype(np.int32)
    rmin = 0.0 if seed % 2 == 0 else 0.3 * rc
    nbin = int(rng.integers(5, 400))
    for same in (True, False):
        h, edges = engine.rdf_frame(pos, typ, molp, ntypes, box, rmin, rc, nbin, same_molecule=same)
        exp, _ = c_oracle.rdf(pos, typ, molp, ntypes, box, rmin, rc, nbin, same)
        assert np.array_equal(h.cpu().numpy(), exp), (seed, same)
