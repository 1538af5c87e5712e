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


@pytest.mark.gpu
@pytest.mark.parametrize("rc", [4.3, 5.0])
def test_crowded_cells_stay_on_the_fast_path(rc):
    """a melt at a large cutoff has ~100 bonds per cell and a few cells far above that; the fast sweep takes
    them in windows of references (no pair-by-pair fallback, no pair-list overflow) and agrees with the all-FP64 sweep."""
    import ctypes as C
    import torch
    from mouse_b200 import _lib, engine
    from mouse_b200.melt import make_melt
    melt = make_melt(600, 101, seed=3)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    ex = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=True, exact=True,
                         private_workspace=True)
    res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=True, private_workspace=True)
    out = (C.c_int64 * 8)()
    _lib.check(_lib.lib().mouse_debug_counters(C.byref(res.grid), melt.n_bonds, engine._ptr(res.ws), out, engine._stream()),
               "debug_counters")
    assert res.grid.fast == 1 and out[4] == 0 and out[3] < 2000      # fast path, pair list far from overflowing
    assert torch.equal(res.count, ex.count)
    assert float((res.mean - ex.mean).abs().max().item()) < 1e-13
    assert res.totals_dict()["n_pairs"] == ex.totals_dict()["n_pairs"]


@pytest.mark.gpu
def test_single_cell_frame_is_ranked_in_parallel():
    """rmax == 0 / no cutoff puts every bond into one cell; the canonical in-cell order comes from
    cell_rank_kernel (one thread per item), not from one warp ranking the cell serially."""
    import time
    import torch
    from mouse_b200 import engine
    from mouse_b200.melt import make_melt
    melt = make_melt(200, 101, seed=1)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, -1.0)
    torch.cuda.synchronize()
    t0 = time.time()
    res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, -1.0)
    torch.cuda.synchronize()
    assert time.time() - t0 < 0.1                 # 20 000 bonds, 4e8 pairs: 6 ms on a B200 (140 ms with the serial ranking)
    t = res.totals_dict()
    assert t["n_pairs"] == melt.n_bonds * (melt.n_bonds - 1)
    assert int(res.count.min().item()) == melt.n_bonds - 1 and int(res.count.max().item()) == melt.n_bonds - 1


@pytest.mark.gpu
def test_very_crowded_grid_takes_the_fast_path_by_itself():
    """343 bonds per cell (rc = 7 sigma on a 60k-bond melt): the library picks the fast sweep (window rows), the result
    equals the all-FP64 sweep."""
    import torch
    from mouse_b200 import engine
    from mouse_b200.melt import make_melt
    melt = make_melt(600, 101, seed=5)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 7.0, same_molecule=False)
    ex = engine.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 7.0, same_molecule=False, exact=True)
    assert res.grid.fast == 1 and res.totals_dict()["n_band"] > 0          # the fast path ran (it has a guard band)
    assert torch.equal(res.count, ex.count)
    assert float((res.mean - ex.mean).abs().max().item()) < 1e-13
