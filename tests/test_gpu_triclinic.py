"""Triclinic boxes on the GPU (BASELINE config 5).  BUILD-DEFINED semantics, parity unpinned (the reference is
orthorhombic only, lammpack_types.py:440-451): the checker is the repo's own restatement (oracle/triclinic_oracle.py,
oracle/c ``*_tri``), itself pinned in tests/test_triclinic_cpu.py.  tilt = 0 through the tilt-aware code paths must
reproduce the orthorhombic results bit for bit."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TILTS = [(1.3, 0.0, 0.0), (2.1, -1.7, 0.9), (-3.0, 2.5, -2.0)]


def _eng():
    from mouse_b200 import engine
    engine.require_cuda()
    return engine


@pytest.mark.parametrize("exact", [False, True])
def test_zero_tilt_through_the_triclinic_paths_is_bit_identical(exact):
    from mouse_b200.melt import make_melt
    eng = _eng()
    melt = make_melt(60, 81, seed=3)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    for same in (True, False):
        a = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.0, same_molecule=same, exact=exact)
        av, ac, asum, acnt = a.vec.clone(), a.ctr.clone(), a.sum.clone(), a.count.clone()
        b = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.0, same_molecule=same, exact=exact,
                         tilt=(0.0, 0.0, 0.0))
        assert b.grid.triclinic == 1 and b.grid.fast == 1
        assert torch.equal(av, b.vec) and torch.equal(ac, b.ctr)
        assert torch.equal(acnt, b.count)
        if exact:
            assert torch.equal(asum, b.sum)
        else:   # the triclinic guard band is three times wider: a few more pairs take the exact FP64 route, whose
            # cos^2 (true divisions) differs from the unit-vector product in the last bit
            torch.testing.assert_close(asum, b.sum, rtol=1e-13, atol=1e-13)
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    h0, _ = eng.rdf_frame(melt.positions, typ, melt.mol, 2, melt.box, 0.0, 2.5, 80, same_molecule=False, exact=exact)
    h1, _ = eng.rdf_frame(melt.positions, typ, melt.mol, 2, melt.box, 0.0, 2.5, 80, same_molecule=False, exact=exact,
                          tilt=(0.0, 0.0, 0.0))
    assert torch.equal(h0, h1)


@pytest.mark.parametrize("tilt", TILTS)
def test_tilted_melt_matches_the_triclinic_oracle(tilt):
    from mouse_b200.melt import make_melt, tilt_melt
    from oracle import c_oracle
    eng = _eng()
    melt = make_melt(120, 61, seed=8)                 # box edge ~20.5: 8 cells of 2.5 per axis
    pos = tilt_melt(melt, tilt)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    vec, ctr = c_oracle.bond_build(pos, melt.bond_atoms, melt.box, tilt=tilt)
    assert np.abs(np.linalg.norm(vec, axis=1) - 0.97).max() < 1e-9
    rng = np.random.default_rng(2)
    nbr = (rng.random(len(vec)) < 0.9).astype(np.uint8)
    ref = (nbr.astype(bool) & (rng.random(len(vec)) < 0.7)).astype(np.uint8)
    for same in (True, False):
        ora = c_oracle.p2(vec, ctr, molb, melt.box, 2.4, same, ref_mask=ref, nbr_mask=nbr, tilt=tilt)
        ok = ora["count"] > 0
        for exact in (False, True):
            res = eng.p2_frame(pos, melt.bond_atoms, molb, melt.box, 2.4, same_molecule=same, nbr_mask=nbr,
                               ref_mask=ref, exact=exact, tilt=tilt)
            assert res.grid.fast == 1 and res.grid.triclinic == 1
            assert np.array_equal(res.vec.cpu().numpy().T, vec) and np.array_equal(res.ctr.cpu().numpy().T, ctr)
            assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"]), (tilt, same, exact)
            np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=1e-14)
        s_ref = 1.5 * ora["mean"][ok].sum() / ok.sum() - 0.5
        assert abs(res.order_parameter() - s_ref) <= 1e-5 * abs(s_ref)
    # ignoring the tilt (what the reference's reader does) gives a different answer: the tilt is really used
    flat = eng.p2_frame(pos, melt.bond_atoms, molb, melt.box, 2.4, same_molecule=True, nbr_mask=nbr, ref_mask=ref)
    assert not np.array_equal(flat.count.cpu().numpy(), ora["count"])
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    for same in (True, False):
        exp, _ = c_oracle.rdf(pos, typ, melt.mol, 2, melt.box, 0.3, 2.4, 90, same, tilt=tilt)
        for exact in (False, True):
            h, _ = eng.rdf_frame(pos, typ, melt.mol, 2, melt.box, 0.3, 2.4, 90, same_molecule=same, exact=exact,
                                 tilt=tilt)
            assert np.array_equal(h.cpu().numpy(), exp), (tilt, same, exact)


def test_small_tilted_box_against_the_all_pairs_restatement():
    """Three cells per axis (every neighbour cell is a periodic image) and a strong tilt; the all-pairs NumPy
    restatement is the checker.  A box too skewed for the FP32 shift falls back to the exact sweep by itself."""
    from mouse_b200.melt import make_melt, tilt_melt
    from oracle import triclinic_oracle as tri
    eng = _eng()
    melt = make_melt(14, 40, seed=4)                  # box edge ~8.7
    molb = melt.mol[melt.bond_atoms[:, 0]]
    for tilt, want_fast in (((0.4, -0.3, 0.2), 1), ((4.0, -3.5, 3.0), 0)):
        pos = tilt_melt(melt, tilt)
        vec, ctr = tri.bond_build(pos, melt.bond_atoms, melt.box, tilt)
        rc = 2.3
        ora = tri.p2_all_pairs(vec, ctr, molb, melt.box, tilt, rc, False)
        res = eng.p2_frame(pos, melt.bond_atoms, molb, melt.box, rc, same_molecule=False, tilt=tilt)
        assert res.grid.fast == want_fast, (tilt, list(res.grid.ncell_axis))
        assert np.array_equal(res.vec.cpu().numpy().T, vec) and np.array_equal(res.ctr.cpu().numpy().T, ctr)
        assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
        ok = ora["count"] > 0
        np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=1e-14)


def test_tilted_trajectory_through_the_pipeline():
    from mouse_b200.melt import jitter, make_melt, tilt_melt
    from mouse_b200.trajectory import FramePipeline
    from oracle import c_oracle
    _eng()
    tilt = (2.1, -1.7, 0.9)
    melt = make_melt(80, 41, seed=6)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    base = tilt_melt(melt, tilt)
    pipe = FramePipeline(melt.bond_atoms, molb, melt.box, 2.0, melt.n_atoms, same_molecule=True, tilt=tilt)
    frames = [base + 0.03 * np.random.default_rng(f).normal(size=base.shape) for f in range(5)]
    done = []
    for f, pos in enumerate(frames):
        done += pipe.submit(pos, tag=f)
    done += pipe.drain()
    pipe.close()
    for d in done:
        vec, ctr = c_oracle.bond_build(frames[d["tag"]], melt.bond_atoms, melt.box, tilt=tilt)
        ora = c_oracle.p2(vec, ctr, molb, melt.box, 2.0, True, tilt=tilt)
        assert d["n_pairs"] == int(ora["count"].sum())
        ok = ora["count"] > 0
        assert abs(d["sum_mean"] - ora["mean"][ok].sum()) <= 1e-9 * ok.sum()
