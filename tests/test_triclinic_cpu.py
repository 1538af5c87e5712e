"""Triclinic boxes (BASELINE config 5): BUILD-DEFINED semantics, parity unpinned - the reference is orthorhombic only
(lammpack_types.py:440-451).  These CPU tests pin the definition itself (oracle/triclinic_oracle.py, oracle/c):
tilt = 0 is bit-identical to the (reference-pinned) orthorhombic oracle, the brick minimum image is the true nearest
image inside the validity range, the C cell-list variant equals the all-pairs restatement."""
import numpy as np
import pytest

from mouse_b200.melt import make_melt
from oracle import c_oracle, p2_oracle, rdf_oracle, triclinic_oracle as tri

TILTS = [(0.0, 0.0, 0.0), (1.3, 0.0, 0.0), (2.1, -1.7, 0.9), (-3.0, 2.5, -2.0)]


def _tilted_melt(n_chains, n_beads, tilt, seed):
    """A melt whose chains were grown in an orthorhombic box, re-wrapped into the tilted cell (bonds crossing the
    tilted faces need the tilt-aware minimum image)."""
    melt = make_melt(n_chains, n_beads, seed=seed)
    unwrapped = melt.positions + melt.images * melt.box        # the chains as grown, before any periodic wrap
    return melt, tri.wrap_into_cell(unwrapped, melt.box, tilt)


@pytest.mark.parametrize("tilt", TILTS)
def test_brick_minimum_image_is_the_nearest_image_inside_half_the_box(tilt):
    rng = np.random.default_rng(3)
    box = np.array([9.0, 7.5, 8.0])
    d = rng.uniform(-25, 25, (4000, 3))
    bx, by, bz = tri.min_image(d[:, 0], d[:, 1], d[:, 2], box, tilt)
    brick2 = bx**2 + by**2 + bz**2
    best, best2 = tri.min_image_bruteforce(d, box, tilt, reach=6)
    half = 0.5 * box.min()
    inside = best2 < half * half
    assert inside.sum() > 500
    # inside the validity range the brick representative IS the nearest image (same lattice vector subtracted)
    np.testing.assert_allclose(np.stack([bx, by, bz], 1)[inside], best[inside], rtol=0, atol=1e-12)
    # outside it the brick representative is never shorter than the nearest image
    assert np.all(brick2 >= best2 - 1e-9)
    assert np.all(np.abs(bx) <= box[0] / 2 + 1e-12) and np.all(np.abs(by) <= box[1] / 2 + 1e-12)


def test_zero_tilt_is_bit_identical_to_the_orthorhombic_oracle():
    melt = make_melt(30, 33, seed=5)
    zero = (0.0, 0.0, 0.0)
    v0, c0 = p2_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    v1, c1 = tri.bond_build(melt.positions, melt.bond_atoms, melt.box, zero)
    v2, c2 = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box, tilt=zero)
    assert np.array_equal(v0, v1) and np.array_equal(c0, c1) and np.array_equal(v0, v2) and np.array_equal(c0, c2)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    for same in (True, False):
        a = c_oracle.p2(v0, c0, molb, melt.box, 1.5, same)
        b = c_oracle.p2(v0, c0, molb, melt.box, 1.5, same, tilt=zero)
        c = tri.p2_all_pairs(v0, c0, molb, melt.box, zero, 1.5, same)
        assert np.array_equal(a["count"], b["count"]) and np.array_equal(a["sum"], b["sum"])
        assert np.array_equal(a["count"], c["count"])
        np.testing.assert_allclose(c["sum"], a["sum"], rtol=1e-13, atol=1e-15)
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    h0, _ = c_oracle.rdf(melt.positions, typ, melt.mol, 2, melt.box, 0.0, 3.0, 60, False)
    h1, _ = c_oracle.rdf(melt.positions, typ, melt.mol, 2, melt.box, 0.0, 3.0, 60, False, tilt=zero)
    h2 = tri.rdf_all_pairs(melt.positions, typ, melt.mol, 2, melt.box, zero, 0.0, 3.0, 60, False)
    h3, _ = rdf_oracle.rdf_total(melt.positions, typ, melt.mol, 2, melt.box, 0.0, 3.0, 60, False) \
        if hasattr(rdf_oracle, "rdf_total") else (h0, None)
    assert np.array_equal(h0, h1) and np.array_equal(h0, h2) and np.array_equal(h0, h3)


@pytest.mark.parametrize("tilt", TILTS[1:])
def test_c_cell_list_variant_equals_the_all_pairs_restatement(tilt):
    melt, pos = _tilted_melt(24, 30, tilt, seed=7)
    vec, ctr = tri.bond_build(pos, melt.bond_atoms, melt.box, tilt)
    vc, cc = c_oracle.bond_build(pos, melt.bond_atoms, melt.box, tilt=tilt)
    assert np.array_equal(vec, vc) and np.array_equal(ctr, cc)
    assert np.abs(np.linalg.norm(vec, axis=1) - 0.97).max() < 1e-9      # every bond found its nearest image
    molb = melt.mol[melt.bond_atoms[:, 0]]
    rng = np.random.default_rng(1)
    nbr = rng.random(len(vec)) < 0.9
    ref = nbr & (rng.random(len(vec)) < 0.7)
    for same in (True, False):
        a = tri.p2_all_pairs(vec, ctr, molb, melt.box, tilt, 1.6, same, ref_mask=ref, nbr_mask=nbr)
        b = c_oracle.p2(vec, ctr, molb, melt.box, 1.6, same, ref_mask=ref, nbr_mask=nbr, tilt=tilt)
        assert np.array_equal(a["count"], b["count"]) and a["count"].sum() > 1000
        np.testing.assert_allclose(a["sum"], b["sum"], rtol=1e-13, atol=1e-15)
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    ha = tri.rdf_all_pairs(pos, typ, melt.mol, 2, melt.box, tilt, 0.2, 2.5, 47, True)
    hb, _ = c_oracle.rdf(pos, typ, melt.mol, 2, melt.box, 0.2, 2.5, 47, True, tilt=tilt)
    assert np.array_equal(ha, hb) and ha.sum() > 1000


def test_tilt_changes_the_answer_and_an_equivalent_lattice_does_not():
    """Sanity of the semantics: the same atoms in a tilted cell give different neighbour sets than with the tilt
    ignored (what the reference's reader would do), while describing the SAME lattice with xy + lx (an equivalent,
    un-reduced tilt) leaves all in-range pairs unchanged as long as the cutoff is inside the validity range."""
    tilt = (2.1, -1.7, 0.9)
    melt, pos = _tilted_melt(24, 30, tilt, seed=11)
    vec, ctr = tri.bond_build(pos, melt.bond_atoms, melt.box, tilt)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    a = c_oracle.p2(vec, ctr, molb, melt.box, 1.5, True, tilt=tilt)
    flat = c_oracle.p2(vec, ctr, molb, melt.box, 1.5, True)
    assert not np.array_equal(a["count"], flat["count"])
    tilt2 = (tilt[0] - melt.box[0], tilt[1], tilt[2])          # b' = b - a spans the same lattice
    b = c_oracle.p2(vec, ctr, molb, melt.box, 1.5, True, tilt=tilt2)
    assert np.array_equal(a["count"], b["count"])
