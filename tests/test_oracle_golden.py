"""Pin the oracle (oracle/p2_oracle.py) against the golden vectors produced by the real
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import p2_oracle as po
from tests.util import P2_CASES, P2_SMALL_CASES, load_golden, p2_masks, reference_dense


@pytest.mark.parametrize("name", P2_CASES)
def test_bond_build_bit_exact(name):
    melt, g = load_golden(name)
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    nbr, _ = p2_masks(melt, g)
    ref = g["ref_npBonds"]
    assert np.array_equal(ref[0].astype(np.int64) - 1, np.flatnonzero(nbr))
    assert np.array_equal(vec[nbr].T, ref[3:6])
    assert np.array_equal(ctr[nbr].T, ref[6:9])
    # mol hashes of the fixtures are collision free (hash8 hazard)
    mol_of_bond = melt.mol[melt.bond_atoms[nbr, 0]]
    assert len(set(zip(mol_of_bond.tolist(), ref[2].tolist()))) == len(set(mol_of_bond.tolist()))


@pytest.mark.parametrize("name", P2_SMALL_CASES)
def test_literal_restatement_bit_exact(name):
    melt, g = load_golden(name)
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    nbr, ref = p2_masks(melt, g)
    num = np.arange(1, melt.n_bonds + 1)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    r = po.p2_literal(vec, ctr, num, molb, melt.box, float(g["case_rcut"]), bool(g["case_same_molecule"]),
                      ref_mask=ref, nbr_mask=nbr)
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    assert np.array_equal(r["count"], count)
    assert np.array_equal(r["nbr_ptr"], ptr)
    assert np.array_equal(r["nbr_idx"], idx)
    assert np.array_equal(r["mean"], mean, equal_nan=True)        # bit-exact
    if "ref_S_unbound" in g:
        with pytest.raises(UnboundLocalError):
            po.order_parameter(r["mean"], r["count"])
    else:
        s = po.order_parameter(r["mean"], r["count"])
        assert np.array_equal(s, g["ref_S"], equal_nan=True)      # bit-exact, sequential sum


@pytest.mark.parametrize("name", [n for n in P2_CASES if n != "kat_rmax_default"])
def test_pairs_restatement(name):
    melt, g = load_golden(name)
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    nbr, ref = p2_masks(melt, g)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    r = po.p2_pairs(vec, ctr, molb, melt.box, float(g["case_rcut"]), bool(g["case_same_molecule"]),
                    ref_mask=ref, nbr_mask=nbr, want_lists=True)
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    assert np.array_equal(r["count"], count)
    if not np.isnan(mean[count > 0]).any():
        assert np.array_equal(r["nbr_ptr"], ptr)
        assert np.array_equal(r["nbr_idx"], idx)
    ok = count > 0
    np.testing.assert_allclose(r["mean"][ok], mean[ok], rtol=1e-13, atol=0, equal_nan=True)
    if "ref_S_unbound" not in g:
        s = po.order_parameter(r["mean"], r["count"], sequential=False)
        if np.isnan(g["ref_S"]):
            assert np.isnan(s)
        else:
            assert abs(s - g["ref_S"]) <= 1e-5 * abs(g["ref_S"])


@pytest.mark.parametrize("name", [n for n in P2_CASES if "histo_args" in load_golden(n)[1]
                                  or "ref_histo_indexerror" in load_golden(n)[1]])
def test_histogram_mode(name):
    melt, g = load_golden(name)
    mean, count, _, _ = reference_dense(g, melt.n_bonds)
    if "ref_histo_indexerror" in g:
        with pytest.raises(IndexError):
            po.p2_histogram(mean, count, -0.5, 1.0, 75)
        return
    smin, smax, sbins = g["histo_args"]
    h = po.p2_histogram(mean, count, float(smin), float(smax), int(sbins))
    assert np.array_equal(np.array(h["values"]), g["ref_histo_values"])
    assert np.array_equal(np.array(h["norm"]), g["ref_histo_norm"])
    assert h["step"] == g["ref_histo_step"]


def test_atom_colouring():
    melt, g = load_golden("melt_290_rc15_same")
    mean, count, _, _ = reference_dense(g, melt.n_bonds)
    types = np.array([melt.atom_type_names[t] for t in melt.atom_type], dtype=object)
    for a, s in po.atom_colouring(mean, count, melt.bond_atoms, melt.n_atoms):
        types[a] = s
    assert np.array_equal(types.astype(str), g["ref_atom_types_after"])


def test_kat_expectations():
    """Analytic known answers hold in the golden vectors themselves."""
    _, g = load_golden("kat_parallel")
    assert g["ref_S"] == 1.0 and "ref_histo_indexerror" in g
    melt, g = load_golden("kat_exact_cutoff")
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    assert count.tolist() == [1, 1, 0, 0]          # exactly rc counted, one ulp beyond not
    assert idx.tolist() == [1, 0]
    melt, g = load_golden("kat_perpendicular")
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    assert count.tolist() == [1, 1, 2, 0]          # the perpendicular in-range pair 0-1 is dropped
    melt, g = load_golden("kat_boundary")
    assert np.all(np.abs(g["ref_npBonds"][3:6]) <= 5.0)
    assert np.abs(g["ref_npBonds"][6:9]).max() >= 5.0   # midpoint left outside the box
    _, g = load_golden("kat_zero_length")
    assert np.isnan(g["ref_S"])


# --------------------------------------------------------------------------- C oracle + RDF
from oracle import c_oracle, rdf_oracle  # noqa: E402
from tests.util import RDF_CASES  # noqa: E402


@pytest.mark.parametrize("name", P2_CASES)
def test_c_oracle_p2_against_golden(name):
    melt, g = load_golden(name)
    vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    vec_np, ctr_np = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    assert np.array_equal(vec, vec_np) and np.array_equal(ctr, ctr_np)
    nbr, ref = p2_masks(melt, g)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    r = c_oracle.p2(vec, ctr, molb, melt.box, float(g["case_rcut"]), bool(g["case_same_molecule"]),
                    ref_mask=ref, nbr_mask=nbr, want_lists=True)
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    assert np.array_equal(r["count"], count)
    if not np.isnan(mean[count > 0]).any():
        assert np.array_equal(r["nbr_ptr"], ptr)
        assert np.array_equal(r["nbr_idx"], idx)
    ok = count > 0
    np.testing.assert_allclose(r["mean"][ok], mean[ok], rtol=1e-13, atol=0, equal_nan=True)


def _rdf_inputs(melt, g):
    names = np.array([melt.atom_type_names[t] for t in melt.atom_type])
    atomtypes = [str(s) for s in g["case_atomtypes"]]
    return names, atomtypes, float(g["case_rmin"]), float(g["case_rmax"]), int(g["case_nbin"]), \
        bool(g["case_same_molecule"])


@pytest.mark.parametrize("name", RDF_CASES)
def test_rdf_numpy_restatement_bit_exact(name):
    melt, g = load_golden(name)
    names, atomtypes, rmin, rmax, nbin, same = _rdf_inputs(melt, g)
    types, data = rdf_oracle.pair_correlation_counts(melt.positions, names, melt.mol, melt.box, atomtypes,
                                                     rmin, rmax, nbin, same)
    assert types == [str(s) for s in g["ref_types"]]
    assert sorted(data) == [str(s) for s in g["ref_keys"]]
    for k, row in zip(g["ref_keys"], g["ref_counts"]):
        assert np.array_equal(data[str(k)], row), k


@pytest.mark.parametrize("name", RDF_CASES)
def test_c_oracle_rdf_against_golden(name):
    melt, g = load_golden(name)
    names, atomtypes, rmin, rmax, nbin, same = _rdf_inputs(melt, g)
    types = [str(s) for s in g["ref_types"]]
    code = np.array([types.index(t) if (not atomtypes or t in atomtypes) else -1 for t in names], np.int32)
    hist, _ = c_oracle.rdf(melt.positions, code, melt.mol, len(types), melt.box, rmin, rmax, nbin, same)
    keys = [types[a] + "_" + types[b] for a in range(len(types)) for b in range(a, len(types))]
    for k, row in zip(g["ref_keys"], g["ref_counts"]):
        assert np.array_equal(hist[keys.index(str(k))], row), k
