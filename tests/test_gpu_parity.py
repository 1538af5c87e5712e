"""Parity of the CUDA path (through the C-ABI) against the golden vectors of the real reference
and against the oracle.  Bit-exact: bond vectors/midpoints, neighbour counts, pair lists, RDF
histograms.  Tolerance 1e-5 relative (north_star) on P2 / S - in practice ~1e-13."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.util import (P2_CASES, RDF_CASES, config_from_melt, load_golden, p2_masks,  # noqa: E402
                        reference_dense)

RTOL = 1e-5       # north_star tolerance for floating-point results
ATOL_MEAN = 1e-14  # rounding-noise floor of an FP64 mean of O(100) terms in [0,1]


def _engine():
    from mouse_b200 import engine
    engine.require_cuda()
    return engine


def _run(melt, g, exact):
    eng = _engine()
    nbr, ref = p2_masks(melt, g)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, float(g["case_rcut"]),
                       same_molecule=bool(g["case_same_molecule"]), nbr_mask=nbr, ref_mask=ref, exact=exact)
    return eng, res


@pytest.mark.parametrize("name", P2_CASES)
def test_bond_build_bit_exact(name):
    melt, g = load_golden(name)
    eng = _engine()
    vec, ctr = eng.bond_build(melt.positions, melt.bond_atoms, melt.box)
    nbr, _ = p2_masks(melt, g)
    assert np.array_equal(vec.cpu().numpy()[:, nbr], g["ref_npBonds"][3:6])
    assert np.array_equal(ctr.cpu().numpy()[:, nbr], g["ref_npBonds"][6:9])


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
@pytest.mark.parametrize("name", P2_CASES)
def test_p2_against_golden(name, exact):
    melt, g = load_golden(name)
    eng, res = _run(melt, g, exact)
    mean, count, ptr, idx = reference_dense(g, melt.n_bonds)
    got_cnt = res.count.cpu().numpy().astype(np.int64)
    assert np.array_equal(got_cnt, count)                                   # bit-exact counts
    ok = count > 0
    got_mean = res.mean.cpu().numpy()
    np.testing.assert_allclose(got_mean[ok], mean[ok], rtol=1e-12, atol=ATOL_MEAN, equal_nan=True)
    assert np.isnan(got_mean[~ok]).all()
    p2_ref, p2_got = 1.5 * mean[ok] - 0.5, 1.5 * got_mean[ok] - 0.5
    np.testing.assert_allclose(p2_got, p2_ref, rtol=RTOL, atol=1e-13, equal_nan=True)
    t = res.totals_dict()
    assert t["n_refs"] == int(ok.sum()) and t["n_pairs"] == int(count.sum())
    if "ref_S_unbound" in g:
        with pytest.raises(UnboundLocalError):
            res.order_parameter()
    elif np.isnan(g["ref_S"]):
        assert np.isnan(res.order_parameter())
    else:
        assert abs(res.order_parameter() - g["ref_S"]) <= RTOL * abs(g["ref_S"])
    if not np.isnan(mean[ok]).any():                                        # bit-exact pair lists
        row_ptr, nbr_idx = eng.p2_pair_lists(res, same_molecule=bool(g["case_same_molecule"]), exact=exact)
        assert np.array_equal(row_ptr.cpu().numpy(), ptr)
        assert np.array_equal(nbr_idx.cpu().numpy().astype(np.int64), idx)


def test_fast_path_is_used_on_config1():
    melt, g = load_golden("config1_9900_rc15")
    _, res = _run(melt, g, False)
    assert res.grid.fast == 1 and min(res.grid.ncell_axis) >= 3


@pytest.mark.parametrize("same", [True, False])
@pytest.mark.parametrize("shape,rc", [((40, 51), 1.5), ((200, 101), 2.5), ((500, 41), 3.0)])
def test_p2_against_c_oracle(shape, rc, same):
    from mouse_b200.melt import make_melt
    from oracle import c_oracle
    melt = make_melt(shape[0], shape[1], seed=11, aspect=(1.0, 1.2, 0.9))
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=same)
    vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    assert np.array_equal(res.vec.cpu().numpy().T, vec) and np.array_equal(res.ctr.cpu().numpy().T, ctr)
    ora = c_oracle.p2(vec, ctr, molb, melt.box, rc, same, want_lists=True)
    assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
    ok = ora["count"] > 0
    np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=ATOL_MEAN)
    row_ptr, nbr_idx = eng.p2_pair_lists(res, same_molecule=same)
    assert np.array_equal(row_ptr.cpu().numpy(), ora["nbr_ptr"])
    assert np.array_equal(nbr_idx.cpu().numpy().astype(np.int64), ora["nbr_idx"])
    # the exact (all-FP64) kernel agrees with the fast one bit-for-bit on counts
    res2 = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=same, exact=True)
    assert np.array_equal(res2.count.cpu().numpy().astype(np.int64), ora["count"])
    np.testing.assert_allclose(res2.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=ATOL_MEAN)


def test_p2_full_size_config2():
    """BASELINE config 2: 1M bonds, rc = 2.5 - counts bit-exact against the C oracle."""
    from mouse_b200.melt import make_melt
    from oracle import c_oracle
    melt = make_melt(10000, 101, seed=0)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    for same in (True, False):
        res = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5, same_molecule=same)
        ora = c_oracle.p2(vec, ctr, molb, melt.box, 2.5, same)
        assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
        ok = ora["count"] > 0
        np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=ATOL_MEAN)
        s_ref = 1.5 * ora["mean"][ok].sum() / ok.sum() - 0.5
        assert abs(res.order_parameter() - s_ref) <= RTOL * abs(s_ref)
        assert res.totals_dict()["n_pairs"] == int(ora["count"].sum())


@pytest.mark.parametrize("same", [True, False])
def test_sparse_cells_use_a_coarser_grid(same):
    """>= 65 536 bonds with few bonds per rc-sized cell: the planner widens the cells to ~10 bonds each
    (mouse_grid_plan); counts, pair sums and S stay those of the reference's cutoff."""
    from mouse_b200.melt import make_melt
    from oracle import c_oracle
    melt = make_melt(1500, 81, seed=4)
    rc = 1.2
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, rc, same_molecule=same)
    assert res.grid.fast == 1 and min(res.grid.cell_w) > 1.5 * rc
    assert 8.0 < melt.n_bonds / res.grid.ncell < 14.0
    vec, ctr = c_oracle.bond_build(melt.positions, melt.bond_atoms, melt.box)
    ora = c_oracle.p2(vec, ctr, molb, melt.box, rc, same)
    assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
    ok = ora["count"] > 0
    np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=ATOL_MEAN)
    s_ref = 1.5 * ora["mean"][ok].sum() / ok.sum() - 0.5
    assert abs(res.order_parameter() - s_ref) <= RTOL * abs(s_ref)
    # the RDF sweep runs on the same (widened) cell list
    typ = np.zeros(melt.n_atoms, np.int32)
    h, _ = eng.rdf_frame(melt.positions, typ, melt.mol, 1, melt.box, 0., rc, 48, same_molecule=same)
    exp, _ = c_oracle.rdf(melt.positions, typ, melt.mol, 1, melt.box, 0., rc, 48, same)
    assert np.array_equal(h.cpu().numpy(), exp) and int(exp.sum()) > 0


def test_deterministic_and_slot_variants():
    from mouse_b200 import _lib
    from mouse_b200.melt import make_melt
    melt = make_melt(100, 101, seed=2)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    base = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5)
    s0, c0 = base.sum.clone(), base.count.clone()
    for slots, warps in ((0, 0), (4, 0), (6, 0), (8, 0), (6, 3)):
        r = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.5, slots=slots, warps=warps)
        assert torch.equal(r.count, c0)
        if slots == 0:
            assert torch.equal(r.sum, s0)            # run-to-run bit-identical
        else:
            torch.testing.assert_close(r.sum, s0, rtol=1e-13, atol=1e-13)


def test_edge_cases():
    eng = _engine()
    box = [10., 10., 10.]
    # empty frame
    res = eng.p2_frame(np.zeros((0, 3)), np.zeros((0, 2), np.int32), np.zeros(0, np.int32), box, 2.0)
    assert res.totals_dict()["n_refs"] == 0
    with pytest.raises(UnboundLocalError):
        res.order_parameter()
    # a single bond has no neighbours
    res = eng.p2_frame(np.array([[0., 0, 0], [1., 0, 0]]), np.array([[0, 1]], np.int32), np.zeros(1, np.int32), box, 2.0)
    assert res.count.cpu().tolist() == [0]
    # no cutoff (rcut < 0): every other bond is a neighbour
    from mouse_b200.melt import make_melt
    from oracle import p2_oracle as po
    melt = make_melt(3, 9, seed=4)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, -1.0)
    assert set(res.count.cpu().tolist()) == {melt.n_bonds - 1}
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    lit = po.p2_literal(vec, ctr, np.arange(1, melt.n_bonds + 1), molb, melt.box, -1.0)
    np.testing.assert_allclose(res.mean.cpu().numpy(), lit["mean"], rtol=1e-12)


# ------------------------------------------------------------------------------- mirror API
def test_mirror_api_order_parameter_and_histogram():
    from mouse_b200 import ordering_functions as of
    melt, g = load_golden("melt_290_rc15_same")
    cfg = config_from_melt(melt)
    s = of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 1.5, mode="average", sameMolecule=True)
    assert isinstance(s, np.float64) and abs(s - g["ref_S"]) <= RTOL * abs(g["ref_S"])
    smin, smax, sbins = g["histo_args"]
    h = of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 1.5, mode="histo", smin=float(smin), smax=float(smax),
                                              sbins=int(sbins), sameMolecule=True)
    assert np.array_equal(np.array(h["values"]), g["ref_histo_values"])
    assert np.array_equal(np.array(h["norm"]), g["ref_histo_norm"]) and h["step"] == g["ref_histo_step"]
    from mouse_b200.lammpack_misc import printHistogram
    assert printHistogram(h, norm=False) == str(g["ref_histo_print"])
    of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 1.5, storeAsAtomtypes=True, sameMolecule=True)
    assert np.array_equal(np.array([a.type for a in cfg.atoms()]), g["ref_atom_types_after"])


def test_mirror_api_selection_and_errors():
    from mouse_b200 import ordering_functions as of
    melt, g = load_golden("ortho_960_types_sel")
    cfg = config_from_melt(melt)
    s = of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 1.8, referenceResTypes=["R0", "R2"], sameMolecule=False)
    assert abs(s - g["ref_S"]) <= RTOL * abs(g["ref_S"])
    melt, g = load_golden("kat_rmax_default")
    s = of.CalculateOrientationOrderParameter(config_from_melt(melt), ["1"])       # rmax == 0 -> half diagonal
    assert abs(s - g["ref_S"]) <= RTOL * abs(g["ref_S"])
    melt, g = load_golden("kat_parallel")
    cfg = config_from_melt(melt)
    assert of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 2.5) == 1.0
    with pytest.raises(IndexError):
        of.CalculateOrientationOrderParameter(cfg, ["1"], 0., 2.5, mode="histo")
    melt, g = load_golden("kat_no_neighbours")
    with pytest.raises(UnboundLocalError):
        of.CalculateOrientationOrderParameter(config_from_melt(melt), ["1"], 0., 1.0)
    melt, g = load_golden("kat_zero_length")
    assert np.isnan(of.CalculateOrientationOrderParameter(config_from_melt(melt), ["1"], 0., 2.0))


def test_mirror_api_per_reference_functions():
    from mouse_b200 import ordering_functions as of
    melt, g = load_golden("melt_980_rc25_diffmol")
    cfg = config_from_melt(melt)
    npb = of.createNumpyBondsArrayFromConfig(cfg, allowedBondTypes=["1"])
    assert npb.shape == g["ref_npBonds"].shape
    assert np.array_equal(npb[0], g["ref_npBonds"][0]) and np.array_equal(npb[3:], g["ref_npBonds"][3:])
    mean, count, _, _ = reference_dense(g, melt.n_bonds)
    for b in (0, 1, 17, 400, 979):
        v = of.calculateAveCosSqForReference(cfg, cfg.bonds()[b], npb, 2.5, sameMolecule=False)
        assert abs(v - mean[b]) <= 1e-12 * abs(mean[b])
    melt, g = load_golden("kat_zero_length")
    cfg = config_from_melt(melt)
    npb = of.createNumpyBondsArrayFromConfig(cfg, allowedBondTypes=["1"])
    with pytest.raises(NameError, match="Zero length"):
        of.calculateAveCosSqForReference(cfg, cfg.bonds()[1], npb, 2.0)
    assert np.isnan(of.calculateAveCosSqForReference(cfg, cfg.bonds()[0], npb, 2.0))
    melt, g = load_golden("kat_no_neighbours")
    cfg = config_from_melt(melt)
    npb = of.createNumpyBondsArrayFromConfig(cfg, allowedBondTypes=["1"])
    with pytest.raises(NameError, match="No values"):
        of.calculateAveCosSqForReference(cfg, cfg.bonds()[0], npb, 1.0)


# ------------------------------------------------------------------------------- RDF
@pytest.mark.parametrize("name", RDF_CASES)
def test_rdf_against_golden(name):
    from mouse_b200 import ordering_functions as of
    melt, g = load_golden(name)
    cfg = config_from_melt(melt)
    atomtypes = [str(s) for s in g["case_atomtypes"]]
    rdfs = of.CalculatePairCorrelationFunctions(cfg, atomtypes, float(g["case_rmin"]), float(g["case_rmax"]),
                                                int(g["case_nbin"]), bool(g["case_same_molecule"]), verbose=False)
    assert sorted(rdfs["data"]) == [str(s) for s in g["ref_keys"]]
    vol = float(np.prod(melt.box))
    for k, row in zip(g["ref_keys"], g["ref_counts"]):
        assert np.array_equal(rdfs["data"][str(k)], row.astype(np.float64)), k       # bit-exact histogram
        assert rdfs["norm"][str(k)] == float(row.sum()) / 2. / melt.box[0] / melt.box[1] / melt.box[2]
    del vol
    from oracle import rdf_oracle
    r, v = rdf_oracle.rdf_tables(float(g["case_rmin"]), float(g["case_rmax"]), int(g["case_nbin"]))
    assert rdfs["r"] == r and rdfs["v"] == v


def test_rdf_single_reference_against_oracle():
    from mouse_b200 import ordering_functions as of
    from oracle import rdf_oracle
    melt, g = load_golden("rdf_1000_2types_5s_1000")
    cfg = config_from_melt(melt)
    npa = of.createNumpyAtomsArrayFromConfig(cfg, allowedAtomTypes=["1"])
    sel = np.flatnonzero(melt.atom_type == 0)
    for a in (0, 5, 999):
        atom = cfg.atoms()[a]
        h, edges = of.calculateRdfForReference(cfg, atom, npa, rmin=0., rmax=5., nbin=1000, sameMolecule=False)
        exp = rdf_oracle.rdf_for_reference(melt.box, melt.positions[a], a + 1, melt.mol[a], sel + 1, melt.mol[sel],
                                           melt.positions[sel], 0., 5., 1000, same_molecule=False)
        assert np.array_equal(h, exp)


def test_rdf_against_c_oracle_larger():
    from mouse_b200.melt import make_melt
    from oracle import c_oracle
    melt = make_melt(200, 100, seed=5, n_atom_types=2)
    eng = _engine()
    for same in (True, False):
        h, edges = eng.rdf_frame(melt.positions, melt.atom_type, melt.mol, 2, melt.box, 0., 5., 1000, same_molecule=same)
        exp, e2 = c_oracle.rdf(melt.positions, melt.atom_type, melt.mol, 2, melt.box, 0., 5., 1000, same)
        assert np.array_equal(edges, e2) and np.array_equal(h.cpu().numpy(), exp)


@pytest.mark.parametrize("rmin,rmax,nbin", [(0., 2.5, 250), (0.9, 3.0, 77), (0., 5., 1000)])
def test_rdf_fast_sweep_variants(rmin, rmax, nbin):
    """Half-shell RDF sweep: bit-exact against the C oracle and against the all-FP64 kernel; un-wrapped input
    coordinates (atoms several box lengths away) and a lower range bound exercise the exact minimum image / bin rule."""
    from mouse_b200.melt import make_melt
    from oracle import c_oracle
    melt = make_melt(150, 100, seed=8, n_atom_types=3)
    eng = _engine()
    pos = melt.positions + melt.images * melt.box        # un-wrapped coordinates
    pos[::7] += 3.0 * melt.box                            # and some atoms far outside the box
    for same in (True, False):
        h, edges = eng.rdf_frame(pos, melt.atom_type, melt.mol, 3, melt.box, rmin, rmax, nbin, same_molecule=same)
        hx, _ = eng.rdf_frame(pos, melt.atom_type, melt.mol, 3, melt.box, rmin, rmax, nbin, same_molecule=same, exact=True)
        exp, e2 = c_oracle.rdf(pos, melt.atom_type, melt.mol, 3, melt.box, rmin, rmax, nbin, same)
        assert np.array_equal(edges, e2)
        assert np.array_equal(hx.cpu().numpy(), exp)
        assert np.array_equal(h.cpu().numpy(), exp)


def test_rdf_fast_sweep_lattice_on_bin_edges():
    """Simple cubic lattice, bin edges placed on lattice distances: every pair sits exactly on an edge."""
    from oracle import c_oracle
    m, a0 = 12, 1.0
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) * a0 - 0.5 * m * a0
    box = np.array([m * a0] * 3)
    typ = np.zeros(len(g), np.int32)
    mol = np.arange(len(g), dtype=np.int32)
    eng = _engine()
    h, edges = eng.rdf_frame(g, typ, mol, 1, box, 0., 3.0, 30)
    exp, _ = c_oracle.rdf(g, typ, mol, 1, box, 0., 3.0, 30, True)
    assert np.array_equal(h.cpu().numpy(), exp) and int(exp.sum()) > 0


@pytest.mark.parametrize("same", [True, False])
def test_rdf_fast_sweep_dense_blobs_overflow_the_hit_queue(same):
    """Tight blobs: nearly every candidate of a home cell is in range, so one pair of reference atoms produces
    more hits than the warp's hit queue holds (the one-reference-at-a-time path), two atom types."""
    from oracle import c_oracle
    rng = np.random.default_rng(17)
    box = np.array([12.0, 12.0, 12.0])
    centres = np.array([[0.3, 0.2, -0.4], [0.9, 0.3, -0.2], [-4.0, 3.0, 5.9], [5.8, -5.9, 0.0]])
    pos = np.concatenate([c + 0.35 * rng.normal(size=(k, 3)) for c, k in zip(centres, (700, 500, 400, 300))]
                         + [rng.uniform(-6, 6, size=(800, 3))])
    pos = pos - np.floor((pos + 6.0) / 12.0) * 12.0
    n = len(pos)
    typ = rng.integers(0, 2, n).astype(np.int32)
    mol = (np.arange(n) // 7).astype(np.int32)
    eng = _engine()
    h, _ = eng.rdf_frame(pos, typ, mol, 2, box, 0., 2.0, 40, same_molecule=same)
    exp, _ = c_oracle.rdf(pos, typ, mol, 2, box, 0., 2.0, 40, same)
    assert np.array_equal(h.cpu().numpy(), exp) and int(exp.sum()) > 200000


def test_cli_matches_reference_output(capsys):
    """calculate_local_ordering.py <golden data file> --bondtypes 1 --max 1.5 --same-molecule prints the
    reference's S for that file (the golden S came from the reference's own Config of the same arrays)."""
    import os
    import calculate_local_ordering as cli
    from tests.util import GOLDEN_DIR
    _, g = load_golden("melt_290_rc15_same")
    cli.main([os.path.join(GOLDEN_DIR, "melt_290.data"), "--bondtypes", "1", "--max", "1.5", "--same-molecule"])
    out = capsys.readouterr().out.strip()
    assert abs(float(out) - float(g["ref_S"])) <= RTOL * abs(float(g["ref_S"]))
    _, g = load_golden("melt_290_rc15_diffmol")
    cli.main([os.path.join(GOLDEN_DIR, "melt_290.data"), "--bondtypes", "1", "--max", "1.5"])
    out = capsys.readouterr().out.strip()
    assert abs(float(out) - float(g["ref_S"])) <= RTOL * abs(float(g["ref_S"]))


def test_cli_out_pdb_matches_the_reference_file(tmp_path, capsys):
    """calculate_local_ordering.py ... --out-pdb: the recoloured PDB (CXX / CmXX atom types from the per-atom mean
    P2, ``ordering_functions.py:207-220``) equals the file the reference wrote, and S is still printed."""
    import os
    import calculate_local_ordering as cli
    from tests.util import GOLDEN_DIR
    out = tmp_path / "coloured.pdb"
    cli.main([os.path.join(GOLDEN_DIR, "melt_290.data"), "--bondtypes", "1", "--max", "1.5", "--same-molecule",
              "--out-pdb", str(out)])
    assert out.read_text() == open(os.path.join(GOLDEN_DIR, "melt_290_coloured.pdb")).read()
    _, g = load_golden("melt_290_rc15_same")
    assert abs(float(capsys.readouterr().out.strip()) - float(g["ref_S"])) <= RTOL * abs(float(g["ref_S"]))


def test_cli_selections_on_arrays_match_the_object_path(capsys):
    """--chains / --subcell: the CLI's array path (ingest + FrameArrays selections) against the same selections made
    on the object model with the mirrors of the reference's select_chains / select_rectangular."""
    import os
    import calculate_local_ordering as cli
    from mouse_b200 import ordering_functions as of
    from mouse_b200.lammpack_misc import read_data_typeselect, select_chains, select_rectangular
    from tests.util import GOLDEN_DIR
    path = os.path.join(GOLDEN_DIR, "melt_290.data")
    cli.main([path, "--bondtypes", "1", "--max", "2.0", "--same-molecule", "--chains", "1", "2", "3",
              "--subcell", "-0.5", "0.35"])
    out = float(capsys.readouterr().out.strip())
    cfg = select_chains(read_data_typeselect(path), ["1", "2", "3"])
    cfg = select_rectangular(cfg, True, -0.5, 0.35)
    ref = of.CalculateOrientationOrderParameter(cfg, ["1"], None, 2.0, sameMolecule=True)
    assert out == float(ref)


# ------------------------------------------------------------------------------- half-shell sweep corner cases
def _bonds_from_midpoints(ctr, vec):
    """Positions / bond table for bonds given by midpoint and vector (two private atoms per bond)."""
    n = len(ctr)
    pos = np.empty((2 * n, 3))
    pos[0::2] = ctr - 0.5 * vec
    pos[1::2] = ctr + 0.5 * vec
    bonds = np.stack([np.arange(0, 2 * n, 2), np.arange(1, 2 * n, 2)], axis=1).astype(np.int32)
    return pos, bonds


def _sweep_counters(res, n):
    """The workspace counters left behind by the last frame (0 zero-length bonds, 1 exact re-decisions, 3 pair-list
    length, 4 pair-list overflow flag)."""
    import ctypes as C
    from mouse_b200 import _lib, engine
    L = _lib.lib()
    ws = res.ws
    out = (C.c_int64 * 8)()
    _lib.check(L.mouse_debug_counters(C.byref(res.grid), n, engine._ptr(ws), out, engine._stream()), "debug_counters")
    return list(out)


def _check_against_c_oracle(pos, bonds, mol, box, rc, same=True, nbr=None, ref=None):
    from oracle import c_oracle
    eng = _engine()
    res = eng.p2_frame(pos, bonds, mol, box, rc, same_molecule=same, nbr_mask=nbr, ref_mask=ref)
    assert res.grid.fast == 1
    vec, ctr = c_oracle.bond_build(pos, bonds, box)
    ora = c_oracle.p2(vec, ctr, mol, box, rc, same, ref_mask=ref, nbr_mask=nbr)
    assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
    ok = ora["count"] > 0
    np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=ATOL_MEAN)
    assert res.totals_dict()["n_pairs"] == int(ora["count"].sum())
    return res, ora


@pytest.mark.parametrize("n_cl", [33, 64, 65, 150, 330, 700])
def test_dense_cluster_oversize_home_cells_and_many_passes(n_cl):
    """One cell holds far more bonds than the fast sweep takes as references at a time: the cell is swept in windows of
    HS_MAXHOME references, each against all chunks of the candidate concatenation; with 330 bonds the home cell itself
    spans two chunks of the landing zone (a window then starts in the chunk that holds its references).  Only
    guard-band pairs may reach the pair list."""
    rng = np.random.default_rng(5)
    box = np.array([24.0, 24.0, 24.0])
    n_bg = 20000                     # 15^3 cells of edge 1.6; the cluster sits in the middle of one of them
    ctr = np.concatenate([rng.uniform(-12, 12, (n_bg, 3)), rng.normal(0.8, 0.2, (n_cl, 3))])
    vec = rng.normal(size=(len(ctr), 3))
    vec *= 0.97 / np.linalg.norm(vec, axis=1)[:, None]
    pos, bonds = _bonds_from_midpoints(ctr, vec)
    mol = (np.arange(len(ctr)) // 7).astype(np.int32)
    for same in (True, False):
        res, ora = _check_against_c_oracle(pos, bonds, mol, box, 1.5, same)
        assert ora["count"].max() > min(140, n_cl - 10)        # the cluster really is dense
        c = _sweep_counters(res, len(bonds))
        assert c[4] == 0 and c[3] < 500


def test_pair_list_overflow_falls_back_to_the_exact_sweep():
    """Four bonds on every site of a cubic lattice, cutoff exactly the lattice constant: 24 neighbours of every bond
    sit exactly on the cutoff, the fast sweep's pair list overflows and the frame is recomputed in FP64.
    Perpendicular bonds exercise the cos^2 == 0 mask at the same time."""
    m, a = 6, 1.0
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) * a - 0.5 * m * a + 0.25
    dirs = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [1.0, 1.0, 0]]) * 0.5
    ctr = np.repeat(g, 4, axis=0)
    vec = np.tile(dirs, (len(g), 1))
    pos, bonds = _bonds_from_midpoints(ctr, vec)
    box = np.array([m * a] * 3)
    mol = np.arange(len(ctr), dtype=np.int32)
    res, ora = _check_against_c_oracle(pos, bonds, mol, box, a, True)
    assert ora["count"].min() >= 1
    assert res.grid.fast == 1
    assert _sweep_counters(res, len(bonds))[4] == 1     # the pair list did overflow


@pytest.mark.parametrize("same", [True, False])
def test_three_cells_per_axis_and_masks(same):
    """Smallest grid the fast sweep accepts (3 x 3 x 3 cells: every neighbour cell is a periodic image of another)
    combined with neighbour / reference selections."""
    from mouse_b200.melt import make_melt
    melt = make_melt(12, 40, seed=9)
    rc = float(melt.box[0]) / 3.3
    rng = np.random.default_rng(1)
    nbr = (rng.random(melt.n_bonds) < 0.8).astype(np.uint8)
    ref = (rng.random(melt.n_bonds) < 0.5).astype(np.uint8) & nbr
    molb = melt.mol[melt.bond_atoms[:, 0]]
    res, _ = _check_against_c_oracle(melt.positions, melt.bond_atoms, molb, melt.box, rc, same, nbr=nbr, ref=ref)
    assert list(res.grid.ncell_axis) == [3, 3, 3]


def test_frame_and_staged_paths_agree():
    """mouse_p2_frame (fused build / finish) and the staged entry points give the same per-bond results."""
    import ctypes as C
    from mouse_b200 import _lib
    from mouse_b200.melt import make_melt
    melt = make_melt(60, 51, seed=3)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    a = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.0)
    b = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.0, hist_args=(-0.5, 1.0, 75))   # staged calls
    assert torch.equal(a.count, b.count) and torch.equal(a.sum, b.sum)
    assert torch.equal(a.vec, b.vec) and torch.equal(a.ctr, b.ctr)
    assert a.totals_dict()["n_pairs"] == b.totals_dict()["n_pairs"] and a.totals_dict()["n_refs"] == b.totals_dict()["n_refs"]
    assert abs(a.order_parameter() - b.order_parameter()) < 1e-13
    assert int(b.hist.sum()) == b.totals_dict()["n_refs"]


def test_staged_sweep_after_a_fused_frame_on_the_same_cell_list():
    """The fused frame leaves its sums in the workspace accumulators; a staged mouse_p2_sweep on the same cell
    list (no rebuild) must start from zero again - twice in a row, for both molecule modes."""
    import ctypes as C
    from mouse_b200 import _lib, engine
    from mouse_b200.melt import make_melt
    melt = make_melt(80, 61, seed=9)
    dev = torch.device("cuda")
    n = melt.n_bonds
    L = _lib.lib()
    grid = _lib.grid_plan(melt.box, 2.0, n)
    assert grid.fast == 1
    pos = torch.as_tensor(melt.positions).to(dev)
    bonds = torch.as_tensor(melt.bond_atoms).to(dev)
    mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
    ws = engine.workspace(L.mouse_workspace_bytes(n, C.byref(grid)), dev)
    vec = torch.empty((3, n), dtype=torch.float64, device=dev)
    ctr = torch.empty_like(vec)
    mean = torch.empty(n, dtype=torch.float64, device=dev)
    tot = torch.zeros(6, dtype=torch.int64, device=dev)
    p, st = engine._ptr, engine._stream()
    out = {}
    for flags in (_lib.MOUSE_SAME_MOLECULE, 0):
        ssum = torch.empty(n, dtype=torch.float64, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        rc = L.mouse_p2_frame(p(pos), p(bonds), p(mol), None, None, n, C.byref(grid), flags, p(ws), ws.numel(),
                              p(vec), p(ctr), p(ssum), p(cnt), p(mean), p(tot), st)
        assert rc == 0
        out[flags] = (ssum.clone(), cnt.clone())
    for _ in range(2):
        for flags in (_lib.MOUSE_SAME_MOLECULE, 0):
            s2 = torch.full((n,), -1.0, dtype=torch.float64, device=dev)
            c2 = torch.full((n,), -1, dtype=torch.int32, device=dev)
            rc = L.mouse_p2_sweep(p(vec), p(ctr), n, C.byref(grid), flags, p(ws), ws.numel(), p(s2), p(c2), st)
            assert rc == 0
            torch.cuda.synchronize()
            assert torch.equal(c2, out[flags][1]) and torch.equal(s2, out[flags][0])


def test_trajectory_histograms_and_rdf_trajectory_single_rank():
    """mode='histo' over a trajectory: the per-frame P2 histograms accumulated by the pipeline equal the staged
    single-frame histograms summed up; the RDF trajectory driver equals the per-frame RDF histograms summed up."""
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import local_ordering_trajectory, pair_correlation_trajectory
    melt = make_melt(40, 51, seed=12)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    frames = [jitter(melt, f, 0.05) for f in range(6)]
    hargs = (-0.5, 1.0, 30)
    s_all, totals = local_ordering_trajectory(frames, melt.bond_atoms, molb, melt.box, 1.8, same_molecule=True,
                                              hist_args=hargs, gather=True)
    exp = np.zeros(30, np.int64)
    for f, pos in enumerate(frames):
        r = eng.p2_frame(pos, melt.bond_atoms, molb, melt.box, 1.8, hist_args=hargs)
        exp += r.hist.cpu().numpy()
        assert s_all[f] == r.order_parameter()
    assert np.array_equal(totals.p2_hist, exp) and exp.sum() == totals.n_refs and totals.n_hist_over == 0
    h = totals.p2_histogram(*hargs)
    assert abs(sum(h["values"]) - 1.0) < 1e-12
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    hist, edges, t2 = pair_correlation_trajectory(frames, typ, melt.mol, 2, melt.box, 0.0, 2.0, 50,
                                                  same_molecule=False)
    exp = sum(eng.rdf_frame(pos, typ, melt.mol, 2, melt.box, 0.0, 2.0, 50, same_molecule=False)[0].cpu().numpy()
              for pos in frames)
    assert np.array_equal(hist, exp) and t2.n_frames == 6 and len(edges) == 51


def _two_rank_worker(rank, world, port, out):
    import torch.distributed as dist
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import local_ordering_trajectory, pair_correlation_trajectory
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    melt = make_melt(40, 51, seed=12)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    frames = [jitter(melt, f, 0.05) for f in range(7)]
    s_all, tot = local_ordering_trajectory(frames, melt.bond_atoms, molb, melt.box, 1.8, hist_args=(-0.5, 1.0, 30),
                                           gather=True, device=torch.device("cuda", rank))
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    hist, _, t2 = pair_correlation_trajectory(frames, typ, melt.mol, 2, melt.box, 0.0, 2.0, 50,
                                              device=torch.device("cuda", rank))
    out[rank] = (s_all.tolist(), tot.p2_hist.tolist(), tot.n_pairs, tot.n_refs, hist.tolist(), t2.n_frames)
    dist.destroy_process_group()


def test_two_ranks_over_nccl_give_the_one_rank_histograms_bit_for_bit():
    """Frames sharded over two GPUs, one NCCL all-reduce of the packed sums + histograms, all-gather of the
    per-frame S: identical (integers: bit for bit) to the single-rank run."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import local_ordering_trajectory, pair_correlation_trajectory
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    melt = make_melt(40, 51, seed=12)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    frames = [jitter(melt, f, 0.05) for f in range(7)]
    s_all, tot = local_ordering_trajectory(frames, melt.bond_atoms, molb, melt.box, 1.8, hist_args=(-0.5, 1.0, 30),
                                           gather=True)
    typ = (np.arange(melt.n_atoms) % 2).astype(np.int32)
    hist, _, _ = pair_correlation_trajectory(frames, typ, melt.mol, 2, melt.box, 0.0, 2.0, 50)
    for r in range(2):
        s_r, p2h, npairs, nrefs, rdfh, nfr = out[r]
        assert s_r == s_all.tolist() and p2h == tot.p2_hist.tolist() and (npairs, nrefs) == (tot.n_pairs, tot.n_refs)
        assert rdfh == hist.tolist() and nfr == 7


def test_captured_graph_replays_the_frame_bit_for_bit():
    """FramePipeline replays one captured CUDA graph per slot (mouse_p2_frame_graph_*): same totals and per-bond
    outputs as the direct launch train, frame after frame; a frame with its own box falls back to direct launches."""
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import FramePipeline
    melt = make_melt(60, 51, seed=3)
    molb = melt.mol[melt.bond_atoms[:, 0]]
    frames = [jitter(melt, f, 0.05) for f in range(7)]
    out = {}
    for use_graph in (True, False):
        pipe = FramePipeline(melt.bond_atoms, molb, melt.box, 2.0, melt.n_atoms, same_molecule=False, depth=2,
                             use_graph=use_graph)
        done, means = [], []
        for f, pos in enumerate(frames):
            done += pipe.submit(pos, tag=f, box=melt.box * 1.01 if f == 4 else None)
        done += pipe.drain()
        if use_graph:
            assert pipe.slots[0]["graph"] is not None
            assert 8 <= pipe.L.mouse_p2_frame_graph_nodes(pipe.slots[0]["graph"]) <= 10
        out[use_graph] = [(d["tag"], d["sum_mean"], d["n_refs"], d["n_pairs"]) for d in done]
        pipe.close()
    assert out[True] == out[False] and [t[0] for t in out[True]] == list(range(7))


def test_trajectory_pipeline_matches_frame_by_frame_calls():
    """local_ordering_trajectory (several frames in flight, own stream / workspace each) gives, frame by frame
    and bit for bit, what separate p2_frame calls give; results come back in submission order."""
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import FramePipeline, local_ordering_trajectory
    melt = make_melt(50, 41, seed=2)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    frames = [melt.positions] + [jitter(melt, f, 0.05) for f in range(1, 8)]
    per_frame, totals = local_ordering_trajectory(frames, melt.bond_atoms, molb, melt.box, 2.0, same_molecule=False)
    assert [f for f, _ in per_frame] == list(range(8))
    n_pairs = 0
    for f, s in per_frame:
        res = eng.p2_frame(frames[f], melt.bond_atoms, molb, melt.box, 2.0, same_molecule=False)
        assert s == res.order_parameter()
        n_pairs += res.totals_dict()["n_pairs"]
    assert totals.n_frames == 8 and totals.n_pairs == n_pairs
    # per-bond outputs of the last submitted frame stay in its slot; pinned, pageable and device inputs agree
    pipe = FramePipeline(melt.bond_atoms, molb, melt.box, 2.0, melt.n_atoms, same_molecule=False, depth=2)
    pinned = torch.from_numpy(frames[3]).pin_memory()
    done = pipe.submit(pinned, tag="a") + pipe.submit(frames[3], tag="b") + pipe.submit(torch.from_numpy(frames[3]).cuda(), tag="c")
    done += pipe.drain()
    assert [d["tag"] for d in done] == ["a", "b", "c"]
    assert len({(d["sum_mean"], d["n_refs"], d["n_pairs"]) for d in done}) == 1
    ref = eng.p2_frame(frames[3], melt.bond_atoms, molb, melt.box, 2.0, same_molecule=False)
    assert torch.equal(pipe.slots[0]["cnt"], ref.count) and torch.equal(pipe.slots[0]["sum"], ref.sum)


def test_dump_trajectory_matches_the_oracle_frame_by_frame():
    """File ingestion row: LAMMPS data + dump -> arrays -> hot path, against the NumPy restatement of the reference
    on the same frames (per-frame box, scaled coordinates)."""
    import os
    from mouse_b200 import ingest
    from mouse_b200.trajectory import local_ordering_from_dump
    from oracle import p2_oracle as po
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    top = ingest.read_lmp_data_arrays(os.path.join(here, "melt_290.data"))
    dump = os.path.join(here, "melt_290.dump")
    per_frame, totals = local_ordering_from_dump(top, dump, ["1"], 1.5, "scaled", same_molecule=True)
    ts, boxes, pos = ingest.read_lmp_dump_arrays(dump, top.atom_num, "scaled")
    assert [t for t, _ in per_frame] == ts.tolist() and totals.n_frames == 3
    for f in range(3):
        vec, ctr = po.bond_build(pos[f], top.bond_atoms, boxes[f])
        lit = po.p2_literal(vec, ctr, top.bond_num, top.bond_mol_hash(), boxes[f], 1.5)
        ok = lit["count"] > 0
        s_ref = 1.5 * lit["mean"][ok].sum() / ok.sum() - 0.5
        assert abs(per_frame[f][1] - s_ref) <= RTOL * abs(s_ref)


def test_cli_over_several_files_and_over_two_ranks(tmp_path, capsys):
    """calculate_local_ordering.py over several frame files: one S per file in command-line order (files are read
    on a host thread while the previous one is on the GPU); under torchrun with two ranks the files are sharded and
    rank 0 prints the same stdout."""
    import subprocess
    import sys
    import calculate_local_ordering as cli
    from mouse_b200.melt import jitter, make_melt, write_lammps_data
    from tests.util import GOLDEN_DIR
    melt = make_melt(12, 30, seed=3)
    files = []
    for f in range(5):
        m2 = make_melt(12, 30, seed=3)
        m2.positions[:] = jitter(melt, f, 0.05) if f else melt.positions
        path = str(tmp_path / ("frame%d.data" % f))
        write_lammps_data(m2, path)
        files.append(path)
    argv = files + ["--bondtypes", "1", "--max", "1.5", "--same-molecule"]
    cli.main(argv)
    out = capsys.readouterr().out
    lines = out.strip().split("\n")
    assert len(lines) == 5
    single = []
    for path in files:
        cli.main([path] + argv[5:])
        single.append(capsys.readouterr().out.strip())
    assert lines == single
    if torch.cuda.device_count() >= 2:
        root = os.path.dirname(GOLDEN_DIR.rstrip("/").rsplit("/tests", 1)[0] + "/x")
        # two ranks the way torchrun starts them (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment); the
        # launcher itself is not used here: its argument parser rejects the reference's "--max" flag as an ambiguous
        # abbreviation of its own options (a real run passes the script's options after "--")
        port = str(29700 + os.getpid() % 200)
        procs = []
        for rank in range(2):
            env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                       MASTER_PORT=port)
            procs.append(subprocess.Popen([sys.executable, os.path.join(root, "calculate_local_ordering.py")] + argv,
                                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=root, env=env))
        outs = [p.communicate(timeout=600) for p in procs]
        assert all(p.returncode == 0 for p in procs), outs[0][1][-2000:] + outs[1][1][-2000:]
        mine = lambda text: [ln for ln in text.strip().split("\n") if ln and not ln.startswith(("W", "NCCL"))]
        assert mine(outs[0][0]) == lines
        assert mine(outs[1][0]) == []            # only rank 0 prints


# ------------------------------------------------------------------------------- next rows: director P2, COM RDF
def test_crystallinity_parameter_against_the_reference():
    """CalculateCrystallinityParameter (ordering_functions.py:230-282): histogram, labels and recoloured atom types
    against golden vectors produced by the reference (tests/golden/make_golden_next.py); 'average' crashes with
    KeyError(0) in the reference and here."""
    import os
    from mouse_b200 import ordering_functions as of
    from mouse_b200.lammpack_misc import read_data_typeselect
    from mouse_b200.lammpack_types import Vector3d
    from tests.util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "next_290.npz"))
    path = os.path.join(GOLDEN_DIR, "melt_290.data")
    for tag, rv in (("mean", Vector3d(0., 0., 0.)), ("z", Vector3d(0., 0., 1.)), ("tilt", Vector3d(0.3, -1.2, 0.5))):
        cfg = read_data_typeselect(path)
        h = of.CalculateCrystallinityParameter(cfg, ['1'], reference_vector=rv, storeAsAtomtypes=True, mode='histo',
                                               smin=-0.5, smax=1.0, sbins=40)
        assert np.array_equal(np.array(h["values"]), g["cryst_%s_values" % tag]), tag
        assert np.array_equal(np.array(h["norm"]), g["cryst_%s_norm" % tag]) and h["step"] == float(g["cryst_%s_step" % tag])
        assert [a.type for a in cfg.atoms()] == [str(s) for s in g["cryst_%s_types" % tag]], tag
    assert str(g["cryst_average_error"]).startswith("KeyError")
    with pytest.raises(KeyError):
        of.CalculateCrystallinityParameter(read_data_typeselect(path), ['1'], mode='average')


def test_com_rdfs_against_the_reference():
    """CalculateComRdfs (ordering_functions.py:284-313) with and without masses."""
    import os
    from mouse_b200 import ordering_functions as of
    from mouse_b200.lammpack_misc import read_data_typeselect
    from tests.util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "next_290.npz"))
    path = os.path.join(GOLDEN_DIR, "melt_290.data")
    for tag in ("nomass", "mass"):
        cfg = read_data_typeselect(path)
        if tag == "mass":
            for k, a in enumerate(cfg.atoms()):
                a.mass = 1.0 + (k % 3) if k % 2 == 0 else 0.0
        r = of.CalculateComRdfs(cfg, 0., 4., 32)
        assert np.array_equal(r["data"]["1"], g["com_%s_data_1" % tag]) and r["norm"]["1"] == g["com_%s_norm_1" % tag]
        assert np.array_equal(np.array(r["r"]), g["com_%s_r" % tag]) and np.array_equal(np.array(r["v"]), g["com_%s_v" % tag])


@pytest.mark.parametrize("exact", [False, True], ids=["fast", "exact"])
def test_outputs_do_not_depend_on_previous_contents(exact):
    """Every per-bond output is (re)written by a frame: reusing the output tensors of an unrelated frame (different
    selections, garbage in the unselected bonds) gives the same bits as fresh tensors."""
    from mouse_b200.melt import make_melt
    melt = make_melt(40, 60, seed=12)
    eng = _engine()
    molb = melt.mol[melt.bond_atoms[:, 0]]
    rng = np.random.default_rng(3)
    nbr = (rng.random(melt.n_bonds) < 0.7).astype(np.uint8)
    ref = ((rng.random(melt.n_bonds) < 0.5) & nbr.astype(bool)).astype(np.uint8)
    fresh = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 1.8, nbr_mask=nbr, ref_mask=ref, exact=exact)
    f_sum, f_cnt, f_mean = fresh.sum.clone(), fresh.count.clone(), fresh.mean.clone()
    dirty = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 2.4, exact=exact)        # everything non-zero
    dirty.sum.fill_(123.0)
    dirty.count.fill_(77)
    dirty.mean.fill_(-5.0)
    again = eng.p2_frame(melt.positions, melt.bond_atoms, molb, melt.box, 1.8, nbr_mask=nbr, ref_mask=ref, exact=exact,
                         out=dirty)
    assert torch.equal(again.count, f_cnt) and torch.equal(again.sum, f_sum)
    assert torch.equal(torch.nan_to_num(again.mean, nan=-1.0), torch.nan_to_num(f_mean, nan=-1.0))
    assert int((again.count[torch.as_tensor(ref == 0, device=again.count.device)] != 0).sum()) == 0


def test_rdf_cli_prints_the_reference_table(capsys):
    """calculate_rdfs.py (same flags as the reference's script): the printed table equals printRdfs of the mirror's
    CalculatePairCorrelationFunctions on the object model, and the histogram behind it is the golden one."""
    import io
    import os
    import calculate_rdfs as cli
    from mouse_b200 import ordering_functions as of
    from mouse_b200.lammpack_misc import printRdfs, read_data_typeselect
    from tests.util import GOLDEN_DIR
    path = os.path.join(GOLDEN_DIR, "melt_290.data")
    cli.main([path, "--rmin", "0.", "--rmax", "3.0", "--nbin", "30", "--same-molecule", "--no-norm"])
    out = capsys.readouterr().out
    rdfs = of.CalculatePairCorrelationFunctions(read_data_typeselect(path), None, 0., 3.0, 30, True, verbose=False)
    buf = io.StringIO()
    printRdfs(rdfs, False, out=buf)
    assert out == buf.getvalue() and out.startswith("#r\tvol\t1_1\t1_AAAAA\tAAAAA_AAAAA\n")
    assert sum(float(line.split("\t")[2]) for line in out.splitlines()[1:]) > 0


def test_crystallinity_and_com_rdf_clis(tmp_path, capsys):
    """calculate_crystallinity.py / calculate_com_rdfs.py (the reference's flags): printed tables equal the
    reference-format printers applied to the golden histogram / to the mirror function; --out-pdb writes the
    recoloured atoms; --mode average fails like the reference."""
    import io
    import os
    import calculate_com_rdfs as com_cli
    import calculate_crystallinity as cli
    from mouse_b200 import ordering_functions as of
    from mouse_b200.lammpack_misc import printHistogram, printRdfs, read_data_typeselect
    from tests.util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "next_290.npz"))
    path = os.path.join(GOLDEN_DIR, "melt_290.data")
    pdb = tmp_path / "c.pdb"
    cli.main([path, "--mode", "histo", "--bins", "40", "--bondtypes", "1", "--reference", "0", "0", "1", "--out-pdb", str(pdb)])
    out = capsys.readouterr().out
    hist = dict(min=-0.5, max=1.0, nbins=40, step=float(g["cryst_z_step"]), values=list(g["cryst_z_values"]),
                norm=list(g["cryst_z_norm"]))
    assert out == printHistogram(hist, norm=True)
    types = [line[12:16].strip() for line in pdb.read_text().splitlines() if line.startswith("ATOM")]
    assert types == [str(t) for t in g["cryst_z_types"]]
    with pytest.raises(KeyError):
        cli.main([path, "--bondtypes", "1"])
    capsys.readouterr()
    com_cli.main([path, "--rmin", "0.", "--rmax", "4.0", "--nbin", "32"])
    out = capsys.readouterr().out
    buf = io.StringIO()
    printRdfs(of.CalculateComRdfs(read_data_typeselect(path), 0., 4., 32), True, out=buf)
    assert out == buf.getvalue() and len(out.splitlines()) == 33
