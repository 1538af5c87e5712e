"""Parity at the sizes BASELINE.json states (configs 3, 4, 5; config 2 is in test_gpu_parity.py): counts / histograms
bit-exact against the OpenMP cell-list C oracle, means within 1e-12, S within the north-star tolerance 1e-5 relative.
The triclinic variant of config 5 is checked against the repo's own restatement (parity unpinned, build-defined)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _eng():
    from mouse_b200 import engine
    engine.require_cuda()
    return engine


def test_config3_trajectory_of_250k_bond_frames_through_the_pipeline():
    """32 jittered frames of the 250k-bond melt through FramePipeline (three frames in flight, captured graph):
    per-frame pair counts and S against the C oracle, per-bond counts of the last frames bit-exact."""
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import FramePipeline
    from mouse_b200.workloads import CONFIG3
    from oracle import c_oracle
    _eng()
    melt = make_melt(CONFIG3["chains"], CONFIG3["beads"], seed=0)
    assert melt.n_bonds == 250_000
    molb = melt.mol[melt.bond_atoms[:, 0]]
    rc = CONFIG3["rcut"]
    pipe = FramePipeline(melt.bond_atoms, molb, melt.box, rc, melt.n_atoms, same_molecule=True)
    frames = [jitter(melt, f, 0.05) for f in range(32)]
    done, per_bond = [], {}
    for f, pos in enumerate(frames):
        done += pipe.submit(pos, tag=f)
    last = pipe.slots[(pipe._next - 1) % pipe.depth]
    done += pipe.drain()
    per_bond[31] = last["cnt"].cpu().numpy().astype(np.int64)
    pipe.close()
    assert [d["tag"] for d in done] == list(range(32))
    for d in done:
        vec, ctr = c_oracle.bond_build(frames[d["tag"]], melt.bond_atoms, melt.box)
        ora = c_oracle.p2(vec, ctr, molb, melt.box, rc, True)
        ok = ora["count"] > 0
        assert d["n_pairs"] == int(ora["count"].sum()) and d["n_refs"] == int(ok.sum())
        s_ref = 1.5 * ora["mean"][ok].sum() / ok.sum() - 0.5
        s = 1.5 * d["sum_mean"] / d["n_refs"] - 0.5
        assert abs(s - s_ref) <= RTOL * abs(s_ref)
        if d["tag"] in per_bond:
            assert np.array_equal(per_bond[d["tag"]], ora["count"])


@pytest.mark.parametrize("same", [True, False])
def test_config4_rdf_of_4m_beads(same):
    """calculate_rdfs' pair histogram: 4 000 000 beads, 0-5 sigma, 1000 bins; int64 histogram bit-exact."""
    from mouse_b200.melt import make_melt
    from mouse_b200.workloads import CONFIG4
    from oracle import c_oracle
    eng = _eng()
    melt = make_melt(CONFIG4["chains"], CONFIG4["beads"], seed=0)
    assert melt.n_atoms == 4_000_000
    typ = np.zeros(melt.n_atoms, np.int32)
    h, edges = eng.rdf_frame(melt.positions, typ, melt.mol, 1, melt.box, CONFIG4["rmin"], CONFIG4["rmax"],
                             CONFIG4["nbin"], same_molecule=same)
    exp, e2 = c_oracle.rdf(melt.positions, typ, melt.mol, 1, melt.box, CONFIG4["rmin"], CONFIG4["rmax"],
                           CONFIG4["nbin"], same)
    assert np.array_equal(edges, e2)
    assert np.array_equal(h.cpu().numpy(), exp) and int(exp.sum()) > 1_500_000_000


@pytest.mark.parametrize("triclinic", [False, True])
def test_config5_8m_bonds_with_selections(triclinic):
    """8M bonds, rc = 3 sigma, bond-type neighbour set, residue-type reference set, other-molecule pairs only;
    orthorhombic (reference semantics) and triclinic (build-defined semantics)."""
    from mouse_b200.workloads import config5
    from oracle import c_oracle
    eng = _eng()
    cfg = config5(triclinic)
    n = len(cfg["bond_atoms"])
    assert n == 8_000_000
    res = eng.p2_frame(cfg["positions"], cfg["bond_atoms"], cfg["molb"], cfg["box"], cfg["rcut"], same_molecule=False,
                       nbr_mask=cfg["nbr_mask"], ref_mask=cfg["ref_mask"], tilt=cfg["tilt"])
    assert res.grid.fast == 1 and res.grid.triclinic == int(triclinic)
    vec, ctr = c_oracle.bond_build(cfg["positions"], cfg["bond_atoms"], cfg["box"], tilt=cfg["tilt"])
    assert np.array_equal(res.vec.cpu().numpy().T, vec) and np.array_equal(res.ctr.cpu().numpy().T, ctr)
    ora = c_oracle.p2(vec, ctr, cfg["molb"], cfg["box"], cfg["rcut"], False, ref_mask=cfg["ref_mask"],
                      nbr_mask=cfg["nbr_mask"], tilt=cfg["tilt"])
    assert np.array_equal(res.count.cpu().numpy().astype(np.int64), ora["count"])
    ok = ora["count"] > 0
    np.testing.assert_allclose(res.mean.cpu().numpy()[ok], ora["mean"][ok], rtol=1e-12, atol=1e-14)
    s_ref = 1.5 * ora["mean"][ok].sum() / ok.sum() - 0.5
    assert abs(res.order_parameter() - s_ref) <= RTOL * abs(s_ref)
    assert res.totals_dict()["n_pairs"] == int(ora["count"].sum())
