"""Ingestion row: the array readers against golden vectors produced by the reference's
own ``read_lmp_data`` / ``return_extra_frames_from_lmp_dump`` (tests/golden/make_golden_ingest.py)."""
import os

import numpy as np
import pytest

from mouse_b200 import ingest
from mouse_b200.frames import FrameArrays
from mouse_b200.lammpack_misc import read_data_typeselect

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA, DUMP = os.path.join(HERE, "melt_290.data"), os.path.join(HERE, "melt_290.dump")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "ingest_290.npz"))


def test_data_file_arrays_match_the_reference_reader(golden):
    fa = ingest.read_lmp_data_arrays(DATA)
    assert np.array_equal(fa.positions, golden["data_pos"]) and np.array_equal(fa.box, golden["data_box"])
    assert np.array_equal(fa.atom_num, golden["data_num"])
    assert fa.atom_mol == [str(s) for s in golden["data_mol"]] and fa.atom_type == [str(s) for s in golden["data_type"]]
    assert np.array_equal(fa.atom_num[fa.bond_atoms], golden["bond_atoms"])          # atom1 < atom2, reference order
    assert np.array_equal(fa.bond_num, golden["bond_num"]) and fa.bond_type == [str(s) for s in golden["bond_type"]]


def test_data_file_arrays_equal_the_object_path():
    fa = ingest.read_lmp_data_arrays(DATA)
    fb = FrameArrays.from_config(read_data_typeselect(DATA))
    assert np.array_equal(fa.positions, fb.positions) and np.array_equal(fa.box, fb.box)
    assert np.array_equal(fa.bond_atoms, fb.bond_atoms) and fa.bond_type == fb.bond_type
    assert fa.atom_mol == fb.atom_mol and np.array_equal(fa.bond_mol_hash(), fb.bond_mol_hash())


@pytest.mark.parametrize("coords", ["scaled", "cut", "uncut"])
def test_dump_frames_match_the_reference_reader(golden, coords):
    fa = ingest.read_lmp_data_arrays(DATA)
    ts, boxes, pos = ingest.read_lmp_dump_arrays(DUMP, fa.atom_num, coords)
    assert ts.tolist() == [0, 1000, 2000] and pos.shape == (3, 300, 3)
    # the reference appends a frame only when it meets the NEXT timestep header: it returns the first two
    ref_pos, ref_box = golden["dump_%s_pos" % coords], golden["dump_%s_box" % coords]
    assert np.array_equal(golden["dump_%s_timestep" % coords], ts[:2])
    assert np.array_equal(pos[:2], ref_pos) and np.array_equal(boxes[:2], ref_box)
    frames = list(ingest.iter_lmp_dump_frames(DUMP, fa.atom_num, coords))
    assert frames[0][4] is not None and frames[0][4].shape == (300, 3)              # image flags are carried along


def test_dump_atom_count_mismatch_raises_like_the_reference():
    with pytest.raises(NameError):
        list(ingest.iter_lmp_dump_frames(DUMP, np.arange(1, 10), "scaled"))
    with pytest.raises(ValueError):
        list(ingest.iter_lmp_dump_frames(DUMP, np.arange(1, 301), "fractional"))


def test_array_selections_equal_the_object_selections():
    """--chains / --subcell on arrays vs the Config-rebuilding mirrors of the reference's selectByAtomFields /
    select_rectangular."""
    from mouse_b200.lammpack_misc import select_chains, select_rectangular
    cfg = read_data_typeselect(DATA)
    fa = ingest.read_lmp_data_arrays(DATA)
    for got, exp in ((fa.select_chains(["1", "3"]), select_chains(cfg, ["1", "3"])),
                     (fa.select_rectangular(True, -0.2, 0.3, False, 0., 1., True, -0.5, 0.1),
                      select_rectangular(cfg, True, -0.2, 0.3, False, 0., 1., True, -0.5, 0.1))):
        ref = FrameArrays.from_config(exp)
        assert got.n_atoms == ref.n_atoms and got.n_bonds == ref.n_bonds and got.n_bonds > 0
        assert np.array_equal(got.positions, ref.positions) and np.array_equal(got.atom_num, ref.atom_num)
        assert np.array_equal(got.bond_atoms, ref.bond_atoms) and np.array_equal(got.bond_num, ref.bond_num)
        assert got.atom_mol == ref.atom_mol and got.bond_type == ref.bond_type


def test_write_pdb_matches_the_reference_file(tmp_path):
    """Config.write_pdb(hide_pbc_bonds=True) of the mirror object model writes, byte for byte, the file the
    reference writes for the same LAMMPS data file (tests/golden/make_golden_pdb.py)."""
    import os
    from mouse_b200.lammpack_misc import read_data_typeselect
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cfg = read_data_typeselect(os.path.join(gold, "melt_290.data"))
    out = tmp_path / "plain.pdb"
    cfg.write_pdb(str(out), hide_pbc_bonds=True)
    assert out.read_text() == open(os.path.join(gold, "melt_290_plain.pdb")).read()
    # without hiding, every bond gets its CONECT record
    cfg.write_pdb(str(out))
    assert sum(l.startswith("CONECT") for l in out.read_text().splitlines()) == cfg.n_bond()
