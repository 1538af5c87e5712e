"""CPU-only checks of the host logic and of the C-ABI library (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mouse_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "mouse_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mouse_[a-z0-9_]+)\s*\(", header))
    assert declared and declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().mouse_version()


def test_grid_plan_host():
    from mouse_b200 import _lib
    g = _lib.grid_plan([105.92, 105.92, 105.92], 2.5, 1_000_000)
    assert list(g.ncell_axis) == [42, 42, 42] and g.fast == 1 and g.rc2 == 2.5**2
    assert all(w >= 2.5 * (1 + 1e-6) for w in g.cell_w)
    g = _lib.grid_plan([10., 10., 10.], 4.0, 100)            # fewer than 3 cells -> exact sweep
    assert list(g.ncell_axis) == [2, 2, 2] and g.fast == 0
    g = _lib.grid_plan([10., 10., 10.], -1.0, 100)           # no cutoff
    assert g.ncell == 1 and g.fast == 0
    g = _lib.grid_plan([1000., 1000., 1000.], 1.0, 100)      # dilute: cells capped near 4 * n
    assert g.ncell <= 4 * 100 + 64 and all(w >= 1.0 for w in g.cell_w)
    # sparse cells (large system, few items per rc-sized cell): cells widened to ~10 items each
    g = _lib.grid_plan([105.57, 105.57, 105.57], 1.5, 999_999)
    assert list(g.ncell_axis) == [47, 47, 47] and g.fast == 1 and 9.0 < 999_999 / g.ncell < 11.0
    g = _lib.grid_plan([22.7, 22.7, 22.7], 1.5, 9_900)       # small system: the fine grid stays
    assert list(g.ncell_axis) == [15, 15, 15]
    with pytest.raises(RuntimeError):
        _lib.grid_plan([0., 1., 1.], 1.0, 10)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mouse_b200 import engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.p2_frame(np.zeros((2, 3)), np.array([[0, 1]], np.int32), None, [5., 5., 5.], 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mouse_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|libmouse_oracle|oracle[/.](c_oracle|p2_oracle|"
                                     r"rdf_oracle|_build|_ref)", src, flags=re.M), os.path.join(dirpath, f)


def test_frame_arrays_masks_and_reader(tmp_path):
    from mouse_b200.frames import FrameArrays
    from mouse_b200.lammpack_misc import read_data_typeselect, select_chains, select_rectangular
    from mouse_b200.melt import make_melt, write_lammps_data
    melt = make_melt(6, 12, seed=3, n_bond_types=2, n_res_types=2)
    path = str(tmp_path / "m.data")
    write_lammps_data(melt, path)
    cfg = read_data_typeselect(path)
    assert cfg.n_atom() == melt.n_atoms and cfg.n_bond() == melt.n_bonds
    fa = FrameArrays.from_config(cfg)
    assert np.array_equal(fa.positions, melt.positions) and np.array_equal(fa.bond_atoms, melt.bond_atoms)
    assert np.allclose(fa.box, melt.box, rtol=0, atol=1e-12)
    m = fa.bond_mask(bond_types=["2"])
    assert np.array_equal(m, melt.bond_type == 1)
    assert fa.bond_mask(bond_types=None).all() and fa.bond_mask(bond_types=[]).all()
    sub = select_chains(cfg, ["1", "3"])
    assert sub.n_atom() == 24 and sub.n_bond() == 22
    sub = select_rectangular(cfg, True, -0.25, 0.25)
    assert 0 < sub.n_atom() < cfg.n_atom()
    assert all(b.atom1.num in sub._atoms_by_num and b.atom2.num in sub._atoms_by_num for b in sub.bonds())


def test_golden_data_file_matches_reference_reader():
    """tests/golden/melt_290.data read by our reader == the arrays the reference was given."""
    from mouse_b200.frames import FrameArrays
    from mouse_b200.lammpack_misc import read_data_typeselect
    from tests.util import GOLDEN_DIR, load_golden
    melt, _ = load_golden("melt_290_rc15_same")
    fa = FrameArrays.from_config(read_data_typeselect(os.path.join(GOLDEN_DIR, "melt_290.data")))
    assert np.array_equal(fa.positions, melt.positions) and np.array_equal(fa.bond_atoms, melt.bond_atoms)
