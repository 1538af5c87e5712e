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


def test_dump_index_frame_ranges_and_threaded_parser(tmp_path):
    """The byte-offset index lets a rank read only its own block of frames; the threaded number parser gives the
    single-threaded values; a shuffled atom order is mapped back through the id column (cached per frame order)."""
    import ctypes as C
    from mouse_b200 import _lib
    fa = ingest.read_lmp_data_arrays(DATA)
    index = ingest.index_lmp_dump(DUMP)
    assert len(index) == 4 and index[0] == 0 and ingest.count_lmp_dump_frames(DUMP, index) == 3
    full = ingest.read_lmp_dump_arrays(DUMP, fa.atom_num, "scaled")
    for lo, hi in ((0, 1), (1, 3), (2, 3), (0, 3)):
        ts, boxes, pos = ingest.read_lmp_dump_arrays(DUMP, fa.atom_num, "scaled", frames=(lo, hi), index=index, threads=4)
        assert np.array_equal(ts, full[0][lo:hi]) and np.array_equal(boxes, full[1][lo:hi])
        assert np.array_equal(pos, full[2][lo:hi])
    # a bigger synthetic dump with the atoms in a different order in every frame, parsed on several threads
    rng = np.random.default_rng(7)
    n, nf = 20000, 3
    num = np.arange(1, n + 1)
    path = tmp_path / "big.dump"
    want = []
    with open(path, "w") as f:
        for k in range(nf):
            xs = rng.random((n, 3))
            want.append(-5.0 + 10.0 * xs)
            order = rng.permutation(n) if k != 1 else np.arange(n)
            f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (100 * k, n))
            f.write("-5.0 5.0\n-5.0 5.0\n-5.0 5.0\nITEM: ATOMS id type xs ys zs\n")
            f.write("".join("%d 1 %r %r %r\n" % (num[i], float(xs[i, 0]), float(xs[i, 1]), float(xs[i, 2])) for i in order))
    for threads in (1, 3, 8):
        ts, boxes, pos = ingest.read_lmp_dump_arrays(str(path), num, "scaled", threads=threads)
        assert ts.tolist() == [0, 100, 200] and np.array_equal(boxes, np.full((3, 3), 10.0))
        for k in range(nf):
            assert np.array_equal(pos[k], want[k])
    text = open(path, "rb").read()
    blk = text[text.index(b"ITEM: ATOMS"):]
    blk = blk[blk.index(b"\n") + 1:blk.index(b"ITEM: TIMESTEP")]
    a, b = np.empty(5 * n + 9), np.empty(5 * n + 9)
    L = _lib.lib()
    assert L.mouse_parse_doubles(blk, len(blk), C.c_void_p(a.ctypes.data), 5 * n) == 5 * n
    assert L.mouse_parse_doubles_mt(blk, len(blk), C.c_void_p(b.ctypes.data), 5 * n, 7) == 5 * n
    assert np.array_equal(a[:5 * n], b[:5 * n])
    bad = blk[:300000] + b" nope " + blk[300000:]
    assert L.mouse_parse_doubles(bad, len(bad), C.c_void_p(a.ctypes.data), 5 * n + 9) == \
        L.mouse_parse_doubles_mt(bad, len(bad), C.c_void_p(b.ctypes.data), 5 * n + 9, 7) < 0


def test_dense_molecule_codes_are_collision_free_and_stable():
    fa = ingest.read_lmp_data_arrays(DATA)
    code = fa.atom_mol_code()
    assert code.dtype == np.int32 and code.min() >= 1          # no empty mol_id in this file
    assert len(set(code.tolist())) == len(set(fa.atom_mol))    # one code per molecule id, no collisions
    first = {}
    for m, c in zip(fa.atom_mol, code):
        assert first.setdefault(m, c) == c
    assert np.array_equal(fa.bond_mol_code(), code[fa.bond_atoms[:, 0]])


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


def test_parse_doubles_is_correctly_rounded():
    """mouse_parse_doubles (the dump reader's number parser): bit-identical to Python float() on fixed, general,
    exponent, long-mantissa and special tokens; reports the offset of a token that is not a number."""
    import ctypes as C
    import random
    from mouse_b200 import _lib
    L = _lib.lib()
    rng = random.Random(7)
    toks = ["0", "-0", "0.0", "-0.000", "1e5", "1E-5", "+3.5", ".5", "5.", "0.000001234", "123456789012345",
            "1234567890123456", "1e22", "1e23", "1e-22", "1e-23", "000123.4500", "9007199254740993", "1e308", "4.9e-324",
            "2.2250738585072014e-308", "0.1", "0.30000000000000004", "nan", "inf", "-inf"]
    for _ in range(20000):
        k = rng.random()
        if k < 0.3:
            toks.append("%.6f" % rng.uniform(-100, 100))
        elif k < 0.5:
            toks.append("%.10g" % rng.uniform(-1e5, 1e5))
        elif k < 0.65:
            toks.append(repr(rng.uniform(-1, 1)))
        elif k < 0.75:
            toks.append("%d" % rng.randint(-10**9, 10**9))
        elif k < 0.9:
            toks.append("%.8e" % (rng.uniform(-1, 1) * 10.0**rng.randint(-30, 30)))
        else:
            toks.append("%.15g" % rng.uniform(-1e3, 1e3))
    text = (" ".join(toks[:5000]) + "\n" + "\t".join(toks[5000:]) + " \r\n").encode()
    out = np.empty(len(toks))
    assert L.mouse_parse_doubles(text, len(text), C.c_void_p(out.ctypes.data), out.size) == len(toks)
    exp = np.array([float(t) for t in toks])
    assert np.array_equal(out.view(np.int64)[~np.isnan(exp)], exp.view(np.int64)[~np.isnan(exp)])
    assert np.array_equal(np.isnan(out), np.isnan(exp))
    bad = b"1.0 2.0 abc 3.0"
    assert L.mouse_parse_doubles(bad, len(bad), C.c_void_p(out.ctypes.data), out.size) == -(bad.index(b"abc") + 1)
    assert L.mouse_parse_doubles(b"1 2 3", 5, C.c_void_p(out.ctypes.data), 2) == -(4 + 1)     # does not fit
    assert L.mouse_parse_doubles(b"  \n", 3, C.c_void_p(out.ctypes.data), 2) == 0


def test_threaded_frame_index_finds_exactly_the_line_starts(tmp_path):
    """mouse_index_dump (memmem over several chunks) against a plain scan: matches inside a line do not count, matches
    that straddle a chunk border are found once, the file size closes the index."""
    rng = np.random.default_rng(5)
    path = tmp_path / "idx.dump"
    want, pos = [], 0
    with open(path, "wb") as f:
        for k in range(400):
            head = b"ITEM: TIMESTEP\n%d\n" % k
            want.append(pos)
            body = b"".join(rng.choice([b"1 2 3\n", b"x ITEM: TIMESTEP\n", b"ITEM: TIMESTE\n", b"0.5 ITEM: ATOMS\n"],
                                       size=int(rng.integers(500, 4000))))
            f.write(head + body)
            pos += len(head) + len(body)
    index = ingest.index_lmp_dump(str(path))
    assert index.dtype == np.int64 and index[-1] == pos and index[:-1].tolist() == want
    import ctypes as C
    from mouse_b200 import _lib
    text = open(path, "rb").read()
    L = _lib.lib()
    for threads in (1, 2, 7, 16):
        out = np.empty(len(want) + 5, np.int64)
        assert L.mouse_index_dump(text, len(text), C.c_void_p(out.ctypes.data), len(out), threads) == len(want)
        assert out[:len(want)].tolist() == want
    out = np.empty(3, np.int64)          # a short array: the count is still complete
    assert L.mouse_index_dump(text, len(text), C.c_void_p(out.ctypes.data), 3, 4) == len(want) and out.tolist() == want[:3]


def _same_frame_arrays(a, b):
    for k in ("positions", "box", "atom_num", "bond_atoms", "bond_num"):
        x, y = getattr(a, k), getattr(b, k)
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), k
    for k in ("atom_type", "atom_mol", "atom_res", "bond_type"):
        assert list(getattr(a, k)) == list(getattr(b, k)), k


def test_table_reader_of_data_files_equals_the_line_reader(tmp_path):
    """read_lmp_data_arrays takes files of the plain LAMMPS shape through mouse_parse_table (threaded, in place) and
    everything else through the line-by-line reader: same arrays, same token strings either way."""
    import glob
    from mouse_b200.melt import make_melt, write_lammps_data
    for path in sorted(glob.glob(os.path.join(os.path.dirname(DATA), "*.data"))):
        _same_frame_arrays(ingest.read_lmp_data_arrays(path), ingest.read_lmp_data_arrays(path, fast=False))
    melt = make_melt(60, 80, seed=4)                       # 4800 atoms: the threaded branch of the table parser
    plain = str(tmp_path / "plain.data")
    write_lammps_data(melt, plain)
    assert ingest._read_lmp_data_fast(plain, True, 5) is not None
    for threads in (1, 3, 8):
        _same_frame_arrays(ingest.read_lmp_data_arrays(plain, threads=threads), ingest.read_lmp_data_arrays(plain, fast=False))
    text = open(plain).read()
    head, rest = text.split("Atoms", 1)
    atoms, bonds = rest.split("Bonds", 1)
    atom_lines = atoms.split("\n")[2:]
    atom_lines = atom_lines[:melt.n_atoms]
    bond_lines = bonds.split("\n")[2:][:melt.n_bonds]

    def build(alines, blines, between="", after="", atoms_head="Atoms # molecular"):
        return head + atoms_head + "\n\n" + "\n".join(alines) + "\n" + between + "\nBonds\n\n" + "\n".join(blines) + "\n" + after

    variants = {
        # shapes the table parser takes
        "velocities": (build(atom_lines, bond_lines, between="\nVelocities\n\n" + "\n".join(
            "%d 0.1 -0.2 0.3" % (i + 1) for i in range(melt.n_atoms)) + "\n"), True),
        "charge": (build([" ".join(ln.split()[:3] + ["-0.5"] + ln.split()[3:]) for ln in atom_lines], bond_lines), True),
        "angles_after": (build(atom_lines, bond_lines, after="\nAngles\n\n1 1 1 2 3\n"), True),
        "negative_mol": (build([" ".join([ln.split()[0], "-7"] + ln.split()[2:]) for ln in atom_lines], bond_lines), True),
        # shapes it declines: the line reader decides (same result, or the same exception)
        "type_label": (build([" ".join(ln.split()[:2] + ["CA"] + ln.split()[3:]) for ln in atom_lines], bond_lines), False),
        "leading_zero": (build([" ".join(ln.split()[:2] + ["01"] + ln.split()[3:]) for ln in atom_lines], bond_lines), False),
        "plus_sign": (build(atom_lines, [" ".join([w[0], "+1"] + w[2:]) for w in (ln.split() for ln in bond_lines)]), False),
        "ragged": (build([ln + (" 9" if k == 17 else "") for k, ln in enumerate(atom_lines)], bond_lines), False),
        "bonds_first": (head + "Bonds\n\n" + "\n".join(bond_lines) + "\n\nAtoms\n\n" + "\n".join(atom_lines) + "\n", False),
        "header_word_later": (build(atom_lines, bond_lines, after="\n5 atoms\n"), False),
    }
    for name, (body, takes) in variants.items():
        path = str(tmp_path / (name + ".data"))
        with open(path, "w") as f:
            f.write(body)
        assert (ingest._read_lmp_data_fast(path, True, 4) is not None) == takes, name
        try:
            want = ingest.read_lmp_data_arrays(path, fast=False)
        except Exception as e:      # noqa: BLE001 - whatever the line reader raises, the default entry raises too
            with pytest.raises(type(e)):
                ingest.read_lmp_data_arrays(path)
            continue
        _same_frame_arrays(ingest.read_lmp_data_arrays(path), want)
