"""Device-side parser of LAMMPS dump frames (mouse_parse_dump_atoms_dev, csrc/dump_parse.cu) against the host parser
(mouse_parse_dump_atoms, pinned to the reference's readers in tests/test_ingest_cpu.py): the same ids, coordinates and
image flags BIT FOR BIT, whatever the number format, the order of the atoms and the white space; frames the device
parser declines come from the host parser."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mouse_b200 import ingest  # noqa: E402


def _write_dump(path, frames, keys="id type xs ys zs ix iy iz", lo=(-5.0, -4.0, -3.0), hi=(5.0, 6.0, 7.0)):
    """frames: list of (timestep, list of text lines)"""
    with open(path, "wb") as f:
        for ts, n, lines, eol in frames:
            head = "ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (ts, n)
            head += "".join("%r %r\n" % (float(a), float(b)) for a, b in zip(lo, hi)) + "ITEM: ATOMS " + keys + "\n"
            f.write(head.encode())
            f.write(eol.join(lines).encode() + eol.encode())


def _frames(path, num, coords, device=None, stats=None):
    out = []
    for ts, box, ctr, pos, img in ingest.iter_lmp_dump_frames(str(path), num, coords, threads=1, device=device,
                                                              stats=stats):
        if isinstance(pos, torch.Tensor):
            pos = pos.cpu().numpy()
            img = None if img is None else img.cpu().numpy()
        out.append((ts, box, pos.copy(), None if img is None else np.array(img)))
    return out


def _same(a, b):
    assert len(a) == len(b)
    for (ta, ba, pa, ia), (tb, bb, pb, ib) in zip(a, b):
        assert ta == tb and np.array_equal(ba, bb)
        assert pa.dtype == pb.dtype == np.float64
        assert np.array_equal(pa.view(np.int64), pb.view(np.int64))          # bit for bit, signed zeros included
        assert (ia is None) == (ib is None) and (ia is None or np.array_equal(ia, ib))


@pytest.mark.parametrize("coords", ["scaled", "cut", "uncut"])
def test_device_parser_is_bit_identical_to_the_host_parser(tmp_path, coords):
    rng = np.random.default_rng(11)
    n = 20000
    num = np.arange(1, n + 1) * 3 - 1            # atom numbers with gaps
    key = {"scaled": "xs ys zs", "cut": "xc yc zc", "uncut": "xu yu zu"}[coords]
    perm1, perm2 = rng.permutation(n), rng.permutation(n)

    def lines(order, fmt, x, img, decorate=False):
        out = []
        for k, i in enumerate(order):
            ln = ("%d 1 " + fmt + " %d %d %d") % (num[i], x[i, 0], x[i, 1], x[i, 2], img[i, 0], img[i, 1], img[i, 2])
            if decorate:
                ln = ["  " + ln, ln + "  ", "\t" + ln.replace(" ", "   "), ln][k % 4]
                if k % 97 == 0:
                    ln += "\n   "                 # a blank line in the middle of the block
            out.append(ln)
        return out

    frames = []
    xs = [rng.random((n, 3)) * (30.0 if coords != "scaled" else 1.0) - (10.0 if coords != "scaled" else 0.0)
          for _ in range(6)]
    xs[2][::7] = 0.0
    xs[2][1::7] *= 1e-7                           # leading zeros behind the point, tiny exponents
    im = rng.integers(-3, 4, (n, 3))
    frames.append((0, n, lines(np.arange(n), "%.8g %.8g %.8g", xs[0], im), "\n"))
    frames.append((10, n, lines(perm1, "%.15g %.15g %.15g", xs[1], im), "\n"))
    frames.append((20, n, lines(perm1, "%.6e %.6e %+.6e", xs[2], im), "\n"))          # same order: cached permutation
    frames.append((30, n, lines(perm2, "%+.10f %.10f %.10f", xs[3], im, decorate=True), "\r\n"))
    frames.append((40, n, lines(perm2, "%r %r %r", xs[4].astype(object), im), "\n"))  # 17 digits: host parser
    xs[5][:, 1] = np.abs(xs[5][:, 1])            # (written behind an explicit minus sign)
    frames.append((50, n, lines(np.arange(n), "%.12g -%.12g %.3f", xs[5], im), "\n"))
    path = tmp_path / "mixed.dump"
    _write_dump(path, frames, keys="id type " + key + " ix iy iz")
    host = _frames(path, num, coords)
    stats = {}
    dev = _frames(path, num, coords, device="cuda", stats=stats)
    _same(host, dev)
    assert stats == {"frames_device": 5, "frames_host": 1}
    assert [f[0] for f in dev] == [0, 10, 20, 30, 40, 50]
    # the scaled transform really is lo + (hi - lo) * x, un-fused (spot check against NumPy)
    if coords == "scaled":
        lo, hi = np.array([-5.0, -4.0, -3.0]), np.array([5.0, 6.0, 7.0])
        x0 = np.array([[float("%.8g" % v) for v in row] for row in xs[0][:50]])
        assert np.array_equal(dev[0][2][:50], lo + (hi - lo) * x0)


def test_device_parser_without_image_columns_and_with_extra_columns(tmp_path):
    rng = np.random.default_rng(3)
    n = 5000
    num = np.arange(1, n + 1)
    x = rng.random((n, 3))
    order = rng.permutation(n)
    lines = ["%d %d %.9g %.7g %.9g %.9g %d" % (2, num[i], x[i, 2], 1.5e-3 * i, x[i, 0], x[i, 1], i % 5) for i in order]
    path = tmp_path / "cols.dump"
    _write_dump(path, [(5, n, lines, "\n"), (6, n, lines[::-1], "\n")], keys="type id zs q xs ys c_extra")
    stats = {}
    _same(_frames(path, num, "scaled"), _frames(path, num, "scaled", device="cuda", stats=stats))
    assert stats == {"frames_device": 2, "frames_host": 0}


def test_frames_the_device_parser_declines_raise_what_the_host_parser_raises(tmp_path):
    n = 300
    num = np.arange(1, n + 1)
    good = ["%d 1 0.5 0.25 0.125" % i for i in num]
    short = list(good)
    short[177] = "178 1 0.5 0.25"                     # a column is missing
    word = list(good)
    word[12] = "13 1 0.5 abc 0.125"                   # not a number
    unknown = list(good)
    unknown[5] = "9999 1 0.5 0.25 0.125"              # an atom number the topology does not have
    for bad in (short, word, unknown, good[:-1]):
        path = tmp_path / "bad.dump"
        _write_dump(path, [(0, n, good, "\n"), (1, n, bad, "\n")], keys="id type xs ys zs")
        errs = []
        for device in (None, "cuda"):
            with pytest.raises(Exception) as e:
                _frames(path, num, "scaled", device=device)
            errs.append(type(e.value))
        assert errs[0] is errs[1]


def test_trajectory_from_a_dump_gives_the_same_S_with_either_parser(tmp_path):
    """local_ordering_from_dump with parse_on='device' (CUDA positions straight into FramePipeline) and 'host'."""
    from mouse_b200.frames import FrameArrays
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import local_ordering_from_dump
    melt = make_melt(40, 51, seed=5)
    top = FrameArrays.from_melt(melt)
    half = melt.box / 2
    rng = np.random.default_rng(0)
    frames = []
    for k in range(7):
        pos = jitter(melt, k, 0.03)
        xs = (pos + half) / melt.box
        order = rng.permutation(melt.n_atoms) if k in (2, 3, 5) else np.arange(melt.n_atoms)
        frames.append((100 * k, melt.n_atoms, ["%d 1 %.8g %.8g %.8g" % (i + 1, xs[i, 0], xs[i, 1], xs[i, 2])
                                              for i in order], "\n"))
    path = tmp_path / "traj.dump"
    _write_dump(path, frames, keys="id type xs ys zs", lo=tuple(-half), hi=tuple(half))
    sd, sh = {}, {}
    dev, tot_d = local_ordering_from_dump(top, str(path), ["1"], 2.0, "scaled", parse_on="device", stats=sd)
    host, tot_h = local_ordering_from_dump(top, str(path), ["1"], 2.0, "scaled", parse_on="host", stats=sh)
    assert sd == {"frames_device": 7, "frames_host": 0} and sh == {}
    assert dev == host and len(dev) == 7
    assert tot_d.n_pairs == tot_h.n_pairs and tot_d.mean_s == tot_h.mean_s


def _melt_dump(tmp_path, nframes=9):
    from mouse_b200.frames import FrameArrays
    from mouse_b200.melt import jitter, make_melt
    melt = make_melt(40, 51, seed=8)
    half = melt.box / 2
    rng = np.random.default_rng(1)
    frames = []
    for k in range(nframes):
        xs = (jitter(melt, k, 0.03) + half) / melt.box
        order = rng.permutation(melt.n_atoms) if k % 3 == 1 else np.arange(melt.n_atoms)
        frames.append((10 * k, melt.n_atoms, ["%d 1 %.9g %.9g %.9g" % (i + 1, xs[i, 0], xs[i, 1], xs[i, 2])
                                             for i in order], "\n"))
    path = str(tmp_path / "ranks.dump")
    _write_dump(path, frames, keys="id type xs ys zs", lo=tuple(-half), hi=tuple(half))
    return FrameArrays.from_melt(melt), path


def _two_rank_dump_worker(rank, world, port, path, out):
    import torch.distributed as dist
    from mouse_b200.frames import FrameArrays
    from mouse_b200.melt import make_melt
    from mouse_b200.trajectory import local_ordering_from_dump
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    top = FrameArrays.from_melt(make_melt(40, 51, seed=8))
    stats = {}
    s_all, tot = local_ordering_from_dump(top, path, ["1"], 2.0, "scaled", gather=True, stats=stats)   # device=None
    out[rank] = (s_all.tolist(), tot.n_pairs, tot.n_refs, tot.n_frames, dict(stats))
    dist.destroy_process_group()


def test_two_ranks_parse_their_own_frames_on_their_own_gpu(tmp_path):
    """A dump sharded over two ranks: every rank seeks to its block of frames, parses it on ITS device (the parser runs
    on a producer thread, whose current device is not the caller's) and the gathered per-frame S is the one-rank list."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from mouse_b200.trajectory import local_ordering_from_dump
    top, path = _melt_dump(tmp_path)
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29700 + os.getpid() % 90
    procs = [ctx.Process(target=_two_rank_dump_worker, args=(r, 2, port, path, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    per_frame, tot = local_ordering_from_dump(top, path, ["1"], 2.0, "scaled", parse_on="host")
    want = [s for _, s in per_frame]
    assert len(want) == 9
    for r in range(2):
        s_r, npairs, nrefs, nfr, stats = out[r]
        assert s_r == want and (npairs, nrefs, nfr) == (tot.n_pairs, tot.n_refs, 9)
        assert stats["frames_host"] == 0 and stats["frames_device"] in (4, 5)


def test_token_grammar_fuzz_device_equals_host_or_declines(tmp_path, monkeypatch):
    """Every spelling of a number the host parser accepts: the device parser gives the same bits or declines the frame
    (then the host parser's values are what the caller sees).  One small frame per spelling mix."""
    monkeypatch.setattr(ingest.DeviceDumpParser, "max_declines", 10 ** 9)
    rng = np.random.default_rng(2024)
    n = 96
    num = np.arange(1, n + 1)

    def spell(v, kind):
        if kind == 0:
            return "%.*g" % (int(rng.integers(1, 16)), v)
        if kind == 1:
            return "%.*e" % (int(rng.integers(0, 15)), v)
        if kind == 2:
            return "%.*f" % (int(rng.integers(0, 20)), v)
        if kind == 3:
            return ("%.*E" % (int(rng.integers(0, 12)), v)).replace("E+", "E").replace("E0", "E")
        if kind == 4:                                   # bare point forms: "5.", ".5", "-.25"
            t = "%.6f" % v
            t = t.rstrip("0")
            return t.replace("0.", ".", 1) if abs(v) < 1 and rng.random() < 0.7 else t
        if kind == 5:
            return "%+.*g" % (int(rng.integers(1, 16)), v)
        if kind == 6:                                   # zeros in many shapes
            return str(rng.choice(["0", "-0", "0.0", "-0.0", "+0", "0e0", "0.000e-5", "-.0", "00", "000.5", "0e30"]))
        if kind == 7:                                   # limits of the exact path and just beyond
            return str(rng.choice(["1e22", "1e23", "1e-22", "1e-23", "123456789012345", "1234567890123456",
                                   "0.123456789012345", "0.1234567890123456", "999999999999999e7", "999999999999999e8",
                                   "1e-300", "4.9e-324", "1.7976931348623157e308", "1e400", "0." + "0" * 25 + "1",
                                   "9007199254740993", "1" + "0" * 22, "0.3e-22"]))
        if kind == 8:
            return repr(float(v))                       # 16-17 digits
        return str(rng.choice(["nan", "inf", "-inf", "1e", "1e+", "--1", "1.2.3", "0x10", "1_0", "1d5", "NaN"]))

    frames, expect_decline = [], []
    for k in range(120):
        if k % 10 == 0:
            kinds = [9, 0, 0]
        elif k % 2:
            kinds = rng.choice([0, 1, 3, 4, 5, 6], size=3)          # spellings inside the exact path: device frames
        else:
            kinds = rng.choice(9, size=3, p=[.2, .15, .15, .1, .1, .1, .08, .07, .05])
        # (15 digits at 1e-9 need 10^-23: outside the exact path - the odd frames stay at magnitudes LAMMPS dumps have)
        x = rng.standard_normal((n, 3)) * 10.0 ** (rng.integers(-2, 4) if k % 2 else rng.integers(-8, 8))
        lines = []
        for i in range(n):
            tok = [spell(x[i, a], int(kinds[a])) for a in range(3)]
            if k % 10 == 0 and i != n // 2:
                tok[0] = "0.5"                          # ONE malformed token in the frame
            lines.append("%d 1 %s %s %s" % (num[i], tok[0], tok[1], tok[2]))
        frames.append((k, n, lines, "\n"))
    path = tmp_path / "fuzz.dump"
    _write_dump(path, frames, keys="id type xu yu zu")
    index = ingest.index_lmp_dump(str(path))
    ndev = nhost = nraise = 0
    for k in range(len(frames)):
        got = []
        for device in (None, "cuda"):
            st = {}
            try:
                fr = [(ts, pos.cpu().numpy() if isinstance(pos, torch.Tensor) else pos)
                      for ts, _, _, pos, _ in ingest.iter_lmp_dump_frames(str(path), num, "uncut", frames=(k, k + 1),
                                                                         index=index, device=device, stats=st)]
                got.append((fr, st))
            except Exception as e:      # noqa: BLE001
                got.append(type(e))
        if isinstance(got[0], type):
            assert got[1] is got[0]
            nraise += 1
            continue
        (host, _), (dev, st) = got
        assert host[0][0] == dev[0][0] == k
        assert np.array_equal(host[0][1].view(np.int64), dev[0][1].view(np.int64)), frames[k][2][:3]
        ndev += st["frames_device"]
        nhost += st["frames_host"]
    assert nraise >= 6 and ndev >= 30 and nhost >= 20, (nraise, ndev, nhost)


def test_rdf_trajectory_from_a_dump_against_the_oracle(tmp_path):
    """pair_correlation_from_dump: histograms summed over the frames of a dump, device parser and host parser, all
    types and a type selection, against the C oracle on the host-parsed frames (integers: bit for bit)."""
    from mouse_b200.frames import FrameArrays
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import pair_correlation_from_dump
    from oracle import c_oracle
    melt = make_melt(30, 40, seed=9)
    top = FrameArrays.from_melt(melt)
    types = [str(1 + (i % 3)) for i in range(melt.n_atoms)]
    top.atom_type = types
    half = melt.box / 2
    frames = []
    for k in range(5):
        xs = (jitter(melt, k, 0.04) + half) / melt.box
        frames.append((k, melt.n_atoms, ["%d %s %.9g %.9g %.9g" % (i + 1, types[i], xs[i, 0], xs[i, 1], xs[i, 2])
                                         for i in range(melt.n_atoms)], "\n"))
    path = str(tmp_path / "rdf.dump")
    _write_dump(path, frames, keys="id type xs ys zs", lo=tuple(-half), hi=tuple(half))
    host_frames = list(ingest.iter_lmp_dump_frames(path, top.atom_num, "scaled"))
    molc = top.atom_mol_code()
    for sel, same in ((None, True), (["1", "3"], False)):
        names = sorted(set(types)) if sel is None else sorted(sel)
        code = {t: k for k, t in enumerate(names)}
        rows = np.array([i for i, t in enumerate(types) if t in code])
        tc = np.array([code[types[i]] for i in rows], np.int32)
        want = 0
        for _, box, _, pos, _ in host_frames:
            h, edges = c_oracle.rdf(pos[rows], tc, molc[rows].astype(np.int64), len(names), box, 0.0, 2.5, 60,
                                    same_molecule=same)
            want = want + h
        for parse_on in ("device", "host"):
            st = {}
            hist, e2, got_names, tot = pair_correlation_from_dump(top, path, 0.0, 2.5, 60, atomtypes=sel,
                                                                  same_molecule=same, parse_on=parse_on, stats=st)
            assert got_names == names and tot.n_frames == 5 and np.array_equal(e2, edges)
            assert hist.dtype == np.int64 and np.array_equal(hist, want)
            assert st == ({"frames_device": 5, "frames_host": 0} if parse_on == "device" else {})
