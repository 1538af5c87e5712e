"""CPU (gloo, world_size 2) tests of the frame sharding and of the final all-reduce packing - the
only multi-rank logic of the path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mouse_b200.trajectory import TrajectoryTotals, all_reduce_totals, gather_per_frame, shard_range

SBINS, RDF_SHAPE = 13, (3, 17)


def test_shard_range_partitions_frames():
    for n in (0, 1, 7, 8, 100, 10_000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_frame(f):
    rng = np.random.default_rng(f)
    n_refs = int(rng.integers(0, 50)) if f % 5 else 0        # every 5th frame has no neighbours at all
    return float(rng.random() * n_refs / 3.0), n_refs, int(rng.integers(0, 1000)) * (n_refs > 0)


def _fake_hists(f):
    rng = np.random.default_rng(1000 + f)
    return rng.integers(0, 2**40, SBINS), rng.integers(0, 2**44, RDF_SHAPE), int(f % 3 == 0)


def _accumulate(frames):
    t = TrajectoryTotals(p2_hist=np.zeros(SBINS, np.int64), rdf_hist=np.zeros(RDF_SHAPE, np.int64))
    per_frame = []
    for f in frames:
        sm, nr, npairs = _fake_frame(f)
        h, r, over = _fake_hists(f)
        t.add_frame(sm, nr, npairs, hist=h, n_hist_over=over)
        t.rdf_hist += r
        per_frame.append((f, 1.5 * sm / nr - 0.5 if nr else float("nan")))
    return t, per_frame


def _worker(rank, world, port, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_frames, rank, world)
    t, per_frame = _accumulate(range(lo, hi))
    red = all_reduce_totals(t, torch.device("cpu"))
    s_all = gather_per_frame(per_frame, n_frames, torch.device("cpu"))
    out[rank] = (red.sum_s, red.n_frames, red.sum_mean, red.n_refs, red.n_pairs, red.n_skipped_frames,
                 red.n_hist_over, red.p2_hist.tolist(), red.rdf_hist.tolist(), s_all.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [1, 23])
def test_all_reduce_totals_gloo_world2(n_frames):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300 + n_frames
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, out)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    serial, per_frame = _accumulate(range(n_frames))
    s_serial = gather_per_frame(per_frame, n_frames, torch.device("cpu"))
    for r in range(world):
        s_s, nf, sm, nr, npairs, nsk, nover, p2h, rdfh, s_all = out[r]
        assert (nf, nr, npairs, nsk, nover) == (serial.n_frames, serial.n_refs, serial.n_pairs,
                                                serial.n_skipped_frames, serial.n_hist_over)
        assert s_s == pytest.approx(serial.sum_s, rel=1e-14, abs=1e-14)
        assert sm == pytest.approx(serial.sum_mean, rel=1e-14, abs=1e-14)
        # the histograms reduce EXACTLY (integers riding in the one float64 buffer of the all-reduce)
        assert p2h == serial.p2_hist.tolist() and rdfh == serial.rdf_hist.tolist()
        # per-frame S in input order on every rank (NaN = frame without a value), as the reference's CLI prints them
        assert np.array_equal(np.array(s_all), s_serial, equal_nan=True)


def test_pack_refuses_totals_too_large_for_an_exact_float64_reduction():
    t = TrajectoryTotals(n_pairs=2**52)
    with pytest.raises(OverflowError):
        t.pack(torch.device("cpu"))
    t = TrajectoryTotals(n_pairs=2**40, p2_hist=np.arange(5))
    buf = t.pack(torch.device("cpu"))
    assert buf.dtype == torch.float64 and buf.numel() == 7 + 5
    back = TrajectoryTotals(p2_hist=np.zeros(5, np.int64)).unpack_into(buf)
    assert back.n_pairs == 2**40 and back.p2_hist.tolist() == [0, 1, 2, 3, 4]


def test_trajectory_histogram_dict_follows_the_reference_rules():
    t = TrajectoryTotals(p2_hist=np.array([1, 3, 0, 4]))
    h = t.p2_histogram(-0.5, 1.0, 4)
    assert h["values"] == [0.125, 0.375, 0.0, 0.5] and h["step"] == 0.5
    t.n_hist_over = 1
    with pytest.raises(IndexError):
        t.p2_histogram(-0.5, 1.0, 4)


def test_totals_mean_s():
    t = TrajectoryTotals()
    t.add_frame(10.0, 30, 500)      # S = 1.5*10/30 - 0.5 = 0
    t.add_frame(20.0, 30, 100)      # S = 0.5
    t.add_frame(0.0, 0, 0)          # skipped frame
    assert t.n_frames == 2 and t.n_skipped_frames == 1 and t.mean_s == pytest.approx(0.25)
