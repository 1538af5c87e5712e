"""CPU (gloo, world_size 2) tests of the frame sharding and of the final all-reduce packing - the
only multi-rank logic of the path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mouse_b200.trajectory import TrajectoryTotals, all_reduce_totals, shard_range


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


def _worker(rank, world, port, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_frames, rank, world)
    t = TrajectoryTotals()
    for f in range(lo, hi):
        t.add_frame(*_fake_frame(f))
    red = all_reduce_totals(t, torch.device("cpu"))
    out[rank] = (red.sum_s, red.n_frames, red.sum_mean, red.n_refs, red.n_pairs, red.n_skipped_frames)
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
    serial = TrajectoryTotals()
    for f in range(n_frames):
        serial.add_frame(*_fake_frame(f))
    for r in range(world):
        s_s, nf, sm, nr, npairs, nsk = out[r]
        assert (nf, nr, npairs, nsk) == (serial.n_frames, serial.n_refs, serial.n_pairs, serial.n_skipped_frames)
        assert s_s == pytest.approx(serial.sum_s, rel=1e-14, abs=1e-14)
        assert sm == pytest.approx(serial.sum_mean, rel=1e-14, abs=1e-14)


def test_totals_mean_s():
    t = TrajectoryTotals()
    t.add_frame(10.0, 30, 500)      # S = 1.5*10/30 - 0.5 = 0
    t.add_frame(20.0, 30, 100)      # S = 0.5
    t.add_frame(0.0, 0, 0)          # skipped frame
    assert t.n_frames == 2 and t.n_skipped_frames == 1 and t.mean_s == pytest.approx(0.25)
