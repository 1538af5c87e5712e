#!/usr/bin/env python3
"""Benchmark of the local-P2 hot path (BASELINE.json metric: ordered in-cutoff pairs/s for local P2
ordering, 1M-bond melt, rc = 2.5 sigma, at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                   the reference's own CPU algorithm, timed here

A "step" is one pass of the hot path (K1 bond build -> K2 cell list -> K3 pair sweep -> K4 reduction)
over one trajectory frame of the workload on every rank (weak scaling: frames are sharded, every GPU
processes its own frames, ONE NCCL all-reduce of the packed sums at the end of the timed region).
Prints ONE JSON line on rank 0.  Besides the contract's keys the line carries `configs`: driver-timed
sub-measurements of the other BASELINE.json configurations (3: 250k-bond trajectory frames, 4: RDF of 4M
beads, 5: 8M bonds with selections, orthorhombic and triclinic).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # the configuration BASELINE.json's metric is quoted on (default)
    "config2": dict(name="config2: single-frame 1M-bond melt (10000 chains x 101 beads), rc=2.5 sigma, "
                         "local P2 per vector + system average", chains=10000, beads=101, rcut=2.5),
    # the frame-sharded trajectory configuration (`--workload config3`; also measured inside `configs`)
    "config3": dict(name="config3: frames of a 250k-bond melt (2500 chains x 101 beads), rc=2.5 sigma, frames sharded "
                         "over the ranks", chains=2500, beads=101, rcut=2.5),
}
WORKLOAD = WORKLOADS["config2"]
METRIC = "pairs/sec for local P2 ordering (1M bonds, rc=2.5σ)"
UNIT = "pairs/s"
W_LANE_INSTR_PER_PAIR = 55.2      # 6.45 x 7 FP32 + 10 FP64 lane-instructions per counted pair (K3)
W5_LANE_INSTR_PER_PAIR = 60.2     # K5: 6.45 x 7 FP32 filter + 15 per counted pair (FP64 d^2, bin, count)
BYTES_PER_BOND = 72.0             # compulsory HBM bytes per bond per frame
N_FRAME_VARIANTS = 4              # distinct frames cycled (working set per frame ~210 MB > 126 MB L2)
LANES_SMALL = int(os.environ.get("MOUSE_BENCH_LANES_SMALL", "6"))             # ... for config 3's 250k-bond frames (the tail of a short sweep is a larger share)
LANES = int(os.environ.get("MOUSE_BENCH_LANES", "4"))                         # frames in flight in the device-resident timed region (own stream + workspace each)
KERNELS_PER_FRAME = 9             # bond_assign, scan_reduce, scan_apply2, scatter_idx, place, items, sweep, fixup, finish+reduce


def static_config(n_bonds):
    """`config` is the same dict in both arms (it names the workload; measured details live in `details`)."""
    return {"workload": WORKLOAD["name"], "bonds_per_frame": int(n_bonds), "rcut": WORKLOAD["rcut"],
            "frames_per_step_per_gpu": 1,
            "l2": "~%d MB touched per frame (L2 is 126 MB), %d distinct frames cycled" % (int(230e-6 * n_bonds),
                                                                                       N_FRAME_VARIANTS)}


def kernel_source_hash():
    """sha256 (16 hex) of the sources of the sweep kernel: profiles/*.json measured on other sources are stale."""
    h = hashlib.sha256()
    for name in ("p2_sweep.cu", "half_shell.cuh", "common.cuh"):
        h.update(open(os.path.join(ROOT, "mouse_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (nvidia_ml_py) on a thread:
    a query costs microseconds and does not stall kernel launches the way a polling `nvidia-smi -lms` child does
    (observed: tens of ms of blocked submission per query); falls back to nvidia-smi when NVML is unavailable."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=0.01):
        self.index, self.samples, self.proc, self.period = index, [], None, period
        self._stop = threading.Event()
        self._thread = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(reasons_fn(h))
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        act = lambda k: "Active" if r & bits[k] else "Not Active"  # noqa: E731
                        self.samples.append([str(sm), str(mx), "%.1f" % pw, act("hw_slowdown"), act("hw_thermal_slowdown"),
                                             act("sw_thermal_slowdown"), act("sw_power_cap")])
                    except Exception:
                        pass
                    self._stop.wait(self.period)
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 7:
                self.samples.append(p)

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = []
        for k, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(s[k].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "how": self.how}


# ------------------------------------------------------------------------------------------- CPU legs
_G = {}   # data shared with forked workers (set before the pool is created)


def _port_worker(refs):
    """Literal NumPy restatement (oracle port) of ordering_functions.py:91-119 for a block of reference bonds."""
    from oracle import p2_oracle as po
    np_bonds, box, vec, ctr, rcut = _G["np_bonds"], _G["box"], _G["vec"], _G["ctr"], _G["rcut"]
    t0 = time.perf_counter()
    pairs = 0
    for i in refs:
        try:
            _, c, _ = po.ave_cos_sq_literal(np_bonds, box, None, vec[i], ctr[i], i + 1, np_bonds[2][i], rcut, True, True)
            pairs += c
        except po.NoValues:
            pass
    return pairs, time.perf_counter() - t0


def _literal_worker(refs):
    """The reference's own calculateAveCosSqForReference (oracle/_ref, unmodified) for a block of reference bonds;
    the number of counted neighbours is read off np.ma.average's argument (oracle/reference_harness._AverageSpy)."""
    from oracle import reference_harness as rh
    of, cfg, np_bonds, rcut = _G["of"], _G["cfg"], _G["np_bonds_ref"], _G["rcut"]
    bonds = cfg.bonds()
    t0 = time.perf_counter()
    pairs = 0
    with rh._AverageSpy() as spy:
        for i in refs:
            try:
                spy.count = None
                of.calculateAveCosSqForReference(cfg, bonds[int(i)], np_bonds, rcut, excludeSelf=True, sameMolecule=True)
                pairs += spy.count or 0
            except NameError:
                pass
    return pairs, time.perf_counter() - t0


def _literal_frame_worker(seed):
    """One whole frame of config 1 through the reference's CalculateOrientationOrderParameter (:180-227)."""
    from mouse_b200.melt import jitter, make_melt
    from oracle import reference_harness as rh
    melt = make_melt(100, 100, seed=0)
    cfg = rh.config_from_melt(melt, jitter(melt, seed, 0.05) if seed else None)
    of = rh.load().of
    t0 = time.perf_counter()
    with rh._AverageSpy() as spy:
        orig = spy._orig
        pairs = [0]

        def counting(a, *args, **kw):
            pairs[0] += int(a.count())
            return orig(a, *args, **kw)
        np.ma.average = counting
        s = of.CalculateOrientationOrderParameter(cfg, ['1'], 0., 1.5, mode='average', sameMolecule=True)
    return float(s), pairs[0], time.perf_counter() - t0


def _port_frame_worker(seed):
    from mouse_b200.melt import jitter, make_melt
    from oracle import p2_oracle as po
    melt = make_melt(100, 100, seed=0)
    pos = jitter(melt, seed, 0.05) if seed else melt.positions
    t0 = time.perf_counter()
    vec, ctr = po.bond_build(pos, melt.bond_atoms, melt.box)
    r = po.p2_literal(vec, ctr, np.arange(1, melt.n_bonds + 1), melt.mol[melt.bond_atoms[:, 0]].astype(np.float64),
                      melt.box, 1.5, True)
    ok = r["count"] > 0
    return float(1.5 * r["mean"][ok].sum() / ok.sum() - 0.5), int(r["count"].sum()), time.perf_counter() - t0


def _reference_available():
    try:
        from oracle import reference_harness as rh
        return rh.available()
    except Exception:
        return False


def config1_literal_leg(cores, frames=3):
    """SURVEY 8(d) CPU baseline item 1: the reference's CalculateOrientationOrderParameter end to end on >= 3 frames
    of the 9 900-bond melt of config 1 (100 chains x 100 beads, rc = 1.5 sigma); one frame per process (the reference
    is single-threaded).  Literal reference code when oracle/_ref (or /root/reference) is there, else the port."""
    import multiprocessing as mp
    literal = _reference_available()
    k = max(frames, min(cores, 8))
    with mp.get_context("fork").Pool(min(cores, k)) as pool:
        t0 = time.perf_counter()
        out = pool.map(_literal_frame_worker if literal else _port_frame_worker, list(range(k)))
        wall = time.perf_counter() - t0
    return {"kind": "reference" if literal else "port", "frames": k, "processes": min(cores, k),
            "s_per_frame_one_core": float(np.mean([o[2] for o in out])), "pairs_per_frame": int(np.mean([o[1] for o in out])),
            "value": float(sum(o[1] for o in out) / wall), "unit": UNIT, "S_frame0": out[0][0],
            "sample": "calculate_local_ordering.py workload of config 1: %d frames (frame 0 + jittered copies) of 100 x 100 "
                      "beads, 9 900 bonds, rc=1.5, CalculateOrientationOrderParameter (ordering_functions.py:180-227)" % k}


def cpu_baseline(melt):
    """Timed on the host cores (rank 0, N=1): (a) the reference algorithm on a bounded sample of reference bonds of
    config 2 (oracle port of ordering_functions.py:91-119; the literal code needs a 1M-object Config: that is what
    `--impl reference` times); (b) the strong cell-list C oracle on the whole frame; (c) config 1 end to end."""
    import multiprocessing as mp
    from oracle import c_oracle, p2_oracle as po
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    n = melt.n_bonds
    molb = melt.mol[melt.bond_atoms[:, 0]]
    np_bonds = po.np_bonds_array(np.arange(1, n + 1), np.zeros(n), molb.astype(np.float64), vec, ctr)
    cores = os.cpu_count() or 1
    refs = np.linspace(0, n - 1, cores * 4).astype(np.int64)
    blocks = [refs[k::cores] for k in range(cores)]
    _G.update(np_bonds=np_bonds, box=melt.box, vec=vec, ctr=ctr, rcut=WORKLOAD["rcut"])
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_port_worker, [refs[:1]] * cores)          # warm the workers
        t0 = time.perf_counter()
        out = pool.map(_port_worker, blocks)
        dt = time.perf_counter() - t0
    lit = sum(o[0] for o in out) / dt
    t0 = time.perf_counter()
    r = c_oracle.p2(vec, ctr, molb, melt.box, WORKLOAD["rcut"], True)
    dt_c = time.perf_counter() - t0
    return {"value": lit, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d evenly spaced reference bonds against all %d bonds, literal all-pairs NumPy passes "
                      "(ordering_functions.py:91-119)" % (len(refs), n),
            "strong_cell_list_c": {"value": float(r["count"].sum()) / dt_c, "unit": UNIT, "cores": cores,
                                   "sample": "whole frame, OpenMP cell-list C oracle (not the reference's algorithm)"},
            "config1_literal": config1_literal_leg(cores)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, on a bounded sample
    of reference bonds of the SAME workload per step.  When the unmodified reference modules are staged
    (oracle/_ref, see oracle/make_ref.py) the timed code IS the reference: its Config, its
    createNumpyBondsArrayFromConfig and its calculateAveCosSqForReference (kind "reference"); otherwise the
    oracle's literal NumPy port of the same lines (kind "port").  pairs/s = counted pairs of the sample / wall."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gc
    import multiprocessing as mp
    from mouse_b200.melt import make_melt
    from oracle import p2_oracle as po
    melt = make_melt(WORKLOAD["chains"], WORKLOAD["beads"], seed=0)
    n = melt.n_bonds
    cores = os.cpu_count() or 1
    per_core = 4                                  # ~55 ms per reference bond at 1M bonds
    steps, warm = max(1, args.steps), max(0, args.warmup)
    literal = _reference_available() and not os.environ.get("MOUSE_BENCH_PORT")
    if literal:
        from oracle import reference_harness as rh
        of = rh.load().of
        cfg = rh.config_from_melt(melt)                                     # the reference's own object model
        np_bonds_ref = of.createNumpyBondsArrayFromConfig(cfg, allowedBondTypes=['1'])
        _G.update(of=of, cfg=cfg, np_bonds_ref=np_bonds_ref, rcut=WORKLOAD["rcut"])
        worker = _literal_worker
        gc.freeze()                                # keep the forked workers from copying the 1M-object heap
    else:
        vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
        molb = melt.mol[melt.bond_atoms[:, 0]].astype(np.float64)
        np_bonds = po.np_bonds_array(np.arange(1, n + 1), np.zeros(n), molb, vec, ctr)
        _G.update(np_bonds=np_bonds, box=melt.box, vec=vec, ctr=ctr, rcut=WORKLOAD["rcut"])
        worker = _port_worker
    total_pairs, total_t = 0, 0.0
    with mp.get_context("fork").Pool(cores) as pool:
        for s in range(warm + steps):
            rng = np.random.default_rng(100 + s)
            refs = rng.choice(n, size=cores * per_core, replace=False)
            blocks = [refs[k::cores] for k in range(cores)]
            t0 = time.perf_counter()
            out = pool.map(worker, blocks)
            dt = time.perf_counter() - t0
            if s >= warm:
                total_pairs += sum(o[0] for o in out)
                total_t += dt
    value = total_pairs / total_t
    sample = "%d randomly drawn reference bonds per step against all %d bonds: %s" % (
        cores * per_core, n,
        "the reference's own calculateAveCosSqForReference on its own Config / npBonds (oracle/_ref, unmodified)"
        if literal else "literal all-pairs NumPy passes (oracle port of ordering_functions.py:91-119)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": static_config(n),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if literal else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
class P2Bench:
    """Device-resident frames of one topology: LANES frames in flight (own stream, workspace and outputs each)."""

    def __init__(self, dev, positions_list, bond_atoms, molb, box, rcut, same_molecule=True, nbr_mask=None, ref_mask=None,
                 tilt=None, lanes=LANES):
        import torch
        from mouse_b200 import _lib, engine
        self.torch, self._lib, self.engine = torch, _lib, engine
        self.dev = dev
        self.frames_dev = [torch.as_tensor(np.ascontiguousarray(p)).to(dev) for p in positions_list]
        self.bonds = torch.as_tensor(np.ascontiguousarray(bond_atoms)).to(dev)
        self.mol = torch.as_tensor(np.ascontiguousarray(molb).astype(np.int32)).to(dev)
        self.nbr = None if nbr_mask is None else torch.as_tensor(np.ascontiguousarray(nbr_mask).astype(np.uint8)).to(dev)
        self.ref = None if ref_mask is None else torch.as_tensor(np.ascontiguousarray(ref_mask).astype(np.uint8)).to(dev)
        self.n = int(self.bonds.shape[0])
        self.L = _lib.lib()
        self.grid = _lib.grid_plan(box, rcut, self.n, tilt=tilt)
        self.flags = (_lib.MOUSE_SAME_MOLECULE if same_molecule else 0) | int(os.environ.get("MOUSE_BENCH_FLAGS", "0"), 0)
        nbytes = self.L.mouse_workspace_bytes(self.n, C.byref(self.grid))
        n = self.n
        self.lanes = []
        for _ in range(lanes):
            self.lanes.append(dict(ws=torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=dev),
                                   vec=torch.empty((3, n), dtype=torch.float64, device=dev),
                                   ctr=torch.empty((3, n), dtype=torch.float64, device=dev),
                                   ssum=torch.empty(n, dtype=torch.float64, device=dev),
                                   cnt=torch.empty(n, dtype=torch.int32, device=dev),
                                   mean=torch.empty(n, dtype=torch.float64, device=dev),
                                   stream=torch.cuda.Stream(device=dev), done=torch.cuda.Event()))

    def ev(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.torch.cuda.current_stream())        # torch creates the CUDA event lazily at the first record
        return e

    def frame(self, k, tot_row, timed=None, lane=None):
        # one C-ABI call per frame: K1+K2 (fused build), K3 (half-shell sweep + exact fixup), K4 (fused finish + reduce);
        # the two events bracket the K3 launches on the frame's stream
        p = self.engine._ptr
        ln = self.lanes[k % len(self.lanes) if lane is None else lane]
        pos = self.frames_dev[k % len(self.frames_dev)]
        ev0 = C.c_void_p(timed[0].cuda_event) if timed is not None else None
        ev1 = C.c_void_p(timed[1].cuda_event) if timed is not None else None
        self._lib.check(self.L.mouse_p2_frame_timed(p(pos), p(self.bonds), p(self.mol), p(self.nbr), p(self.ref), self.n,
                                                    C.byref(self.grid), self.flags, p(ln["ws"]), ln["ws"].numel(),
                                                    p(ln["vec"]), p(ln["ctr"]), p(ln["ssum"]), p(ln["cnt"]), p(ln["mean"]),
                                                    p(tot_row), ev0, ev1, C.c_void_p(ln["stream"].cuda_stream)),
                        "p2_frame")

    def overlapped(self, steps, tot, first=0, timed=None):
        """Records `steps` frames over the lanes between two events on the current stream; returns the events."""
        torch = self.torch
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ln in self.lanes:
            ln["stream"].wait_event(e0)
        for k in range(steps):
            self.frame(first + k, tot[first + k], None if timed is None else timed[k])
        for ln in self.lanes:
            ln["done"].record(ln["stream"])
            stream.wait_event(ln["done"])
        return e0, e1

    def serial(self, count, tot):
        """`count` frames serialised on lane 0 with the K3 events: (ms per frame, ms per sweep, pairs per frame)."""
        torch = self.torch
        evs = [(self.ev(), self.ev()) for _ in range(count)]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ln0 = self.lanes[0]["stream"]
        ln0.wait_stream(torch.cuda.current_stream())
        s0.record(ln0)
        for k, tev in enumerate(evs):
            self.frame(k, tot[k], tev, lane=0)
        s1.record(ln0)
        torch.cuda.synchronize()
        return (s0.elapsed_time(s1) / count, float(np.mean([a.elapsed_time(b) for a, b in evs])),
                float(tot[:count, 2].sum().item()) / count)


def measure_small_config(dev, positions_list, bond_atoms, molb, box, rcut, steps, warm, **kw):
    """ms per frame (LANES in flight, device resident), serial frame / sweep times and pairs per frame."""
    import torch
    b = P2Bench(dev, positions_list, bond_atoms, molb, box, rcut, **kw)
    tot = torch.zeros((warm + steps, 6), dtype=torch.int64, device=dev)
    for k in range(warm):
        b.frame(k, tot[k])
    torch.cuda.synchronize()
    e0, e1 = b.overlapped(steps, tot, first=warm)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tot_s = torch.zeros((min(steps, 20), 6), dtype=torch.int64, device=dev)
    serial_ms, sweep_ms, pairs = b.serial(min(steps, 20), tot_s)
    t = tot_s[0].cpu()
    s = 1.5 * float(t.view(torch.float64)[0]) / max(int(t[1]), 1) - 0.5
    return dict(ms_per_frame=ms, serial_frame_ms=serial_ms, sweep_ms=sweep_ms, pairs_per_frame=pairs,
                pairs_per_s=pairs / (ms * 1e-3), frames_per_s=1e3 / ms, n_refs=int(t[1]), S=s,
                cell_grid=list(b.grid.ncell_axis), fast=int(b.grid.fast), triclinic=int(b.grid.triclinic))


def measure_dump_ingest(dev, melt, rcut, nframes=48):
    """File ingestion row on config 3's frame size: a LAMMPS dump (scaled coordinates, %.8g - LAMMPS' own default is
    %g) of `nframes` frames is written to the temp directory, then read back (a) by the device-side parser - raw text
    over PCIe, positions born on the GPU -, (b) by the host parser on one and on all usable threads, and (c) end to
    end through local_ordering_from_dump (index + parse + K1-K4 + totals).  Host wall clock: this leg is host-bound."""
    import io
    import tempfile
    import time
    import torch
    from mouse_b200 import ingest
    from mouse_b200.frames import FrameArrays
    from mouse_b200.trajectory import local_ordering_from_dump
    top = FrameArrays.from_melt(melt)
    half = melt.box / 2
    xs = (melt.positions + half) / melt.box
    body = io.BytesIO()
    np.savetxt(body, np.column_stack([np.arange(1, melt.n_atoms + 1), np.ones(melt.n_atoms), xs, melt.images]),
               fmt="%d %d %.8g %.8g %.8g %d %d %d")
    body = body.getvalue()
    path = os.path.join(tempfile.gettempdir(), "mouse_b200_bench_%d.dump" % os.getpid())
    try:
        with open(path, "wb") as f:
            for k in range(nframes):
                f.write(("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (1000 * k, melt.n_atoms)).encode())
                f.write("".join("%.10g %.10g\n" % (-half[a], half[a]) for a in range(3)).encode())
                f.write(b"ITEM: ATOMS id type xs ys zs ix iy iz\n")
                f.write(body)
        t0 = time.perf_counter()
        index = ingest.index_lmp_dump(path)
        index_ms = 1e3 * (time.perf_counter() - t0)
        cores = max(1, min(16, os.cpu_count() or 1))

        def sweep(skip=0, **kw):
            it = ingest.iter_lmp_dump_frames(path, top.atom_num, "scaled", index=index, **kw)
            for _ in range(skip):           # (steady state: behind the one-off pinning / first-frame id lookup)
                next(it)
            t0 = time.perf_counter()
            n = sum(1 for _ in it)
            torch.cuda.synchronize()
            return 1e3 * (time.perf_counter() - t0) / n

        stats = {}
        sweep(device=dev)                                    # warms the pinned buffers and the allocator
        dev_ms = sweep(device=dev, stats=stats)
        dev_steady_ms = sweep(skip=4, device=dev)
        host1_ms = sweep(frames=(0, 3), threads=1)
        hostn_ms = sweep(threads=cores)
        e2e = {}
        for parse_on in ("device", "host"):
            local_ordering_from_dump(top, path, ["1"], rcut, "scaled", device=dev, parse_on=parse_on)
            t0 = time.perf_counter()
            per_frame, tot = local_ordering_from_dump(top, path, ["1"], rcut, "scaled", device=dev, parse_on=parse_on)
            e2e[parse_on] = (1e3 * (time.perf_counter() - t0) / tot.n_frames, per_frame[0][1], tot.n_pairs)
        assert e2e["device"][1:] == e2e["host"][1:]          # same S, same pair count whichever parser fed the frames
        return dict(workload="LAMMPS dump of %d frames x %d atoms (%.1f MB of text per frame), scaled coordinates, image "
                             "flags; host wall clock" % (nframes, melt.n_atoms, len(body) / 1e6),
                    text_mb_per_frame=len(body) / 1e6, index_ms_per_frame=index_ms / nframes,
                    device_parse_ms_per_frame=dev_ms, device_parse_ms_per_frame_steady_state=dev_steady_ms,
                    device_parser_frames=stats,
                    host_parse_ms_per_frame_1_thread=host1_ms, host_parse_ms_per_frame_all_threads=hostn_ms,
                    host_threads=cores, trajectory_ms_per_frame_device_parser=e2e["device"][0],
                    trajectory_ms_per_frame_host_parser=e2e["host"][0], S_frame0=e2e["device"][1],
                    how="device parser: page cache -> pinned buffer (pread threads, next frame prefetched) -> H2D of the "
                        "raw text -> 4 kernels (dump_parse.cu) -> CUDA positions -> FramePipeline; bit-identical to the "
                        "host parser (tests/test_gpu_ingest.py)")
    finally:
        if os.path.exists(path):
            os.remove(path)


def bench_other_configs(dev, world, rank, sms, f_clk, dist):
    """Driver-timed sub-measurements of BASELINE configs 3, 4 and 5 (inputs resident in HBM, CUDA events)."""
    import torch
    from mouse_b200 import _lib, engine
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.workloads import CONFIG3, CONFIG4, config5
    out = {}
    peak_issue = sms * 128 * f_clk * 1e6
    # ---- config 3: frames of the 250k-bond melt, every rank its own frames (weak scaling over the frame shards)
    melt = make_melt(CONFIG3["chains"], CONFIG3["beads"], seed=0)
    frames = [jitter(melt, rank * 8 + f, 0.02) for f in range(8)]
    r = measure_small_config(dev, frames, melt.bond_atoms, melt.mol[melt.bond_atoms[:, 0]], melt.box, CONFIG3["rcut"],
                             steps=200, warm=20, lanes=LANES_SMALL)
    t = torch.tensor([r["ms_per_frame"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out["config3"] = dict(workload="250k-bond melt frames (2500 x 101 beads), rc=2.5, frames sharded over %d rank(s), %d in "
                                   "flight per GPU" % (world, LANES_SMALL), bonds_per_frame=melt.n_bonds, ms_per_frame=ms,
                          frames_per_s=world * 1e3 / ms, pairs_per_s=world * r["pairs_per_frame"] / (ms * 1e-3),
                          serial_frame_ms=r["serial_frame_ms"], sweep_ms=r["sweep_ms"], pairs_per_frame=r["pairs_per_frame"],
                          trajectory_10k_frames_s=10000 * ms * 1e-3 / world, cell_grid=r["cell_grid"], S=r["S"],
                          roofline_frac_sweep=r["pairs_per_frame"] / (r["sweep_ms"] * 1e-3) * W_LANE_INSTR_PER_PAIR / peak_issue)
    if world > 1:
        out["note"] = "configs 4 and 5 are single-frame workloads (replicas only): measured at N=1"
        return out
    del frames
    try:        # (a host-side leg with a 0.6 GB temporary file: never the reason for a bench run to fail)
        import shutil
        import tempfile
        free = shutil.disk_usage(tempfile.gettempdir()).free
        if free > (4 << 30):
            out["config3_from_dump"] = measure_dump_ingest(dev, melt, CONFIG3["rcut"], nframes=48)
        elif free > (1 << 30):
            out["config3_from_dump"] = measure_dump_ingest(dev, melt, CONFIG3["rcut"], nframes=12)
        else:
            out["config3_from_dump"] = {"skipped": "less than 1 GB free in the temp directory"}
    except Exception as e:      # noqa: BLE001
        out["config3_from_dump"] = {"error": "%s: %s" % (type(e).__name__, e)}
    # ---- config 5: 8M bonds, rc = 3 sigma, bond-type neighbour set, residue-type references, other-molecule pairs
    for key, tri in (("config5_orthorhombic", False), ("config5_triclinic", True)):
        cfg = config5(tri)
        r = measure_small_config(dev, [cfg["positions"]], cfg["bond_atoms"], cfg["molb"], cfg["box"], cfg["rcut"],
                                 steps=6, warm=2, same_molecule=False, nbr_mask=cfg["nbr_mask"], ref_mask=cfg["ref_mask"],
                                 tilt=cfg["tilt"], lanes=1)
        out[key] = dict(workload="8M bonds (80000 x 101 beads), rc=3.0, neighbour set = bond type 1 (4M bonds), references = "
                                 "residue types R0/R1, other-molecule pairs only%s" % (
                                     "; triclinic cell (build-defined minimum image, parity unpinned)" if tri else ""),
                        bonds=len(cfg["bond_atoms"]), tilt=cfg["tilt"], ms_per_frame=r["ms_per_frame"],
                        sweep_ms=r["sweep_ms"], pairs_per_frame=r["pairs_per_frame"], pairs_per_s=r["pairs_per_s"],
                        n_refs=r["n_refs"], cell_grid=r["cell_grid"], fast=r["fast"], S=r["S"],
                        roofline_frac_sweep=r["pairs_per_frame"] / (r["sweep_ms"] * 1e-3) * W_LANE_INSTR_PER_PAIR / peak_issue)
        del cfg
        torch.cuda.empty_cache()
    # ---- config 4: RDF pair histogram of 4M beads, 0-5 sigma, 1000 bins (K5)
    melt = make_melt(CONFIG4["chains"], CONFIG4["beads"], seed=0)
    L = _lib.lib()
    n = melt.n_atoms
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=CONFIG4["nbin"], range=(CONFIG4["rmin"],
                                                                                               CONFIG4["rmax"])))
    grid = _lib.grid_plan(melt.box, float(edges[-1]), n)
    ws = torch.empty(L.mouse_workspace_bytes(n, C.byref(grid)) + 4096, dtype=torch.uint8, device=dev)
    pos = torch.as_tensor(melt.positions).to(dev)
    typ = torch.zeros(n, dtype=torch.int32, device=dev)
    mol = torch.as_tensor(melt.mol.astype(np.int32)).to(dev)
    hist = torch.empty((1, CONFIG4["nbin"]), dtype=torch.int64, device=dev)
    p, st = engine._ptr, engine._stream()
    reps = 5
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps + 1)]
    for row in ev:
        for e in row:
            e.record()
    for k in range(reps + 1):
        ev[k][0].record()
        _lib.check(L.mouse_rdf_frame_timed(p(pos), p(typ), p(mol), n, 1, C.byref(grid), _lib.MOUSE_SAME_MOLECULE,
                                           edges.ctypes.data_as(C.POINTER(C.c_double)), CONFIG4["nbin"], p(ws), ws.numel(),
                                           p(hist), C.c_void_p(ev[k][1].cuda_event), C.c_void_p(ev[k][2].cuda_event), st),
                   "mouse_rdf_frame")
        ev[k][3].record()
        torch.cuda.synchronize()
    frame_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][3]) for k in range(1, reps + 1)]))
    sweep_ms = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(1, reps + 1)]))
    pairs = int(hist.sum().item())
    ach = pairs / (sweep_ms * 1e-3) * W5_LANE_INSTR_PER_PAIR
    out["config4"] = dict(workload="RDF pair histogram, 4 000 000 beads (40000 x 100), 0-5 sigma, 1000 bins, one type, "
                                   "sameMolecule=True", atoms=n, ms_per_frame=frame_ms, sweep_ms=sweep_ms,
                          pairs_per_frame=pairs, pairs_per_s=pairs / (frame_ms * 1e-3), cell_grid=list(grid.ncell_axis),
                          roofline={"bound": "fp_issue", "kernel": "rdf_sweep_half_kernel", "achieved": ach / 1e12,
                                    "peak": peak_issue / 1e12, "unit": "Tlane-instr/s", "frac": ach / peak_issue,
                                    "traffic": None,
                                    "how": "ordered pairs in range / K5 launch time x 60.2 lane-instr per pair (6.45 "
                                           "candidates x 7 FP32 filter + 15 per counted pair: FP64 d^2, bin, count) over "
                                           "SMs x 128 lanes x f_clk"})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mouse_b200 import engine
    from mouse_b200.melt import jitter, make_melt
    from mouse_b200.trajectory import FramePipeline, TrajectoryTotals, all_reduce_totals

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    engine.require_cuda()
    steps, warm = max(1, args.steps), max(3, args.warmup)
    melt = make_melt(WORKLOAD["chains"], WORKLOAD["beads"], seed=0)
    # the CPU-baseline leg forks worker processes: run it before this process creates a CUDA context
    cpu_line = cpu_baseline(melt) if (world == 1 and not args.no_cpu_baseline) else None
    torch.cuda.set_device(local_rank)
    numa_bound = engine.bind_host_to_gpu(local_rank) if (world > 1 and not os.environ.get("MOUSE_BENCH_NOBIND")) else False
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, n_atoms, rcut = melt.n_bonds, melt.n_atoms, WORKLOAD["rcut"]
    molb = melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)
    # frames of the trajectory: base melt + per-frame jitter; every rank owns different frames
    frames_np = [np.ascontiguousarray(jitter(melt, rank * N_FRAME_VARIANTS + f, 0.02)) for f in range(N_FRAME_VARIANTS)]
    frames_host = [torch.from_numpy(f).pin_memory() for f in frames_np]
    # LANES frames in flight, each on its own stream with its own workspace and per-bond outputs (frames are
    # independent): the HBM-side kernels and the tail of the persistent sweep of one frame overlap the sweep of
    # the next.  Lane 0 alone (serialised frames) gives the per-kernel durations of the roofline object.
    b = P2Bench(dev, frames_np, melt.bond_atoms, molb, melt.box, rcut)
    grid = b.grid
    tot = torch.zeros((warm + steps, 6), dtype=torch.int64, device=dev)
    tot_serial = torch.zeros((100, 6), dtype=torch.int64, device=dev)
    sweep_ev = [(b.ev(), b.ev()) for _ in range(steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region (value)
    sampler = ClockSampler(local_rank, period=float(os.environ.get("MOUSE_BENCH_CLOCK_PERIOD", "0.01")))
    if rank == 0 and not os.environ.get("MOUSE_BENCH_NOSAMPLER"):
        sampler.start()          # NVML start-up perturbs the GPU for a few ms: start well before timing
        time.sleep(1.0)
    # warm-up: at least `warm` frames AND at least 0.3 s of back-to-back frames, so that the timed region measures the
    # steady state of a trajectory run (clocks / power state settled), not the first milliseconds after idle
    t_spin = time.perf_counter() + 0.3
    k = 0
    while k < warm or time.perf_counter() < t_spin:
        b.frame(k % warm, tot[k % warm])
        k += 1
        if k % 16 == 0:
            torch.cuda.synchronize()
    warm_done = k
    # the tail of the timed region (sum of the per-frame totals + the path's only collective) is warmed up too:
    # the first call of a torch kernel / NCCL collective pays one-off module loading and communicator set-up
    packed = torch.zeros(6, dtype=torch.float64, device=dev)

    def pack_totals(rows):
        # ONE buffer for the path's only collective: the f64 sum of the per-frame sums and the integer counts as
        # float64 (exact below 2^53, trajectory.TrajectoryTotals.pack)
        tsum = rows.sum(dim=0)
        packed[0] = rows[:, 0].view(torch.float64).sum()
        packed[1:] = tsum[1:].to(torch.float64)
        if world > 1:
            dist.all_reduce(packed)

    pack_totals(tot[:warm])      # exactly the ops of the timed region's tail, run once before it
    barrier()
    n_before = len(sampler.samples)
    stream = torch.cuda.current_stream()
    e0, e1 = b.overlapped(steps, tot, first=warm, timed=sweep_ev)
    pack_totals(tot[warm:])      # the path's only collective: ONE all-reduce of the packed sums and counts
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    sweep_ms_overlapped = float(np.mean([a.elapsed_time(bb) for a, bb in sweep_ev]))
    # per-kernel durations: the same frames once more, serialised on one lane (in the overlapped region above the
    # kernels of up to LANES frames share the SMs, so an in-stream duration there is not a kernel time)
    n_serial = min(steps, 100)
    serial_ms, sweep_ms, pairs_serial = b.serial(n_serial, tot_serial)
    if os.environ.get("MOUSE_BENCH_DIAG"):
        sys.stderr.write("overlapped frame %.4f ms; serial frame %.4f ms, sweep %.4f ms (in-stream, overlapped: %.4f ms)\n" % (
            ms / steps, serial_ms, sweep_ms, sweep_ms_overlapped))
    pairs_local = int(tot[warm:, 2].sum().item())
    pairs_all = int(round(float(packed[2].item())))
    value = pairs_all / (ms * 1e-3)

    # ---- end-to-end through the host-buffer API (pinned host positions in, totals out, every step)
    pipe = FramePipeline(melt.bond_atoms, molb, melt.box, rcut, n_atoms, same_molecule=True, device=dev,
                         depth=int(os.environ.get("MOUSE_PIPE_DEPTH", "3")))
    for k in range(warm):
        pipe.submit(frames_host[k % N_FRAME_VARIANTS])
    pipe.drain()
    pipe.host_submit_s, pipe.frames_submitted = 0.0, 0
    barrier()
    t0 = time.perf_counter()
    done = []
    for k in range(steps):
        done += pipe.submit(frames_host[(warm + k) % N_FRAME_VARIANTS])
    done += pipe.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    host_submit_us = 1e6 * pipe.host_submit_s / max(pipe.frames_submitted, 1)
    graph_nodes = pipe.L.mouse_p2_frame_graph_nodes(pipe.slots[0]["graph"]) if pipe.slots[0].get("graph") else 0
    e2e_pairs = sum(d["n_pairs"] for d in done)
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    e2e_p = torch.tensor([e2e_pairs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_p)
    # ---- the ceiling of that path on this box: the same pinned buffers copied to the same device buffers, no kernels,
    # on every rank at once (what the host side of the box delivers to N GPUs streaming 24 MB frames)
    barrier()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(pipe.copy_stream):
        c0.record(pipe.copy_stream)
        for k in range(steps):
            pipe.slots[k % pipe.depth]["pos"].copy_(frames_host[k % N_FRAME_VARIANTS], non_blocking=True)
        c1.record(pipe.copy_stream)
    barrier()
    copy_t = torch.tensor([c0.elapsed_time(c1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(copy_t, op=dist.ReduceOp.MAX)
    h2d_ceiling = world * steps * pipe.h2d_bytes_per_frame / float(copy_t.item()) / 1e9
    pipe.close()
    if rank == 0 and len(sampler.samples) - n_before < 3:
        # very short runs: keep the kernels busy a little longer so that the clocks are sampled under load
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            b.frame(0, tot[0])
        torch.cuda.synchronize()
    if rank == 0:
        sampler.samples = sampler.samples[n_before:]
    clocks = sampler.stop() if rank == 0 else None
    totals = TrajectoryTotals()
    for d in done:
        totals.add_frame(d["sum_mean"], d["n_refs"], d["n_pairs"])
    totals = all_reduce_totals(totals, dev)

    props = torch.cuda.get_device_properties(dev)
    sms = props.multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    f_clk = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    if world > 1:      # every rank needs the same clock figure for the sub-measurements' fractions
        fc = torch.tensor([f_clk if rank == 0 else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(fc)
        f_clk = float(fc.item())
    configs = None
    if not args.no_configs:
        del b
        torch.cuda.empty_cache()
        configs = bench_other_configs(dev, world, rank, sms, f_clk, dist)

    if rank == 0:
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the sweep kernel (ncu --set full capture),
            # valid only for the kernel sources it was measured on
            prof = json.load(open(os.path.join(ROOT, "profiles", "r2_sweep_half.json")))
            if prof.get("kernel_source_sha16") == kernel_source_hash():
                traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
                traffic_src = prof["source"]
            else:
                traffic_src = "profiles/r2_sweep_half.json is stale (kernel sources changed since the capture)"
        except (OSError, KeyError, ValueError):
            pass
        peak_issue = sms * 128 * f_clk * 1e6                       # lane-instructions / s
        sweep_pairs_s = pairs_serial / (sweep_ms * 1e-3)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        build_ms = serial_ms - sweep_ms
        e2e_ms = 1e3 * float(e2e_t.item()) / steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (packed FP32 cutoff filter with exact FP64 re-check, FP64 cos^2, 2^-47 fixed-point sums)",
            "data": "synthetic", "config": static_config(n),
            "details": {"warmup_frames_actual": warm_done, "steps_timed": steps, "pairs_per_frame": pairs_local // steps,
                        "cell_grid": list(grid.ncell_axis), "frames_per_s": world * steps / (ms * 1e-3),
                        "parallelism": "frames sharded, %d rank(s), ONE NCCL all-reduce of the packed sums and counts; %d "
                                       "frames in flight per GPU (own stream / workspace each)" % (world, LANES),
                        "S_mean": totals.mean_s},
            "e2e": {"value": float(e2e_p.item()) / float(e2e_t.item()), "unit": UNIT,
                    "h2d_bytes_per_step": pipe.h2d_bytes_per_frame, "d2h_bytes_per_step": pipe.d2h_bytes_per_frame,
                    "ms_per_step": e2e_ms, "host_submit_us": host_submit_us, "graph_nodes_per_frame": graph_nodes,
                    "h2d_achieved_gbs": world * pipe.h2d_bytes_per_frame / (e2e_ms * 1e-3) / 1e9,
                    "h2d_ceiling_gbs": h2d_ceiling,
                    "h2d_ceiling_how": "the same pinned frame buffers copied to the same device buffers on every rank "
                                       "at once, no kernels (CUDA events, max over ranks)",
                    "note": "FramePipeline.submit(): pinned host positions -> ONE H2D copy -> K1-K4 replayed from a "
                            "captured CUDA graph -> 48-byte totals -> D2H, %d frames in flight (copy of frame k+1 and its "
                            "HBM-side kernels overlap the sweep of frame k); topology resident%s" % (
                                pipe.depth, "; host process bound to the GPU's CPU set" if numa_bound else "")},
            "gpu_launches": steps * KERNELS_PER_FRAME + (1 if world > 1 else 0),
            "clocks": clocks,
            "roofline": {
                "bound": "fp_issue", "kernel": "p2_sweep_half_kernel (+ p2_fixup_kernel)",
                "achieved": sweep_pairs_s * W_LANE_INSTR_PER_PAIR / 1e12,
                "peak": peak_issue / 1e12, "unit": "Tlane-instr/s",
                "frac": sweep_pairs_s * W_LANE_INSTR_PER_PAIR / peak_issue,
                "traffic": traffic, "traffic_source": traffic_src, "kernel_source_sha16": kernel_source_hash(),
                "how": "SURVEY 8(d): ordered pairs/s x 55.2 lane-instr/pair (full-shell bound; the kernel is half-shell: "
                       "every unordered pair is decided once and counted for both partners) over SMs x 128 lanes x f_clk "
                       "(%d SMs, %.0f MHz %s); sweep %.3f ms/launch: CUDA events around the K3 launches of %d frames "
                       "serialised on one stream right after the timed region (in the timed region up to %d frames "
                       "share the SMs; the in-stream duration there is %.3f ms)"
                       % (sms, f_clk, "sampled under load" if clocks and clocks.get("sm_mhz") else "max clock", sweep_ms,
                          n_serial, LANES, sweep_ms_overlapped),
                "sweep_ms": sweep_ms, "sweep_ms_in_timed_region": sweep_ms_overlapped, "serial_frame_ms": serial_ms,
                "sweep_pairs_per_s": sweep_pairs_s,
                "hbm": {"bound": "hbm", "stage": "K1+K2+K4 build/reduce stages", "achieved": BYTES_PER_BOND * n /
                        (max(build_ms, 1e-6) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": BYTES_PER_BOND * n / (max(build_ms, 1e-6) * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "ms": build_ms, "algorithmic_bytes": BYTES_PER_BOND * n}},
        }
        if configs is not None:
            line["configs"] = configs
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-measurements of configs 3, 4 and 5")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU (the driver does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        sys.exit(subprocess.call(cmd + sys.argv[1:]))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
