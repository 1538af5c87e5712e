#!/usr/bin/env python3
"""Benchmark of the local-P2 hot path (BASELINE.json metric: ordered in-cutoff pairs/s for local P2
ordering, 1M-bond melt, rc = 2.5 sigma, at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W          this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference ...                   the reference's own CPU algorithm, timed here

A "step" is one pass of the hot path (K1 bond build -> K2 cell list -> K3 pair sweep -> K4 reduction)
over one trajectory frame of the workload on every rank (weak scaling: frames are sharded, every GPU
processes its own frames, one NCCL all-reduce of the packed sums at the end of the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
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
    # the frame-sharded trajectory configuration (not a bench line of the driver; `--workload config3`)
    "config3": dict(name="config3: frames of a 250k-bond melt (2500 chains x 101 beads), rc=2.5 sigma, frames sharded "
                         "over the ranks", chains=2500, beads=101, rcut=2.5),
}
WORKLOAD = WORKLOADS["config2"]
METRIC = "pairs/sec for local P2 ordering (1M bonds, rc=2.5σ)"
UNIT = "pairs/s"
W_LANE_INSTR_PER_PAIR = 55.2      # 6.45 x 7 FP32 + 10 FP64 lane-instructions per counted pair
BYTES_PER_BOND = 72.0             # compulsory HBM bytes per bond per frame
N_FRAME_VARIANTS = 4              # distinct frames cycled (working set per frame ~210 MB > 126 MB L2)
LANES = 3                         # frames in flight in the device-resident timed region (own stream + workspace each)


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


# ------------------------------------------------------------------------------------------- reference arm
_G = {}   # arrays shared with forked workers (set before the pool is created)


def _ref_worker(refs):
    """Literal NumPy restatement of ordering_functions.py:91-119 for a block of reference bonds."""
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


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (all-pairs NumPy passes per reference bond,
    ordering_functions.py:91-119, 196-214) through the oracle port, all host threads, on a bounded
    sample of reference bonds of the SAME workload; pairs/s = counted pairs of the sample / wall time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from mouse_b200.melt import make_melt
    from oracle import p2_oracle as po
    melt = make_melt(WORKLOAD["chains"], WORKLOAD["beads"], seed=0)
    vec, ctr = po.bond_build(melt.positions, melt.bond_atoms, melt.box)
    n = melt.n_bonds
    molb = melt.mol[melt.bond_atoms[:, 0]].astype(np.float64)
    np_bonds = po.np_bonds_array(np.arange(1, n + 1), np.zeros(n), molb, vec, ctr)
    cores = os.cpu_count() or 1
    per_core = 4                                  # ~55 ms per reference bond at 1M bonds
    steps, warm = max(1, args.steps), max(0, args.warmup)
    total_pairs, total_t = 0, 0.0
    _G.update(np_bonds=np_bonds, box=melt.box, vec=vec, ctr=ctr, rcut=WORKLOAD["rcut"])
    with mp.get_context("fork").Pool(cores) as pool:
        for s in range(warm + steps):
            rng = np.random.default_rng(100 + s)
            refs = rng.choice(n, size=cores * per_core, replace=False)
            blocks = [refs[k::cores] for k in range(cores)]
            t0 = time.perf_counter()
            out = pool.map(_ref_worker, blocks)
            dt = time.perf_counter() - t0
            if s >= warm:
                total_pairs += sum(o[0] for o in out)
                total_t += dt
    value = total_pairs / total_t
    sample = "%d evenly drawn reference bonds per step against all %d bonds (literal all-pairs NumPy passes)" % (
        cores * per_core, n)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- CPU baseline leg
def cpu_baseline(melt):
    """Oracle timed on the host cores (rank 0, N=1): (a) the literal reference algorithm on a bounded
    sample of reference bonds; (b) the strong cell-list C oracle on the whole frame."""
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
        pool.map(_ref_worker, [refs[:1]] * cores)          # warm the workers
        t0 = time.perf_counter()
        out = pool.map(_ref_worker, blocks)
        dt = time.perf_counter() - t0
    lit = sum(o[0] for o in out) / dt
    t0 = time.perf_counter()
    r = c_oracle.p2(vec, ctr, molb, melt.box, WORKLOAD["rcut"], True)
    dt_c = time.perf_counter() - t0
    return {"value": lit, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d evenly spaced reference bonds against all %d bonds, literal all-pairs NumPy passes "
                      "(ordering_functions.py:91-119)" % (len(refs), n),
            "strong_cell_list_c": {"value": float(r["count"].sum()) / dt_c, "unit": UNIT, "cores": cores,
                                   "sample": "whole frame, OpenMP cell-list C oracle (not the reference's algorithm)"}}


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mouse_b200 import _lib, engine
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
    # frames of the trajectory: base melt + per-frame jitter; every rank owns different frames
    frames_host = [torch.from_numpy(np.ascontiguousarray(jitter(melt, rank * N_FRAME_VARIANTS + f, 0.02))).pin_memory()
                   for f in range(N_FRAME_VARIANTS)]
    frames_dev = [f.to(dev) for f in frames_host]
    bonds = torch.as_tensor(melt.bond_atoms).to(dev)
    mol = torch.as_tensor(melt.mol[melt.bond_atoms[:, 0]].astype(np.int32)).to(dev)
    L = _lib.lib()
    grid = _lib.grid_plan(melt.box, rcut, n)
    # LANES frames in flight, each on its own stream with its own workspace and per-bond outputs (frames are
    # independent): the HBM-side kernels and the tail of the persistent sweep of one frame overlap the sweep of
    # the next.  Lane 0 alone (serialised frames) gives the per-kernel durations of the roofline object.
    nbytes = L.mouse_workspace_bytes(n, C.byref(grid))
    lanes = []
    for _ in range(LANES):
        lanes.append(dict(ws=torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=dev),
                          vec=torch.empty((3, n), dtype=torch.float64, device=dev),
                          ctr=torch.empty((3, n), dtype=torch.float64, device=dev),
                          ssum=torch.empty(n, dtype=torch.float64, device=dev),
                          cnt=torch.empty(n, dtype=torch.int32, device=dev),
                          mean=torch.empty(n, dtype=torch.float64, device=dev),
                          stream=torch.cuda.Stream(device=dev), done=torch.cuda.Event()))
    tot = torch.zeros((warm + steps, 6), dtype=torch.int64, device=dev)
    tot_serial = torch.zeros((100, 6), dtype=torch.int64, device=dev)
    p, flags = engine._ptr, _lib.MOUSE_SAME_MOLECULE
    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    sweep_ev = [(ev(), ev()) for _ in range(steps)]
    serial_ev = [(ev(), ev()) for _ in range(min(steps, 100))]

    for e0_, e1_ in sweep_ev + serial_ev:   # create the CUDA events (torch creates them lazily at the first record)
        e0_.record(stream)
        e1_.record(stream)

    def frame(k, timed=None, lane=None, out=None):
        # one C-ABI call per frame: K1+K2 (fused build), K3 (half-shell sweep + exact fixup), K4 (fused finish + reduce);
        # the two events bracket the K3 launches on the frame's stream
        ln = lanes[k % LANES if lane is None else lane]
        pos = frames_dev[k % N_FRAME_VARIANTS]
        ev0 = C.c_void_p(timed[0].cuda_event) if timed is not None else None
        ev1 = C.c_void_p(timed[1].cuda_event) if timed is not None else None
        _lib.check(L.mouse_p2_frame_timed(p(pos), p(bonds), p(mol), None, None, n, C.byref(grid), flags, p(ln["ws"]),
                                          ln["ws"].numel(), p(ln["vec"]), p(ln["ctr"]), p(ln["ssum"]), p(ln["cnt"]),
                                          p(ln["mean"]), p(tot[k] if out is None else out[k]), ev0, ev1,
                                          C.c_void_p(ln["stream"].cuda_stream)),
                   "p2_frame")

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
        frame(k % warm)
        k += 1
        if k % 16 == 0:
            torch.cuda.synchronize()
    warm_done = k
    # the tail of the timed region (sum of the per-frame totals + the path's only collective) is warmed up too:
    # the first call of a torch kernel / NCCL collective pays one-off module loading and communicator set-up
    _w = tot[:warm].sum(dim=0)
    if world > 1:
        _wf = tot[:warm, 0].view(torch.float64).sum().reshape(1)
        dist.all_reduce(_wf)
        dist.all_reduce(_w)
    barrier()
    n_before = len(sampler.samples)
    e0, e1 = ev(), ev()
    e0.record(stream)
    for ln in lanes:
        ln["stream"].wait_event(e0)
    for k in range(steps):
        frame(warm + k, sweep_ev[k])
    for ln in lanes:
        ln["done"].record(ln["stream"])
        stream.wait_event(ln["done"])
    totals = TrajectoryTotals()
    # the path's only collective: all-reduce of the packed order-parameter sums (device tensors)
    tsum = tot[warm:].sum(dim=0)
    if world > 1:
        fsum = tot[warm:, 0].view(torch.float64).sum().reshape(1)
        dist.all_reduce(fsum)
        dist.all_reduce(tsum)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    sweep_ms_overlapped = float(np.mean([a.elapsed_time(b) for a, b in sweep_ev]))
    # per-kernel durations: the same frames once more, serialised on one lane (in the overlapped region above the
    # kernels of up to LANES frames share the SMs, so an in-stream duration there is not a kernel time)
    s0, s1 = ev(), ev()
    ln0 = lanes[0]["stream"]
    s0.record(ln0)
    for k, tev in enumerate(serial_ev):
        frame(k, tev, lane=0, out=tot_serial)
    s1.record(ln0)
    torch.cuda.synchronize()
    serial_ms = s0.elapsed_time(s1) / len(serial_ev)
    sweep_ms = float(np.mean([a.elapsed_time(b) for a, b in serial_ev]))
    if os.environ.get("MOUSE_BENCH_DIAG"):
        sys.stderr.write("overlapped frame %.4f ms; serial frame %.4f ms, sweep %.4f ms (in-stream, overlapped: %.4f ms)\n" % (
            ms / steps, serial_ms, sweep_ms, sweep_ms_overlapped))
    tl = tot[warm:].cpu()
    pairs_local = int(tl[:, 2].sum())
    pairs_all = int(tsum[2].item()) if world > 1 else pairs_local
    value = pairs_all / (ms * 1e-3)

    # ---- end-to-end through the host-buffer API (pinned host positions in, totals out, every step)
    pipe = FramePipeline(melt.bond_atoms, melt.mol[melt.bond_atoms[:, 0]].astype(np.int32), melt.box, rcut, n_atoms,
                         same_molecule=True, device=dev)
    for k in range(warm):
        pipe.submit(frames_host[k % N_FRAME_VARIANTS])
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    done = []
    for k in range(steps):
        done += pipe.submit(frames_host[(warm + k) % N_FRAME_VARIANTS])
    done += pipe.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_pairs = sum(d["n_pairs"] for d in done)
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    e2e_p = torch.tensor([e2e_pairs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_p)
    if rank == 0 and len(sampler.samples) - n_before < 3:
        # very short runs: keep the kernels busy a little longer so that the clocks are sampled under load
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            frame(0)
        torch.cuda.synchronize()
    if rank == 0:
        sampler.samples = sampler.samples[n_before:]
    clocks = sampler.stop() if rank == 0 else None
    for d in done:
        totals.add_frame(d["sum_mean"], d["n_refs"], d["n_pairs"])
    totals = all_reduce_totals(totals, dev)

    if rank == 0:
        props = torch.cuda.get_device_properties(dev)
        sms = props.multi_processor_count
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        f_clk = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the sweep kernel (ncu --set full capture)
            prof = json.load(open(os.path.join(ROOT, "profiles", "r1_sweep_half_v6.json")))
            traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
            traffic_src = prof["source"]
        except (OSError, KeyError, ValueError):
            pass
        peak_issue = sms * 128 * f_clk * 1e6                       # lane-instructions / s
        pairs_serial = float(tot_serial[:len(serial_ev), 2].sum().item()) / len(serial_ev)
        sweep_pairs_s = pairs_serial / (sweep_ms * 1e-3)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        build_ms = serial_ms - sweep_ms
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm_done,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (packed FP32 cutoff filter with exact FP64 re-check, FP64 cos^2, 2^-47 fixed-point sums)", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "bonds_per_frame": n, "frames_per_step_per_gpu": 1,
                       "pairs_per_frame": pairs_local // steps, "cell_grid": list(grid.ncell_axis),
                       "l2": "~%d MB touched per frame (L2 is 126 MB), %d distinct frames cycled"
                             % (int(230e-6 * n), N_FRAME_VARIANTS),
                       "frames_per_s": world * steps / (ms * 1e-3),
                       "parallelism": "frames sharded, %d rank(s), one NCCL all-reduce of packed sums; %d frames in "
                                      "flight per GPU (own stream / workspace each)" % (world, LANES),
                       "S_mean": totals.mean_s},
            "e2e": {"value": float(e2e_p.item()) / float(e2e_t.item()), "unit": UNIT,
                    "h2d_bytes_per_step": pipe.h2d_bytes_per_frame, "d2h_bytes_per_step": pipe.d2h_bytes_per_frame,
                    "ms_per_step": 1e3 * float(e2e_t.item()) / steps,
                    "note": "FramePipeline.submit(): pinned host positions -> H2D -> K1-K4 -> 48-byte totals -> D2H, "
                            "%d frames in flight (copy of frame k+1 and its HBM-side kernels overlap the sweep of frame k); "
                            "topology resident%s" % (pipe.depth, "; host process bound to the GPU's CPU set" if numa_bound else "")},
            "gpu_launches": steps * 9 + (2 if world > 1 else 0),
            "clocks": clocks,
            "roofline": {
                "bound": "fp_issue", "kernel": "p2_sweep_half_kernel (+ p2_fixup_kernel)",
                "achieved": sweep_pairs_s * W_LANE_INSTR_PER_PAIR / 1e12,
                "peak": peak_issue / 1e12, "unit": "Tlane-instr/s",
                "frac": sweep_pairs_s * W_LANE_INSTR_PER_PAIR / peak_issue,
                "traffic": traffic, "traffic_source": traffic_src,
                "how": "SURVEY 8(d): ordered pairs/s x 55.2 lane-instr/pair (full-shell bound; the kernel is half-shell: "
                       "every unordered pair is decided once and counted for both partners) over SMs x 128 lanes x f_clk "
                       "(%d SMs, %.0f MHz %s); sweep %.3f ms/launch: CUDA events around the K3 launches of %d frames "
                       "serialised on one stream right after the timed region (in the timed region up to %d frames "
                       "share the SMs; the in-stream duration there is %.3f ms)"
                       % (sms, f_clk, "sampled under load" if clocks and clocks.get("sm_mhz") else "max clock", sweep_ms,
                          len(serial_ev), LANES, sweep_ms_overlapped),
                "sweep_ms": sweep_ms, "sweep_ms_in_timed_region": sweep_ms_overlapped, "serial_frame_ms": serial_ms,
                "sweep_pairs_per_s": sweep_pairs_s,
                "hbm": {"bound": "hbm", "stage": "K1+K2+K4 build/reduce stages", "achieved": BYTES_PER_BOND * n /
                        (max(build_ms, 1e-6) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": BYTES_PER_BOND * n / (max(build_ms, 1e-6) * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "ms": build_ms, "algorithmic_bytes": BYTES_PER_BOND * n}},
        }
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
