"""Trajectory driver: frames are independent units (the reference loops over frame files with no
carried state, ``calculate_local_ordering.py:39-76``), so they are block-sharded over the ranks of
one node - one process per GPU - and the only communication is ONE final all-reduce of a packed
buffer of order-parameter sums and counts.  ``torch.distributed`` is the
plumbing (NCCL on GPUs, gloo in the CPU tests of the sharding / reduction logic)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, engine


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous block of frames of ``rank``: [r*F/G, (r+1)*F/G)."""
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


@dataclass
class TrajectoryTotals:
    """What the final all-reduce carries: the order-parameter sums, the counts and the histograms of a trajectory
    (reference semantics being reduced: ``ordering_functions.py:203-206, 221-227`` for S and the P2 histogram,
    ``:162-176`` for the RDF counts)."""
    sum_s: float = 0.0        # sum over frames of S_frame
    n_frames: int = 0         # frames that produced an S
    sum_mean: float = 0.0     # sum over frames of sum_i m_i
    n_refs: int = 0           # sum over frames of i_s
    n_pairs: int = 0          # ordered in-cutoff pairs counted
    n_skipped_frames: int = 0  # frames without any reference that has neighbours
    n_hist_over: int = 0      # P2 values the reference's updateHistogram would raise IndexError on
    p2_hist: np.ndarray | None = None    # int64 [sbins]: per-vector P2 histogram counts summed over frames
    rdf_hist: np.ndarray | None = None   # int64 [npairs, nbin]: RDF pair counts summed over frames

    # Everything rides in ONE float64 buffer: [sum_s, sum_mean | n_frames, n_refs, n_pairs, n_skipped, n_hist_over |
    # p2_hist... | rdf_hist...].  The integer fields are stored as float64, which adds integers exactly (and in any
    # order) as long as every partial sum stays below 2^53 - checked when packing.
    _N_SCALARS = 7

    def pack(self, device):
        ints = [self.n_frames, self.n_refs, self.n_pairs, self.n_skipped_frames, self.n_hist_over]
        parts = [np.array([self.sum_s, self.sum_mean], np.float64), np.array(ints, np.float64)]
        for h in (self.p2_hist, self.rdf_hist):
            if h is not None:
                parts.append(np.asarray(h, np.int64).ravel().astype(np.float64))
        buf = np.concatenate(parts)
        if np.abs(buf[2:]).max() >= 2.0**53 / 64:      # every field keeps head-room for up to 64 ranks
            raise OverflowError("TrajectoryTotals: integer totals too large for an exact float64 all-reduce")
        return torch.from_numpy(buf).to(device)

    def unpack_into(self, buf):
        b = buf.cpu().numpy()
        self.sum_s, self.sum_mean = float(b[0]), float(b[1])
        (self.n_frames, self.n_refs, self.n_pairs, self.n_skipped_frames,
         self.n_hist_over) = (int(round(v)) for v in b[2:self._N_SCALARS])
        off = self._N_SCALARS
        if self.p2_hist is not None:
            k = self.p2_hist.size
            self.p2_hist = np.rint(b[off:off + k]).astype(np.int64).reshape(self.p2_hist.shape)
            off += k
        if self.rdf_hist is not None:
            k = self.rdf_hist.size
            self.rdf_hist = np.rint(b[off:off + k]).astype(np.int64).reshape(self.rdf_hist.shape)
        return self

    def add_frame(self, sum_mean, n_refs, n_pairs, hist=None, n_hist_over=0):
        if n_refs > 0:
            self.sum_s += 1.5 * sum_mean / n_refs - 0.5
            self.n_frames += 1
        else:
            self.n_skipped_frames += 1
        self.sum_mean += sum_mean
        self.n_refs += int(n_refs)
        self.n_pairs += int(n_pairs)
        self.n_hist_over += int(n_hist_over)
        if hist is not None:
            h = np.asarray(hist, np.int64)
            self.p2_hist = h.copy() if self.p2_hist is None else self.p2_hist + h

    def add_rdf_frame(self, hist):
        h = np.asarray(hist, np.int64)
        self.rdf_hist = h.copy() if self.rdf_hist is None else self.rdf_hist + h
        self.n_frames += 1

    @property
    def mean_s(self):
        return self.sum_s / self.n_frames if self.n_frames else float("nan")

    def p2_histogram(self, smin, smax, sbins):
        """The reference's histogram dict (``ordering_functions.py:187-193, 225-227``) of the whole trajectory:
        counts summed over frames (and ranks), then ``normHistogram``; raises ``IndexError`` if any frame did."""
        from .lammpack_misc import createHistogram, normHistogram
        if self.n_hist_over > 0:
            raise IndexError("list index out of range")
        h = createHistogram(smin, smax, sbins)
        for i in range(sbins):
            sleft, sright = smin + i * h["step"], smin + (i + 1) * h["step"]
            h["norm"][i] = 1. / ((2. / 3. * (sleft + 0.5))**0.5 + (2. / 3. * (sright + 0.5))**0.5)
        h["values"] = [float(v) for v in (self.p2_hist if self.p2_hist is not None else np.zeros(sbins))]
        normHistogram(h)
        return h


def _dist_state():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def all_reduce_totals(t: TrajectoryTotals, device) -> TrajectoryTotals:
    """The single collective of the path: ONE sum-all-reduce of the packed totals (sums, counts and histogram
    bins in one float64 buffer; the integer fields reduce exactly).  Every rank must carry histograms of the same
    shape (or none)."""
    dist, _, world = _dist_state()
    if world == 1:
        return t
    buf = t.pack(device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return t.unpack_into(buf)


def gather_per_frame(per_frame, n_frames, device):
    """The reference prints one S per frame in input order (``calculate_local_ordering.py:39-76``): every rank's
    (frame index, S) block is all-gathered (F x 8 bytes) and put back in frame order.  Returns a float64 array of
    ``n_frames`` values (NaN = the frame produced no S) on every rank."""
    dist, rank, world = _dist_state()
    out = np.full(n_frames, np.nan)
    if world == 1:
        for f, s in per_frame:
            out[f] = s
        return out
    width = max(((r + 1) * n_frames) // world - (r * n_frames) // world for r in range(world))
    mine = torch.full((max(width, 1),), float("nan"), dtype=torch.float64, device=device)
    lo, _ = shard_range(n_frames, rank, world)
    for f, s in per_frame:
        mine[f - lo] = float(s)
    blocks = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blocks, mine)
    for r, blk in enumerate(blocks):
        a, b = shard_range(n_frames, r, world)
        out[a:b] = blk.cpu().numpy()[:b - a]
    return out


class FramePipeline:
    """Host-buffer entry point of the hot path for one topology: ``submit(positions_host)`` copies a
    frame's positions from pinned host memory, runs K1-K4 and reads back the 48-byte totals.
    ``depth`` frames are in flight, each on its own compute stream with its own workspace and per-bond
    outputs: the copy of frame k+1 overlaps the kernels of frame k, and the HBM-side kernels / the tail of
    the persistent sweep of one frame overlap the sweep of the next (frames are independent,
    ``calculate_local_ordering.py:39-76``).  Per-bond results of a frame live in ``slot["sum"/"cnt"/"mean"]``
    until the slot is reused."""

    def __init__(self, bond_atoms, mol, box, rcut, n_atoms, same_molecule=True, nbr_mask=None, ref_mask=None,
                 device=None, depth=3, use_graph=True, tilt=None, hist_args=None):
        engine.require_cuda()
        self.device = torch.device(device or "cuda")
        self.n_atoms = int(n_atoms)
        self.bonds = engine.as_dev(bond_atoms, torch.int32, self.device)
        self.n = int(self.bonds.shape[0])
        self.mol = engine.as_dev(mol, torch.int32, self.device) if mol is not None else \
            torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.nbr = engine.as_dev(nbr_mask, torch.uint8, self.device)
        self.ref = engine.as_dev(ref_mask, torch.uint8, self.device)
        self.box = [float(b) for b in box]
        self.rcut = float(rcut)
        self.tilt = None if tilt is None else [float(t) for t in tilt]
        self.grid = _lib.grid_plan(self.box, self.rcut, self.n, tilt=self.tilt)
        self.flags = (_lib.MOUSE_SAME_MOLECULE if same_molecule else 0) | (_lib.MOUSE_NO_CUTOFF if rcut < 0 else 0)
        L = _lib.lib()
        self.L = L
        nbytes = L.mouse_workspace_bytes(self.n, C.byref(self.grid))
        self.depth = depth
        self.use_graph = bool(use_graph)     # frames with the pipeline's own box / cutoff replay a captured CUDA graph
        self.hist_args = None if hist_args is None else (float(hist_args[0]), float(hist_args[1]), int(hist_args[2]))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                pos=torch.empty((self.n_atoms, 3), dtype=torch.float64, device=self.device),
                totals=torch.zeros(6, dtype=torch.int64, device=self.device),
                totals_host=torch.zeros(6, dtype=torch.int64).pin_memory(),
                copied=torch.cuda.Event(), done=torch.cuda.Event(), busy=False,
                stream=torch.cuda.Stream(device=self.device),
                ws=torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=self.device),
                vec=torch.empty((3, self.n), dtype=torch.float64, device=self.device),
                ctr=torch.empty((3, self.n), dtype=torch.float64, device=self.device),
                sum=torch.empty(self.n, dtype=torch.float64, device=self.device),
                cnt=torch.empty(self.n, dtype=torch.int32, device=self.device),
                mean=torch.empty(self.n, dtype=torch.float64, device=self.device)))
            assert self.slots[-1]["ws"].data_ptr() % 256 == 0
            if self.hist_args is not None:       # per-vector P2 histogram of the frame (mode='histo', :187-193, 203)
                self.slots[-1]["hist"] = torch.zeros(self.hist_args[2], dtype=torch.int64, device=self.device)
                self.slots[-1]["hist_host"] = torch.zeros(self.hist_args[2], dtype=torch.int64).pin_memory()
        self._next = 0
        self._pending = []
        self.h2d_bytes_per_frame = self.n_atoms * 3 * 8
        self.d2h_bytes_per_frame = 48
        self.kernels_per_frame = 8   # bond_assign, scan, scatter_idx, canon, sweep, fixup, exact check, finish+reduce
        self.host_submit_s = 0.0     # wall time spent inside submit() (enqueue cost of the host side)
        self.frames_submitted = 0

    def submit(self, positions_host, tag=None, box=None, rcut=None):
        """Enqueue one frame.  ``positions_host``: (N_a,3) f64 - a pinned CPU tensor is copied asynchronously as it
        is; a NumPy array / pageable tensor is staged through the slot's own pinned buffer; a CUDA tensor is
        copied on the device.  Returns the finished results (list of totals dicts, submission order, each with
        the ``tag`` it was submitted with) that had to be collected to free a slot.  ``box`` / ``rcut``: this
        frame's own box and cutoff (NPT dumps): the cell grid is re-planned for the frame, the slot's workspace
        grows if the new grid needs it."""
        import time
        t_enter = time.perf_counter()
        out = []
        s = self.slots[self._next]
        if s["busy"]:
            out.append(self._collect(s))
            t_enter = time.perf_counter()      # waiting for the GPU is not enqueue cost
        s["tag"] = tag
        grid, flags = self.grid, self.flags
        own_grid = box is not None or rcut is not None
        if own_grid:
            rc = self.rcut if rcut is None else float(rcut)
            grid = _lib.grid_plan(self.box if box is None else [float(b) for b in box], rc, self.n, tilt=self.tilt)
            flags = (self.flags & ~_lib.MOUSE_NO_CUTOFF) | (_lib.MOUSE_NO_CUTOFF if rc < 0 else 0)
            need = self.L.mouse_workspace_bytes(self.n, C.byref(grid))
            if s["ws"].numel() < need:
                s["done"].synchronize()
                if s.get("graph") is not None:      # the captured graph names the old buffer
                    self.L.mouse_p2_frame_graph_destroy(s["graph"])
                    s["graph"] = None
                s["ws"] = torch.empty(int(need * 1.1) + 4096, dtype=torch.uint8, device=self.device)
                assert s["ws"].data_ptr() % 256 == 0
        s["grid"] = grid     # keeps the ctypes struct alive until the call below has returned
        p = engine._ptr
        src = positions_host
        if not isinstance(src, torch.Tensor):
            src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64))
        if src.device.type == "cpu" and not src.is_pinned():
            if "stage" not in s:
                s["stage"] = torch.empty((self.n_atoms, 3), dtype=torch.float64).pin_memory()
            s["stage"].copy_(src.reshape(self.n_atoms, 3))
            src = s["stage"]
        with torch.cuda.stream(self.copy_stream):
            if src.device.type == "cuda":       # (the allocator must not hand the block out again before the copy ran)
                src.record_stream(self.copy_stream)
            s["pos"].copy_(src.reshape(self.n_atoms, 3), non_blocking=True)
            s["copied"].record(self.copy_stream)
        with torch.cuda.stream(s["stream"]):
            s["stream"].wait_event(s["copied"])
            st = C.c_void_p(s["stream"].cuda_stream)
            args = (p(s["pos"]), p(self.bonds), p(self.mol), p(self.nbr), p(self.ref), self.n, C.byref(grid), flags,
                    p(s["ws"]), s["ws"].numel(), p(s["vec"]), p(s["ctr"]), p(s["sum"]), p(s["cnt"]), p(s["mean"]),
                    p(s["totals"]))
            if self.use_graph and not own_grid and self.n > 0:
                if s.get("graph") is None:      # recorded once per slot: the slot's buffers never change
                    h = C.c_void_p()
                    _lib.check(self.L.mouse_p2_frame_graph_create(*args, C.byref(h)), "mouse_p2_frame_graph_create")
                    s["graph"] = h
                _lib.check(self.L.mouse_p2_frame_graph_launch(s["graph"], st), "mouse_p2_frame_graph_launch")
            else:
                _lib.check(self.L.mouse_p2_frame(*args, st), "mouse_p2_frame")
            if self.hist_args is not None:
                smin, smax, sbins = self.hist_args
                _lib.check(self.L.mouse_p2_reduce(p(s["sum"]), p(s["cnt"]), self.n, C.byref(grid), p(s["ws"]),
                                                  s["ws"].numel(), p(s["mean"]), p(s["totals"]), p(s["hist"]), smin, smax,
                                                  sbins, st), "mouse_p2_reduce")
                s["hist_host"].copy_(s["hist"], non_blocking=True)
            s["totals_host"].copy_(s["totals"], non_blocking=True)
            s["done"].record(s["stream"])
        s["busy"] = True
        self._pending.append(s)
        self._next = (self._next + 1) % self.depth
        self.host_submit_s += time.perf_counter() - t_enter
        self.frames_submitted += 1
        return out

    def _collect(self, s):
        s["done"].synchronize()
        s["busy"] = False
        if s in self._pending:
            self._pending.remove(s)
        t = s["totals_host"]
        out = dict(sum_mean=float(t.view(torch.float64)[0]), n_refs=int(t[1]), n_pairs=int(t[2]),
                   n_zero_len=int(t[3]), n_hist_over=int(t[4]), tag=s.get("tag"))
        if self.hist_args is not None:
            out["hist"] = s["hist_host"].numpy().copy()
        return out

    def drain(self):
        return [self._collect(s) for s in list(self._pending)]

    def close(self):
        """Waits for the frames in flight and frees the captured graphs."""
        self.drain()
        for s in self.slots:
            if s.get("graph") is not None:
                self.L.mouse_p2_frame_graph_destroy(s["graph"])
                s["graph"] = None

    def __del__(self):
        try:
            for s in self.slots:
                if s.get("graph") is not None:
                    self.L.mouse_p2_frame_graph_destroy(s["graph"])
                    s["graph"] = None
        except Exception:
            pass


def _frame_s(t):
    return np.float64(1.5 * t["sum_mean"] / t["n_refs"] - 0.5) if t["n_refs"] else np.float64("nan")


def local_ordering_trajectory(frames_positions, bond_atoms, mol, box, rcut, same_molecule=True, nbr_mask=None,
                              ref_mask=None, device=None, hist_args=None, tilt=None, gather=False):
    """Per-frame S of this rank's shard plus the trajectory totals reduced over all ranks.
    ``frames_positions``: sequence of (N_a,3) float64 arrays / tensors (the WHOLE trajectory, every
    rank indexes its own block).  ``hist_args`` = (smin, smax, sbins) also accumulates the per-vector P2 histogram of
    every frame (``mode='histo'``), reduced with the sums in the same all-reduce.  Returns (list of (frame_index, S)
    for this rank, TrajectoryTotals); with ``gather=True`` the first item is instead the float64 array of ALL frames'
    S in frame order on every rank (the order the reference's CLI prints them in)."""
    dist, rank, world = _dist_state()
    lo, hi = shard_range(len(frames_positions), rank, world)
    device = torch.device(device or "cuda")
    totals = TrajectoryTotals()
    if hist_args is not None:
        totals.p2_hist = np.zeros(int(hist_args[2]), np.int64)
    per_frame = []
    if hi > lo:
        # fixed topology and box: the frames go through the pipeline (several frames in flight, no per-frame sync)
        n_atoms = int(np.asarray(frames_positions[lo].shape)[0]) if hasattr(frames_positions[lo], "shape") \
            else len(frames_positions[lo])
        pipe = FramePipeline(bond_atoms, mol, box, rcut, n_atoms, same_molecule=same_molecule, nbr_mask=nbr_mask,
                             ref_mask=ref_mask, device=device, hist_args=hist_args, tilt=tilt)
        done = []
        for f in range(lo, hi):
            done += pipe.submit(frames_positions[f], tag=f)
        done += pipe.drain()
        pipe.close()
        for t in done:
            totals.add_frame(t["sum_mean"], t["n_refs"], t["n_pairs"], t.get("hist"), t.get("n_hist_over", 0))
            per_frame.append((t["tag"], _frame_s(t)))
    totals = all_reduce_totals(totals, device)
    if gather:
        return gather_per_frame(per_frame, len(frames_positions), device), totals
    return per_frame, totals


def pair_correlation_trajectory(frames_positions, type_code, mol, ntypes, box, rmin, rmax, nbin, same_molecule=True,
                                device=None, tilt=None):
    """RDF pair histograms of a trajectory (``CalculatePairCorrelationFunctions`` per frame,
    ``ordering_functions.py:122-177``): the frames are block-sharded over the ranks like
    ``local_ordering_trajectory``, each rank sums the int64 ``[npairs, nbin]`` counts of its frames on the device, and
    the single all-reduce adds them over the ranks (integers: exact, identical to a one-rank run bit for bit).
    Returns (hist int64 [npairs, nbin] of ALL frames, edges, TrajectoryTotals with ``n_frames``)."""
    dist, rank, world = _dist_state()
    lo, hi = shard_range(len(frames_positions), rank, world)
    device = torch.device(device or "cuda")
    npairs = int(ntypes) * (int(ntypes) + 1) // 2
    acc = torch.zeros((npairs, int(nbin)), dtype=torch.int64, device=device)
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=int(nbin), range=(rmin, rmax)), np.float64)
    totals = TrajectoryTotals()
    tc = engine.as_dev(type_code, torch.int32, device)
    mc = engine.as_dev(mol, torch.int32, device) if mol is not None else None
    for f in range(lo, hi):
        h, _ = engine.rdf_frame(frames_positions[f], tc, mc, ntypes, box, rmin, rmax, nbin,
                                same_molecule=same_molecule, device=device, tilt=tilt)
        acc += h
        totals.n_frames += 1
    totals.rdf_hist = acc.cpu().numpy()
    totals = all_reduce_totals(totals, device)
    return totals.rdf_hist, edges, totals


def pair_correlation_from_dump(topology, dump_path, rmin, rmax, nbin, atomtypes=None, coords_type="scaled",
                               same_molecule=True, device=None, parse_on="device", parse_threads=None, stats=None):
    """RDF pair histograms summed over the frames of a LAMMPS dump (``CalculatePairCorrelationFunctions`` per frame,
    ``ordering_functions.py:122-177``, fed by ``return_extra_frames_from_lmp_dump``, ``lammpack_types.py:344-398``).
    ``topology``: ``FrameArrays`` of the data file; ``atomtypes``: the atom types taken into account (None / empty:
    all, ``:134-139``) - they get the type codes 0.. in sorted order, and the histogram rows are the reference's
    ``"ti_tj"`` pairs in that order.  Every frame carries its own box: the grid is planned per frame for the last bin
    edge.  Frames are block-sharded over the ranks by the dump's byte-offset index, parsed on the rank's GPU
    (``parse_on="device"``) or by the host parser, histograms are summed on the device and all-reduced once.
    Returns (hist int64 [npairs, nbin], edges, type names, TrajectoryTotals with ``n_frames``)."""
    import os
    from . import ingest
    dist, rank, world = _dist_state()
    device = torch.device(device or "cuda")
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    index = ingest.index_lmp_dump(dump_path)
    lo, hi = shard_range(len(index) - 1, rank, world)
    types = np.asarray([str(t) for t in topology.atom_type], dtype=object)
    wanted = sorted(set(types.tolist()) if not atomtypes else set(str(t) for t in atomtypes) & set(types.tolist()))
    code_of = {t: k for k, t in enumerate(wanted)}
    keep = np.array([t in code_of for t in types.tolist()], bool)
    rows = np.nonzero(keep)[0]
    ntypes = max(len(wanted), 1)
    tcode = torch.as_tensor(np.array([code_of[t] for t in types[rows].tolist()], np.int32)).to(device)
    mcode = torch.as_tensor(np.ascontiguousarray(topology.atom_mol_code()[rows], np.int32)).to(device)
    rows_dev = None if keep.all() else torch.as_tensor(rows).to(device)
    npairs = ntypes * (ntypes + 1) // 2
    acc = torch.zeros((npairs, int(nbin)), dtype=torch.int64, device=device)
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=int(nbin), range=(rmin, rmax)), np.float64)
    totals = TrajectoryTotals()
    if parse_threads is None:
        parse_threads = max(1, min(16, (os.cpu_count() or 1) // max(world, 1)))
    if hi > lo and len(rows):
        for _, box, _, pos, _ in ingest.iter_lmp_dump_frames(dump_path, topology.atom_num, coords_type, frames=(lo, hi),
                                                             index=index, threads=parse_threads,
                                                             device=device if parse_on == "device" else None,
                                                             stats=stats):
            pos = engine.as_dev(pos, torch.float64, device)
            if rows_dev is not None:
                pos = pos.index_select(0, rows_dev)
            h, _ = engine.rdf_frame(pos, tcode, mcode, ntypes, box, rmin, rmax, nbin, same_molecule=same_molecule,
                                    device=device)
            acc += h
            totals.n_frames += 1
    totals.rdf_hist = acc.cpu().numpy()
    totals = all_reduce_totals(totals, device)
    return totals.rdf_hist, edges, wanted, totals


def local_ordering_from_dump(topology, dump_path, bondtypes=None, rcut=0., coords_type="scaled",
                             reference_res_types=None, same_molecule=True, device=None, gather=False,
                             parse_threads=None, hist_args=None, parse_on="device", stats=None):
    """Trajectory driver over files: ``topology`` is a ``FrameArrays`` (e.g.
    ``ingest.read_lmp_data_arrays(data_file)``), the frames come from a LAMMPS dump.  Every frame carries its own
    box, so the cell grid is re-planned per frame; frames are block-sharded over the ranks exactly like
    ``local_ordering_trajectory`` and reduced with the same single all-reduce.  Selections follow the reference:
    neighbour set = bonds passing ``bondtypes``, reference set = those also passing ``reference_res_types``
    (``ordering_functions.py:196-199``); ``rcut == 0`` means half the box diagonal (``:185-186``).

    Ingestion keeps up with the GPU: the dump is indexed once (byte offset of every ``ITEM: TIMESTEP``), every rank
    seeks to its own block of frames and parses nothing else, a producer thread parses frame k+1 (multi-threaded C
    number parser, GIL released) while frame k is on the GPU, and frames are handed to the pipeline one at a time -
    the host never holds more than a few frames whatever the length of the trajectory.  ``parse_on="device"`` (default)
    ships the raw text of the ATOMS block to the GPU and parses it there (``ingest.DeviceDumpParser``: same values bit
    for bit; frames with tokens outside the number parser's exact fast path fall to the host parser one by one);
    ``"host"`` keeps the parsing on ``parse_threads`` host threads.  ``stats`` (dict) receives the frame counts."""
    import os
    import queue
    import threading
    from . import ingest
    dist, rank, world = _dist_state()
    device = torch.device(device or "cuda")
    if device.type == "cuda" and device.index is None:      # resolved HERE: the producer thread has its own current device
        device = torch.device("cuda", torch.cuda.current_device())
    index = ingest.index_lmp_dump(dump_path)
    n_frames = len(index) - 1
    lo, hi = shard_range(n_frames, rank, world)
    nbr_mask = topology.bond_mask(bond_types=bondtypes)
    ref_mask = topology.bond_mask(bond_types=bondtypes, res_types=reference_res_types)
    mol = topology.bond_mol_code()
    if parse_threads is None:
        parse_threads = max(1, min(16, (os.cpu_count() or 1) // max(world, 1)))
    totals = TrajectoryTotals()
    if hist_args is not None:
        totals.p2_hist = np.zeros(int(hist_args[2]), np.int64)
    per_frame = []
    if hi > lo:
        q = queue.Queue(maxsize=3)

        def produce():
            try:
                for item in ingest.iter_lmp_dump_frames(dump_path, topology.atom_num, coords_type, frames=(lo, hi),
                                                        index=index, threads=parse_threads,
                                                        device=device if parse_on == "device" else None, stats=stats):
                    q.put(item)
                q.put(None)
            except BaseException as e:      # hand the error to the consumer
                q.put(e)

        threading.Thread(target=produce, daemon=True).start()
        pipe, done, f = None, [], lo
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            timestep, box, _, pos, _ = item
            rc = float(rcut) if rcut != 0. else engine.half_diagonal(box)
            if pipe is None:
                pipe = FramePipeline(topology.bond_atoms, mol, box, rc, pos.shape[0], same_molecule=same_molecule,
                                     nbr_mask=nbr_mask, ref_mask=ref_mask, device=device, hist_args=hist_args)
            same_grid = np.array_equal(box, pipe.box) and rc == pipe.rcut      # NVT dumps replay the captured graph
            done += pipe.submit(pos, tag=(f, int(timestep)), box=None if same_grid else box,
                                rcut=None if same_grid else rc)
            f += 1
        if pipe is not None:
            done += pipe.drain()
            pipe.close()
        for t in done:
            totals.add_frame(t["sum_mean"], t["n_refs"], t["n_pairs"], t.get("hist"), t.get("n_hist_over", 0))
            per_frame.append((t["tag"], _frame_s(t)))
    totals = all_reduce_totals(totals, device)
    if gather:
        return gather_per_frame([(tag[0], s) for tag, s in per_frame], n_frames, device), totals
    return [(tag[1], s) for tag, s in per_frame], totals
