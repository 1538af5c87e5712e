"""Device-side driver of the hot path: owns workspaces and output tensors (PyTorch = device
memory + streams only) and calls the C-ABI of ``libmouse_b200.so``.  No CPU fallback: every
function here needs a CUDA device and raises otherwise."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mouse_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


_ws_cache: dict = {}
_ws_generation: dict = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """256-byte aligned scratch buffer, grown on demand and reused per device.  Every hand-out bumps the
    device's generation counter: a result that still needs the cell list it was computed on (``p2_pair_lists``)
    remembers the generation and refuses to run once another call has reused the buffer."""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
        assert buf.data_ptr() % 256 == 0
        _ws_cache[key] = buf
    _ws_generation[key] = _ws_generation.get(key, 0) + 1
    return buf


def workspace_generation(device) -> int:
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    return _ws_generation.get(key, 0)


def bind_host_to_gpu(device_index: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root), so that the
    pinned staging buffers allocated afterwards are first-touched next to the device.  Matters when several
    ranks stream frames over PCIe at the same time; returns False (and changes nothing) when NVML or
    sched_setaffinity is not available or the mask would be empty."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        have = os.sched_getaffinity(0)
        cpus &= have
        if not cpus or cpus == have:     # no topology information (e.g. a VM that reports every CPU for every GPU)
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def as_dev(x, dtype, device):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device)


@dataclass
class P2Result:
    """Per-frame outputs (all device tensors, original bond order)."""
    vec: torch.Tensor      # (3, N_b) f64  bond vectors    = npBonds rows 3-5
    ctr: torch.Tensor      # (3, N_b) f64  bond midpoints  = npBonds rows 6-8
    sum: torch.Tensor      # (N_b,) f64    sum of cos^2 over counted neighbours
    count: torch.Tensor    # (N_b,) i32    counted neighbours (0 = skipped / not a reference)
    mean: torch.Tensor     # (N_b,) f64    mean cos^2 (NaN when count == 0)
    totals: torch.Tensor   # (6,) i64 view of mouse_p2_totals_t
    grid: _lib.Grid
    hist: torch.Tensor | None = None
    ws: torch.Tensor | None = None   # the workspace holding this frame's cell list ...
    ws_generation: int = -1          # ... and the shared buffer's generation when it was built (-1: private buffer)

    def totals_dict(self):
        t = self.totals.cpu()
        return dict(sum_mean=float(t.view(torch.float64)[0]), n_refs=int(t[1]), n_pairs=int(t[2]),
                    n_zero_len=int(t[3]), n_hist_over=int(t[4]), n_band=int(t[5]))

    def order_parameter(self):
        """S = 1.5 * sum_i m_i / i_s - 0.5 (``ordering_functions.py:221-224``)."""
        t = self.totals_dict()
        if t["n_refs"] == 0:
            raise UnboundLocalError("cannot access local variable 's' where it is not associated with a value")
        return np.float64(1.5 * t["sum_mean"] / t["n_refs"] - 0.5)


def half_diagonal(box) -> float:
    """``config.box().length() / 2.`` (``ordering_functions.py:185-186``, ``vector3d.py:60-61``)."""
    x, y, z = float(box[0]), float(box[1]), float(box[2])
    return math.sqrt(x * x + y * y + z * z) / 2.


def p2_frame(positions, bond_atoms, mol, box, rcut, same_molecule=True, nbr_mask=None, ref_mask=None,
             exact=False, hist_args=None, device=None, out: P2Result | None = None, slots=0, warps=0,
             private_workspace=False, force_fast=False, tilt=None) -> P2Result:
    """K1 -> K4 for one frame.  ``positions`` (N_a,3) f64, ``bond_atoms`` (N_b,2) i32 (atom1 ->
    atom2 direction, 0-based), ``mol`` (N_b,) i32 molecule code of atom1 of each bond,
    masks (N_b,) uint8/bool or None.  ``rcut < 0`` means no cutoff (reference ``:102-103``).
    ``tilt`` = (xy, xz, yz): triclinic box (build-defined semantics, see ``include/mouse_b200.h``).
    ``slots`` / ``warps`` / ``force_fast``: tuning fields of the C-ABI flags (tests); ``private_workspace``: the result owns its
    workspace (needed when ``p2_pair_lists`` is not called right away)."""
    require_cuda()
    device = torch.device(device or "cuda")
    pos = as_dev(positions, torch.float64, device)
    bonds = as_dev(bond_atoms, torch.int32, device)
    n = int(bonds.shape[0])
    molc = as_dev(mol, torch.int32, device) if mol is not None else torch.zeros(n, dtype=torch.int32, device=device)
    nbr = as_dev(nbr_mask, torch.uint8, device)
    ref = as_dev(ref_mask, torch.uint8, device)
    grid = _lib.grid_plan(box, float(rcut), n, tilt=tilt)
    L = _lib.lib()
    nbytes = L.mouse_workspace_bytes(n, C.byref(grid))
    if private_workspace:
        ws, gen = torch.empty(int(nbytes) + 4096, dtype=torch.uint8, device=device), -1
    else:
        ws = workspace(nbytes, device)
        gen = workspace_generation(device)
    if out is None or out.sum.shape[0] != n:
        out = P2Result(vec=torch.empty((3, n), dtype=torch.float64, device=device),
                       ctr=torch.empty((3, n), dtype=torch.float64, device=device),
                       sum=torch.empty(n, dtype=torch.float64, device=device),
                       count=torch.empty(n, dtype=torch.int32, device=device),
                       mean=torch.empty(n, dtype=torch.float64, device=device),
                       totals=torch.zeros(6, dtype=torch.int64, device=device), grid=grid)
    out.grid, out.ws, out.ws_generation = grid, ws, gen
    flags = (_lib.MOUSE_SAME_MOLECULE if same_molecule else 0) | (_lib.MOUSE_FORCE_EXACT if exact else 0) | \
        (_lib.MOUSE_NO_CUTOFF if rcut < 0 else 0) | _lib.MOUSE_SLOTS(slots) | _lib.MOUSE_WARPS(warps) | \
        (_lib.MOUSE_FORCE_FAST if force_fast else 0)
    st = _stream()
    if hist_args is None:
        _lib.check(L.mouse_p2_frame(_ptr(pos), _ptr(bonds), _ptr(molc), _ptr(nbr), _ptr(ref), n, C.byref(grid),
                                    flags, _ptr(ws), ws.numel(), _ptr(out.vec), _ptr(out.ctr), _ptr(out.sum),
                                    _ptr(out.count), _ptr(out.mean), _ptr(out.totals), st), "mouse_p2_frame")
    else:
        smin, smax, sbins = hist_args
        out.hist = torch.zeros(int(sbins), dtype=torch.int64, device=device)
        box3 = (C.c_double * 3)(*[float(b) for b in box])
        if tilt is None:
            _lib.check(L.mouse_bond_build(_ptr(pos), _ptr(bonds), n, box3, _ptr(out.vec), _ptr(out.ctr), st),
                       "mouse_bond_build")
        else:
            tilt3 = (C.c_double * 3)(*[float(t) for t in tilt])
            _lib.check(L.mouse_bond_build_triclinic(_ptr(pos), _ptr(bonds), n, box3, tilt3, _ptr(out.vec),
                                                    _ptr(out.ctr), st), "mouse_bond_build_triclinic")
        _lib.check(L.mouse_cell_build(_ptr(out.vec), _ptr(out.ctr), _ptr(molc), _ptr(nbr), _ptr(ref), n,
                                      C.byref(grid), _ptr(ws), ws.numel(), st), "mouse_cell_build")
        _lib.check(L.mouse_p2_sweep(_ptr(out.vec), _ptr(out.ctr), n, C.byref(grid), flags, _ptr(ws), ws.numel(),
                                    _ptr(out.sum), _ptr(out.count), st), "mouse_p2_sweep")
        _lib.check(L.mouse_p2_reduce(_ptr(out.sum), _ptr(out.count), n, C.byref(grid), _ptr(ws), ws.numel(),
                                     _ptr(out.mean), _ptr(out.totals), _ptr(out.hist), float(smin), float(smax),
                                     int(sbins), st),
                   "mouse_p2_reduce")
    return out


def p2_pair_lists(res: P2Result, same_molecule=True, exact=False, rcut=None):
    """Counted neighbour lists of the frame ``res`` was computed on (same ``same_molecule`` / ``rcut`` arguments).
    The cell list lives in ``res.ws``: either a private buffer (``p2_frame(private_workspace=True)``) or the shared
    per-device one, in which case this must be the next engine call on the device - anything in between reuses the
    buffer and this function raises instead of reading a foreign cell list.  Returns (row_ptr int64 (N_b+1,),
    nbr_idx int32) with rows sorted ascending (sorting is plumbing done with torch)."""
    require_cuda()
    n = int(res.sum.shape[0])
    device = res.sum.device
    total = int(res.count.sum().item()) if res.totals_dict()["n_zero_len"] == 0 else None
    if total is None:
        raise ValueError("pair lists are undefined when zero-length bonds poison the frame")
    L = _lib.lib()
    ws = res.ws
    if ws is None or (res.ws_generation >= 0 and res.ws_generation != workspace_generation(device)):
        raise RuntimeError("p2_pair_lists: the shared workspace was reused after this frame was computed; "
                           "call p2_frame(..., private_workspace=True) to keep the frame's cell list")
    row_ptr = torch.empty(n + 1, dtype=torch.int64, device=device)
    nbr = torch.empty(max(total, 1), dtype=torch.int32, device=device)
    flags = (_lib.MOUSE_SAME_MOLECULE if same_molecule else 0) | (_lib.MOUSE_FORCE_EXACT if exact else 0) | \
        (_lib.MOUSE_NO_CUTOFF if (rcut is not None and rcut < 0) else 0)
    _lib.check(L.mouse_p2_pair_lists(_ptr(res.vec), _ptr(res.ctr), n, C.byref(res.grid), flags, _ptr(ws), ws.numel(),
                                     _ptr(res.count), _ptr(row_ptr), _ptr(nbr), max(total, 1), _stream()),
               "mouse_p2_pair_lists")
    nbr = nbr[:total]
    rows = torch.repeat_interleave(torch.arange(n, device=device), res.count.long())
    key = rows * (2 ** 31) + nbr.long()
    nbr = (torch.sort(key).values % (2 ** 31)).int()
    return row_ptr, nbr


def bond_build(positions, bond_atoms, box, device=None, tilt=None):
    """K1 alone: (vec, ctr) as (3, N_b) f64 device tensors."""
    require_cuda()
    device = torch.device(device or "cuda")
    pos = as_dev(positions, torch.float64, device)
    bonds = as_dev(bond_atoms, torch.int32, device)
    n = int(bonds.shape[0])
    vec = torch.empty((3, n), dtype=torch.float64, device=device)
    ctr = torch.empty((3, n), dtype=torch.float64, device=device)
    box3 = (C.c_double * 3)(*[float(b) for b in box])
    if tilt is None:
        _lib.check(_lib.lib().mouse_bond_build(_ptr(pos), _ptr(bonds), n, box3, _ptr(vec), _ptr(ctr), _stream()),
                   "mouse_bond_build")
    else:
        tilt3 = (C.c_double * 3)(*[float(t) for t in tilt])
        _lib.check(_lib.lib().mouse_bond_build_triclinic(_ptr(pos), _ptr(bonds), n, box3, tilt3, _ptr(vec), _ptr(ctr),
                                                         _stream()), "mouse_bond_build_triclinic")
    return vec, ctr


def rdf_frame(positions, type_code, mol, ntypes, box, rmin, rmax, nbin, same_molecule=True, device=None, exact=False,
              tilt=None):
    """K5: int64 histogram [ntypes*(ntypes+1)/2, nbin] over ordered pairs, NumPy bin edges.  ``exact`` forces the
    all-FP64 one-thread-per-atom kernel (the default picks the half-shell sweep whenever the grid allows it)."""
    require_cuda()
    device = torch.device(device or "cuda")
    pos = as_dev(positions, torch.float64, device)
    n = int(pos.shape[0])
    tcode = as_dev(type_code, torch.int32, device)
    molc = as_dev(mol, torch.int32, device) if mol is not None else torch.zeros(n, dtype=torch.int32, device=device)
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=int(nbin), range=(rmin, rmax)),
                                 dtype=np.float64)
    rcell = float(edges[-1])
    grid = _lib.grid_plan(box, rcell, n, tilt=tilt)
    L = _lib.lib()
    ws = workspace(L.mouse_workspace_bytes(n, C.byref(grid)), device)
    npairs = ntypes * (ntypes + 1) // 2
    hist = torch.empty((npairs, int(nbin)), dtype=torch.int64, device=device)
    flags = (_lib.MOUSE_SAME_MOLECULE if same_molecule else 0) | (_lib.MOUSE_FORCE_EXACT if exact else 0)
    _lib.check(L.mouse_rdf_frame(_ptr(pos), _ptr(tcode), _ptr(molc), n, int(ntypes), C.byref(grid), flags,
                                 edges.ctypes.data_as(C.POINTER(C.c_double)), int(nbin), _ptr(ws), ws.numel(),
                                 _ptr(hist), _stream()), "mouse_rdf_frame")
    torch.cuda.current_stream().synchronize()   # `edges` is host memory read asynchronously
    return hist, edges
