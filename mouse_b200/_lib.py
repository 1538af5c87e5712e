"""ctypes binding of ``libmouse_b200.so`` (the C-ABI declared in ``include/mouse_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails the
import / call raises.  PyTorch is only used by the callers for device memory and streams;
this module passes raw pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOUSE_B200_LIB") or os.path.join(_PKG, "libmouse_b200.so")   # env: tuning builds only

MOUSE_SAME_MOLECULE = 1
MOUSE_FORCE_EXACT = 2
MOUSE_NO_CUTOFF = 4
MOUSE_INCLUDE_SELF = 8
MOUSE_FORCE_FAST = 16


def MOUSE_SLOTS(n: int) -> int:
    """Tuning field of ``flags``: register-resident candidate slots per lane of the fast P2 sweep (0 = automatic)."""
    return (int(n) & 0xff) << 8


def MOUSE_WARPS(n: int) -> int:
    """Tuning field of ``flags``: resident sweep warps per SM (0 = as many as fit)."""
    return (int(n) & 0xff) << 16


class Grid(C.Structure):
    """``mouse_grid_t``"""
    _fields_ = [("box", C.c_double * 3), ("rcut", C.c_double), ("rc2", C.c_double),
                ("cell_w", C.c_double * 3), ("ncell_axis", C.c_int32 * 3), ("ncell", C.c_int32),
                ("fast", C.c_int32), ("guard", C.c_float), ("tilt", C.c_double * 3), ("triclinic", C.c_int32),
                ("reserved", C.c_int32)]


class P2Totals(C.Structure):
    """``mouse_p2_totals_t``"""
    _fields_ = [("sum_mean", C.c_double), ("n_refs", C.c_int64), ("n_pairs", C.c_int64),
                ("n_zero_len", C.c_int64), ("n_hist_over", C.c_int64), ("n_band", C.c_int64)]


EXPORTS = [
    "mouse_version", "mouse_last_error", "mouse_grid_plan", "mouse_grid_plan_triclinic", "mouse_bond_build",
    "mouse_bond_build_triclinic",
    "mouse_workspace_bytes", "mouse_cell_build", "mouse_p2_sweep", "mouse_p2_pair_lists", "mouse_p2_reduce",
    "mouse_p2_frame", "mouse_p2_frame_timed", "mouse_p2_frame_graph_create", "mouse_p2_frame_graph_launch",
    "mouse_p2_frame_graph_nodes", "mouse_p2_frame_graph_destroy", "mouse_rdf_frame_timed", "mouse_rdf_frame", "mouse_single_ref_workspace_bytes", "mouse_p2_single_ref",
    "mouse_rdf_single_ref", "mouse_director_p2", "mouse_atom_colour", "mouse_debug_counters", "mouse_parse_doubles", "mouse_parse_doubles_mt", "mouse_parse_dump_atoms",
    "mouse_parse_dump_ws_bytes", "mouse_parse_dump_atoms_dev", "mouse_index_dump", "mouse_parse_table",
]


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc")]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("mouse_b200: %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_double
    gp = C.POINTER(Grid)
    L.mouse_version.restype = C.c_char_p
    L.mouse_last_error.restype = C.c_char_p
    L.mouse_grid_plan.argtypes = [C.POINTER(C.c_double), dbl, dbl, i64, gp]
    L.mouse_grid_plan_triclinic.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), dbl, dbl, i64, gp]
    L.mouse_bond_build.argtypes = [vp, vp, i64, C.POINTER(C.c_double), vp, vp, vp]
    L.mouse_bond_build_triclinic.argtypes = [vp, vp, i64, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp, vp]
    L.mouse_workspace_bytes.argtypes = [i64, gp]
    L.mouse_workspace_bytes.restype = C.c_size_t
    L.mouse_cell_build.argtypes = [vp, vp, vp, vp, vp, i64, gp, vp, C.c_size_t, vp]
    L.mouse_p2_sweep.argtypes = [vp, vp, i64, gp, u32, vp, C.c_size_t, vp, vp, vp]
    L.mouse_p2_pair_lists.argtypes = [vp, vp, i64, gp, u32, vp, C.c_size_t, vp, vp, vp, i64, vp]
    L.mouse_p2_reduce.argtypes = [vp, vp, i64, gp, vp, C.c_size_t, vp, vp, vp, dbl, dbl, i32, vp]
    L.mouse_p2_frame.argtypes = [vp, vp, vp, vp, vp, i64, gp, u32, vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    L.mouse_p2_frame_timed.argtypes = [vp, vp, vp, vp, vp, i64, gp, u32, vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mouse_p2_frame_graph_create.argtypes = [vp, vp, vp, vp, vp, i64, gp, u32, vp, C.c_size_t, vp, vp, vp, vp, vp, vp,
                                              C.POINTER(C.c_void_p)]
    L.mouse_p2_frame_graph_launch.argtypes = [vp, vp]
    L.mouse_p2_frame_graph_nodes.argtypes = [vp]
    L.mouse_p2_frame_graph_destroy.argtypes = [vp]
    L.mouse_rdf_frame.argtypes = [vp, vp, vp, i64, i32, gp, u32, C.POINTER(C.c_double), i32, vp, C.c_size_t, vp, vp]
    L.mouse_rdf_frame_timed.argtypes = [vp, vp, vp, i64, i32, gp, u32, C.POINTER(C.c_double), i32, vp, C.c_size_t, vp, vp,
                                        vp, vp]
    L.mouse_single_ref_workspace_bytes.restype = C.c_size_t
    L.mouse_p2_single_ref.argtypes = [vp, i64, C.POINTER(C.c_double), C.POINTER(C.c_double), dbl, dbl, u32, vp, vp, vp]
    L.mouse_rdf_single_ref.argtypes = [vp, i64, C.POINTER(C.c_double), C.POINTER(C.c_double), u32,
                                       C.POINTER(C.c_double), i32, vp, vp, vp]
    L.mouse_director_p2.argtypes = [vp, vp, i64, C.POINTER(C.c_double), vp, vp, dbl, dbl, i32, vp]
    L.mouse_atom_colour.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
    L.mouse_debug_counters.argtypes = [gp, i64, vp, vp, vp]
    L.mouse_parse_doubles.argtypes = [C.c_char_p, i64, vp, i64]
    L.mouse_parse_doubles.restype = i64
    L.mouse_parse_doubles_mt.argtypes = [C.c_char_p, i64, vp, i64, i32]
    L.mouse_parse_doubles_mt.restype = i64
    L.mouse_parse_dump_atoms.argtypes = [vp, i64, i64, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, i32]
    L.mouse_parse_dump_atoms.restype = i64
    L.mouse_parse_table.argtypes = [vp, i64, i64, i32, i32, vp, vp, vp, vp, i32]
    L.mouse_parse_table.restype = i64
    L.mouse_index_dump.argtypes = [vp, i64, vp, i64, i32]
    L.mouse_index_dump.restype = i64
    L.mouse_parse_dump_ws_bytes.argtypes = [i64, i64]
    L.mouse_parse_dump_ws_bytes.restype = C.c_size_t
    L.mouse_parse_dump_atoms_dev.argtypes = [vp, i64, i64, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                             C.c_size_t, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().mouse_last_error().decode()))


def grid_plan(box, rcut: float, n_items: int, rc2: float | None = None, tilt=None) -> Grid:
    """Host-side cell grid.  ``rc2`` defaults to the Python expression ``rcut**2`` so that the
    threshold bits equal the reference's (``ordering_functions.py:102``).  ``tilt`` = (xy, xz, yz) plans a triclinic
    box (build-defined semantics: the reference is orthorhombic only, ``lammpack_types.py:440-451``)."""
    g = Grid()
    b = (C.c_double * 3)(float(box[0]), float(box[1]), float(box[2]))
    r2 = float(rcut)**2 if rc2 is None else float(rc2)
    if tilt is None:
        check(lib().mouse_grid_plan(b, float(rcut), r2, int(n_items), C.byref(g)), "mouse_grid_plan")
    else:
        t = (C.c_double * 3)(float(tilt[0]), float(tilt[1]), float(tilt[2]))
        check(lib().mouse_grid_plan_triclinic(b, t, float(rcut), r2, int(n_items), C.byref(g)),
              "mouse_grid_plan_triclinic")
    return g
