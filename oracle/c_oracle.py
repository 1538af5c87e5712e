"""TEST INFRASTRUCTURE ONLY - ctypes front-end of the plain-C oracle (oracle/c/mouse_oracle.c).

Used by tests/, smoke() and bench.py's CPU-baseline leg as the CHECKER for sizes where the
Theta(N^2) NumPy restatement is intractable (10^5 - 10^6 bonds).  Never imported by the
product package.  Pinned against the golden vectors in tests/test_oracle_golden.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmouse_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "mouse_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        p = C.c_void_p
        _lib.mo_bond_build.argtypes = [p, p, C.c_int64, p, p, p]
        _lib.mo_p2.argtypes = [p, p, p, p, p, C.c_int64, p, C.c_double, C.c_double, C.c_int, p, p, p,
                               C.POINTER(C.c_void_p)]
        _lib.mo_p2.restype = C.c_int
        _lib.mo_rdf.argtypes = [p, p, p, C.c_int64, C.c_int, p, C.c_double, p, C.c_int64, C.c_int, p]
        _lib.mo_rdf.restype = C.c_int
        _lib.mo_free.argtypes = [p]
        # triclinic variants (build-defined semantics, parity unpinned: the reference is orthorhombic only)
        _lib.mo_bond_build_tri.argtypes = [p, p, C.c_int64, p, p, p, p]
        _lib.mo_p2_tri.argtypes = [p, p, p, p, p, C.c_int64, p, p, C.c_double, C.c_double, C.c_int, p, p, p,
                                   C.POINTER(C.c_void_p)]
        _lib.mo_p2_tri.restype = C.c_int
        _lib.mo_rdf_tri.argtypes = [p, p, p, C.c_int64, C.c_int, p, p, C.c_double, p, C.c_int64, C.c_int, p]
        _lib.mo_rdf_tri.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def bond_build(positions, bond_atoms, box, tilt=None):
    """``tilt`` = (xy, xz, yz) selects the triclinic (build-defined) variant."""
    pos = np.ascontiguousarray(positions, np.float64)
    b = np.ascontiguousarray(bond_atoms, np.int32)
    box = np.ascontiguousarray(box, np.float64)
    vec = np.empty((len(b), 3), np.float64)
    ctr = np.empty((len(b), 3), np.float64)
    if tilt is None:
        lib().mo_bond_build(_ptr(pos), _ptr(b), len(b), _ptr(box), _ptr(vec), _ptr(ctr))
    else:
        t = np.ascontiguousarray(tilt, np.float64)
        lib().mo_bond_build_tri(_ptr(pos), _ptr(b), len(b), _ptr(box), _ptr(t), _ptr(vec), _ptr(ctr))
    return vec, ctr


def p2(vec, ctr, mol, box, rcut, same_molecule=True, ref_mask=None, nbr_mask=None, want_lists=False,
       threads=None, tilt=None):
    """Per-bond (sum cos^2, count[, neighbour lists]) with the reference's arithmetic."""
    n = len(vec)
    vec = np.ascontiguousarray(vec, np.float64)
    ctr = np.ascontiguousarray(ctr, np.float64)
    mol = np.ascontiguousarray(mol, np.int64)
    box = np.ascontiguousarray(box, np.float64)
    rm = np.ones(n, np.uint8) if ref_mask is None else np.ascontiguousarray(ref_mask, np.uint8)
    nm = np.ones(n, np.uint8) if nbr_mask is None else np.ascontiguousarray(nbr_mask, np.uint8)
    ssum = np.zeros(n, np.float64)
    cnt = np.zeros(n, np.int64)
    ptr = np.zeros(n + 1, np.int64) if want_lists else None
    idx_p = C.c_void_p()
    if threads is not None:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    if tilt is None:
        rc = lib().mo_p2(_ptr(vec), _ptr(ctr), _ptr(mol), _ptr(rm), _ptr(nm), n, _ptr(box), float(rcut),
                         float(rcut)**2, int(bool(same_molecule)), _ptr(ssum), _ptr(cnt),
                         _ptr(ptr) if want_lists else None, C.byref(idx_p))
    else:
        t = np.ascontiguousarray(tilt, np.float64)
        rc = lib().mo_p2_tri(_ptr(vec), _ptr(ctr), _ptr(mol), _ptr(rm), _ptr(nm), n, _ptr(box), _ptr(t), float(rcut),
                             float(rcut)**2, int(bool(same_molecule)), _ptr(ssum), _ptr(cnt),
                             _ptr(ptr) if want_lists else None, C.byref(idx_p))
    if rc:
        raise MemoryError("mo_p2 failed")
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, ssum / np.maximum(cnt, 1), np.nan)
    out = dict(sum=ssum, count=cnt, mean=mean)
    if want_lists:
        m = int(ptr[-1])
        buf = (C.c_int64 * max(m, 1)).from_address(idx_p.value)
        out["nbr_ptr"] = ptr
        out["nbr_idx"] = np.array(buf[:m], dtype=np.int64)
        lib().mo_free(idx_p)
    return out


def rdf(positions, type_code, mol, ntypes, box, rmin, rmax, nbin, same_molecule=True, tilt=None):
    """int64 histogram [ntypes*(ntypes+1)/2, nbin] over ordered pairs; bin edges are NumPy's."""
    pos = np.ascontiguousarray(positions, np.float64)
    t = np.ascontiguousarray(type_code, np.int32)
    mol = np.ascontiguousarray(mol, np.int64)
    box = np.ascontiguousarray(box, np.float64)
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=int(nbin), range=(rmin, rmax)))
    hist = np.zeros((ntypes * (ntypes + 1) // 2, int(nbin)), np.int64)
    if tilt is None:
        rc = lib().mo_rdf(_ptr(pos), _ptr(t), _ptr(mol), len(pos), int(ntypes), _ptr(box), float(edges[-1]),
                          _ptr(edges), int(nbin), int(bool(same_molecule)), _ptr(hist))
    else:
        tl = np.ascontiguousarray(tilt, np.float64)
        rc = lib().mo_rdf_tri(_ptr(pos), _ptr(t), _ptr(mol), len(pos), int(ntypes), _ptr(box), _ptr(tl),
                              float(edges[-1]), _ptr(edges), int(nbin), int(bool(same_molecule)), _ptr(hist))
    if rc:
        raise MemoryError("mo_rdf failed")
    return hist, edges
