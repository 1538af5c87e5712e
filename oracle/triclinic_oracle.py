"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the BUILD-DEFINED triclinic semantics.

PARITY UNPINNED: the reference is orthorhombic only (``lammpack_types.py:440-451`` reads ``xlo xhi`` / ``ylo yhi`` /
``zlo zhi`` and drops the ``xy xz yz`` line; ``lammpack_pdb_functions.py:71-77`` drops the CRYST1 angles), so there is
nothing of the reference to compare a tilted box with.  BASELINE.json config 5 names a triclinic box; this file
defines what the library computes for it and is pinned three ways (tests/test_triclinic_cpu.py):

* with tilt = 0 every function is bit-identical to the orthorhombic oracle (which IS pinned to the reference);
* ``min_image`` equals the brute-force nearest image over 5 x 5 x 5 lattice translations whenever that image is
  shorter than half the smallest box edge (the same validity range as the reference's ``dr - L*around(dr/L)``);
* the C oracle's triclinic variants (cell list in fractional coordinates) agree with the all-pairs functions here.

Convention (LAMMPS): a = (lx, 0, 0), b = (xy, ly, 0), c = (xz, yz, lz); ``box = (lx, ly, lz)``, ``tilt = (xy, xz, yz)``.
The minimum image of a displacement is its representative in the brick |dx| <= lx/2, |dy| <= ly/2, |dz| <= lz/2,
found axis by axis from z to x with the reference's own step ``dr - L * np.around(dr / L)``
(``ordering_functions.py:98-100``); the tilt components of the subtracted lattice vector are removed from the lower
axes.  Nothing in ``mouse_b200`` imports this module.
"""
from __future__ import annotations

import numpy as np


def min_image(dx, dy, dz, box, tilt):
    """Brick representative of the displacement(s) (dx, dy, dz); un-fused, in this order."""
    lx, ly, lz = (float(v) for v in box)
    xy, xz, yz = (float(v) for v in tilt)
    nz = np.around(dz / lz)
    dz = dz - lz * nz
    dy = dy - yz * nz
    dx = dx - xz * nz
    ny = np.around(dy / ly)
    dy = dy - ly * ny
    dx = dx - xy * ny
    nx = np.around(dx / lx)
    dx = dx - lx * nx
    return dx, dy, dz


def min_image_bruteforce(d, box, tilt, reach=2):
    """Shortest image of each displacement d (n,3) over (2 reach + 1)^3 lattice translations (float64, plain)."""
    lx, ly, lz = (float(v) for v in box)
    xy, xz, yz = (float(v) for v in tilt)
    a, b, c = np.array([lx, 0., 0.]), np.array([xy, ly, 0.]), np.array([xz, yz, lz])
    d = np.asarray(d, np.float64)
    best = np.full(len(d), np.inf)
    out = np.zeros_like(d)
    r = range(-reach, reach + 1)
    for i in r:
        for j in r:
            for k in r:
                t = d - (i * a + j * b + k * c)
                n2 = (t * t).sum(1)
                m = n2 < best
                best[m] = n2[m]
                out[m] = t[m]
    return out, best


def bond_build(positions, bond_atoms, box, tilt):
    """Bond vector + midpoint: the ``np.remainder`` form of ``lammpack_types.py:587-599`` applied axis by axis from z
    to x; a shift by n lattice vectors along an axis also removes n times their tilt components from the lower axes.
    Returns (vec, ctr), each (N_b, 3); the midpoint is not wrapped (as in the reference)."""
    p = np.asarray(positions, np.float64)
    b = np.asarray(bond_atoms)
    lx, ly, lz = (float(v) for v in box)
    xy, xz, yz = (float(v) for v in tilt)
    p1, p2 = p[b[:, 0]], p[b[:, 1]]
    x = p2 - p1
    xc = (p2 + p1) / 2.
    wx, wy, wz = x[:, 0], x[:, 1], x[:, 2]
    tz = np.remainder(wz + lz / 2., lz) - lz / 2.
    nz = np.around((wz - tz) / lz)
    wy = wy - yz * nz
    wx = wx - xz * nz
    ty = np.remainder(wy + ly / 2., ly) - ly / 2.
    ny = np.around((wy - ty) / ly)
    wx = wx - xy * ny
    tx = np.remainder(wx + lx / 2., lx) - lx / 2.
    x_t = np.stack([tx, ty, tz], axis=1)
    return x_t, xc + (x_t - x) / 2.


def p2_all_pairs(vec, ctr, mol, box, tilt, rcut, same_molecule=True, ref_mask=None, nbr_mask=None):
    """All-pairs restatement of ``calculateAveCosSqForReference`` (``ordering_functions.py:91-119``) with the triclinic
    minimum image in place of ``:98-100``.  Returns dict(sum, count, mean) over all bonds (original order)."""
    n = len(vec)
    ref_mask = np.ones(n, bool) if ref_mask is None else np.asarray(ref_mask, bool)
    nbr_mask = np.ones(n, bool) if nbr_mask is None else np.asarray(nbr_mask, bool)
    mol = np.asarray(mol)
    bx, by, bz = vec[:, 0], vec[:, 1], vec[:, 2]
    ssum, cnt = np.zeros(n), np.zeros(n, np.int64)
    for i in range(n):
        if not (ref_mask[i] and nbr_mask[i]):
            continue
        bRx, bRy, bRz = vec[i]
        if bRx == 0. and bRy == 0. and bRz == 0.:
            continue
        dx, dy, dz = min_image(ctr[:, 0] - ctr[i, 0], ctr[:, 1] - ctr[i, 1], ctr[:, 2] - ctr[i, 2], box, tilt)
        rsq = dx**2 + dy**2 + dz**2
        ok = nbr_mask & (np.arange(n) != i)
        if rcut >= 0.:
            ok &= rsq <= rcut**2
        if not same_molecule:
            ok &= mol != mol[i]
        with np.errstate(invalid="ignore", divide="ignore"):
            cs = (bRx * bx + bRy * by + bRz * bz)**2 / (bRx * bRx + bRy * bRy + bRz * bRz) / (bx**2 + by**2 + bz**2)
        cs = np.where(ok, cs, 0.)
        keep = cs != 0.
        ssum[i], cnt[i] = cs[keep].sum(), keep.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, ssum / np.maximum(cnt, 1), np.nan)
    return dict(sum=ssum, count=cnt, mean=mean)


def rdf_all_pairs(positions, type_code, mol, ntypes, box, tilt, rmin, rmax, nbin, same_molecule=True):
    """All-pairs restatement of ``calculateRdfForReference`` summed as in ``CalculatePairCorrelationFunctions``
    (``ordering_functions.py:66-88, 162-173``) with the triclinic minimum image; int64 [npairs, nbin]."""
    pos = np.asarray(positions, np.float64)
    t = np.asarray(type_code)
    mol = np.asarray(mol)
    n = len(pos)
    hist = np.zeros((ntypes * (ntypes + 1) // 2, int(nbin)), np.int64)
    for i in range(n):
        if t[i] < 0:
            continue
        dx, dy, dz = min_image(pos[:, 0] + -1. * pos[i, 0], pos[:, 1] + -1. * pos[i, 1], pos[:, 2] + -1. * pos[i, 2],
                               box, tilt)
        d = np.sqrt(dx**2 + dy**2 + dz**2)
        ok = (t >= 0) & (np.arange(n) != i) & (d != 0.)
        if not same_molecule:
            ok &= mol != mol[i]
        for tj in range(ntypes):
            sel = ok & (t == tj)
            a, b = min(t[i], tj), max(t[i], tj)
            slot = a * ntypes - a * (a - 1) // 2 + (b - a)
            h, _ = np.histogram(d[sel], bins=int(nbin), range=(rmin, rmax))
            hist[slot] += h
    return hist


def wrap_into_cell(positions, box, tilt):
    """Positions wrapped into the triclinic cell (fractional coordinates in [0, 1)), for building test inputs."""
    p = np.asarray(positions, np.float64)
    lx, ly, lz = (float(v) for v in box)
    xy, xz, yz = (float(v) for v in tilt)
    sz = p[:, 2] / lz
    sy = (p[:, 1] - yz * sz) / ly
    sx = (p[:, 0] - xy * sy - xz * sz) / lx
    s = np.stack([sx, sy, sz], axis=1)
    s -= np.floor(s)
    return np.stack([s[:, 0] * lx + s[:, 1] * xy + s[:, 2] * xz, s[:, 1] * ly + s[:, 2] * yz, s[:, 2] * lz], axis=1)
