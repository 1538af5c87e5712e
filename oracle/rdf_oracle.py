"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the reference's RDF pair histogram.

Restates ``calculateRdfForReference`` (``ordering_functions.py:66-88``) and the accumulation
loop of ``CalculatePairCorrelationFunctions`` (``ordering_functions.py:122-177``) with the
counts accumulated onto ZEROS (the reference starts from uninitialised memory at ``:155``;
SURVEY.md §8(a) a7 - intended behaviour).  Pinned against tests/golden/rdf_*.npz, which were
produced by calling the reference's own ``calculateRdfForReference`` per atom."""
from __future__ import annotations

import numpy as np


def rdf_for_reference(box, ref_pos, ref_num, ref_mol, num, mol, pos, rmin, rmax, nbin, same_molecule=True):
    """Literal per-reference histogram (``ordering_functions.py:66-88``): returns counts or None
    when the reference raises NameError (nothing left after masking)."""
    Lx, Ly, Lz = (float(v) for v in box)
    if rmax < 0.:
        rmax = min(Lx, Ly, Lz) / 2.
    drx = np.add(pos[:, 0], -1. * ref_pos[0])
    dry = np.add(pos[:, 1], -1. * ref_pos[1])
    drz = np.add(pos[:, 2], -1. * ref_pos[2])
    drxT = np.add(drx, -1. * Lx * np.around(drx / Lx))
    dryT = np.add(dry, -1. * Ly * np.around(dry / Ly))
    drzT = np.add(drz, -1. * Lz * np.around(drz / Lz))
    notSelf = num != ref_num
    notSame = (mol != ref_mol) if not same_molecule else 1.
    d = notSelf * notSame * np.sqrt(drxT**2 + dryT**2 + drzT**2)
    dm = np.ma.masked_equal(d, 0.)
    if dm.count() > 0:
        return np.histogram(dm.compressed(), bins=nbin, range=(rmin, rmax))[0]
    return None


def pair_keys(type_names):
    """Sorted type list with the reference's pseudo type 'AAAAA' (``:136-137``) and the
    ``"ti_tj"`` keys for i <= j (``:152-156``)."""
    types = sorted(list(type_names) + ["AAAAA"])
    keys = [str(types[i]) + "_" + str(types[j]) for i in range(len(types)) for j in range(i, len(types))]
    return types, keys


def pair_correlation_counts(positions, atom_type_names_per_atom, mol, box, atomtypes, rmin, rmax, nbin,
                            same_molecule=True):
    """Theta(N^2) accumulation of ``ordering_functions.py:162-173`` -> {key: int64[nbin]}."""
    names = np.asarray(atom_type_names_per_atom)
    present = []
    for t in names.tolist():
        if (atomtypes is None or len(atomtypes) == 0 or t in atomtypes) and t not in present:
            present.append(t)
    types, keys = pair_keys(present)
    data = {k: np.zeros(nbin, np.int64) for k in keys}
    num = np.arange(1, len(names) + 1)
    sel = {t: np.flatnonzero(names == t) for t in types}
    for i in range(len(names)):
        t1 = names[i]
        if t1 not in types:
            continue
        for t2 in types:
            s = sel[t2]
            if len(s) == 0:
                continue
            h = rdf_for_reference(box, positions[i], num[i], mol[i], num[s], mol[s], positions[s], rmin, rmax,
                                  nbin, same_molecule)
            if h is not None:
                data["_".join(sorted([t1, t2]))] += h
    return types, data


def rdf_tables(rmin, rmax, nbin):
    """``r`` (bin centres) and ``v`` (shell volumes) of ``ordering_functions.py:140-149``."""
    r = [rmin + (rmax - rmin) * (i + 0.5) / nbin for i in range(nbin)]
    v = []
    for i in range(nbin):
        r1 = rmin + (rmax - rmin) * i / nbin
        r2 = rmin + (rmax - rmin) * (i + 1) / nbin
        v.append(4. / 3. * np.pi * (r2**3 - r1**3))
    return r, v
