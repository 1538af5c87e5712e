"""TEST INFRASTRUCTURE ONLY - CPU restatement (NumPy) of the reference's local-P2 path.

Nothing in the product path (``mouse_b200``) imports this file; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Parity pin: the reference ships no tests or golden vectors, so the pin is
the reference itself, executed in the build container by ``tests/golden/make_golden.py``
(through ``oracle/reference_harness.py``); its outputs are committed under
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks every function here
bit-for-bit against them (per-vector mean cos^2, neighbour counts, neighbour index lists,
S, histograms).

Every function cites the reference lines it restates (paths relative to the reference
root).  The expressions, dtypes and evaluation order are kept identical so that the
"literal" functions are bit-exact against the reference on the same NumPy/libm.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- a1/a2/a3
def bond_build(positions, bond_atoms, box):
    """Bond vector + midpoint with minimum image, vectorised.

    Restates ``vector_and_center_pbc_trim`` method 1 (``lammpack_types.py:587-599``) as
    called by ``Config.bond_vector_by_num`` (``lammpack_types.py:508-515``) for every bond
    in ``createNumpyBondsArrayFromConfig`` (``ordering_functions.py:29``):
        x = p2 - p1; xc = (p2 + p1) / 2.
        x_t = remainder(x + L/2., L) - L/2.;  xc_t = xc + (x_t - x) / 2.
    Returns (vec, ctr), each (N_b, 3) float64.  The midpoint is NOT wrapped into the box.
    """
    p = np.asarray(positions, dtype=np.float64)
    b = np.asarray(bond_atoms)
    box = np.asarray(box, dtype=np.float64)
    p1 = p[b[:, 0]]
    p2 = p[b[:, 1]]
    x = p2 - p1
    xc = (p2 + p1) / 2.
    x_t = np.remainder(x + box / 2., box) - box / 2.
    xc_t = xc + (x_t - x) / 2.
    return x_t, xc_t


def np_bonds_array(num, res_hash, mol_hash, vec, ctr):
    """The (9, N_b) float64 array of ``createNumpyBondsArrayFromConfig``
    (``ordering_functions.py:41``): rows num, res hash, mol hash, bx,by,bz, rx,ry,rz."""
    return np.array((np.asarray(num), np.asarray(res_hash), np.asarray(mol_hash),
                     vec[:, 0], vec[:, 1], vec[:, 2], ctr[:, 0], ctr[:, 1], ctr[:, 2]))


class NoValues(Exception):
    """Stands for the reference's ``NameError("No values to calculate")`` /
    ``NameError("Zero length of reference bond ...")`` (``ordering_functions.py:117,119``)."""


# --------------------------------------------------------------------------- a4 literal
def ave_cos_sq_literal(np_bonds, box, ref_col, ref_vec, ref_ctr, ref_num, ref_mol_hash,
                       rcut, exclude_self=True, same_molecule=True, empty_mol_hash=None):
    """One reference bond against ALL columns of ``np_bonds``: literal restatement of
    ``calculateAveCosSqForReference`` (``ordering_functions.py:91-119``), same NumPy
    expressions in the same order (full-length passes, ``np.around``, ``np.ma.average``).

    ``ref_vec``/``ref_ctr`` are what ``config.bond_vector_by_num`` returns for the reference
    bond (np.float64 scalars, so that ``bRef.x**2`` goes through the scalar ``pow`` exactly
    as in the reference).  Returns (mean, count, counted_column_indices)."""
    bRx, bRy, bRz = (np.float64(v) for v in ref_vec)
    cRx, cRy, cRz = (np.float64(v) for v in ref_ctr)
    Lx, Ly, Lz = (float(v) for v in box)
    if bRx != 0. or bRy != 0. or bRz != 0.:
        bx, by, bz = np_bonds[3], np_bonds[4], np_bonds[5]
        drx = np_bonds[6] - cRx
        dry = np_bonds[7] - cRy
        drz = np_bonds[8] - cRz
        drxTrim = drx - Lx * np.around(drx / Lx)
        dryTrim = dry - Ly * np.around(dry / Ly)
        drzTrim = drz - Lz * np.around(drz / Lz)
        rsq = drxTrim**2 + dryTrim**2 + drzTrim**2
        if rcut >= 0.:
            inRange = rsq <= rcut**2
        else:
            inRange = 1.
        if exclude_self:
            notSelf = np_bonds[0] != ref_num
        else:
            notSelf = 1.
        if not same_molecule:
            if empty_mol_hash is not None and np.sum(np_bonds[2] == empty_mol_hash) > 0:
                raise NoValues("empty mol ids")
            notSame = np_bonds[2] != ref_mol_hash
        else:
            notSame = 1.
        cosSq = inRange * notSelf * notSame * (bRx * bx + bRy * by + bRz * bz)**2 \
            / (bRx**2 + bRy**2 + bRz**2) / (bx**2 + by**2 + bz**2)
        masked = np.ma.masked_equal(cosSq, 0.)
        if masked.count() > 0:
            idx = np.flatnonzero(~np.ma.getmaskarray(masked))
            return np.ma.average(masked), int(masked.count()), idx
        raise NoValues("No values to calculate")
    raise NoValues("Zero length of reference bond")


def p2_literal(vec, ctr, num, mol_hash, box, rcut, same_molecule=True, ref_mask=None,
               nbr_mask=None, empty_mol_hash=None):
    """Theta(N^2) driver: restates the loop of ``CalculateOrientationOrderParameter``
    (``ordering_functions.py:196-214``) in terms of column indices.

    ``nbr_mask`` selects the columns that make up ``npBonds`` (bonds passing ``bondtypes``,
    ``:196``); ``ref_mask`` the reference bonds (``bondtypes`` and ``referenceResTypes``,
    ``:198-199``).  Returns dict(mean, count, nbr_ptr, nbr_idx) over ALL bonds (in original
    bond order); bonds that are not references or were skipped (NameError) have
    count 0 and mean NaN.  ``nbr_idx`` holds ORIGINAL bond indices."""
    n = vec.shape[0]
    ref_mask = np.ones(n, bool) if ref_mask is None else np.asarray(ref_mask, bool)
    nbr_mask = np.ones(n, bool) if nbr_mask is None else np.asarray(nbr_mask, bool)
    cols = np.flatnonzero(nbr_mask)
    np_bonds = np_bonds_array(np.asarray(num)[cols], np.zeros(len(cols)), np.asarray(mol_hash)[cols],
                              vec[cols], ctr[cols])
    mean = np.full(n, np.nan)
    count = np.zeros(n, np.int64)
    ptr = np.zeros(n + 1, np.int64)
    idx = []
    for i in range(n):
        ptr[i + 1] = ptr[i]
        if not (ref_mask[i] and nbr_mask[i]):
            continue
        try:
            m, c, k = ave_cos_sq_literal(np_bonds, box, None, vec[i], ctr[i], num[i], mol_hash[i], rcut,
                                         True, same_molecule, empty_mol_hash)
        except NoValues:
            continue
        mean[i], count[i] = m, c
        idx.append(cols[k])
        ptr[i + 1] += len(k)
    return dict(mean=mean, count=count, nbr_ptr=ptr,
                nbr_idx=np.concatenate(idx) if idx else np.zeros(0, np.int64))


# --------------------------------------------------------------------------- a4 on candidate pairs
def pair_terms(vec, ctr, box, i, j):
    """The per-pair quantities of ``ordering_functions.py:95-101,112`` for index arrays
    (i = reference, j = neighbour): rsq (exact reference evaluation order), the un-fused
    dot product and the normalised cos^2 (without the masks)."""
    Lx, Ly, Lz = (float(v) for v in box)
    drx = ctr[j, 0] - ctr[i, 0]
    dry = ctr[j, 1] - ctr[i, 1]
    drz = ctr[j, 2] - ctr[i, 2]
    drxT = drx - Lx * np.around(drx / Lx)
    dryT = dry - Ly * np.around(dry / Ly)
    drzT = drz - Lz * np.around(drz / Lz)
    rsq = drxT**2 + dryT**2 + drzT**2
    dot = vec[i, 0] * vec[j, 0] + vec[i, 1] * vec[j, 1] + vec[i, 2] * vec[j, 2]
    nr = vec[i, 0]**2 + vec[i, 1]**2 + vec[i, 2]**2
    nj = vec[j, 0]**2 + vec[j, 1]**2 + vec[j, 2]**2
    with np.errstate(invalid="ignore", divide="ignore"):
        cos2 = dot**2 / nr / nj
    return rsq, dot, cos2


def candidate_pairs(ctr, box, rcut, chunk=200_000):
    """Ordered candidate pairs (i, j), i != j, within rcut*(1+1e-9) under the minimum image,
    from a periodic KD-tree on a WRAPPED COPY of the midpoints (the exact ``<=`` test is
    applied afterwards with the reference's expressions).  Yields (i, j) array chunks."""
    from scipy.spatial import cKDTree
    box = np.asarray(box, np.float64)
    w = np.mod(ctr, box)
    w = np.where(w >= box, w - box, w)
    tree = cKDTree(w, boxsize=box)
    r = rcut * (1. + 1e-9) + 1e-12
    n = len(ctr)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lists = tree.query_ball_point(w[s:e], r, workers=-1)
        lens = np.fromiter((len(l) for l in lists), np.int64, e - s)
        i = np.repeat(np.arange(s, e, dtype=np.int64), lens)
        j = np.fromiter((k for l in lists for k in l), np.int64, int(lens.sum()))
        keep = i != j
        yield i[keep], j[keep]


def p2_pairs(vec, ctr, mol_hash, box, rcut, same_molecule=True, ref_mask=None, nbr_mask=None,
             want_lists=False):
    """Cell/tree-accelerated restatement (SURVEY.md §8(c) "Tractability"): the candidate set
    comes from a KD-tree, every kept pair is decided by the reference's own expressions
    (``rsq <= rcut**2`` in float64, ``cos^2 != 0`` mask, ``ordering_functions.py:102-113``).
    Requires rcut < min(box)/2 (otherwise use ``p2_literal``).  Sums are taken over the
    counted neighbours in ascending neighbour order.  Returns dict(sum, count, mean[, nbr_ptr,
    nbr_idx]) over all bonds; non-reference / skipped bonds have count 0 and mean NaN.
    A zero-length NEIGHBOUR bond poisons every mean with NaN in the reference
    (0/0) - this function reproduces that."""
    n = vec.shape[0]
    ref_mask = np.ones(n, bool) if ref_mask is None else np.asarray(ref_mask, bool)
    nbr_mask = np.ones(n, bool) if nbr_mask is None else np.asarray(nbr_mask, bool)
    ref_mask = ref_mask & nbr_mask            # the driver only visits bonds passing bondtypes too
    zero_len = (vec[:, 0] == 0.) & (vec[:, 1] == 0.) & (vec[:, 2] == 0.)
    rc2 = float(rcut)**2
    ssum = np.zeros(n, np.float64)
    cnt = np.zeros(n, np.int64)
    pi, pj = [], []
    for i, j in candidate_pairs(ctr, box, rcut):
        keep = ref_mask[i] & nbr_mask[j] & ~zero_len[i] & ~zero_len[j]
        if not same_molecule:
            keep &= mol_hash[i] != mol_hash[j]
        i, j = i[keep], j[keep]
        rsq, dot, cos2 = pair_terms(vec, ctr, box, i, j)
        hit = (rsq <= rc2) & (cos2 != 0.)
        i, j, cos2 = i[hit], j[hit], cos2[hit]
        order = np.lexsort((j, i))
        i, j, cos2 = i[order], j[order], cos2[order]
        np.add.at(ssum, i, cos2)
        np.add.at(cnt, i, 1)
        if want_lists:
            pi.append(i)
            pj.append(j)
    nz = np.flatnonzero(zero_len & nbr_mask)
    if len(nz):                                # NaN poisoning by zero-length neighbours
        live = ref_mask & ~zero_len
        cnt[live] += len(nz)
        ssum[live] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(cnt > 0, ssum / np.maximum(cnt, 1), np.nan)
    out = dict(sum=ssum, count=cnt, mean=mean)
    if want_lists:
        i = np.concatenate(pi) if pi else np.zeros(0, np.int64)
        j = np.concatenate(pj) if pj else np.zeros(0, np.int64)
        order = np.lexsort((j, i))
        i, j = i[order], j[order]
        ptr = np.zeros(n + 1, np.int64)
        np.add.at(ptr, i + 1, 1)
        out["nbr_ptr"] = np.cumsum(ptr)
        out["nbr_idx"] = j
    return out


# --------------------------------------------------------------------------- a5 reductions
def order_parameter(mean, count, sequential=True):
    """``S = 1.5 * (sum_i m_i) / i_s - 0.5`` over references with >= 1 counted neighbour
    (``ordering_functions.py:205-206, 221-224``).  ``sequential`` reproduces the reference's
    left-to-right Python accumulation bit-for-bit."""
    vals = mean[count > 0]
    i_s = len(vals)
    if i_s == 0:
        raise UnboundLocalError("no reference bond produced a value")
    if sequential:
        acc = 0.
        for v in vals.tolist():
            acc += v
    else:
        acc = float(np.sum(vals))
    return np.float64(1.5 * acc / i_s - 0.5)


def create_histogram(minvalue, maxvalue, nbins):
    """``createHistogram`` (``lammpack_misc.py:177-183``): labels use step=(max-min)/(nbins-1)."""
    if not isinstance(nbins, int):
        raise NameError("Number of bins (" + str(nbins) + ") in the histogram must be integer")
    if nbins > 1:
        step = (maxvalue - minvalue) / (nbins - 1)
    elif nbins == 1:
        step = maxvalue - minvalue
    else:
        raise NameError("Invalid number of bins in the histogram (" + str(nbins) + ")")
    return {"min": minvalue, "max": maxvalue, "nbins": nbins, "step": step,
            "values": [0.] * nbins, "norm": [0.] * nbins}


def p2_histogram(mean, count, smin=-0.5, smax=1.0, sbins=75):
    """'histo' mode of ``CalculateOrientationOrderParameter`` (``ordering_functions.py:187-193,
    203, 225-227``) with ``updateHistogram``/``normHistogram`` (``lammpack_misc.py:185-203``):
    bin = int(nbins*(v-min)/(max-min)) (Python list indexing: negative wraps, == nbins raises
    IndexError), values normalised by their sum, ``norm`` = 1/(cos_left + cos_right)."""
    h = create_histogram(smin, smax, sbins)
    for i in range(sbins):
        sleft, sright = smin + i * h["step"], smin + (i + 1) * h["step"]
        cl, cr = (2. / 3. * (sleft + 0.5))**0.5, (2. / 3. * (sright + 0.5))**0.5
        h["norm"][i] = 1. / (cl + cr)
    for m, c in zip(mean.tolist(), count.tolist()):
        if c > 0:
            v = 1.5 * m - 0.5
            ibin = int(len(h["values"]) * (v - h["min"]) / (h["max"] - h["min"]))
            h["values"][ibin] += 1
    hs = 0.
    for v in h["values"]:
        hs += v
    if hs > 0.:
        h["values"] = [v / hs for v in h["values"]]
    return h


def atom_colouring(mean, count, bond_atoms, n_atoms):
    """``storeAsAtomtypes`` (``ordering_functions.py:207-220``): per atom, mean P2 of its
    reference bonds that produced a value -> type string "C%02d" / "Cm%02d".
    Returns list of (atom_index, type_string) for atoms that got a value."""
    vals = [[] for _ in range(n_atoms)]
    for b in range(len(mean)):
        if count[b] > 0:
            p2 = 1.5 * mean[b] - 0.5
            vals[int(bond_atoms[b, 0])].append(p2)
            vals[int(bond_atoms[b, 1])].append(p2)
    out = []
    for a, v in enumerate(vals):
        if v:
            mv = np.mean(v)
            out.append((a, "C" + "%02d" % int(mv * 100) if mv >= 0. else "Cm" + "%02d" % int(mv * -100)))
    return out
