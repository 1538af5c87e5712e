"""Drop-in mirror of the reference's ``ordering_functions.py`` call surface.

Same function names, argument meaning, return types and error behaviour as the reference
(mglagolev/mouse), but every array computation runs in hand-written sm_100a kernels reached
through the C-ABI of ``libmouse_b200.so``.  ``config`` may be a reference-style ``Config``
(duck-typed: ``atoms() / bonds() / box()``) or a ``mouse_b200.frames.FrameArrays``.

There is no CPU fallback: without the CUDA extension or a device these functions raise.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np
import torch

from . import _lib, engine
from .frames import FrameArrays, hash8
from .lammpack_misc import createHistogram, normHistogram

__all__ = ["hash8", "createNumpyBondsArrayFromConfig", "createNumpyAtomsArrayFromConfig",
           "calculateRdfForReference", "calculateAveCosSqForReference", "CalculatePairCorrelationFunctions",
           "CalculateOrientationOrderParameter", "CalculateCrystallinityParameter", "CalculateComRdfs"]


def _arrays(config) -> FrameArrays:
    return config if isinstance(config, FrameArrays) else FrameArrays.from_config(config)


class DeviceBackedArray(np.ndarray):
    """NumPy array (what the reference returns) that remembers its device copy, so that the
    per-reference functions do not re-upload it on every call."""
    _dev = None

    def device_tensor(self):
        if self._dev is None or self._dev.shape != self.shape:
            self._dev = torch.as_tensor(np.ascontiguousarray(self), dtype=torch.float64).cuda()
        return self._dev


def _device_copy(a):
    if isinstance(a, DeviceBackedArray):
        return a.device_tensor()
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


# --------------------------------------------------------------------------------------- a3 / a8
def createNumpyBondsArrayFromConfig(config, allowedBondTypes=[], allowedResTypes=[], allowedMolIds=[],
                                    allowedBondNums=[]):
    """(9, N_b) float64 array ``[num, hash8(res_type), hash8(mol_id), bx, by, bz, rx, ry, rz]`` of the
    bonds passing the filters - reference ``ordering_functions.py:14-41``.  The bond vectors and
    midpoints (rows 3-8, ``lammpack_types.py:587-599``) come from the ``bond_build`` kernel."""
    fa = _arrays(config)
    mask = fa.bond_mask(allowedBondTypes, allowedResTypes, allowedMolIds, allowedBondNums)
    sel = np.flatnonzero(mask)
    engine.require_cuda()
    vec, ctr = engine.bond_build(fa.positions, fa.bond_atoms[sel], fa.box)
    dev = torch.empty((9, len(sel)), dtype=torch.float64, device=vec.device)
    dev[0] = torch.as_tensor(fa.bond_num[sel].astype(np.float64)).to(vec.device)
    dev[1] = torch.as_tensor(fa.bond_res_hash()[sel].astype(np.float64)).to(vec.device)
    dev[2] = torch.as_tensor(fa.bond_mol_hash()[sel].astype(np.float64)).to(vec.device)
    dev[3:6] = vec
    dev[6:9] = ctr
    out = dev.cpu().numpy().view(DeviceBackedArray)
    out._dev = dev
    return out


def createNumpyAtomsArrayFromConfig(config, allowedAtomTypes=[], allowedResTypes=[], allowedMolIds=[],
                                    allowedAtomNums=[]):
    """(6, N) float64 array ``[num, hash8(res_type), hash8(mol_id), x, y, z]`` - reference
    ``ordering_functions.py:44-63`` (its py3 ``map`` bug is not reproduced: the intended array is
    returned).  Pure selection, no arithmetic."""
    fa = _arrays(config)
    m = np.ones(fa.n_atoms, bool)
    for values, allowed in ((fa.atom_type, allowedAtomTypes), (fa.atom_res, allowedResTypes),
                            (fa.atom_mol, allowedMolIds), (fa.atom_num.tolist(), allowedAtomNums)):
        if len(allowed) != 0:
            s = set(allowed)
            m &= np.fromiter((v in s for v in values), bool, fa.n_atoms)
    sel = np.flatnonzero(m)
    out = np.empty((6, len(sel)), np.float64)
    out[0] = fa.atom_num[sel]
    out[1] = fa.atom_res_hash()[sel]
    out[2] = fa.atom_mol_hash()[sel]
    out[3:6] = fa.positions[sel].T
    return out.view(DeviceBackedArray)


# --------------------------------------------------------------------------------------- a1 on the host
def _bond_vector(p1, p2, box):
    """``vector_and_center_pbc_trim`` (``lammpack_types.py:587-599``) for ONE bond, scalar host
    arithmetic in the reference's operation order (used for the reference bond of the
    per-reference call, ``ordering_functions.py:92``)."""
    vec, ctr = [], []
    for a in range(3):
        L = float(box[a])
        x = p2[a] - p1[a]
        xc = (p2[a] + p1[a]) / 2.
        xt = np.remainder(x + L / 2., L) - L / 2.
        vec.append(np.float64(xt))
        ctr.append(np.float64(xc + (xt - x) / 2.))
    return vec, ctr


def _box3(config):
    if isinstance(config, FrameArrays):
        return [float(v) for v in config.box]
    b = config.box()
    return [float(b.x), float(b.y), float(b.z)]


_sr_ws = {}


def _single_ref_ws(device):
    key = str(device)
    if key not in _sr_ws:
        n = _lib.lib().mouse_single_ref_workspace_bytes()
        _sr_ws[key] = (torch.zeros(n + 256, dtype=torch.uint8, device=device),
                       torch.zeros(2, dtype=torch.float64, device=device))
    return _sr_ws[key]


# --------------------------------------------------------------------------------------- a4
def calculateAveCosSqForReference(config, refBond, npBonds, rcut, excludeSelf=True, sameMolecule=True):
    """Mean cos^2 between ``refBond`` and every column of ``npBonds`` within ``rcut`` -
    reference ``ordering_functions.py:91-119``; raises ``NameError`` exactly where it does."""
    box = _box3(config)
    p1 = (refBond.atom1.pos.x, refBond.atom1.pos.y, refBond.atom1.pos.z)
    p2 = (refBond.atom2.pos.x, refBond.atom2.pos.y, refBond.atom2.pos.z)
    b, c = _bond_vector(p1, p2, box)
    if b[0] != 0. or b[1] != 0. or b[2] != 0.:
        if rcut >= 0.:
            rc2 = rcut**2
        else:
            rc2 = 0.
        dev = _device_copy(npBonds)
        n = int(dev.shape[1])
        mol_hash = hash8(refBond.atom1.mol_id)
        if not sameMolecule:
            if n and bool((dev[2] == float(hash8(''))).any().item()):
                raise NameError("We should omit the atoms belonging to the same molecules, but some mol_ids are empty")
        ws, out = _single_ref_ws(dev.device)
        ref = (C.c_double * 8)(float(b[0]), float(b[1]), float(b[2]), float(c[0]), float(c[1]), float(c[2]),
                               float(refBond.num), float(mol_hash))
        flags = (_lib.MOUSE_SAME_MOLECULE if sameMolecule else 0) | (0 if excludeSelf else _lib.MOUSE_INCLUDE_SELF)
        _lib.check(_lib.lib().mouse_p2_single_ref(C.c_void_p(dev.data_ptr()), n, ref, (C.c_double * 3)(*box),
                                                  float(rcut), float(rc2), flags, C.c_void_p(ws.data_ptr()),
                                                  C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "mouse_p2_single_ref")
        s, cnt = out.cpu().tolist()
        if cnt > 0:
            return np.float64(s) / np.float64(cnt)
        raise NameError("No values to calculate")
    raise NameError("Zero length of reference bond " + str(refBond.num))


# --------------------------------------------------------------------------------------- a6
def calculateRdfForReference(config, refAtom, npAtoms, rmin=0., rmax=-1., nbin=100, excludeSelf=True,
                             sameMolecule=True):
    """``np.histogram`` of the distances between ``refAtom`` and the columns of ``npAtoms`` -
    reference ``ordering_functions.py:66-88``.  Returns ``(counts int64[nbin], edges float64[nbin+1])``."""
    box = _box3(config)
    if rmax < 0.:
        rmax = min(box) / 2.
    dev = _device_copy(npAtoms)
    n = int(dev.shape[1]) if dev.dim() == 2 else 0
    if not sameMolecule and n and bool((dev[2] == float(hash8(''))).any().item()):
        raise NameError("We should omit the atoms belonging to the same molecules, but some mol_ids are empty")
    edges = np.ascontiguousarray(np.histogram_bin_edges(np.zeros(0), bins=nbin, range=(rmin, rmax)), np.float64)
    nbin = len(edges) - 1
    ws, _ = _single_ref_ws(dev.device if n else torch.device("cuda"))
    hist = torch.zeros(nbin + 1, dtype=torch.int64, device=ws.device)
    if n:
        ref = (C.c_double * 5)(float(refAtom.pos.x), float(refAtom.pos.y), float(refAtom.pos.z),
                               float(refAtom.num) if refAtom.num is not None else -1.,
                               float(hash8(refAtom.mol_id)))
        flags = (_lib.MOUSE_SAME_MOLECULE if sameMolecule else 0) | (0 if excludeSelf else _lib.MOUSE_INCLUDE_SELF)
        _lib.check(_lib.lib().mouse_rdf_single_ref(C.c_void_p(dev.data_ptr()), n, ref, (C.c_double * 3)(*box), flags,
                                                   edges.ctypes.data_as(C.POINTER(C.c_double)), nbin,
                                                   C.c_void_p(ws.data_ptr()), C.c_void_p(hist.data_ptr()),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "mouse_rdf_single_ref")
    h = hist.cpu().numpy()
    if h[nbin] > 0:
        return h[:nbin].copy(), edges
    raise NameError("No values to calculate")


# --------------------------------------------------------------------------------------- a7
def CalculatePairCorrelationFunctions(config, atomtypes=[], rmin=0., rmax=0., nbin=100, sameMolecule=True,
                                      verbose=True):
    """All-pairs RDF histograms per atom-type pair - reference ``ordering_functions.py:122-177``
    with the counts accumulated onto zeros (the reference starts from uninitialised memory,
    ``:155``).  Returns ``{'r', 'v', 'data': {"ti_tj": float64[nbin]}, 'norm': {"ti_tj": float}}``."""
    fa = _arrays(config)
    box = _box3(config)
    if rmax < 0.:
        rmax = min(box) / 2.
    types = []
    for t in fa.atom_type:
        if atomtypes is None or len(atomtypes) == 0 or (t in atomtypes):
            if t not in types:
                types.append(t)
    types.append('AAAAA')
    types.sort()
    rdfs = {}
    rdfs['r'] = [rmin + (rmax - rmin) * (i + 0.5) / nbin for i in range(nbin)]
    volumes = []
    for i in range(nbin):
        r1 = rmin + (rmax - rmin) * i / nbin
        r2 = rmin + (rmax - rmin) * (i + 1) / nbin
        volumes.append(4. / 3. * np.pi * (r2**3 - r1**3))
    rdfs['v'] = volumes
    code_of = {t: k for k, t in enumerate(types)}
    tcode = np.fromiter((code_of.get(t, -1) for t in fa.atom_type), np.int32, fa.n_atoms)
    # molecule codes: dense and collision-free (0 = empty mol_id); the reference compares hash8(mol_id), whose
    # collisions (and per-process salt) would silently drop inter-molecular pairs - not reproduced
    if not sameMolecule and bool(np.any((fa.atom_mol_code() == 0) & (tcode >= 0))):
        hist = np.zeros((len(types) * (len(types) + 1) // 2, nbin), np.int64)   # every reference raises NameError
    else:
        h, _ = engine.rdf_frame(fa.positions, tcode, fa.atom_mol_code(), len(types), box, rmin,
                                rmax, nbin, same_molecule=sameMolecule)
        hist = h.cpu().numpy()
    rdfdata, norm = {}, {}
    slot = 0
    volume = box[0] * box[1] * box[2]
    for i in range(len(types)):
        for j in range(i, len(types)):
            key = str(types[i]) + '_' + str(types[j])
            rdfdata[key] = hist[slot].astype(np.float64)
            norm[key] = float(np.sum(hist[slot])) / 2. / box[0] / box[1] / box[2]
            slot += 1
    del volume
    rdfs['data'] = rdfdata
    rdfs['norm'] = norm
    if verbose:
        sys.stderr.write(str(rdfs))
    return rdfs


# --------------------------------------------------------------------------------------- a5
def CalculateOrientationOrderParameter(config, bondtypes, rmin=0., rmax=0., storeAsAtomtypes=False, mode='average',
                                       smin=-0.5, smax=1.0, sbins=75, referenceResTypes=[], sameMolecule=True):
    """Local orientational order parameter S = <3cos^2(theta)-1>/2 over neighbouring bond pairs -
    reference ``ordering_functions.py:180-227``.  ``rmin`` is accepted and ignored, as in the
    reference.  ``mode='average'`` returns ``np.float64``; ``mode='histo'`` returns the
    reference's histogram dict; ``storeAsAtomtypes`` rewrites ``atom.type`` to ``CXX``/``CmXX``."""
    fa = _arrays(config)
    box = _box3(config)
    if rmax == 0.:
        rmax = engine.half_diagonal(box)
    if rmax is None:
        raise TypeError("'>=' not supported between instances of 'NoneType' and 'float'")
    if mode == 'histo':
        s_hist = createHistogram(smin, smax, sbins)
        for i in range(sbins):
            sleft, sright = smin + i * s_hist["step"], smin + (i + 1) * s_hist["step"]
            costhetaleft, costhetaright = (2. / 3. * (sleft + 0.5))**0.5, (2. / 3. * (sright + 0.5))**0.5
            s_hist["norm"][i] = 1. / (costhetaleft + costhetaright)
    nbr_mask = fa.bond_mask(bond_types=bondtypes)
    ref_mask = fa.bond_mask(bond_types=bondtypes, res_types=referenceResTypes)
    # molecule codes: dense and collision-free (0 = empty mol_id); the reference compares hash8(mol_id), whose
    # collisions (and per-process salt) would silently drop inter-molecular pairs - not reproduced
    mol = fa.bond_mol_code()
    all_skipped = False
    if not sameMolecule and bool(np.any(nbr_mask & (mol == 0))):
        all_skipped = True   # every reference raises NameError (:107-109) and is skipped (:214)
        ref_mask = np.zeros_like(ref_mask)
    res = engine.p2_frame(fa.positions, fa.bond_atoms, mol, box, rmax, same_molecule=sameMolecule,
                          nbr_mask=nbr_mask, ref_mask=ref_mask,
                          hist_args=(smin, smax, sbins) if mode == 'histo' else None)
    if storeAsAtomtypes and not isinstance(config, FrameArrays):
        _store_as_atomtypes(config, fa, res)
    if mode == 'average':
        return res.order_parameter()
    elif mode == 'histo':
        t = res.totals_dict()
        if t["n_hist_over"] > 0:
            raise IndexError("list index out of range")
        s_hist["values"] = [float(v) for v in res.hist.cpu().tolist()]
        normHistogram(s_hist)
        return s_hist


def atom_colour_codes(fa, res=None, values=None, selected=None):
    """Per-atom colour codes of ``storeAsAtomtypes`` (``ordering_functions.py:207-220``) computed on the device: a CSR
    adjacency atom -> incident bonds (built once per topology with vectorised NumPy, bonds ascending) feeds a
    segmented reduction of the per-bond P2 (``mouse_atom_colour``).  Returns an int32 array: code >= 0 -> ``C%02d``,
    code < 0 -> ``Cm%02d`` of ``-code - 1``, INT32_MIN -> atom without a value.  Either ``res`` (a ``P2Result``: value
    = 1.5 * mean - 0.5 of the bonds with a count) or ``values`` / ``selected`` (device tensors: per-bond value and
    int32 0/1 selection, the director P2 of ``CalculateCrystallinityParameter``)."""
    key = "_csr"
    if key not in fa._cache:
        ends = fa.bond_atoms.astype(np.int64).ravel()                    # (b0.a1, b0.a2, b1.a1, ...)
        order = np.argsort(ends, kind="stable")                          # per atom: ascending bond index
        row_ptr = np.zeros(fa.n_atoms + 1, np.int64)
        np.cumsum(np.bincount(ends, minlength=fa.n_atoms), out=row_ptr[1:])
        fa._cache[key] = (row_ptr, (order // 2).astype(np.int32))
    row_ptr, bond_of = fa._cache[key]
    val, sel, direct = (res.mean, res.count, 0) if res is not None else (values, selected, 1)
    dev = val.device
    code = torch.empty(fa.n_atoms, dtype=torch.int32, device=dev)
    rp, bo = torch.as_tensor(row_ptr).to(dev), torch.as_tensor(bond_of).to(dev)
    _lib.check(_lib.lib().mouse_atom_colour(C.c_void_p(val.data_ptr()), C.c_void_p(sel.data_ptr()),
                                            C.c_void_p(rp.data_ptr()), C.c_void_p(bo.data_ptr()), fa.n_atoms, direct,
                                            C.c_void_p(code.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mouse_atom_colour")
    return code.cpu().numpy()


_COLOUR_NAMES = None


def _store_as_atomtypes(config, fa, res=None, values=None, selected=None):
    """``ordering_functions.py:207-220``: per atom, mean P2 of its reference bonds that produced a value ->
    ``atom.type = "C%02d"`` / ``"Cm%02d"``.  The reduction runs on the device (``atom_colour_codes``); the host only
    turns the integer codes into strings (one table look-up per atom)."""
    global _COLOUR_NAMES
    if _COLOUR_NAMES is None:
        _COLOUR_NAMES = {c: "C%02d" % c for c in range(0, 201)}
        _COLOUR_NAMES.update({-c - 1: "Cm%02d" % c for c in range(0, 201)})
    code = atom_colour_codes(fa, res, values, selected)
    atoms = config.atoms()
    for a in np.flatnonzero(code != np.iinfo(np.int32).min):
        c = int(code[a])
        atoms[a].type = _COLOUR_NAMES.get(c) or (("C%02d" % c) if c >= 0 else ("Cm%02d" % (-c - 1)))


# --------------------------------------------------------------------------------------- next rows
def CalculateCrystallinityParameter(config, bondtypes, reference_vector=None, storeAsAtomtypes=False, mode='average',
                                    smin=-0.5, smax=1.0, sbins=75):
    """P2 of every selected bond against one director - reference ``ordering_functions.py:230-282``.
    ``reference_vector`` (anything with ``.x .y .z``; ``None`` or zero length = the mean bond vector, ``:242-251``).
    ``mode='histo'`` returns the reference's histogram dict (``normHistogram(..., normInternalNorm=True)``) and, with
    ``storeAsAtomtypes``, rewrites ``atom.type`` to ``CXX`` / ``CmXX``.  ``mode='average'`` raises ``KeyError(0)``
    exactly like the reference does (it fabricates a ``Bond`` with ``num == 0`` that ``bond_vector_by_num`` cannot find,
    ``:257-261`` - SURVEY.md section 2)."""
    fa = _arrays(config)
    box = _box3(config)
    if mode == 'histo':
        s_hist = createHistogram(smin, smax, sbins)
        for i in range(sbins):
            sleft, sright = smin + i * s_hist["step"], smin + (i + 1) * s_hist["step"]
            costhetaleft, costhetaright = (2. / 3. * (sleft + 0.5))**0.5, (2. / 3. * (sright + 0.5))**0.5
            s_hist["norm"][i] = 1. / (costhetaleft + costhetaright)
    mask = np.ones(fa.n_bonds, bool) if bondtypes is None else np.fromiter(
        (t in bondtypes for t in fa.bond_type), bool, fa.n_bonds)              # `bond.type in bondtypes` (:241)
    vec, _ = engine.bond_build(fa.positions, fa.bond_atoms, box)                # K1 on the device, (3, N_b)
    rv = (0., 0., 0.) if reference_vector is None else (float(reference_vector.x), float(reference_vector.y),
                                                        float(reference_vector.z))
    if rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2] == 0.:
        # ave_b = ((b_1 + b_2) + b_3) + ... scaled by 1/i_b (:243-249): a running sum in bond order
        sel = vec[:, torch.as_tensor(mask, device=vec.device)].cpu().numpy()
        if sel.shape[1] > 0:
            tot = np.cumsum(sel, axis=1)[:, -1]
            k = 1. / float(sel.shape[1])
            rv = (float(tot[0] * k), float(tot[1] * k), float(tot[2] * k))
    if mode == 'average':
        raise KeyError(0)
    elif mode == 'histo':
        dev = vec.device
        out_s = torch.empty(fa.n_bonds, dtype=torch.float64, device=dev)
        hist = torch.zeros(int(sbins) + 1, dtype=torch.int64, device=dev)
        m = torch.as_tensor(mask.astype(np.uint8), device=dev)
        _lib.check(_lib.lib().mouse_director_p2(engine._ptr(vec), engine._ptr(m), fa.n_bonds, (C.c_double * 3)(*rv),
                                                engine._ptr(out_s), engine._ptr(hist), float(smin), float(smax),
                                                int(sbins), engine._stream()), "mouse_director_p2")
        h = hist.cpu().numpy()
        if h[int(sbins)] > 0:
            raise IndexError("list index out of range")
        s_hist["values"] = [float(v) for v in h[:int(sbins)]]
        if storeAsAtomtypes and not isinstance(config, FrameArrays):
            _store_as_atomtypes(config, fa, values=out_s, selected=m.to(torch.int32))      # (:274-281)
        normHistogram(s_hist, normInternalNorm=True)
        return s_hist


def CalculateComRdfs(config, rmin, rmax, nbin):
    """Radial distribution of every atom type around the centre of mass - reference ``ordering_functions.py:284-313``:
    coordinates un-cut with the image flags (``convertCoordsCutToUncut``, ``lammpack_types.py:228-232``), centre =
    ``Config.com()`` (``:234-248``: masses are used when any atom has one, but the sum is divided by the NUMBER of
    atoms with mass), one ``calculateRdfForReference(..., excludeSelf=False)`` per type.  Needs the object model
    (image flags and masses live on the atoms)."""
    from .lammpack_types import Atom, Vector3d
    atoms = config.atoms()
    box = _box3(config)
    n = len(atoms)
    pos = np.array([[a.pos.x + box[0] * a.pbc.x, a.pos.y + box[1] * a.pbc.y, a.pos.z + box[2] * a.pbc.z] for a in atoms],
                   np.float64).reshape(n, 3)
    masses = np.array([a.mass for a in atoms], np.float64)
    with_mass = int((masses > 0.).sum())
    w = masses if with_mass > 0 else np.ones(n)
    denom = with_mass if with_mass > 0 else n
    com = np.cumsum(pos * w[:, None], axis=0)[-1] * (1. / float(denom)) if n else np.zeros(3)   # running sum, atom order
    center = Atom()
    center.pos = Vector3d(float(com[0]), float(com[1]), float(com[2]))
    center.num = None
    types = [a.type for a in atoms]
    rdfs = {'r': [rmin + (rmax - rmin) * (i + 0.5) / nbin for i in range(nbin)],
            'v': [4. / 3. * np.pi * ((rmin + (rmax - rmin) * (i + 1) / nbin)**3 - (rmin + (rmax - rmin) * i / nbin)**3)
                  for i in range(nbin)]}
    nums = np.array([float(a.num) for a in atoms])
    resh = np.array([float(hash8(a.res_type)) for a in atoms])
    molh = np.array([float(hash8(a.mol_id)) for a in atoms])
    data, norm = {}, {}
    for atype in set(types):
        sel = np.fromiter((t == atype for t in types), bool, n)
        np_atoms = np.array([nums[sel], resh[sel], molh[sel], pos[sel, 0], pos[sel, 1], pos[sel, 2]])
        rdf = calculateRdfForReference(config, center, np_atoms, rmin=rmin, rmax=rmax, nbin=nbin, excludeSelf=False,
                                       sameMolecule=True)
        data[atype] = rdf[0]
        norm[atype] = np.sum(rdf[0])
    rdfs['data'], rdfs['norm'] = data, norm
    return rdfs
