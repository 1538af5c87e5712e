"""TEST INFRASTRUCTURE ONLY - runs the UNMODIFIED reference (mglagolev/mouse) in-process.

Works where ``/root/reference`` exists (the build container) or where ``oracle/make_ref.py`` has staged the
eight unmodified modules of the path under ``oracle/_ref/`` (git-ignored, travels to the GPU box).  Nothing under
``tests -m gpu`` or ``smoke()`` imports this module; ``bench.py`` uses it only for the CPU legs that time the
literal reference (``--impl reference``, ``cpu_baseline``).  It is used by ``tests/golden/make_golden.py`` to
produce the committed golden vectors and by the CPU test-suite (skipped when the reference is absent) to re-pin the
NumPy restatement in ``oracle/p2_oracle.py`` against the real thing.

Shims, none of which touches reference source:
  * a stub ``natsort`` module (``natsorted = sorted``): ``lammpack_pdb_functions.py:3``
    imports it at module load, the LAMMPS/tensor path never calls it;
  * ``ordering_functions.map`` shadowed with a list-returning map so that
    ``createNumpyAtomsArrayFromConfig`` (``ordering_functions.py:61-63``) works on py3;
  * RDF counts are accumulated onto zeros by calling ``calculateRdfForReference``
    per atom ourselves (``ordering_functions.py:155`` starts from uninitialised memory).
Run with ``PYTHONHASHSEED=0`` so that ``hash8`` (``ordering_functions.py:10-11``) is
reproducible.
"""
from __future__ import annotations

import builtins
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_DIR = os.environ.get("MOUSE_REFERENCE_DIR") or (
    "/root/reference" if os.path.isfile("/root/reference/ordering_functions.py") else _STAGED)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "ordering_functions.py"))


_mods = None


def load():
    """Import the reference modules (once) and return a namespace."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    if "natsort" not in sys.modules:
        stub = types.ModuleType("natsort")
        stub.natsorted = sorted
        sys.modules["natsort"] = stub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import ordering_functions as of   # noqa: E402  (reference module)
    import lammpack_types as lt       # noqa: E402
    import lammpack_misc as lm        # noqa: E402
    import vector3d as v3             # noqa: E402
    of.map = lambda f, xs: list(builtins.map(f, xs))
    _mods = types.SimpleNamespace(of=of, lt=lt, lm=lm, v3=v3)
    return _mods


def config_from_melt(melt, positions=None):
    """Build a reference ``Config`` in memory exactly as ``read_lmp_data`` would
    (``lammpack_lmp_functions.py:3-33``: atom.num 1-based, mol_id/type strings,
    bond atoms sorted ascending, bond.num = running index from 1)."""
    m = load()
    pos = melt.positions if positions is None else positions
    cfg = m.lt.Config()
    cfg.set_box(m.v3.Vector3d(float(melt.box[0]), float(melt.box[1]), float(melt.box[2])))
    mol_names = melt.mol_id_strings()
    for i in range(melt.n_atoms):
        a = m.lt.Atom()
        a.id = str(i + 1)
        a.num = i + 1
        a.mol_id = mol_names[i]
        a.type = melt.atom_type_names[melt.atom_type[i]]
        a.res_type = melt.res_type_names[melt.res[i]]
        a.res_num = int(melt.mol[i]) + 1
        a.pos = m.v3.Vector3d(float(pos[i, 0]), float(pos[i, 1]), float(pos[i, 2]))
        a.pbc = m.v3.Vector3d(int(melt.images[i, 0]), int(melt.images[i, 1]), int(melt.images[i, 2]))
        cfg.insert_atom(a)
    for b in range(melt.n_bonds):
        bond = m.lt.Bond()
        a1, a2 = sorted((int(melt.bond_atoms[b, 0]) + 1, int(melt.bond_atoms[b, 1]) + 1))
        bond.atom1 = cfg.atom_by_num(a1)
        bond.atom2 = cfg.atom_by_num(a2)
        bond.type = melt.bond_type_names[melt.bond_type[b]]
        bond.num = cfg.n_bond() + 1
        cfg.insert_bond(bond)
    return cfg


class _AverageSpy:
    """Records what ``np.ma.average`` saw in ``calculateAveCosSqForReference``
    (``ordering_functions.py:113-115``): the count and the unmasked index list."""

    def __init__(self):
        self.count = None
        self.indices = None
        self._orig = np.ma.average

    def __enter__(self):
        def spy(a, *args, **kw):
            self.count = int(a.count())
            self.indices = np.flatnonzero(~np.ma.getmaskarray(a)).astype(np.int64)
            return self._orig(a, *args, **kw)
        np.ma.average = spy
        return self

    def __exit__(self, *exc):
        np.ma.average = self._orig
        return False


def p2_per_vector(cfg, bondtypes, rcut, same_molecule=True, reference_res_types=None):
    """Literal reference, one call of ``calculateAveCosSqForReference`` per reference
    bond, mirroring the driver loop ``ordering_functions.py:196-214``.

    Returns dict: npBonds (9,N) as built by the reference, ``ref_index`` (index into the
    npBonds columns of each reference bond that produced a value), ``mean`` (per reference
    mean cos^2), ``count``, ``nbr_ptr``/``nbr_idx`` (CSR of counted neighbour columns)."""
    m = load()
    of = m.of
    np_bonds = of.createNumpyBondsArrayFromConfig(cfg, allowedBondTypes=bondtypes)
    col_of_num = {int(n): k for k, n in enumerate(np_bonds[0])}
    ref_index, mean, count, ptr, idx = [], [], [], [0], []
    skipped = []
    rrt = reference_res_types
    with _AverageSpy() as spy:
        for bond in cfg.bonds():
            if (bondtypes is None or bond.type in bondtypes or len(bondtypes) == 0) and \
               (rrt is None or (bond.atom1.res_type in rrt and bond.atom2.res_type in rrt) or len(rrt) == 0):
                try:
                    spy.count = None
                    v = of.calculateAveCosSqForReference(cfg, bond, np_bonds, rcut,
                                                         excludeSelf=True, sameMolecule=same_molecule)
                except NameError:
                    skipped.append(col_of_num[bond.num])
                    continue
                ref_index.append(col_of_num[bond.num])
                mean.append(float(v))
                count.append(spy.count)
                idx.append(spy.indices)
                ptr.append(ptr[-1] + len(spy.indices))
    return dict(npBonds=np_bonds, ref_index=np.asarray(ref_index, np.int64),
                mean=np.asarray(mean, np.float64), count=np.asarray(count, np.int64),
                nbr_ptr=np.asarray(ptr, np.int64),
                nbr_idx=(np.concatenate(idx) if idx else np.zeros(0, np.int64)),
                skipped=np.asarray(skipped, np.int64))


def rdf_counts(cfg, atomtypes, rmin, rmax, nbin, same_molecule=True):
    """Intended behaviour of ``CalculatePairCorrelationFunctions``
    (``ordering_functions.py:122-177``) with the counts accumulated onto zeros: returns
    (types, {pair_key: int64[nbin]}) using the reference's own per-atom function."""
    m = load()
    of = m.of
    types = []
    for atom in cfg.atoms():
        if atomtypes is None or len(atomtypes) == 0 or atom.type in atomtypes:
            if atom.type not in types:
                types.append(atom.type)
    types.append("AAAAA")
    types.sort()
    by_type = {t: of.createNumpyAtomsArrayFromConfig(cfg, allowedAtomTypes=[t]) for t in types}
    data = {}
    for i in range(len(types)):
        for j in range(i, len(types)):
            data[str(types[i]) + "_" + str(types[j])] = np.zeros(nbin, np.int64)
    for atom in cfg.atoms():
        if atom.type in types:
            for t2 in types:
                key = "_".join(sorted([atom.type, t2]))
                try:
                    h = of.calculateRdfForReference(cfg, atom, by_type[t2], rmin=rmin, rmax=rmax,
                                                    nbin=nbin, sameMolecule=same_molecule)
                    data[key] += h[0]
                except (NameError, IndexError):
                    pass
    return types, data
