"""Host-side helpers with the reference's names (``lammpack_misc.py``): file-type dispatch,
selections, the dict-based histogram and the printers used by the CLIs.  Pure host logic -
no array arithmetic lives here."""
from __future__ import annotations

import sys


def determine_data_type(filename):
    """File type from extension, else from the first line (reference ``lammpack_misc.py:8-39``)."""
    low = filename.lower()
    if low.endswith((".ent", ".pdb")):
        return "pdb"
    if low.endswith(".gro"):
        return "gro"
    with open(filename, "r") as f:
        header = f.readline()
    h = header.lower()
    if h.startswith("lammps data file") or h.startswith("lammps description"):
        return "lammps_data"
    if h[:6] in ("model ", "atom  ", "hetatm", "conect"):
        return "pdb"
    if h.startswith("item:"):
        return "lammps_dump"
    return "unknown"


def read_data_typeselect(filename, ftype=None, options={}):
    """Read a frame according to its file type (reference ``lammpack_misc.py:42-66``).  Only the
    LAMMPS-data reader is on the accelerated path; other formats raise ``NameError`` like an
    unknown type does in the reference."""
    from .lammpack_types import Config
    if ftype is None or ftype == "auto":
        ftype = determine_data_type(filename)
    if ftype == "lammps_data":
        config = Config()
        config.read_lmp_data(filename, **options.get(ftype, {}))
        return config
    raise NameError("No handler function for file type " + ftype)


def selectByAtomFields(config, selectionCriteria):
    """New Config with the atoms whose fields match, plus the bonds whose BOTH atoms survive
    (reference ``lammpack_misc.py:96-117``)."""
    from .lammpack_types import Config
    new = Config()
    new.set_box(config.box())
    new.set_box_center(config.box_center())
    atoms = config.atoms()
    for name, allowed in selectionCriteria.items():
        allowed = set(allowed)
        atoms = [a for a in atoms if getattr(a, name) in allowed]
    for a in atoms:
        new.insert_atom(a)
    kept = {a.num for a in atoms}
    for b in config.bonds():
        if b.atom1.num in kept and b.atom2.num in kept:
            new.insert_bond(b)
    return new


def select_chains(config, chainList):
    return selectByAtomFields(config, {"mol_id": chainList})


def select_rectangular(config, cutx=False, xrelmin=0., xrelmax=1., cuty=False, yrelmin=0., yrelmax=1.,
                       cutz=False, zrelmin=0., zrelmax=1.):
    """Atoms inside a sub-cell given in box-relative coordinates, and the bonds fully inside
    (reference ``lammpack_misc.py:124-144``)."""
    from .lammpack_types import Config
    new = Config()
    new.set_box(config.box())
    new.set_box_center(config.box_center())
    box = config.box()
    kept = set()
    for a in config.atoms():
        xr, yr, zr = a.pos.x / box.x, a.pos.y / box.y, a.pos.z / box.z
        if (not cutx or xrelmin <= xr <= xrelmax) and (not cuty or yrelmin <= yr <= yrelmax) and \
           (not cutz or zrelmin <= zr <= zrelmax):
            new.insert_atom(a)
            kept.add(a.num)
    for b in config.bonds():
        if b.atom1.num in kept and b.atom2.num in kept:
            new.insert_bond(b)
    return new


class Histogram(dict):
    """The reference's histogram record (``lammpack_misc.py:177-212``) - a plain dict with the keys ``min, max,
    nbins, step, values, norm`` that callers index directly - plus the operations on it as methods.  Two quirks of
    the reference are part of the contract: the bin of a value is ``int(nbins * (v - min) / (max - min))`` used as a
    Python list index (negative wraps around, ``v == max`` raises ``IndexError``), while the printed abscissa is
    ``min + step * i`` with ``step = (max - min) / (nbins - 1)``."""

    def __init__(self, lo, hi, nbins):
        if not isinstance(nbins, int):
            raise NameError("Number of bins (" + str(nbins) + ") in the histogram must be integer")
        if nbins < 1:
            raise NameError("Invalid number of bins in the histogram (" + str(nbins) + ")")
        span = hi - lo
        super().__init__(min=lo, max=hi, nbins=nbins, step=span / (nbins - 1) if nbins > 1 else span,
                         values=[0.] * nbins, norm=[0.] * nbins)

    def bin_of(self, value):
        return int(len(self["values"]) * (value - self["min"]) / (self["max"] - self["min"]))

    def add(self, value, weight=1):
        self["values"][self.bin_of(value)] += weight

    def normalise(self, norm_too=False):
        for key in ("values", "norm") if norm_too else ("values",):
            total = sum(self[key], 0.)          # left to right, like the reference's running sum
            if total > 0.:
                self[key] = [v / total for v in self[key]]

    def lines(self, norm=False):
        weights = self["norm"] if norm else [1] * self["nbins"]
        rows = ((self["min"] + self["step"] * i, self["values"][i] / w if norm else self["values"][i])
                for i, w in zip(range(self["nbins"]), weights))
        return "".join(str(x) + "\t" + str(v) + "\n" for x, v in rows)


def createHistogram(minvalue, maxvalue, nbins):
    """Reference ``lammpack_misc.py:177-183``."""
    return Histogram(minvalue, maxvalue, nbins)


def updateHistogram(value, histogram, weight=1):
    """Reference ``lammpack_misc.py:185-189``."""
    Histogram.add(histogram, value, weight)


def normHistogram(histogram, normInternalNorm=False):
    """Reference ``lammpack_misc.py:191-203``."""
    Histogram.normalise(histogram, normInternalNorm)


def printHistogram(histogram, norm=False):
    """Reference ``lammpack_misc.py:205-212``: ``"<x>\\t<value>\\n"`` per bin."""
    return Histogram.lines(histogram, norm)


def printRdfs(rdfs, norm=True, out=None):
    """Reference ``lammpack_misc.py:214-233`` / ``calculate_rdfs.py:9-28``: g = data / norm / v."""
    out = out or sys.stdout
    keys = sorted(rdfs['data'])
    out.write("#r\tvol")
    for key in keys:
        out.write("\t" + str(key))
    out.write("\n")
    for n in range(len(rdfs['r'])):
        r, v = rdfs['r'][n], rdfs['v'][n]
        out.write(str(r))
        out.write("\t" + str(v))
        for key in keys:
            if norm:
                out.write("\t" + str(rdfs['data'][key][n] / rdfs['norm'][key] / v))
            else:
                out.write("\t" + str(rdfs['data'][key][n]))
        out.write("\n")
