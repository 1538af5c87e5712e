"""Frame ingestion straight into arrays: LAMMPS data and LAMMPS dump files become the
tensors the kernels consume WITHOUT building per-atom Python objects.

Semantics follow the reference's readers line by line:

* data file - ``Config.read_lmp_data`` (``lammpack_types.py:400-464``) with ``AtomFromLmpDataLine`` /
  ``BondFromLmpDataLine`` (``lammpack_lmp_functions.py:3-33``): header counts ``N atoms`` / ``M bonds``; box edge
  = ``abs(hi - lo)`` of the ``xlo xhi`` lines (tilt factors are ignored, as in the reference); an atom line is
  ``id mol type [q] x y z ix iy iz`` with the coordinates and image flags taken from the END of the line; bond
  atoms are sorted ascending (atom1 = lower atom number) and bonds are renumbered 1.. in file order; the
  ``Atoms`` section has to precede ``Bonds``.
* dump file - ``Config.return_extra_frames_from_lmp_dump`` (``lammpack_types.py:344-398``): per ``ITEM:
  TIMESTEP`` a frame; ``BOX BOUNDS`` gives box edge ``hi - lo`` and the origin; ``ATOMS`` columns are looked up
  by keyword; ``coords_type`` "scaled" -> ``lo + (hi - lo) * xs``, "cut" -> ``lo + xc``, "uncut" -> ``xu``;
  positions are assigned by atom ``id``; the atom count must equal the topology's.

The parsing itself is NumPy block conversion (one ``split`` per section instead of one per line).
"""
from __future__ import annotations

import itertools

import numpy as np

from .frames import FrameArrays


def _block(lines, ncol_min):
    """Whitespace-separated block of a LAMMPS section -> per-COLUMN token lists ``cols[j][row]``, addressed from
    the front (``cols[j]``) or from the end of the line (``cols[j - ncol]``).  One ``str.split`` over the whole
    block when every line has the same number of tokens (every file LAMMPS writes); ragged blocks fall back to
    one split per line and are addressed per row through ``_RaggedCols``."""
    if not lines:
        return _Cols([], 0, 0)
    ncol = len(lines[0].split())
    if ncol < ncol_min:
        raise ValueError("short line in LAMMPS section: %r" % lines[0])
    tokens = "\n".join(lines).split()
    if len(tokens) == ncol * len(lines) and all(len(ln.split()) == ncol for ln in lines[::61]) \
            and len(lines[-1].split()) == ncol:
        # same total, same width on the first, the last and every 61st line: a block LAMMPS wrote
        return _Cols(tokens, ncol, len(lines))
    rows = [ln.split() for ln in lines]
    for r in rows:
        if len(r) < ncol_min:
            raise ValueError("short line in LAMMPS section: %r" % " ".join(r))
    return _RaggedCols(rows)


class _Cols:
    def __init__(self, tokens, ncol, nrow):
        self.tokens, self.ncol, self.nrow = tokens, ncol, nrow

    def __len__(self):
        return self.nrow

    def col(self, j):
        """Column ``j`` (negative: counted from the end of the line) as a list of token strings."""
        if self.ncol == 0:
            return []
        return self.tokens[j % self.ncol::self.ncol]


class _RaggedCols:
    def __init__(self, rows):
        self.rows, self.nrow = rows, len(rows)

    def __len__(self):
        return self.nrow

    def col(self, j):
        return [r[j] for r in self.rows]


def read_lmp_data_arrays(path, bonds=True) -> FrameArrays:
    """LAMMPS data file -> ``FrameArrays`` (same content as ``FrameArrays.from_config(Config.read_lmp_data)``)."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    natom = nbond = 0
    box = np.zeros(3)
    atom_rows = bond_rows = None
    k = 0
    while k < len(lines):
        w = lines[k].split()
        k += 1
        if not w:
            continue
        if w[0] == "Atoms":
            k += 1                                   # the blank line after the section header
            atom_rows = _block(lines[k:k + natom], 9)
            k += natom
        elif w[0] == "Bonds":
            if atom_rows is None:
                raise KeyError("Bonds section before Atoms (the reference reader fails the same way)")
            k += 1
            bond_rows = _block(lines[k:k + nbond], 4)
            k += nbond
        elif len(w) >= 4 and w[2] in ("xlo", "ylo", "zlo") and w[3] == w[2][0] + "hi":
            box["xyz".index(w[2][0])] = abs(float(w[1]) - float(w[0]))
        elif len(w) >= 2 and w[1] == "atoms":
            natom = int(w[0])
        elif len(w) >= 2 and w[1] == "bonds":
            nbond = int(w[0])
    if atom_rows is None:
        atom_rows = _block([], 9)
    n = len(atom_rows)
    num = np.array(atom_rows.col(0), dtype=np.int64).reshape(n)
    # coordinates (and image flags) are taken from the END of the line (lammpack_lmp_functions.py:22-27)
    pos = np.empty((n, 3), np.float64)
    for a, j in enumerate((-6, -5, -4)):
        pos[:, a] = np.array(atom_rows.col(j), dtype=np.float64).reshape(n)
    mol = atom_rows.col(1)
    typ = atom_rows.col(2)
    if bonds and bond_rows is not None and len(bond_rows):
        m = len(bond_rows)
        a = np.stack([np.array(bond_rows.col(2), dtype=np.int64), np.array(bond_rows.col(3), dtype=np.int64)], 1).reshape(m, 2)
        a.sort(axis=1)                               # atom1 = lower atom number (lammpack_lmp_functions.py:8-11)
        order = np.argsort(num, kind="stable")
        where = np.searchsorted(num[order], a.ravel())
        if np.any(where >= n) or np.any(num[order][np.minimum(where, n - 1)] != a.ravel()):
            raise KeyError("bond refers to an unknown atom number")
        ba = order[where].reshape(m, 2).astype(np.int32)
        btype = bond_rows.col(1)
        bnum = np.arange(1, m + 1, dtype=np.int64)   # renumbered in file order (:12)
    else:
        ba, btype, bnum = np.zeros((0, 2), np.int32), [], np.zeros(0, np.int64)
    return FrameArrays(positions=pos, box=box, atom_num=num, atom_type=typ, atom_res=[""] * n, atom_mol=mol,
                       bond_atoms=ba, bond_num=bnum, bond_type=btype)


_COORD_KEYS = {"scaled": ("xs", "ys", "zs"), "cut": ("xc", "yc", "zc"), "uncut": ("xu", "yu", "zu")}


def _parse_numbers(block: bytes, count: int):
    """``count`` whitespace-separated numbers of a text block as float64 (correctly rounded, the same values
    ``float()`` gives): the C parser of ``libmouse_b200.so`` (``mouse_parse_doubles``), NumPy's text parser when the
    library is not built.  ``None`` when the block does not hold exactly ``count`` numbers."""
    try:
        import ctypes as C
        from . import _lib
        out = np.empty(max(count, 1), np.float64)
        got = _lib.lib().mouse_parse_doubles(block, len(block), C.c_void_p(out.ctypes.data), count)
        return out[:count] if got == count else None
    except (ImportError, OSError):
        data = np.fromstring(block.decode(), dtype=np.float64, sep=" ") if block.strip() else np.zeros(0)
        return data if data.size == count else None


def iter_lmp_dump_frames(path, atom_num, coords_type="scaled"):
    """Yield ``(timestep, box[3], box_center[3], positions[N,3], images[N,3] or None)`` per frame of a LAMMPS dump,
    with the rows ordered like ``atom_num`` (the topology's atom numbers)."""
    if coords_type not in _COORD_KEYS:
        raise ValueError("coords_type must be scaled, cut or uncut")
    atom_num = np.asarray(atom_num, np.int64)
    order = np.argsort(atom_num, kind="stable")
    sorted_num = atom_num[order]
    n = len(atom_num)
    with open(path, "rb") as f:
        timestep, natom, lo, hi = 0, n, np.zeros(3), np.zeros(3)
        while True:
            head = f.readline().decode()
            item = head[6:-1]
            if len(item) == 0:
                break
            if item == "TIMESTEP":
                timestep = int(f.readline())
            elif item == "NUMBER OF ATOMS":
                natom = int(f.readline())
                if natom != n:
                    raise NameError("Error reading extra frame for a configuration with %d from a dump file with %d "
                                    "atoms\n" % (n, natom))
            elif item[:10] == "BOX BOUNDS":
                for a in range(3):
                    w = f.readline().split()
                    lo[a], hi[a] = float(w[0]), float(w[1])
            elif item.split()[0] == "ATOMS":
                keys = item.split()[1:]
                block = b"".join(itertools.islice(f, natom))
                data = _parse_numbers(block, natom * len(keys))
                if data is None:                        # non-numeric columns or a short block: token by token
                    data = np.array([ln.split() for ln in block.decode().split("\n") if ln.strip()], dtype=np.float64)
                data = data.reshape(natom, len(keys))
                ids = data[:, keys.index("id")].astype(np.int64)
                where = np.searchsorted(sorted_num, ids)
                if np.any(where >= n) or np.any(sorted_num[np.minimum(where, n - 1)] != ids):
                    raise KeyError("dump refers to an unknown atom number")
                row = order[where]
                c = data[:, [keys.index(kk) for kk in _COORD_KEYS[coords_type]]]
                if coords_type == "scaled":
                    xyz = lo + (hi - lo) * c
                elif coords_type == "cut":
                    xyz = lo + c
                else:
                    xyz = c
                pos = np.empty((n, 3), np.float64)
                pos[row] = xyz
                img = None
                if all(kk in keys for kk in ("ix", "iy", "iz")):
                    img = np.zeros((n, 3), np.int32)
                    img[row] = data[:, [keys.index("ix"), keys.index("iy"), keys.index("iz")]].astype(np.int32)
                yield timestep, hi - lo, (lo + hi) / 2., pos, img


def read_lmp_dump_arrays(path, atom_num, coords_type="scaled"):
    """All frames of a dump: ``(timesteps[F], boxes[F,3], positions[F,N,3])`` - the trajectory tensors of
    ``mouse_b200.trajectory``."""
    ts, boxes, pos = [], [], []
    for t, box, _, p, _ in iter_lmp_dump_frames(path, atom_num, coords_type):
        ts.append(t)
        boxes.append(box.copy())
        pos.append(p)
    f = len(ts)
    return (np.array(ts, np.int64), np.array(boxes, np.float64).reshape(f, 3),
            np.array(pos, np.float64).reshape(f, len(atom_num), 3))
