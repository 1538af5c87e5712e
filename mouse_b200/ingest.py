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


def _int_tokens(values):
    """int64 column parsed from CANONICAL integer tokens -> the list of token strings (one str() per distinct value)."""
    u, inv = np.unique(values, return_inverse=True)
    return np.array([str(v) for v in u.tolist()], dtype=object)[inv].tolist()


def _header_items(text):
    """Counts and box of the header lines of a data file (the rules of the line loop in read_lmp_data_arrays)."""
    natom = nbond = 0
    box = np.zeros(3)
    for ln in text.split("\n"):
        w = ln.split()
        if not w:
            continue
        if len(w) >= 4 and w[2] in ("xlo", "ylo", "zlo") and w[3] == w[2][0] + "hi":
            box["xyz".index(w[2][0])] = abs(float(w[1]) - float(w[0]))
        elif len(w) >= 2 and w[1] == "atoms":
            natom = int(w[0])
        elif len(w) >= 2 and w[1] == "bonds":
            nbond = int(w[0])
    return natom, nbond, box


def _section_start(data, k):
    """``k`` = offset of the newline in front of a section keyword: the offset of the section's first data line (the
    keyword line and ONE more line - the blank one - are skipped, like the line loop does), or -1."""
    e = data.find(b"\n", k + 1)
    if e < 0:
        return -1
    e = data.find(b"\n", e + 1)
    return -1 if e < 0 else e + 1


def _read_lmp_data_fast(path, bonds, threads):
    """The file as LAMMPS writes it - header, ``Atoms`` (then possibly ``Velocities``), ``Bonds``, every section a table
    of canonical integers and decimal numbers - read with ``mouse_parse_table`` (threaded, in place).  Returns None for
    anything else (type labels, ragged or blank lines inside a section, sections out of order, header-like lines
    elsewhere): the line-by-line reader below then decides."""
    import ctypes as C
    try:
        from . import _lib
        L = _lib.lib()
    except (ImportError, OSError):
        return None
    with open(path, "rb") as f:
        data = f.read()
    ka = data.find(b"\nAtoms")
    if ka < 0 or data.find(b"\nBonds", 0, ka) >= 0 or data.startswith(b"Atoms") or data.startswith(b"Bonds"):
        return None
    try:
        natom, nbond, box = _header_items(data[:ka + 1].decode())
    except (UnicodeDecodeError, ValueError):
        return None
    if natom <= 0:
        return None
    base = np.frombuffer(data, dtype=np.uint8)

    def table(start, nrows, ncol, cols, kinds):
        outs = [np.empty(nrows, np.int64 if kd else np.float64) for kd in kinds]
        arr_cols = (C.c_int32 * len(cols))(*cols)
        arr_kinds = (C.c_int32 * len(cols))(*kinds)
        arr_outs = (C.c_void_p * len(cols))(*[o.ctypes.data for o in outs])
        used = C.c_int64(0)
        got = L.mouse_parse_table(C.c_void_p(base.ctypes.data + start), len(data) - start, nrows, ncol, len(cols),
                                  arr_cols, arr_kinds, arr_outs, C.byref(used), threads)
        return (outs, start + used.value) if got == nrows else (None, start)

    def quiet(chunk):       # nothing the line loop would take for a header item or a section (the tables themselves
        # hold canonical numbers only and are not searched)
        return not any(key in chunk for key in (b"xlo", b"ylo", b"zlo", b" atoms", b"\tatoms", b" bonds", b"\tbonds",
                                                b"\nAtoms", b"\nBonds"))

    start = _section_start(data, ka)
    if start < 0:
        return None
    eol = data.find(b"\n", start)
    ncol = len(data[start:eol if eol >= 0 else len(data)].split())
    if ncol < 9:
        return None
    outs, after = table(start, natom, ncol, [0, 1, 2, ncol - 6, ncol - 5, ncol - 4], [1, 1, 1, 0, 0, 0])
    if outs is None:
        return None
    num, mol_i, typ_i = outs[0], outs[1], outs[2]
    pos = np.stack(outs[3:6], axis=1)
    bond_cols, tail = None, after
    kb = data.find(b"\nBonds", after - 1)
    if kb >= 0:
        if not quiet(data[after - 1:kb]):
            return None
        if nbond > 0:
            bstart = _section_start(data, kb)
            if bstart < 0:
                return None
            eol = data.find(b"\n", bstart)
            bcol = len(data[bstart:eol if eol >= 0 else len(data)].split())
            if bcol < 4:
                return None
            bond_cols, tail = table(bstart, nbond, bcol, [1, 2, 3], [1, 1, 1])
            if bond_cols is None:
                return None
        else:
            tail = kb + 1
    if not quiet(data[max(tail - 1, kb + 1):]):
        return None
    n = natom
    if bonds and bond_cols is not None:
        m = nbond
        a = np.stack([bond_cols[1], bond_cols[2]], 1)
        a.sort(axis=1)                               # atom1 = lower atom number (lammpack_lmp_functions.py:8-11)
        order = np.argsort(num, kind="stable")
        where = np.searchsorted(num[order], a.ravel())
        if np.any(where >= n) or np.any(num[order][np.minimum(where, n - 1)] != a.ravel()):
            raise KeyError("bond refers to an unknown atom number")
        ba = order[where].reshape(m, 2).astype(np.int32)
        btype = _int_tokens(bond_cols[0])
        bnum = np.arange(1, m + 1, dtype=np.int64)   # renumbered in file order (:12)
    else:
        ba, btype, bnum = np.zeros((0, 2), np.int32), [], np.zeros(0, np.int64)
    return FrameArrays(positions=pos, box=box, atom_num=num, atom_type=_int_tokens(typ_i), atom_res=[""] * n,
                       atom_mol=_int_tokens(mol_i), bond_atoms=ba, bond_num=bnum, bond_type=btype)


def read_lmp_data_arrays(path, bonds=True, fast=True, threads=None) -> FrameArrays:
    """LAMMPS data file -> ``FrameArrays`` (same content as ``FrameArrays.from_config(Config.read_lmp_data)``).
    ``fast``: files of the plain shape LAMMPS writes go through the library's threaded table parser
    (``_read_lmp_data_fast``: the same arrays and token strings, about six times sooner); everything else, and
    ``fast=False``, through the line-by-line reader below."""
    if fast:
        import os
        got = _read_lmp_data_fast(path, bonds, threads if threads else max(1, min(8, os.cpu_count() or 1)))
        if got is not None:
            return got
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


def _parse_numbers(block: bytes, count: int, threads: int = 1):
    """``count`` whitespace-separated numbers of a text block as float64 (correctly rounded, the same values
    ``float()`` gives): the C parser of ``libmouse_b200.so`` (``mouse_parse_doubles`` / ``_mt`` on ``threads`` host
    threads - ctypes releases the GIL for the call), NumPy's text parser when the library is not built.  ``None``
    when the block does not hold exactly ``count`` numbers."""
    try:
        import ctypes as C
        from . import _lib
        out = np.empty(max(count, 1), np.float64)
        if threads > 1:
            got = _lib.lib().mouse_parse_doubles_mt(block, len(block), C.c_void_p(out.ctypes.data), count, int(threads))
        else:
            got = _lib.lib().mouse_parse_doubles(block, len(block), C.c_void_p(out.ctypes.data), count)
        return out[:count] if got == count else None
    except (ImportError, OSError):
        data = np.fromstring(block.decode(), dtype=np.float64, sep=" ") if block.strip() else np.zeros(0)
        return data if data.size == count else None


def index_lmp_dump(path):
    """Byte offset of every ``ITEM: TIMESTEP`` line of a LAMMPS dump, plus the file size as the last entry
    (int64 array of F + 1 values): built with one sequential scan of the memory-mapped file, it lets every rank
    seek straight to its own frames instead of parsing the whole trajectory."""
    import mmap
    import os
    size = os.path.getsize(path)
    offs = []
    if size:
        with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
            try:                        # the library's threaded memmem scan (GIL released)
                import ctypes as C
                from . import _lib
                L = _lib.lib()
                buf = np.frombuffer(mm, dtype=np.uint8)
                try:
                    threads = max(1, min(16, os.cpu_count() or 1))
                    out = np.empty(max(16, size // 4096), np.int64)
                    got = L.mouse_index_dump(C.c_void_p(buf.ctypes.data), size, C.c_void_p(out.ctypes.data), len(out),
                                             threads)
                    if got > len(out):
                        out = np.empty(got, np.int64)
                        got = L.mouse_index_dump(C.c_void_p(buf.ctypes.data), size, C.c_void_p(out.ctypes.data),
                                                 len(out), threads)
                finally:
                    del buf
                return np.concatenate([out[:got], np.array([size], np.int64)])
            except (ImportError, OSError, AttributeError):
                pass
            k = mm.find(b"ITEM: TIMESTEP")
            while k >= 0:
                if k == 0 or mm[k - 1] == 0x0a:
                    offs.append(k)
                k = mm.find(b"ITEM: TIMESTEP", k + 14)
    return np.array(offs + [size], np.int64)


class _RowMap:
    """atom id -> row of the topology, cached between frames: dumps normally list the atoms in the same order in
    every frame, and most often in the topology's own order (then no permutation is applied at all)."""

    def __init__(self, atom_num):
        self.atom_num = np.asarray(atom_num, np.int64)
        self.order = np.argsort(self.atom_num, kind="stable")
        self.sorted_num = self.atom_num[self.order]
        self.last_ids, self.row, self.identity = None, None, False

    def rows(self, ids):
        if self.last_ids is not None and np.array_equal(ids, self.last_ids):
            return self.row, self.identity
        n = len(self.atom_num)
        where = np.searchsorted(self.sorted_num, ids)
        if np.any(where >= n) or np.any(self.sorted_num[np.minimum(where, n - 1)] != ids):
            raise KeyError("dump refers to an unknown atom number")
        self.row = self.order[where]
        self.identity = bool(np.array_equal(self.row, np.arange(n)))
        self.last_ids = ids.copy()
        return self.row, self.identity


class DeviceDumpParser:
    """The ATOMS block of a dump frame parsed ON THE GPU (``mouse_parse_dump_atoms_dev``, ``csrc/dump_parse.cu``):
    the raw text goes from the page cache into a pinned staging buffer (two of them: the next frame is read by a few
    ``pread`` threads while this one is on the GPU) and over PCIe as it is, four small kernels
    turn it into ids / positions / image flags with the host parser's arithmetic (bit-identical values), and the
    positions come back as a CUDA tensor that ``FramePipeline.submit`` takes without another host round trip.  The
    id -> topology-row permutation of the previous frame stays on the device; the ids of a frame are compared with
    the previous frame's there, and only a frame that lists the atoms in a different order costs a host lookup.
    ``parse`` returns None for a frame the device parser does not take (tokens outside the exact fast path of the
    number parser, wrong line or column counts): the caller then uses the host parser, which is authoritative."""

    max_declines = 3       # frames declined in a row after which the text is no longer shipped (17-digit dumps)

    def __init__(self, rowmap: "_RowMap", device, fd=None, read_threads=None):
        import torch
        from . import _lib
        self.torch, self.L = torch, _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DeviceDumpParser needs a CUDA device")
        self.rowmap = rowmap
        self.n = len(rowmap.atom_num)
        # with the file descriptor the text is read straight into the pinned buffer (pread on a few threads: no page
        # faults of a mapping, the GIL is released); without it the block is copied from the caller's mapping
        self.fd, self.pool = fd, None
        if read_threads is None:       # (measured on the B200 box: 9 / 18 / 30 / 37 GB/s with 1 / 2 / 4 / 8 threads)
            import os
            read_threads = max(2, min(8, (os.cpu_count() or 4) // 4))
        if fd is not None and read_threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(read_threads)
        self.read_threads = max(1, int(read_threads))
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream(device=self.device)
            self.status = torch.zeros(4, dtype=torch.int64, device=self.device)
            self.status_host = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.text = self.ws = None
        # two pinned staging buffers: while the GPU works on the frame in one of them, the read threads fill the other
        # with the next frame (``next_range``, set by the iterator: the byte range of the frame after this one)
        self.stages, self.cur = [None, None], 0
        self.next_range, self.ahead = None, None
        self.ids = [torch.empty(self.n, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.have_ids = False          # ids[0] holds the previous frame's ids (line order)
        self.row_dev = None            # int32 permutation line -> topology row (None: identity)
        self.frames_device = self.frames_host = 0
        self.declined_in_a_row = 0     # a dump written with 17 digits per number never takes the device path: stop
                                       # shipping its text after three frames

    def _room(self, nbytes):
        torch = self.torch
        if self.text is None or self.text.numel() < nbytes:
            self.text = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        need = int(self.L.mouse_parse_dump_ws_bytes(nbytes, self.n))
        if self.ws is None or self.ws.numel() < need:
            self.ws = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=self.device)

    def _stage(self, i, nbytes):
        st = self.stages[i]
        if st is None or st[0].numel() < nbytes:
            t = self.torch.empty(int(nbytes * 1.25) + 4096, dtype=self.torch.uint8).pin_memory()
            st = self.stages[i] = (t, t.numpy())
        return st

    def _read(self, i, buf, b0, nbytes, wait=True):
        """File bytes [b0, b0 + nbytes) -> staging buffer i; wait=False returns the futures of the read threads."""
        import os
        view_np = self._stage(i, nbytes)[1]
        if self.fd is None:
            np.copyto(view_np[:nbytes], buf[b0:b0 + nbytes])
            return []
        view = memoryview(view_np)

        def chunk(a, b):
            while a < b:
                got = os.preadv(self.fd, [view[a:b]], b0 + a)
                if got <= 0:
                    raise IOError("short read of the dump file")
                a += got

        parts = self.read_threads if self.pool is not None and nbytes >= (1 << 20) else 1
        if self.pool is None:
            chunk(0, nbytes)
            return []
        cuts = [nbytes * k // parts for k in range(parts + 1)]
        futs = [self.pool.submit(chunk, cuts[k], cuts[k + 1]) for k in range(parts)]
        if wait:
            for fut in futs:
                fut.result()
            return []
        return futs

    def _prefetch(self):
        if self.next_range is None or self.pool is None or self.ahead is not None:
            return
        start, stop = self.next_range
        self.next_range = None
        if stop > start:
            i = 1 - self.cur
            self.ahead = (int(start), int(stop), i, self._read(i, None, int(start), int(stop - start), wait=False))

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
            self.pool = None

    def _run(self, nbytes, ncol, col_id, cxyz, cimg, mode, lo, hi, pos, img, expect):
        import ctypes as C
        from . import _lib
        vp = C.c_void_p
        rc = self.L.mouse_parse_dump_atoms_dev(
            vp(self.text.data_ptr()), nbytes, self.n, ncol, col_id, cxyz, cimg, mode, vp(lo.ctypes.data),
            vp(hi.ctypes.data), vp(self.row_dev.data_ptr()) if self.row_dev is not None else None,
            vp(self.ids[0].data_ptr()) if expect else None, vp(self.ids[1].data_ptr()), vp(pos.data_ptr()),
            vp(img.data_ptr()) if img is not None else None, vp(self.ws.data_ptr()), self.ws.numel(),
            vp(self.status.data_ptr()), vp(self.stream.cuda_stream))
        _lib.check(rc, "mouse_parse_dump_atoms_dev")
        self.status_host.copy_(self.status, non_blocking=True)
        self.stream.synchronize()
        return [int(v) for v in self.status_host]

    def parse(self, buf, b0, stop, keys, ci, coords_type, lo, hi):
        import ctypes as C
        torch = self.torch
        nbytes = int(stop - b0)
        if nbytes <= 0 or nbytes >= 2 ** 31 or self.declined_in_a_row >= self.max_declines:
            self.frames_host += 1
            self.next_range = None
            return None
        has_img = all(kk in keys for kk in ("ix", "iy", "iz"))
        cxyz = (C.c_int32 * 3)(*ci)
        cimg = (C.c_int32 * 3)(*[keys.index(kk) for kk in ("ix", "iy", "iz")]) if has_img else None
        mode = {"scaled": 0, "cut": 1, "uncut": 2}[coords_type]
        lo = np.ascontiguousarray(lo, np.float64)
        hi = np.ascontiguousarray(hi, np.float64)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._room(nbytes)
            off = 0
            if self.ahead is not None and self.ahead[0] <= b0 and self.ahead[1] == stop:
                start, _, self.cur, futs = self.ahead        # read while the previous frame was on the GPU
                self.ahead = None
                for fut in futs:
                    fut.result()
                off = int(b0 - start)
            else:
                if self.ahead is not None:                   # (a range nobody asked for: let it finish, drop it)
                    for fut in self.ahead[3]:
                        fut.result()
                    self.ahead = None
                self._read(self.cur, buf, b0, nbytes)         # page cache -> pinned, the only host pass
            self.text[:nbytes].copy_(self.stages[self.cur][0][off:off + nbytes], non_blocking=True)
            self._prefetch()
            pos = torch.empty((self.n, 3), dtype=torch.float64, device=self.device)
            img = torch.empty((self.n, 3), dtype=torch.int32, device=self.device) if has_img else None
            args = (nbytes, len(keys), keys.index("id"), cxyz, cimg, mode, lo, hi, pos, img)
            rows, bad, slow, changed = self._run(*args, expect=self.have_ids)
            if rows != self.n or bad != 2 ** 63 - 1 or slow:
                self.frames_host += 1
                self.declined_in_a_row += 1
                return None
            self.declined_in_a_row = 0
            if changed or not self.have_ids:
                # first frame, or the atoms are listed in another order than in the previous frame: look the ids up
                # on the host once, keep the permutation on the device
                row, identity = self.rowmap.rows(self.ids[1].cpu().numpy())
                old = self.row_dev
                self.row_dev = None if identity else torch.from_numpy(row.astype(np.int32)).to(self.device)
                if not (identity and old is None):                     # the first pass scattered with the old map
                    rows, bad, slow, _ = self._run(*args, expect=False)
            self.ids.reverse()          # this frame's ids become "the previous frame's"
            self.have_ids = True
        self.frames_device += 1
        return pos, img


def _parse_dump_header(head: str, n: int):
    lines = head.split("\n")
    timestep, natom, lo, hi = 0, n, np.zeros(3), np.zeros(3)
    j = 0
    while j < len(lines):
        item = lines[j][6:].strip() if lines[j].startswith("ITEM:") else ""
        if item == "TIMESTEP":
            timestep = int(lines[j + 1])
            j += 1
        elif item == "NUMBER OF ATOMS":
            natom = int(lines[j + 1])
            j += 1
            if natom != n:
                raise NameError("Error reading extra frame for a configuration with %d from a dump file with %d "
                                "atoms\n" % (n, natom))
        elif item[:10] == "BOX BOUNDS":
            for a in range(3):
                w = lines[j + 1 + a].split()
                lo[a], hi[a] = float(w[0]), float(w[1])
            j += 3
        j += 1
    return timestep, natom, lo, hi


def _parse_dump_frame(buf, start: int, stop: int, rowmap: _RowMap, coords_type: str, threads: int, device_parser=None):
    """One frame = the bytes [start, stop) of ``buf`` (the memory-mapped dump as a uint8 array), from its ``ITEM:
    TIMESTEP`` line up to the next one -> (timestep, box, centre, pos, img).  The ATOMS block goes to
    ``mouse_parse_dump_atoms`` in place (no copy of the text, ids / coordinates / image flags written directly,
    the reference's coordinate transform applied in C); blocks that function cannot take (non-numeric columns,
    library not built) go through the token-by-token path.  With a ``device_parser`` the block is parsed on the GPU
    and ``pos`` / ``img`` are CUDA tensors; frames it declines take the host path (NumPy arrays)."""
    n = len(rowmap.atom_num)
    head_end = min(stop, start + 4096)
    head_bytes = bytes(buf[start:head_end])
    k = head_bytes.find(b"ITEM: ATOMS")
    while k < 0 and head_end < stop:                  # (a header longer than 4 KiB: keep looking)
        head_end = min(stop, head_end + 65536)
        head_bytes = bytes(buf[start:head_end])
        k = head_bytes.find(b"ITEM: ATOMS")
    if k < 0:
        raise ValueError("dump frame without an ATOMS section")
    eol = head_bytes.find(b"\n", k)
    if eol < 0:
        raise ValueError("dump frame with a truncated ATOMS header")
    keys = head_bytes[k:eol].decode().split()[2:]
    timestep, natom, lo, hi = _parse_dump_header(head_bytes[:k].decode(), n)
    b0 = start + eol + 1
    ci = [keys.index(kk) for kk in _COORD_KEYS[coords_type]]
    has_img = all(kk in keys for kk in ("ix", "iy", "iz"))
    if device_parser is not None and natom == n:
        got = device_parser.parse(buf, b0, stop, keys, ci, coords_type, lo, hi)
        if got is not None:
            return timestep, hi - lo, (lo + hi) / 2., got[0], got[1]
    ids = xyz = iv = None
    try:
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        ids = np.empty(natom, np.int64)
        xyz = np.empty((natom, 3), np.float64)
        iv = np.empty((natom, 3), np.int32) if has_img else None
        cxyz = (C.c_int32 * 3)(*ci)
        cimg = (C.c_int32 * 3)(*[keys.index(kk) for kk in ("ix", "iy", "iz")]) if has_img else None
        mode = {"scaled": 0, "cut": 1, "uncut": 2}[coords_type]
        got = L.mouse_parse_dump_atoms(C.c_void_p(buf.ctypes.data + b0), stop - b0, natom, len(keys), keys.index("id"),
                                       cxyz, cimg, mode, C.c_void_p(lo.ctypes.data), C.c_void_p(hi.ctypes.data),
                                       C.c_void_p(ids.ctypes.data), C.c_void_p(xyz.ctypes.data),
                                       C.c_void_p(iv.ctypes.data) if has_img else None, int(threads))
        if got != natom:
            ids = None
    except (ImportError, OSError):
        ids = None
    if ids is None:                         # token by token (what the reference does, one float() per token)
        block = bytes(buf[b0:stop])
        data = _parse_numbers(block, natom * len(keys), threads)
        if data is None:
            rows = [ln.split() for ln in block.decode().split("\n") if ln.strip()][:natom]
            data = np.array(rows, dtype=np.float64)
        data = data.reshape(natom, len(keys))
        ids = data[:, keys.index("id")].astype(np.int64)
        c = data[:, ci]
        if coords_type == "scaled":
            xyz = lo + (hi - lo) * c
        elif coords_type == "cut":
            xyz = lo + c
        else:
            xyz = np.array(c)
        iv = data[:, [keys.index("ix"), keys.index("iy"), keys.index("iz")]].astype(np.int32) if has_img else None
    row, identity = rowmap.rows(ids)
    if identity:
        pos, img = np.ascontiguousarray(xyz), iv
    else:
        pos = np.empty((n, 3), np.float64)
        pos[row] = xyz
        img = None
        if has_img:
            img = np.zeros((n, 3), np.int32)
            img[row] = iv
    return timestep, hi - lo, (lo + hi) / 2., pos, img


def iter_lmp_dump_frames(path, atom_num, coords_type="scaled", frames=None, index=None, threads=1, device=None,
                         stats=None):
    """Yield ``(timestep, box[3], box_center[3], positions[N,3], images[N,3] or None)`` per frame of a LAMMPS dump,
    with the rows ordered like ``atom_num`` (the topology's atom numbers).  ``frames`` = (lo, hi) restricts the
    iteration to that range of frames: with the byte-offset ``index`` (``index_lmp_dump``; built here when not
    given) the reader seeks to frame ``lo`` and never touches the others - this is how the ranks of a sharded
    trajectory each read their own block.  ``threads``: host threads of the number parser.  One frame is in memory
    at a time.  ``device`` (a CUDA device): the ATOMS blocks are parsed on the GPU (``DeviceDumpParser``) and
    positions / images are CUDA tensors there - except for frames the device parser declines, which come from the host
    parser as NumPy arrays with the same values; ``stats`` (a dict) receives the count of each kind."""
    if coords_type not in _COORD_KEYS:
        raise ValueError("coords_type must be scaled, cut or uncut")
    if index is None:
        index = index_lmp_dump(path)
    nframes = len(index) - 1
    lo_f, hi_f = (0, nframes) if frames is None else (max(0, int(frames[0])), min(nframes, int(frames[1])))
    rowmap = _RowMap(atom_num)
    if hi_f <= lo_f:
        return
    dev_parser = None
    import mmap
    with open(path, "rb") as f:
        if device is not None:
            dev_parser = DeviceDumpParser(rowmap, device, fd=f.fileno())
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        buf = np.frombuffer(mm, dtype=np.uint8)       # the parser reads the mapped text in place
        try:
            for k in range(lo_f, hi_f):
                if dev_parser is not None:
                    dev_parser.next_range = (int(index[k + 1]), int(index[k + 2])) if k + 1 < hi_f else None
                yield _parse_dump_frame(buf, int(index[k]), int(index[k + 1]), rowmap, coords_type, threads, dev_parser)
        finally:
            if dev_parser is not None:
                dev_parser.close()
                if stats is not None:
                    stats["frames_device"] = dev_parser.frames_device
                    stats["frames_host"] = dev_parser.frames_host
            del buf
            try:
                mm.close()
            except BufferError:      # a traceback still holds a view of the mapping: the collector unmaps it later
                pass


def count_lmp_dump_frames(path, index=None):
    return len(index_lmp_dump(path) if index is None else index) - 1


def read_lmp_dump_arrays(path, atom_num, coords_type="scaled", frames=None, index=None, threads=1):
    """Frames of a dump (all, or the range ``frames``): ``(timesteps[F], boxes[F,3], positions[F,N,3])`` - the
    trajectory tensors of ``mouse_b200.trajectory``.  Materialises the frames: use ``iter_lmp_dump_frames`` to
    stream."""
    ts, boxes, pos = [], [], []
    for t, box, _, p, _ in iter_lmp_dump_frames(path, atom_num, coords_type, frames, index, threads):
        ts.append(t)
        boxes.append(box.copy())
        pos.append(p)
    f = len(ts)
    return (np.array(ts, np.int64), np.array(boxes, np.float64).reshape(f, 3),
            np.array(pos, np.float64).reshape(f, len(atom_num), 3))
