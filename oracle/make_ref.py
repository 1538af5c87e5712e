#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY - stages the UNMODIFIED reference modules of the hot path under ``oracle/_ref/``.

The reference (mglagolev/mouse) is a flat directory of pure-Python modules without a build system; its path through
``ordering_functions.py`` needs eight of them (plus the radii table one of them reads at import time).  This recipe copies exactly those files, byte for byte, from
``/root/reference`` (present only in the build container) to ``oracle/_ref/`` - a git-ignored directory that is NOT
gpurun-ignored, so it travels to the GPU box like a built ``.so`` and ``bench.py --impl reference`` can time the
literal reference code there (``cpu_baseline.kind = "reference"``).  Nothing is copied into the git history, nothing
is modified; the shims the code needs under Python 3 live in ``oracle/reference_harness.py``.
Run by ``__graft_entry__.build()`` whenever ``/root/reference`` exists."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MOUSE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["ordering_functions.py", "lammpack_types.py", "lammpack_misc.py", "lammpack_lmp_functions.py",
         "lammpack_pdb_functions.py", "gro_functions.py", "clustering_functions.py", "vector3d.py",
         "radii.txt"]          # (read at import time by lammpack_types.py:565-570)


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return False
    os.makedirs(DST, exist_ok=True)
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as man:
        man.write("# unmodified copies from %s (git-ignored; see oracle/make_ref.py)\n" % SRC)
        for name in FILES:
            shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
            man.write("%s  %s\n" % (hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest(), name))
    return True


if __name__ == "__main__":
    ok = stage()
    print("oracle/_ref %s" % ("staged" if ok else "not staged: %s is absent" % SRC))
    sys.exit(0)
