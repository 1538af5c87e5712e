#!/usr/bin/env python3
"""Golden PDB files written by the UNMODIFIED reference from the committed ``melt_290.data``:
``melt_290_plain.pdb`` = ``Config.write_pdb(..., hide_pbc_bonds=True)`` right after reading, and
``melt_290_coloured.pdb`` = the same after ``CalculateOrientationOrderParameter(..., storeAsAtomtypes=True)``
with the arguments of ``calculate_local_ordering.py melt_290.data --bondtypes 1 --max 1.5 --same-molecule
--out-pdb out.pdb`` (``calculate_local_ordering.py:66-74``).

    PYTHONHASHSEED=0 python tests/golden/make_golden_pdb.py        (needs /root/reference)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_harness as rh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    m = rh.load()
    data = os.path.join(OUT, "melt_290.data")
    cfg = m.lt.Config()
    cfg.read_lmp_data(data)
    cfg.write_pdb(os.path.join(OUT, "melt_290_plain.pdb"), hide_pbc_bonds=True)
    s = m.of.CalculateOrientationOrderParameter(cfg, ['1'], 0., 1.5, storeAsAtomtypes=True, sameMolecule=True)
    cfg.write_pdb(os.path.join(OUT, "melt_290_coloured.pdb"), hide_pbc_bonds=True)
    print("S =", repr(s))


if __name__ == "__main__":
    main()
