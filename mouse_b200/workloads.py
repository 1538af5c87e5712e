"""The synthetic inputs of the BASELINE.json configurations, shared by ``bench.py`` and the
full-size parity tests so that both run exactly the same systems.  NumPy only."""
from __future__ import annotations

import numpy as np

from .melt import make_melt, tilt_melt

CONFIG2 = dict(chains=10000, beads=101, rcut=2.5)            # 1M bonds, rc = 2.5 sigma (the bench headline)
CONFIG3 = dict(chains=2500, beads=101, rcut=2.5)             # 250k-bond frames of a trajectory
CONFIG4 = dict(chains=40000, beads=100, rmin=0.0, rmax=5.0, nbin=1000)     # RDF, 4M beads
CONFIG5 = dict(chains=80000, beads=101, rcut=3.0, tilt_frac=(0.25, -0.15, 0.10))   # 8.0M bonds, selections, tilt


def config5(triclinic: bool, chains: int | None = None):
    """Config 5: a PLA-like melt with two bond types (the backbone type is the neighbour set) and three residue types
    (the references are the backbone bonds of the first two), other-molecule pairs only.  ``triclinic``: the same
    chains wrapped into a tilted cell (tilt factors = fixed fractions of the box edge).  Returns a dict with
    positions, bond_atoms, molb, box, tilt (or None), nbr_mask, ref_mask, rcut."""
    melt = make_melt(chains or CONFIG5["chains"], CONFIG5["beads"], seed=1, n_bond_types=2, n_res_types=3)
    nbr = (melt.bond_type == 0).astype(np.uint8)
    resb = melt.res[melt.bond_atoms[:, 0]]
    ref = (nbr.astype(bool) & (resb != 2)).astype(np.uint8)
    tilt, pos = None, melt.positions
    if triclinic:
        fx, fy, fz = CONFIG5["tilt_frac"]
        tilt = (fx * float(melt.box[0]), fy * float(melt.box[0]), fz * float(melt.box[1]))
        pos = tilt_melt(melt, tilt)
    return dict(melt=melt, positions=pos, bond_atoms=melt.bond_atoms, molb=melt.mol[melt.bond_atoms[:, 0]],
                box=melt.box, tilt=tilt, nbr_mask=nbr, ref_mask=ref, rcut=CONFIG5["rcut"])
