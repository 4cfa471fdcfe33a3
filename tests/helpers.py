import os

import numpy as np

from oracle import mps_np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIDES = ("D", "DR", "UR", "U", "UL", "DL")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def to_oracle_mps(m):
    r = mps_np.MPS(m.N)
    r.A = list(m.A)
    r.nr_mantissa, r.nr_exp = m.nr_mantissa, m.nr_exp
    return r


def golden_mps(g, side):
    sites = []
    k = 0
    while f"{side}_site{k}" in g:
        sites.append(g[f"{side}_site{k}"])
        k += 1
    r = mps_np.MPS(len(sites))
    r.A = sites
    r.nr_mantissa, r.nr_exp = float(g[f"{side}_nr"][0]), int(g[f"{side}_nr"][1])
    return r


def dense_rel_diff(a_mps, b_mps, with_factor=True):
    """relative L2 distance of the densely contracted MPSs up to a global phase (the SVD/QR gauge freedom
    cancels inside the contraction; a global phase can survive in the scalar bookkeeping)."""
    a = mps_np.mps_to_dense(a_mps, with_factor)
    b = mps_np.mps_to_dense(b_mps, with_factor)
    ph = np.vdot(a, b)
    ph = ph / abs(ph)
    return float(np.linalg.norm(a * ph - b) / np.linalg.norm(a))


def _unit_sites(m):
    """the same state with every site scaled to unit Frobenius norm (long chains under- / overflow otherwise)."""
    r = mps_np.MPS(m.N)
    r.A = [np.asarray(a) / np.linalg.norm(a) for a in m.A]
    return r


def overlap_defect(a_mps, b_mps):
    a_mps, b_mps = _unit_sites(a_mps), _unit_sites(b_mps)
    ab = mps_np.mps_inner_product(a_mps, b_mps, True)
    aa = mps_np.mps_inner_product(a_mps, a_mps, True)
    bb = mps_np.mps_inner_product(b_mps, b_mps, True)
    return float(1 - abs(ab) / np.sqrt(abs(aa) * abs(bb)))
