"""TEST INFRASTRUCTURE (CPU oracle) -- an INDEPENDENT numpy statement of the two-site reduced density matrix of
src/libs/ITE.py:683-756 (``rho_ij``, mps_env branch): explicit ``numpy.einsum`` index strings written from the reference's tensor-network picture, in a
different contraction order from kagomeperiodicbp_b200/ite.py (double-layer site tensor first, then the environment tensors) and
sharing no code with it.  Only tests/ may import this module.

Network (ITE.py:700-745): T_i[p, s, k1..kn] and T_j[q, s, l1..lm] share the bond s; the environment is a closed ring of
n + m tensors E[L, ket, bra, R] -- the first n attached to the legs of T_i in order, the last m to the legs of T_j --
whose bra legs carry conj(T).  rho[p, p*, q, q*] is the contraction with the physical legs left open, divided by its trace.
"""
from __future__ import annotations

import numpy as np


def _half(T, envs):
    """double-layer site tensor first (ket and bra together), then the environment tensors of its legs one by one:
    -> H[p, p*, s, s*, L_first, R_last]"""
    n = T.ndim - 2
    assert n == 3 and len(envs) == 3, "Kagome sites: one shared bond + three environment legs"
    X = np.einsum("psabc,PSABC->pPsSaAbBcC", T, np.conj(T))
    X = np.einsum("pPsSaAbBcC,LaAx->pPsSbBcCLx", X, np.asarray(envs[0]))
    X = np.einsum("pPsSbBcCLx,xbBy->pPsScCLy", X, np.asarray(envs[1]))
    return np.einsum("pPsScCLy,ycCz->pPsSLz", X, np.asarray(envs[2]))


def rho_ij_einsum(Ti, Tj, mps_env):
    Ti, Tj = np.asarray(Ti), np.asarray(Tj)
    Hi = _half(Ti, mps_env[:3])                      # [p, p*, s, s*, L0, R2]
    Hj = _half(Tj, mps_env[3:])                      # [q, q*, s, s*, L3 (= R2), R5 (= L0)]
    rho = np.einsum("pPsSLz,qQsSzL->pPqQ", Hi, Hj)
    return rho / np.einsum("iijj->", rho)
