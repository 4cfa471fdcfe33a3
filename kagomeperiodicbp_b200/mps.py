"""Host-side MPS container returned by the accelerated path.  Field-compatible with the reference's
``bmpslib.mps`` (src/libs/bmpslib.py:214-232: ``A``, ``Corder``, ``Ps``, ``nr_mantissa``, ``nr_exp``, ``N``,
``mtype``) so callers that index ``.A[i]``, copy, pickle or read the (mantissa, exp10) scale keep
working.  It holds data only -- all numerics run on the device.
"""
from __future__ import annotations

import math

import numpy as np


class MPS:
    def __init__(self, N: int, mtype: str = "MPS"):
        self.N = N
        self.A: list = [None] * N
        self.Corder: list = [None] * N
        self.Ps = [1] * N
        self.nr_mantissa = 1.0
        self.nr_exp = 0
        self.mtype = mtype

    def set_site(self, mat, i, Corder=None):
        self.A[i] = np.array(mat, copy=True)
        self.Corder[i] = Corder

    def copy(self, full: bool = False, mode: str | None = None) -> "MPS":
        new = MPS(self.N, self.mtype)
        new.nr_mantissa, new.nr_exp = self.nr_mantissa, self.nr_exp
        new.Corder = list(self.Corder)
        new.Ps = list(self.Ps)
        deep = full or (mode is not None and mode != "list")
        new.A = [a.copy() for a in self.A] if deep else list(self.A)
        return new

    def overall_factor(self):
        return self.nr_mantissa * 10 ** self.nr_exp

    def reset_nr(self):
        self.nr_mantissa = 1
        self.nr_exp = 0

    def set_log_scale(self, ln_scale: float):
        """store exp(ln_scale) as (mantissa in [1, 10), exp10) -- what update_A0_norm accumulates
        (src/libs/bmpslib.py:359-375)."""
        l10 = ln_scale / math.log(10.0)
        e = math.floor(l10)
        self.nr_exp = int(e)
        self.nr_mantissa = 10.0 ** (l10 - e)

    def reduceDiter(self, maxD, nr_bulk=False, max_iter=10, err=1e-6):
        """QR-only iterative compression on the device, in place (src/libs/bmpslib.py:989-1364; see reduce_iter.py)."""
        from . import reduce_iter
        reduce_iter.reduceDiter(reduce_iter.backend(), self, maxD, nr_bulk=nr_bulk, max_iter=max_iter, err=err)

    def mps_shape(self) -> str:
        return " ".join(f"A_{i}{tuple(a.shape)}" for i, a in enumerate(self.A))

    def maxD(self) -> int:
        return max([self.A[0].shape[0]] + [a.shape[2] for a in self.A])

    @staticmethod
    def from_sites(sites, Corder=None, ln_scale: float = 0.0) -> "MPS":
        m = MPS(len(sites))
        m.A = [np.array(s, dtype=np.complex128) for s in sites]
        if Corder is not None:
            m.Corder = list(Corder)
        m.set_log_scale(ln_scale)
        return m
