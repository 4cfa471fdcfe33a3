"""Boundary MPS whose sites are device tensors: the product-side counterpart of the reference's
``bmpslib.mps`` (src/libs/bmpslib.py:214) for the operations the block-BP path uses.  Every method
only *records* device ops into a ``Program``; shapes and canonical tags are tracked on the host so
that the truncation schedule (which sites get QR'd, RQ'd, SVD-truncated after each swallow) is the
reference's, decision for decision:

  left_canonical_QR   src/libs/bmpslib.py:553-595
  right_canonical     src/libs/bmpslib.py:688-822   (SVD where the left bond exceeds maxD, else RQ)
  reduceD (mode MC)   src/libs/bmpslib.py:868-959   (minimal window search on shapes + tags)
  update_A0_norm      src/libs/bmpslib.py:359-375   ((mantissa, exp10) kept as one log slot on device)
"""
from __future__ import annotations

import numpy as np

from .program import DT, Program

SLOT_LOGNORM, SLOT_TRUNC, SLOT_NONFINITE = 0, 1, 2


class DevMPS:
    def __init__(self, prog: Program, N: int, slot_lognorm: int = SLOT_LOGNORM, slot_trunc: int = SLOT_TRUNC):
        self.p = prog
        self.N = N
        self.A: list = [None] * N
        self.Corder: list = [None] * N
        self.slot_lognorm = slot_lognorm
        self.slot_trunc = slot_trunc

    def set_site(self, t: DT, i: int, Corder=None):
        assert t.ndim == 3
        self.A[i] = t
        self.Corder[i] = Corder

    def set_lists(self, A, Corder):
        self.A, self.Corder = list(A), list(Corder)
        self.N = len(self.A)

    def shapes(self):
        return [a.shape for a in self.A]

    def update_A0_norm(self):
        a0 = self.p.copy(self.A[0])
        self.p.normalize_(a0, self.slot_lognorm)
        self.A[0] = a0          # tag unchanged

    def left_canonical_QR(self, i0=None, i1=None):
        if self.N < 2:
            return
        i0 = 0 if i0 is None else i0
        i1 = self.N - 2 if i1 is None else min(i1, self.N - 2)
        p = self.p
        for i in range(i0, i1 + 1):
            if self.Corder[i] == "L":
                continue
            D1, d, D2 = self.A[i].shape
            Q, R = p.qr(self.A[i].reshape(D1 * d, D2))
            self.set_site(Q.reshape(D1, d, Q.shape[1]), i, "L")
            self.set_site(p.tensordot(R, self.A[i + 1], ([1], [0])), i + 1)

    def right_canonical(self, maxD=None, i0=None, i1=None, nr_bulk=False):
        if self.N < 2:
            return
        maxD = 10000000 if maxD is None else maxD
        i0 = 1 if i0 is None else i0
        i1 = self.N - 1 if i1 is None else i1
        p = self.p
        for i in range(i1, i0 - 1, -1):
            D1, d, D2 = self.A[i].shape
            M = self.A[i].reshape(D1, d * D2)
            if D1 > maxD:
                keep = min(min(D1, d * D2), maxD)
                US, Vh = p.svd_trunc(M, keep, nr_bulk, self.slot_lognorm, self.slot_trunc)
                self.set_site(Vh.reshape(keep, d, D2), i, "R")
                self.set_site(p.tensordot(self.A[i - 1], US, ([2], [0])), i - 1)
            else:
                if self.Corder[i] == "R":
                    continue
                Lm, Q = p.lq(M)
                if nr_bulk:
                    p.normalize_(Lm, self.slot_lognorm)
                self.set_site(Q.reshape(Q.shape[0], d, D2), i, "R")
                self.set_site(p.tensordot(self.A[i - 1], Lm, ([2], [0])), i - 1)
        if nr_bulk:
            self.update_A0_norm()

    def reduceD(self, maxD: int, nr_bulk=False):
        if self.N < 3:
            return
        sh = self.shapes()

        def oversized(i):
            a, b = sh[i], sh[i + 1]
            return a[2] > min(maxD, a[0] * a[1], b[2] * b[1])

        need = [i for i in range(self.N - 1) if oversized(i)]
        if not need:
            return
        iD0, iD1 = need[0], need[-1]
        i0 = 0
        for i0 in range(iD0 + 1):
            if self.Corder[i0] != "L":
                break
        self.left_canonical_QR(i0, iD1)
        i1 = self.N - 1
        for i1 in range(self.N - 1, iD1 - 1, -1):
            if self.Corder[i1] != "R":
                break
        self.right_canonical(maxD, i0=iD0 + 1, i1=i1, nr_bulk=nr_bulk)


def add_two_mps(p: Program, A: DevMPS, alpha, B: DevMPS, beta, sign_slot_beta=-1) -> DevMPS:
    """block-diagonal sum alpha*A + beta*B (src/libs/bmpslib.py:2781-2864); if ``sign_slot_beta`` >= 0 the
    coefficient beta is multiplied on the device by sign(slots[sign_slot_beta])."""
    assert A.N == B.N
    N = A.N
    out = DevMPS(p, N, A.slot_lognorm, A.slot_trunc)
    for i in range(N):
        a, b = A.A[i], B.A[i]
        DLa, d, DRa = a.shape
        DLb, d2, DRb = b.shape
        assert d == d2
        if i == 0:
            s = p.zeros((1, d, DRa + DRb))
            p.embed(s, (0, 0, 0), a, alpha)
            p.embed(s, (0, 0, DRa), b, beta, sign_slot_beta)
        elif i == N - 1:
            s = p.zeros((DLa + DLb, d, 1))
            p.embed(s, (0, 0, 0), a)
            p.embed(s, (DLa, 0, 0), b)
        else:
            s = p.zeros((DLa + DLb, d, DRa + DRb))
            p.embed(s, (0, 0, 0), a)
            p.embed(s, (DLa, 0, DRa), b)
        out.set_site(s, i)
    return out


def inner_product(p: Program, A: DevMPS, B: DevMPS) -> DT:
    """<A|B> with B conjugated, site tensors only (src/libs/bmpslib.py:2172-2211, 2288-2301).  Returns a
    1-element device tensor."""
    C = None
    for i in range(A.N):
        a, b = A.A[i], B.A[i]
        if C is None:
            assert a.shape[0] == 1 and b.shape[0] == 1
            C = p.tensordot(a.reshape(a.shape[1], a.shape[2]), b.reshape(b.shape[1], b.shape[2]), ([0], [0]), conj_b=True)
        else:
            C1 = p.tensordot(C, a, ([0], [0]))
            C = p.tensordot(C1, b, ([0, 1], [0, 1]), conj_b=True)
    assert C.size == 1
    return C
