"""Boundary MPS whose sites are device tensors: the product-side counterpart of the reference's
``bmpslib.mps`` (src/libs/bmpslib.py:214) for the operations the block-BP path uses.  Every method
only *records* device ops into a ``Program``; shapes and canonical tags are tracked on the host so
that the truncation schedule (which sites get QR'd, RQ'd, SVD-truncated after each swallow) is the
reference's, decision for decision:

  left_canonical_QR   src/libs/bmpslib.py:553-595
  right_canonical     src/libs/bmpslib.py:688-822   (SVD where the left bond exceeds maxD, else RQ)
  reduceD (mode MC)   src/libs/bmpslib.py:868-959   (minimal window search on shapes + tags)
  update_A0_norm      src/libs/bmpslib.py:359-375   ((mantissa, exp10) kept as one log slot on device)

What is NOT the reference's is how the sweeps are realised.  ``merge_T`` re-expands every swallowed
tensor into sites through *reshaped identities* (src/libs/bubblecon.py:390-483), and the reference then
runs dense (chi D^2) x (chi D^2) QR / RQ factorisations on those identities.  Here identity sites stay
symbolic:

  Id(DL, d, DR)          reshaped identity, 'L' type (DL*d == DR) or 'R' type (DL == d*DR)
  Kron('L', R, d)        R (x) I_d : an 'L' identity that absorbed the factor R[k, DL] from its left
  Kron('R', R, d)        I_d (x) R : an 'R' identity that absorbed the factor R[DR, k] from its right

QR of an 'L' identity is the identity; QR of Kron('L', R, d) is Id . (R (x) I_d), so the factor just moves on
as a small GEMM on a split index; the same holds mirrored for RQ.  And the one factorisation that is
genuinely dense -- the last QR of the left sweep followed by the truncating SVD on the next site -- is
done as ONE SVD of the two-site product (Q R X = (Q U) S V^H: same singular values, same truncated state,
same tags), so no (chi D^2)^2 QR is ever executed.  All of this only changes the gauge inside exact
factorisations; every truncation is the same Schmidt truncation as in the reference (parity-tested to
1e-10 against the oracle, which does it the literal way).
"""
from __future__ import annotations

from .program import DT, Program

SLOT_LOGNORM, SLOT_TRUNC, SLOT_NONFINITE = 0, 1, 2


class Id:
    """reshaped-identity site (src/libs/bubblecon.py:345-382)."""
    __slots__ = ("shape", "kind")

    def __init__(self, DL, d, DR):
        self.shape = (int(DL), int(d), int(DR))
        self.kind = "L" if DL * d == DR else "R"
        assert DL * d == DR or DL == d * DR


class Kron:
    """side 'L': tensor[k, p, (l, p')] = R[k, l] delta(p, p')   shape (k, d, DL*d)
       side 'R': tensor[(p', r), p, k] = delta(p, p') R[r, k]   shape (d*DR, d, k)"""
    __slots__ = ("side", "R", "d")

    def __init__(self, side, R: DT, d):
        self.side, self.R, self.d = side, R, int(d)

    @property
    def shape(self):
        if self.side == "L":
            return (self.R.shape[0], self.d, self.R.shape[1] * self.d)
        return (self.d * self.R.shape[0], self.d, self.R.shape[1])


def dense(p: Program, s) -> DT:
    if isinstance(s, DT):
        return s
    if isinstance(s, Id):
        DL, d, DR = s.shape
        n = DR if s.kind == "L" else DL
        return p.eye(n, n).reshape(DL, d, DR)
    if s.side == "L":      # R[k, DL] . eye[DL, d, DL*d]
        k, DL = s.R.shape
        eye = p.eye(DL * s.d, DL * s.d).reshape(DL, s.d * DL * s.d)
        return p.matmul(s.R, eye, k, s.d * DL * s.d, DL).reshape(k, s.d, DL * s.d)
    DR, k = s.R.shape       # eye[d*DR, d, DR] . R[DR, k]
    eye = p.eye(s.d * DR, s.d * DR).reshape(s.d * DR * s.d, DR)
    return p.matmul(eye, s.R, s.d * DR * s.d, k, DR).reshape(s.d * DR, s.d, k)


def absorb_left(p: Program, F: DT, s):
    """tensordot(F[k, DL], site, ([1], [0]))"""
    k = F.shape[0]
    if isinstance(s, Id):
        DL, d, DR = s.shape
        if s.kind == "L":
            return Kron("L", F, d)
        return F.reshape(k, d, DR)
    if isinstance(s, Kron):
        if s.side == "L":
            k0, DL = s.R.shape
            return Kron("L", p.matmul(F, s.R, k, DL, k0), s.d)
        DR, k2 = s.R.shape
        return p.matmul(F.reshape(k * s.d, DR), s.R, k * s.d, k2, DR).reshape(k, s.d, k2)
    return p.tensordot(F, s, ([1], [0]))


def absorb_right(p: Program, s, F: DT):
    """tensordot(site, F[DR, k], ([2], [0]))"""
    k = F.shape[1]
    if isinstance(s, Id):
        DL, d, DR = s.shape
        if s.kind == "R":
            return Kron("R", F, d)
        return F.reshape(DL, d, k)
    if isinstance(s, Kron):
        if s.side == "R":
            DR, k0 = s.R.shape
            return Kron("R", p.matmul(s.R, F, DR, k, k0), s.d)
        k0, DL = s.R.shape
        return p.matmul(s.R, F.reshape(DL, s.d * k), k0, s.d * k, DL).reshape(k0, s.d, k)
    return p.tensordot(s, F, ([2], [0]))


def two_site_matrix(p: Program, a, b) -> DT:
    """matrix [DLa*da, db*DRb] of the contraction of two neighbouring sites over their shared bond."""
    DLa, da, Dm = a.shape
    Dm2, db, DRb = b.shape
    assert Dm == Dm2
    rows, cols = DLa * da, db * DRb
    if isinstance(a, Id) and a.kind == "L":
        return dense(p, b).reshape(rows, cols)
    if isinstance(b, Id) and b.kind == "R":
        return dense(p, a).reshape(rows, cols)
    if isinstance(a, Kron) and a.side == "L":          # (R (x) I_d) . B
        k, DL = a.R.shape
        B = dense(p, b).reshape(DL, a.d * cols)
        return p.matmul(a.R, B, k, a.d * cols, DL).reshape(rows, cols)
    if isinstance(b, Kron) and b.side == "R":          # A . (I_d (x) R)
        DR, k = b.R.shape
        A = dense(p, a).reshape(rows * b.d, DR)
        return p.matmul(A, b.R, rows * b.d, k, DR).reshape(rows, cols)
    return p.matmul(dense(p, a).reshape(rows, Dm), dense(p, b).reshape(Dm, cols), rows, cols, Dm)


class DevMPS:
    def __init__(self, prog: Program, N: int, slot_lognorm: int = SLOT_LOGNORM, slot_trunc: int = SLOT_TRUNC):
        self.p = prog
        self.N = N
        self.A: list = [None] * N
        self.Corder: list = [None] * N
        self.slot_lognorm = slot_lognorm
        self.slot_trunc = slot_trunc

    def set_site(self, t, i: int, Corder=None):
        assert len(t.shape) == 3
        self.A[i] = t
        self.Corder[i] = Corder

    def set_lists(self, A, Corder):
        self.A, self.Corder = list(A), list(Corder)
        self.N = len(self.A)

    def shapes(self):
        return [a.shape for a in self.A]

    def site(self, i: int) -> DT:
        """dense device tensor of site i (materialises symbolic sites)."""
        self.A[i] = dense(self.p, self.A[i])
        return self.A[i]

    def dense_sites(self):
        return [self.site(i) for i in range(self.N)]

    def update_A0_norm(self):
        a0 = self.p.copy(self.site(0))
        self.p.normalize_(a0, self.slot_lognorm)
        self.A[0] = a0          # tag unchanged

    # ------------------------------------------------------------------------------------------
    def _qr_site(self, i: int):
        """left-canonicalise site i, push the remainder into site i+1 (tag of i+1 is reset)."""
        p, s = self.p, self.A[i]
        if isinstance(s, Id) and s.kind == "L":
            self.Corder[i] = "L"                  # QR(I) = I . I
            self.Corder[i + 1] = None
            return
        if isinstance(s, Kron) and s.side == "L":
            nxt = self.A[i + 1]
            k, DL = s.R.shape
            if isinstance(nxt, DT) or (isinstance(nxt, Id) and nxt.kind == "R"):
                # (R (x) I_d) = I . (R (x) I_d): this site becomes an identity, the factor moves on as R on a split index
                B = dense(p, nxt)
                Dn, d2, D2 = B.shape
                assert Dn == DL * s.d
                new = p.matmul(s.R, B.reshape(DL, s.d * d2 * D2), k, s.d * d2 * D2, DL).reshape(k * s.d, d2, D2)
                self.set_site(Id(k, s.d, k * s.d), i, "L")
                self.set_site(new, i + 1)
                return
        D1, d, D2 = s.shape
        Q, R = p.qr(dense(p, s).reshape(D1 * d, D2))
        self.set_site(Q.reshape(D1, d, Q.shape[1]), i, "L")
        self.set_site(absorb_left(p, R, self.A[i + 1]), i + 1)

    def left_canonical_QR(self, i0=None, i1=None):
        if self.N < 2:
            return
        i0 = 0 if i0 is None else i0
        i1 = self.N - 2 if i1 is None else min(i1, self.N - 2)
        for i in range(i0, i1 + 1):
            if self.Corder[i] == "L":
                continue
            self._qr_site(i)

    def _rq_site(self, i: int, nr_bulk: bool):
        """right-canonicalise site i without truncation, push the remainder into site i-1."""
        p, s = self.p, self.A[i]
        if isinstance(s, Id) and s.kind == "R":
            self.Corder[i] = "R"                  # RQ(I) = I . I
            self.Corder[i - 1] = None
            return
        if isinstance(s, Kron) and s.side == "R" and isinstance(self.A[i - 1], DT):
            DR, k = s.R.shape                     # (I_d (x) R) = (I_d (x) R) . I
            prev = self.A[i - 1]
            D0, d0, Dp = prev.shape
            assert Dp == s.d * DR
            new = p.matmul(prev.reshape(D0 * d0 * s.d, DR), s.R, D0 * d0 * s.d, k, DR).reshape(D0, d0, s.d * k)
            self.set_site(Id(s.d * k, s.d, k), i, "R")
            self.set_site(new, i - 1)
            return
        D1, d, D2 = s.shape
        Lm, Q = p.lq(dense(p, s).reshape(D1, d * D2))
        if nr_bulk:
            p.normalize_(Lm, self.slot_lognorm)
        self.set_site(Q.reshape(Q.shape[0], d, D2), i, "R")
        self.set_site(absorb_right(p, self.A[i - 1], Lm), i - 1)

    def right_canonical(self, maxD=None, i0=None, i1=None, nr_bulk=False, two_site_at=None):
        """``two_site_at`` = i: the QR of site i-1 was deferred; factorise the product of sites (i-1, i) in one SVD."""
        if self.N < 2:
            return
        maxD = 10000000 if maxD is None else maxD
        i0 = 1 if i0 is None else i0
        i1 = self.N - 1 if i1 is None else i1
        p = self.p
        for i in range(i1, i0 - 1, -1):
            D1, d, D2 = self.A[i].shape
            if two_site_at == i:
                a = self.A[i - 1]
                Da, da, _ = a.shape
                P = two_site_matrix(p, a, self.A[i])
                r = min(Da * da, D1)                       # the bond the reference's QR would have left
                keep = min(min(r, d * D2), maxD)
                US, Vh = p.svd_trunc(P, keep, nr_bulk, self.slot_lognorm, self.slot_trunc)
                self.set_site(Vh.reshape(keep, d, D2), i, "R")
                self.set_site(US.reshape(Da, da, keep), i - 1)
            elif D1 > maxD:
                M = dense(p, self.A[i]).reshape(D1, d * D2)
                keep = min(min(D1, d * D2), maxD)
                US, Vh = p.svd_trunc(M, keep, nr_bulk, self.slot_lognorm, self.slot_trunc)
                self.set_site(Vh.reshape(keep, d, D2), i, "R")
                self.set_site(absorb_right(p, self.A[i - 1], US), i - 1)
            else:
                if self.Corder[i] == "R":
                    continue
                self._rq_site(i, nr_bulk)
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
        # the left sweep up to iD1-1 is the reference's; its last QR (site iD1) is deferred when the reference
        # would follow it with a truncating SVD of site iD1+1:  r = min(D1*d, bond) is the bond that QR leaves,
        # and the SVD happens iff r > maxD
        self.left_canonical_QR(i0, iD1 - 1)
        a = self.A[iD1].shape
        defer = self.Corder[iD1] != "L" and min(a[0] * a[1], a[2]) > maxD
        if defer:
            self.Corder[iD1] = "L"        # the reference's tag while it looks for i1 (overwritten by the two-site step)
        else:
            self.left_canonical_QR(iD1, iD1)
        i1 = self.N - 1
        for i1 in range(self.N - 1, iD1 - 1, -1):
            if self.Corder[i1] != "R":
                break
        if defer and i1 < iD1 + 1:
            i1 = iD1 + 1
        self.right_canonical(maxD, i0=iD0 + 1, i1=i1, nr_bulk=nr_bulk, two_site_at=(iD1 + 1) if defer else None)


def add_two_mps(p: Program, A: DevMPS, alpha, B: DevMPS, beta, sign_slot_beta=-1) -> DevMPS:
    """block-diagonal sum alpha*A + beta*B (src/libs/bmpslib.py:2781-2864); if ``sign_slot_beta`` >= 0 the
    coefficient beta is multiplied on the device by sign(slots[sign_slot_beta])."""
    assert A.N == B.N
    N = A.N
    out = DevMPS(p, N, A.slot_lognorm, A.slot_trunc)
    for i in range(N):
        a, b = A.site(i), B.site(i)
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
        a, b = A.site(i), B.site(i)
        if C is None:
            assert a.shape[0] == 1 and b.shape[0] == 1
            C = p.tensordot(a.reshape(a.shape[1], a.shape[2]), b.reshape(b.shape[1], b.shape[2]), ([0], [0]), conj_b=True)
        else:
            C1 = p.tensordot(C, a, ([0], [0]))
            C = p.tensordot(C1, b, ([0, 1], [0, 1]), conj_b=True)
    assert C.size == 1
    return C
