"""ORACLE (test infrastructure, not product): numpy restatement of the MPS class the
reference's boundary-MPS contractor is built on.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  Parity status: PINNED -- every function here is checked against the
real reference (imported in the build container, exact-SVD branch) by ``tools/make_golden.py``;
the resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py``.

Each function cites the reference lines it restates (paths relative to /root/reference/).
The SVD is ``numpy.linalg.svd`` = the reference's ``svd_emthod == "svd"`` branch
(src/libs/bmpslib.py:2874-2875); the default ``rsvd`` comes from the un-vendored, un-pinned
third-party package ``quimb`` and is random at the 1e-5 level, so it cannot be a parity target.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import rq as _scipy_rq


def fexp(f) -> int:
    """base-10 exponent (src/libs/bmpslib.py:1746-1747)."""
    return int(math.floor(math.log10(abs(f)))) if f != 0 else 0


def fman(f):
    """base-10 mantissa (src/libs/bmpslib.py:1749-1750)."""
    return f / 10 ** fexp(f)


class MPS:
    """List of [DL, d, DR] site tensors + per-site canonical tag + (mantissa, exp10) overall scale.
    Mirrors ``bmpslib.mps`` (src/libs/bmpslib.py:214-232)."""

    def __init__(self, N: int):
        self.N = N
        self.A: list = [None] * N
        self.Corder: list = [None] * N
        self.nr_mantissa = 1.0
        self.nr_exp = 0

    # -- bookkeeping ---------------------------------------------------------------------------
    def set_site(self, mat, i, Corder=None):
        """always copies and resets the tag unless given (src/libs/bmpslib.py:501-505)."""
        self.A[i] = np.array(mat, copy=True)
        self.Corder[i] = Corder

    def set_lists(self, A, Corder):
        """replace the site lists, N follows (src/libs/bmpslib.py:245-274)."""
        self.A = list(A)
        self.Corder = list(Corder)
        self.N = len(self.A)

    def copy(self, full: bool = False) -> "MPS":
        """(src/libs/bmpslib.py:416-450)"""
        new = MPS(self.N)
        new.nr_mantissa, new.nr_exp = self.nr_mantissa, self.nr_exp
        new.Corder = list(self.Corder)
        new.A = [a.copy() for a in self.A] if full else list(self.A)
        return new

    def overall_factor(self):
        """(src/libs/bmpslib.py:381-382)"""
        return self.nr_mantissa * 10 ** self.nr_exp

    def reset_nr(self):
        """(src/libs/bmpslib.py:390-398)"""
        self.nr_mantissa = 1
        self.nr_exp = 0

    def update_A0_norm(self):
        """push |A[0]| into (mantissa, exp) (src/libs/bmpslib.py:359-374)."""
        nr = np.linalg.norm(self.A[0])
        self.set_site(self.A[0] / nr, 0, self.Corder[0])
        self.nr_mantissa *= fman(nr)
        self.nr_exp += fexp(nr)
        if abs(self.nr_mantissa) >= 10 or abs(self.nr_mantissa) < 1:
            e = fexp(self.nr_mantissa)
            self.nr_mantissa = fman(self.nr_mantissa)
            self.nr_exp += e

    def shapes(self):
        return [a.shape for a in self.A]

    # -- canonicalisation ----------------------------------------------------------------------
    def left_canonical_QR(self, i0=None, i1=None):
        """QR sweep left->right over [i0, i1], skipping sites tagged 'L'
        (src/libs/bmpslib.py:553-595)."""
        if self.N < 2:
            return
        i0 = 0 if i0 is None else i0
        i1 = self.N - 2 if i1 is None else min(i1, self.N - 2)
        for i in range(i0, i1 + 1):
            if self.Corder[i] == "L":
                continue
            D1, d, D2 = self.A[i].shape
            Q, R = np.linalg.qr(self.A[i].reshape(D1 * d, D2))
            self.set_site(Q.reshape(D1, d, Q.shape[1]), i, "L")
            self.set_site(np.tensordot(R, self.A[i + 1], axes=([1], [0])), i + 1)

    def right_canonical(self, maxD=None, eps=None, i0=None, i1=None, nr_bulk=False):
        """sweep right->left over [i0, i1]: truncating SVD where the left bond exceeds maxD, RQ
        otherwise; returns the summed relative truncation error (src/libs/bmpslib.py:688-822)."""
        if self.N < 2:
            return
        maxD = 10000000 if maxD is None else maxD
        i0 = 1 if i0 is None else i0
        i1 = self.N - 1 if i1 is None else i1
        overall = 1.0
        trunc_err = 0.0
        for i in range(i1, i0 - 1, -1):
            s = np.sum(self.A[i])
            if np.isnan(s) or np.isinf(s):
                raise FloatingPointError(f"site {i} holds nan/inf")  # reference: print + exit(1), :711-717
            D1, d, D2 = self.A[i].shape
            M = self.A[i].reshape(D1, d * D2)
            err = 0.0
            if D1 > maxD or eps is not None:
                U, S, V = np.linalg.svd(M, full_matrices=False)
                if nr_bulk:
                    nrS = np.linalg.norm(S)
                    S = S / nrS
                    overall *= nrS
                keep = min(len(S) if eps is None else int(np.sum(S > S[0] * eps)), maxD)
                err = math.sqrt(np.sum(S[keep:] ** 2) / np.sum(S ** 2))
                self.set_site(V[:keep].reshape(keep, d, D2), i, "R")
                US = U[:, :keep] * S[:keep]
                self.set_site(np.tensordot(self.A[i - 1], US, axes=([2], [0])), i - 1)
            else:
                if self.Corder[i] == "R":
                    continue
                R, Q = _scipy_rq(M, mode="economic")
                if nr_bulk:
                    nrR = np.linalg.norm(R)
                    R = R / nrR
                    overall *= nrR
                self.set_site(Q.reshape(Q.shape[0], d, D2), i, "R")
                self.set_site(np.tensordot(self.A[i - 1], R, axes=([2], [0])), i - 1)
            trunc_err += err
        if nr_bulk:
            self.set_site(self.A[0] * overall, 0, self.Corder[0])
            self.update_A0_norm()
        return trunc_err

    def reduceD(self, maxD, eps=None, nr_bulk=False):
        """mode 'MC' only (the one bubblecon uses): minimal window needing truncation, QR on its left
        part, SVD/RQ on its right part (src/libs/bmpslib.py:868-959)."""
        if self.N < 3:
            return

        def oversized(i):
            a, b = self.A[i], self.A[i + 1]
            return a.shape[2] > min(maxD, a.shape[0] * a.shape[1], b.shape[2] * b.shape[1])

        need = [i for i in range(self.N - 1) if oversized(i)]
        if not need:
            return 0
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
        return self.right_canonical(maxD, eps, i0=iD0 + 1, i1=i1, nr_bulk=nr_bulk)


# -- free functions ------------------------------------------------------------------------------
def update_C_left(C, A, B, conjB=False):
    """transfer-matrix step left->right (src/libs/bmpslib.py:2172-2211)."""
    Bc = np.conj(B) if conjB else B
    if C is None:
        return np.tensordot(A[0], Bc[0], axes=([0], [0]))
    C1 = np.tensordot(C, A, axes=([0], [0]))
    return np.tensordot(C1, Bc, axes=([0, 1], [0, 1]))


def update_C_right(C, A, B, conjB=False):
    """transfer-matrix step right->left (src/libs/bmpslib.py:2225-2264)."""
    Bc = np.conj(B) if conjB else B
    if C is None:
        return np.tensordot(A[:, :, 0], Bc[:, :, 0], axes=([1], [1]))
    C1 = np.tensordot(A, C, axes=([2], [0]))
    return np.tensordot(C1, Bc, axes=([1, 2], [1, 2]))


def mps_inner_product(A: MPS, B: MPS, conjB=False):
    """<A|B> including both overall factors (src/libs/bmpslib.py:2288-2301)."""
    C = None
    for i in range(A.N):
        C = update_C_left(C, A.A[i], B.A[i], conjB)
    fb = np.conj(B.overall_factor()) if conjB else B.overall_factor()
    return C[0, 0] * A.overall_factor() * fb


def add_two_MPSs(mpsA: MPS, alpha, mpsB: MPS, beta) -> MPS:
    """block-diagonal MPS sum alpha*A + beta*B; the scalars go into site 0 only; overall factors
    of the inputs are NOT carried (src/libs/bmpslib.py:2781-2864)."""
    assert mpsA.N == mpsB.N
    N = mpsA.N
    out = MPS(N)
    for i in range(N):
        a, b = mpsA.A[i], mpsB.A[i]
        DLa, d, DRa = a.shape
        DLb, _, DRb = b.shape
        dt = (a[0, 0, 0] + b[0, 0, 0]).dtype
        if i == 0:
            s = np.zeros([1, d, DRa + DRb], dtype=dt)
            s[:, :, :DRa] = alpha * a
            s[:, :, DRa:] = beta * b
        elif i == N - 1:
            s = np.zeros([DLa + DLb, d, 1], dtype=dt)
            s[:DLa] = a
            s[DLa:] = b
        else:
            s = np.zeros([DLa + DLb, d, DRa + DRb], dtype=dt)
            s[:DLa, :, :DRa] = a
            s[DLa:, :, DRa:] = b
        out.set_site(s, i)
    return out


def mps_distance(m1: MPS, m2: MPS) -> float:
    """1 - |<m1|m2>| clipped at 0 (src/tensor_networks/mps.py:48-74)."""
    d = 1 - abs(mps_inner_product(m1, m2, True))
    return 0.0 if d < 0 else float(d)


def hermitize_a_message(mpA: MPS) -> MPS:
    """0.5*(M + M^dagger) on the fused (ket,bra) legs, recompressed to the largest left bond
    (src/libs/ITE.py:116-185)."""
    N = mpA.N
    mpB = MPS(N)
    Dmax = 0
    for i in range(N):
        DL, d2, DR = mpA.A[i].shape
        d = int(math.sqrt(d2))
        t = np.conj(mpA.A[i].reshape(DL, d, d, DR).transpose(0, 2, 1, 3)).reshape(DL, d2, DR)
        mpB.set_site(t, i)
        Dmax = max(Dmax, DL)
    mpC = add_two_MPSs(mpA, 0.5, mpB, 0.5)
    mpC.reduceD(Dmax)
    return mpC


def init_mps_quantum(D_list, random=False, rng=None) -> MPS:
    """product-state initial message: vectorised identity (UQ) or random |v><v| (RQ) per site,
    embedded with bond D^2 by slicing, then left-canonical QR and end-normalisation
    (src/tensor_networks/mps.py:77-156).  ``rng`` replaces the reference's global np.random."""
    N = len(D_list)
    mp = MPS(N)
    for i, D in enumerate(D_list):
        D2, D3 = D * D, D * D * D
        if random:
            rs = np.random if rng is None else rng
            a = rs.normal(size=[D3, D3]) + 1j * rs.normal(size=[D3, D3])
            a /= np.linalg.norm(a)
            kb = a @ np.conj(a.T)
        else:
            kb = np.eye(D3)
            kb /= np.linalg.norm(kb)
        kb = kb.reshape([D] * 6).transpose([0, 3, 1, 4, 2, 5]).reshape([D2, D2, D2])
        if i == 0:
            kb = kb[0, :, :].reshape([1, D2, D2])
        if i == N - 1:
            kb = kb[:, :, 0].reshape([kb.shape[0], D2, 1])
        mp.set_site(kb, i)
    mp.left_canonical_QR()
    mp.set_site(mp.A[N - 1] / np.linalg.norm(mp.A[N - 1]), N - 1)
    return mp


def mps_to_dense(mp: MPS, with_factor: bool = True):
    """helper for tests: contract an MPS into the dense vector over its physical legs."""
    v = mp.A[0]
    v = v.reshape(v.shape[0], -1, v.shape[2])
    out = v
    for i in range(1, mp.N):
        out = np.tensordot(out, mp.A[i], axes=([out.ndim - 1], [0]))
    out = out.reshape(out.shape[0], -1, out.shape[-1])
    assert out.shape[0] == 1 and out.shape[2] == 1
    out = out[0, :, 0]
    return out * mp.overall_factor() if with_factor else out
