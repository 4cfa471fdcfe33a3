"""Dense linear algebra of the ITE step as device programs: the backend object ``B`` that ``ite.py`` / ``edge_env.py``
are written against.  Every call compiles (once per shape signature) and runs a small tensor program through the C ABI
(include/kbp.h): DMMA ZGEMM for tensordot / scaling / linear combinations, Householder QR, Jacobi SVD, and the Hermitian
eigensolver built on the Jacobi SVD of the shifted matrix.  Arrays cross the boundary as host numpy complex128 (the S3/S4
seams of the reference are per-call host interfaces); reshape / transpose / slicing are host data movement.

There is no CPU fallback: constructing ``DeviceBackend`` without libkbp.so or without a GPU raises ``EngineUnavailable``.
"""
from __future__ import annotations

import math

import numpy as np

from .program import Program, _prod
from .runtime import Compiled, get_engine

_EIG_SHIFT = 1.5


class DeviceBackend:
    def __init__(self, engine_key="ite", device: int = 0):
        self.eng = get_engine(engine_key, device)
        self._cache: dict = {}
        self.calls = 0

    # ---------------------------------------------------------------- plumbing
    def _run(self, key, build, inputs):
        """build(p, [DT...]) -> list of output DTs; returns list of ndarrays (+ slots)."""
        comp = self._cache.get(key)
        if comp is None:
            p = Program(8)
            dts = [(f"i{k}", p.input(f"i{k}", a.shape if a.ndim else (1,))) for k, a in enumerate(inputs)]
            outs = build(p, [t for _, t in dts])
            comp = Compiled(p, dts, [(f"o{k}", t) for k, t in enumerate(outs)])
            self._cache[key] = comp
        o, sl, rc = comp.run(self.eng, [{f"i{k}": a for k, a in enumerate(inputs)}], soft_errors=(-4,))
        self.calls += 1
        return [o[0][f"o{k}"] for k in range(len(comp.out_layout))], sl[0]

    @staticmethod
    def _c(a):
        return np.ascontiguousarray(a, dtype=np.complex128)

    # ---------------------------------------------------------------- data movement (host)
    def reshape(self, a, shape):
        return np.reshape(np.asarray(a), shape)

    def transpose(self, a, perm):
        return np.transpose(np.asarray(a), perm)

    # ---------------------------------------------------------------- device ops
    def tensordot(self, a, b, axes, conj_a=False, conj_b=False):
        a, b = self._c(a), self._c(b)
        if isinstance(axes, int):
            assert axes == 0
            out = self.tensordot(a.reshape(-1, 1), b.reshape(1, -1), ([1], [0]), conj_a, conj_b)
            return out.reshape(a.shape + b.shape)
        ax = (tuple(int(x) for x in axes[0]), tuple(int(x) for x in axes[1]))
        if a.size == 0 or b.size == 0:
            return np.tensordot(a, b, axes=ax)
        key = ("td", a.shape, b.shape, ax, conj_a, conj_b)
        (c,), _ = self._run(key, lambda p, t: [p.tensordot(t[0], t[1], ax, conj_a=conj_a, conj_b=conj_b)], [a, b])
        return c

    def scale(self, a, s):
        a = self._c(a)
        out = self.tensordot(a.reshape(-1, 1), np.array([[s]], dtype=np.complex128), ([1], [0]))
        return out.reshape(a.shape)

    def lincomb(self, a, alpha, b, beta):
        """alpha a + beta b as one GEMM  [a | b] [[alpha I], [beta I]]."""
        a, b = self._c(a), self._c(b)
        shape = a.shape
        a2, b2 = a.reshape(-1, shape[-1]), b.reshape(-1, shape[-1])
        r, c = a2.shape
        coef = np.concatenate([alpha * np.eye(c), beta * np.eye(c)], axis=0).astype(np.complex128)

        def build(p, t):
            z = p.zeros((r, 1, 2 * c))
            p.embed(z, (0, 0, 0), t[0].reshape(r, 1, c))
            p.embed(z, (0, 0, c), t[1].reshape(r, 1, c))
            return [p.matmul(z.reshape(r, 2 * c), t[2], r, c, 2 * c)]
        (out,), _ = self._run(("lc", r, c), build, [a2, b2, coef])
        return out.reshape(shape)

    def hermitize(self, m):
        m = self._c(m)
        return self.lincomb(m, 0.5, np.conj(m.T), 0.5)

    def norm(self, a) -> float:
        a = self._c(a).reshape(-1)
        if a.size == 0:
            return 0.0

        def build(p, t):
            x = p.copy(t[0])
            p.normalize_(x, 0)
            return [x]
        _, sl = self._run(("nrm", a.size), build, [a])
        v = float(sl[0])
        return math.exp(v) if np.isfinite(v) and v != 0.0 else float(np.linalg.norm(a) == 0.0 and 0.0 or math.exp(v))

    def qr(self, m):
        m = self._c(m)
        (q, r), _ = self._run(("qr", m.shape), lambda p, t: list(p.qr(t[0])), [m])
        return q, r

    def _svd_raw(self, m):
        m = self._c(m)
        k = min(m.shape)

        def build(p, t):
            us, vh = p.svd_trunc(t[0], k, False, 0, 1, warm=False)
            g = p.matmul(us, us, k, k, m.shape[0], 2, 0)            # US^H US: its diagonal holds s^2
            return [us, vh, g]
        (us, vh, g), _ = self._run(("svd", m.shape), build, [m])
        s = np.sqrt(np.maximum(np.real(np.diag(g)), 0.0))
        return us, s, vh

    def svd(self, m):
        us, s, vh = self._svd_raw(m)
        inv = np.where(s > 0, 1.0 / np.where(s > 0, s, 1.0), 0.0)
        u = self.tensordot(us, np.diag(inv).astype(np.complex128), ([1], [0]))
        return u, s, vh

    def eigh(self, h):
        """Hermitian eigen-decomposition, eigenvalues ascending: Jacobi SVD of the positive definite H / |H|_F + 1.5 I."""
        h = self._c(h)
        n = h.shape[0]
        nrm = self.norm(h)
        if not nrm > 0.0:
            return np.zeros(n), np.eye(n, dtype=np.complex128)
        m = self.lincomb(h, 1.0 / nrm, np.eye(n, dtype=np.complex128), _EIG_SHIFT)
        us, s, vh = self._svd_raw(m)
        w = (s - _EIG_SHIFT) * nrm
        order = np.argsort(w, kind="stable")
        return w[order], np.conj(vh[order]).T
