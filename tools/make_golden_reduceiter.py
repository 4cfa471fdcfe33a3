"""Generate tests/golden/reduce_iter.npz by running the UNMODIFIED reference `bmpslib.mps.reduceDiter`
(src/libs/bmpslib.py:989-1364) on seeded MPSs in the build container.

    python tools/make_golden_reduceiter.py
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import ref_env  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (bond dims D_0..D_N, physical dim, maxD, nr_bulk, max_iter, err, number of leading sites made left-canonical first)
CASES = {
    "plain":        ([1, 4, 12, 12, 12, 4, 1], 3, 6, False, 10, 1e-6, 0),
    "bulk":         ([1, 4, 12, 14, 12, 4, 1], 3, 5, True, 10, 1e-6, 0),
    "bulk_lcanon":  ([1, 3, 9, 12, 12, 9, 3, 1], 3, 6, True, 10, 1e-8, 2),
    "fixed_rounds": ([1, 4, 10, 10, 4, 1], 2, 4, False, 3, 0.0, 0),
    "one_round":    ([1, 4, 10, 10, 4, 1], 2, 4, True, 1, 1e-6, 0),
    "nothing":      ([1, 2, 4, 4, 2, 1], 2, 8, True, 10, 1e-6, 0),
    "open_ends":    ([3, 8, 12, 12, 8, 2], 2, 6, True, 12, 1e-7, 1),
    "kagome_like":  ([1, 16, 64, 64, 64, 16, 1], 4, 32, True, 6, 1e-6, 0),
}


def build(bmpslib, dims, d, n_left, rng):
    n = len(dims) - 1
    mp = bmpslib.mps(n)
    for i in range(n):
        a = rng.normal(size=(dims[i], d, dims[i + 1])) + 1j * rng.normal(size=(dims[i], d, dims[i + 1]))
        mp.set_site(a * (0.7 + 0.1 * i), i)
    if n_left > 0:
        mp.left_canonical_QR(0, n_left - 1)
    return mp


def main():
    ref_env.setup()
    from libs import bmpslib
    out = {}
    for name, (dims, d, maxD, nr_bulk, max_iter, err, n_left) in CASES.items():
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        mp = build(bmpslib, dims, d, n_left, rng)
        n = mp.N
        for i in range(n):
            out[f"{name}/in{i}"] = np.array(mp.A[i])
        out[f"{name}/in_corder"] = np.array([c if c is not None else "" for c in mp.Corder])
        out[f"{name}/params"] = np.array([maxD, int(nr_bulk), max_iter, err, d, n_left], dtype=np.float64)
        mp.reduceDiter(maxD, nr_bulk=nr_bulk, max_iter=max_iter, err=err)
        for i in range(n):
            out[f"{name}/out{i}"] = np.array(mp.A[i])
        out[f"{name}/out_corder"] = np.array([c if c is not None else "" for c in mp.Corder])
        out[f"{name}/nr"] = np.array([mp.nr_mantissa, mp.nr_exp], dtype=np.float64)
        print(name, "->", [a.shape for a in mp.A], mp.Corder, mp.nr_mantissa, mp.nr_exp)
    np.savez_compressed(os.path.join(GOLD, "reduce_iter.npz"), **out)
    print("wrote", os.path.join(GOLD, "reduce_iter.npz"))


if __name__ == "__main__":
    main()
