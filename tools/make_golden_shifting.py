"""Fixtures for the periodic lattice shifting of the non-repeated block (src/lattices/triangle.py:965-1036, 1138-1200;
src/tensor_networks/tensor_network.py:484-519) and for calc_measurement_non_unit_cell_kagome_tn (src/algo/measurements.py:245-321),
produced by the UNMODIFIED reference.

    python tools/make_golden_shifting.py   ->  tests/golden/shifting.npz
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


def main():
    ref_env.setup()
    from libs import bmpslib
    _orig = bmpslib._perf_svd
    bmpslib._perf_svd = lambda m, svd_emthod="svd", check_result=False: _orig(m, "svd")
    from enums import MessageModel
    from lattices import triangle as tri
    from lattices.directions import LatticeDirection
    from lattices.kagome import KagomeLattice
    from tensor_networks.tensor_network import KagomeTNArbitrary
    out = {}
    for N in (2, 3):
        perms = [list(p) for p in tri.all_periodic_lattice_shifting_permutation(N)]
        out[f"N{N}_all"] = np.array(perms, dtype=np.int64)
        for d in LatticeDirection.all_in_counter_clockwise_order():
            out[f"N{N}_dir_{d}"] = np.array(tri.shift_periodically_in_direction(N, d), dtype=np.int64)
        lat = KagomeLattice(N)
        out[f"N{N}_triangles"] = np.array([[t.up.index, t.left.index, t.right.index] for t in lat.triangles], dtype=np.int64)
        print(f"N={N}: {len(perms)} shifts; triangles[0..2] = {out[f'N{N}_triangles'][:3].tolist()}")
    # the shifted tensor lists of one arbitrary block (pure data movement) + the measurement
    D, N, d = 2, 2, 2
    rng = np.random.default_rng(78)
    tensors = []
    for _ in range(21):
        t = rng.normal(size=(d, D, D, D, D)) + 1j * rng.normal(size=(d, D, D, D, D))
        tensors.append(t / np.linalg.norm(t))
    tn = KagomeTNArbitrary([t.copy() for t in tensors])
    for i, t in enumerate(tensors):
        out[f"site{i}"] = t
    shifted = list(tn.all_lattice_shifting_options())
    # which original tensor sits on every site after each shift
    src = []
    for s in shifted:
        src.append([next(j for j, t in enumerate(tensors) if np.array_equal(t, x)) for x in s.tensors])
    out["shift_sources"] = np.array(src, dtype=np.int64)
    from algo.measurements import calc_measurement_non_unit_cell_kagome_tn
    # the reference draws the bottom-up direction of the core reduction at random from {U, DL, DR} (kagome_to_core.py:177-179):
    # fixed to U for a reproducible fixture, without touching the reference's files
    import algo.tn_reduction.kagome_to_core as ktc
    from lattices.directions import BlockSide
    _basic = ktc._basic_data
    ktc._basic_data = lambda tn_, direction: _basic(tn_, BlockSide.U if direction is None else direction)
    cfg = ref_env.quiet_config(D, N)
    cfg.bp.init_msg = MessageModel("UQ")
    cfg.bp.msg_diff_terminate = 1e-10
    cfg.bp.damping = 0.1
    e = calc_measurement_non_unit_cell_kagome_tn(KagomeTNArbitrary([t.copy() for t in tensors]), config=cfg, print_=False)
    out["measurement"] = np.array([float(np.real(e))])
    out["cfg"] = np.array([cfg.bp.trunc_dim, cfg.contraction.trunc_dim, 1e-10, 0.1], dtype=float)
    print("calc_measurement_non_unit_cell_kagome_tn ->", e)
    np.savez_compressed(os.path.join(GOLD, "shifting.npz"), **out)
    print("wrote shifting.npz")


if __name__ == "__main__":
    main()
