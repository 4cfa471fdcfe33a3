"""Fixture for the non-repeated Kagome block (KagomeTNArbitrary, src/tensor_networks/tensor_network.py:400-519): 21 independent
random site tensors (D=2, N=2), the UNMODIFIED reference's block BP from uniform messages (exact-SVD branch) and the six
mode-A edge energies from its Core -> Mode -> Edge reduction.

    python tools/make_golden_arbitrary.py   ->  tests/golden/arbitrary_D2_N2.npz
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
    from algo.belief_propagation import belief_propagation
    from algo.tn_reduction import reduce_core_to_mode, reduce_full_kagome_to_core, reduce_mode_to_edge
    from containers import UpdateEdge
    from enums import UpdateMode
    from lattices.directions import BlockSide
    from libs.ITE import rho_ij
    from physics.hamiltonians import heisenberg_afm
    from tensor_networks.tensor_network import KagomeTNArbitrary
    D, N, d = 2, 2, 2
    rng = np.random.default_rng(77)
    n_sites = 21
    tensors = []
    for _ in range(n_sites):
        t = rng.normal(size=(d, D, D, D, D)) + 1j * rng.normal(size=(d, D, D, D, D))
        tensors.append(t / np.linalg.norm(t))
    tn = KagomeTNArbitrary([t.copy() for t in tensors])
    tn.deal_cell_flavors()          # up / left / right corner of every upper triangle -> A / B / C (what measure_* does first)
    cfg = ref_env.quiet_config(D, N)
    cfg.bp.msg_diff_terminate = 1e-10
    cfg.bp.damping = 0.1
    tn.connect_uniform_messages()
    msgs, stats = belief_propagation(tn, tn.messages, cfg.bp)
    print(f"BP {stats.iterations} iterations, error {stats.final_error:.2e}")
    h = np.asarray(heisenberg_afm())
    out = {f"site{i}": t for i, t in enumerate(tensors)}
    out["h"] = h
    out["cfg"] = np.array([cfg.bp.trunc_dim, cfg.contraction.trunc_dim, stats.iterations, 1e-10, 0.1, stats.final_error], dtype=float)
    for side, m in msgs.items():
        for k, a in enumerate(m.mps.A):
            out[f"msg_{side}_{k}"] = a
    core = reduce_full_kagome_to_core(tn, cfg.contraction, direction=BlockSide.U)
    mode_tn = reduce_core_to_mode(core, UpdateMode.A)
    names, energies = [], []
    for e in UpdateEdge.all_options():
        et = reduce_mode_to_edge(mode_tn, e, cfg.contraction, arange_legs=False)
        et.rearrange_tensors_and_legs_into_canonical_order()
        t1, t2, env = et.edge_and_environment()
        rdm = rho_ij(t1, t2, mps_env=env)
        names.append(str(e).replace("(", "").replace(")", "").replace(", ", "").replace(" ", ""))
        energies.append(float(np.real(np.dot(rdm.flatten(), h.flatten()))))
        out[f"rdm_{names[-1]}"] = rdm
    out["edges"] = np.array(names)
    out["edge_energies"] = np.array(energies)
    print(dict(zip(names, energies)))
    np.savez_compressed(os.path.join(GOLD, "arbitrary_D2_N2.npz"), **out)
    print("wrote arbitrary_D2_N2.npz")


if __name__ == "__main__":
    main()
