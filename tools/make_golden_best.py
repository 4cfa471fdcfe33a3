"""Known-answer fixture from the reference's own data: its best D=2 unit cell of the Kagome Heisenberg AFM
(data/unit_cells/best/"D=2 energy=-0.4046412208223448.dat", a dill pickle of BestUnitCellData) measured by the UNMODIFIED
reference (exact-SVD branch) in the build container: BP to convergence, Core -> Mode A -> six edges, two-site RDM energies.

    python tools/make_golden_best.py       ->  tests/golden/best_D2.npz
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
BEST = "/root/reference/data/unit_cells/best/D=2 energy=-0.4046412208223448.dat"


def main():
    ref_env.setup()
    import dill
    from libs import bmpslib
    _orig = bmpslib._perf_svd
    bmpslib._perf_svd = lambda m, svd_emthod="svd", check_result=False: _orig(m, "svd")
    from algo.belief_propagation import belief_propagation
    from algo.tn_reduction import reduce_core_to_mode, reduce_full_kagome_to_core, reduce_mode_to_edge
    from containers import UpdateEdge
    from enums import MessageModel, UpdateMode
    from lattices.directions import BlockSide
    from libs.ITE import rho_ij
    from physics.hamiltonians import heisenberg_afm
    from tensor_networks.construction import kagome_tn_from_unit_cell

    h = np.asarray(heisenberg_afm())
    for D, fname, Ns in ((2, "D=2 energy=-0.4046412208223448.dat", (2, 3)), (3, "D=3 energy=-0.41420215914597786.dat", (2,))):
        with open(os.path.join(os.path.dirname(BEST), fname), "rb") as f:
            best = dill.load(f)
        uc = best.unit_cell
        out = {"A": np.asarray(uc.A), "B": np.asarray(uc.B), "C": np.asarray(uc.C), "h": h, "file_energy": np.array([float(best.mean_energy)])}
        for N in Ns:
            cfg = ref_env.quiet_config(D, N)
            cfg.bp.init_msg = MessageModel("UQ")
            cfg.bp.msg_diff_terminate = 1e-10
            tn = kagome_tn_from_unit_cell(uc, cfg.dims)
            tn.connect_uniform_messages()
            msgs, stats = belief_propagation(tn, tn.messages, cfg.bp)
            core = reduce_full_kagome_to_core(tn, cfg.contraction, direction=BlockSide.U)
            mode_tn = reduce_core_to_mode(core, UpdateMode.A)
            energies = []
            for e in UpdateEdge.all_options():
                et = reduce_mode_to_edge(mode_tn, e, cfg.contraction, arange_legs=False)
                et.rearrange_tensors_and_legs_into_canonical_order()
                t1, t2, env = et.edge_and_environment()
                rdm = rho_ij(t1, t2, mps_env=env)
                energies.append(float(np.real(np.dot(rdm.flatten(), h.flatten()))))
            out[f"N{N}_edge_energies"] = np.array(energies)
            out[f"N{N}_edges"] = np.array([str(e).replace("(", "").replace(")", "").replace(", ", "").replace(" ", "") for e in UpdateEdge.all_options()])
            out[f"N{N}_cfg"] = np.array([cfg.bp.trunc_dim, cfg.contraction.trunc_dim, stats.iterations, 1e-10, cfg.bp.damping if cfg.bp.damping is not None else -1.0], dtype=float)
            print(f"D={D} N={N}: BP {stats.iterations} iterations (err {stats.final_error:.2e}), chi_bp {cfg.bp.trunc_dim}, chi {cfg.contraction.trunc_dim}, "
                  f"energy per site {sum(energies) / 3:+.10f}  (file: {float(best.mean_energy):+.10f})")
        np.savez_compressed(os.path.join(GOLD, f"best_D{D}.npz"), **out)
        print("wrote", os.path.join(GOLD, f"best_D{D}.npz"))


if __name__ == "__main__":
    main()
