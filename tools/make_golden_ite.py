"""Generate the ITE-step fixtures and the topology tables of the 21-node core network by running the UNMODIFIED
reference (exact-SVD branch) in the build container.  Nothing here runs on the GPU box.

    python tools/make_golden_ite.py

Writes
  kagomeperiodicbp_b200/core_tables.json
      data only (no code): for every update mode the 13 nodes of the ModeTN as the reference orders them (name, member
      core/env indices, edge names, leg angles, ket flag), and for the six (mode, edge) cases that the reference closes with a
      truncated ToEdge boundary contraction (src/algo/tn_reduction/mode_to_edge.py:166-218) the swallow order, the bubble
      angle and where the four MPS sites attach.  This is topology of the fixed 9-site core: independent of D, N and data.
  tests/golden/ite_D{D}_N2.npz
      seeded unit cell, the reference's converged BP messages, the 12 environment tensors of the core, and per (mode, edge):
      two-site RDM, edge energy, eigenvalues of N_red, and the updated pair tensor after apply_2local_gate
      (src/libs/ITE.py:1761) contracted over its shared bond.  For D = 2 also the inputs (T_i, T_j, mps_env) of the S4 seam.
"""
from __future__ import annotations

import json
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
SEED = 1234
DT = 0.01


def main():
    ref_env.setup()
    from libs import bmpslib
    _orig = bmpslib._perf_svd
    bmpslib._perf_svd = lambda m, svd_emthod="svd", check_result=False: _orig(m, "svd")
    import algo.contract_tensor_network as ctn
    from algo.belief_propagation import belief_propagation
    from algo.imaginary_time_evolution._tn_update import get_imaginary_time_evolution_operator
    from algo.tn_reduction import reduce_core_to_mode, reduce_full_kagome_to_core, reduce_mode_to_edge
    from containers import UpdateEdge
    from enums import MessageModel, UpdateMode
    from lattices.directions import BlockSide
    from libs.ITE import apply_2local_gate, rho_ij
    from physics.hamiltonians import heisenberg_afm
    from tensor_networks.construction import kagome_tn_from_unit_cell

    calls = []
    orig_bubblecon = ctn.bubblecon

    def spy(T_list, edges_list, angles_list, bubble_angle, swallow_order, **kw):
        calls.append(dict(edges=[list(map(str, e)) for e in edges_list], angles=[[float(a) for a in x] for x in angles_list],
                          bubble_angle=float(bubble_angle), order=[int(v) for v in swallow_order],
                          kets=[bool(k) for k in kw.get("ket_tensors")], D_trunc=kw.get("D_trunc")))
        return orig_bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, **kw)

    tables = {}
    for D in (2, 3):
        N = 2
        uc = ref_env.seeded_unit_cell(D, SEED)
        cfg = ref_env.quiet_config(D, N)
        cfg.bp.init_msg = MessageModel("UQ")
        cfg.bp.msg_diff_terminate = 1e-10
        tn = kagome_tn_from_unit_cell(uc, cfg.dims)
        msgs, stats = belief_propagation(tn, None, cfg.bp)
        print(f"D={D}: BP {stats.iterations} iterations, error {stats.final_error:.2e}")
        out = {"A": uc.A, "B": uc.B, "C": uc.C, "chi": cfg.contraction.trunc_dim, "chi_bp": cfg.bp.trunc_dim, "dt": DT}
        for side, m in msgs.items():
            for k, a in enumerate(m.mps.A):
                out[f"msg_{side}_{k}"] = a
        core = reduce_full_kagome_to_core(tn, cfg.contraction, direction=BlockSide.U)
        for n in core.nodes[9:]:
            out[f"core_env_{n.name}"] = n.tensor
        from libs.ITE import g_from_exp_h
        h = heisenberg_afm()
        g = g_from_exp_h(h, DT)
        out["h"], out["g"] = np.asarray(h), np.asarray(g)
        if D == 2:
            tables["core"] = [dict(index=n.index, name=n.name, ket=bool(n.is_ket), edges=list(map(str, n.edges)),
                                   angles=[float(a) for a in n.angles], pos=[float(x) for x in n.pos]) for n in core.nodes]
            tables["modes"] = {}
        ctn.bubblecon = spy
        for mode in (UpdateMode.A, UpdateMode.B, UpdateMode.C):
            mode_tn = reduce_core_to_mode(core, mode)
            if D == 2:
                tables["modes"][mode.name] = dict(
                    nodes=[dict(name=n.name, ket=bool(n.is_ket), edges=list(map(str, n.edges)), angles=[float(a) for a in n.angles])
                           for n in mode_tn.nodes],
                    center=int(mode_tn.center_node.index), edges={})
            for e in UpdateEdge.all_options():
                calls.clear()
                et = reduce_mode_to_edge(mode_tn, e, cfg.contraction, arange_legs=False)
                perms = et.rearrange_tensors_and_legs_into_canonical_order()
                t1, t2, env = et.edge_and_environment()
                key = f"{mode.name}_{e.first.name}{e.second.name}" if hasattr(e, "first") else f"{mode.name}_{str(e)}"
                key = key.replace("(", "").replace(")", "").replace(", ", "").replace(" ", "")
                rdm = rho_ij(t1, t2, mps_env=env)
                energy = np.dot(rdm.flatten(), np.asarray(h).flatten())
                t1n, t2n, eig = apply_2local_gate(g=g, Dmax=D, Ti=t1, Tj=t2, mps_env=env)
                t1n, t2n = t1n / np.linalg.norm(t1n), t2n / np.linalg.norm(t2n)
                pair = np.tensordot(t1n, t2n, axes=([1], [1]))
                if mode is UpdateMode.A:
                    from algo.measurements import compute_negativity_of_rdm, expectation_values_with_rdm
                    ev = expectation_values_with_rdm(rdm, force_real=True)
                    out[f"expect_{key}"] = np.array([[ev[x][0], ev[x][1]] for x in ("x", "y", "z")], dtype=float)
                    out[f"negativity_{key}"] = np.array([compute_negativity_of_rdm(rdm)], dtype=float)
                out[f"rdm_{key}"] = rdm
                out[f"energy_{key}"] = np.array([energy.real, energy.imag])
                out[f"pair_{key}"] = pair
                out[f"eig_{key}"] = np.asarray(eig)
                if D == 2:
                    out[f"in_ti_{key}"], out[f"in_tj_{key}"] = t1, t2
                    for k, x in enumerate(env):
                        out[f"in_env{k}_{key}"] = x
                    tables["modes"][mode.name]["edges"][key.split("_")[1]] = dict(
                        nodes=[dict(name=n.name, edges=list(map(str, n.edges)), index=int(n.index)) for n in et.nodes],
                        perms={k: [int(x) for x in v] for k, v in perms.items()},
                        bubblecon=(calls[0] if calls else None))
                print(f"  D={D} mode {mode.name} edge {key}: energy {energy.real:+.10f}  env bonds {[x.shape[0] for x in env]}"
                      f"{'  [ToEdge bubblecon]' if calls else ''}")
        ctn.bubblecon = orig_bubblecon
        # the other two bottom-up directions the reference may draw (kagome_to_core.py:68, 178-179)
        for dirn in (BlockSide.DL, BlockSide.DR):
            core_d = reduce_full_kagome_to_core(tn, cfg.contraction, direction=dirn)
            for n in core_d.nodes[9:]:
                out[f"core_env_{dirn}_{n.name}"] = n.tensor
            mode_tn = reduce_core_to_mode(core_d, UpdateMode.A)
            for e in UpdateEdge.all_options():
                et = reduce_mode_to_edge(mode_tn, e, cfg.contraction, arange_legs=False)
                et.rearrange_tensors_and_legs_into_canonical_order()
                t1, t2, env = et.edge_and_environment()
                key = str(e).replace("(", "").replace(")", "").replace(", ", "").replace(" ", "")
                rdm = rho_ij(t1, t2, mps_env=env)
                out[f"rdm_{dirn}_A_{key}"] = rdm
                out[f"energy_{dirn}_A_{key}"] = np.array([np.dot(rdm.flatten(), np.asarray(h).flatten()).real])
            print(f"  D={D} direction {dirn}: done")
        np.savez_compressed(os.path.join(GOLD, f"ite_D{D}_N2.npz"), **out)
    with open(os.path.join(ROOT, "kagomeperiodicbp_b200", "core_tables.json"), "w") as f:
        json.dump(tables, f, indent=1)
    print("done")


if __name__ == "__main__":
    main()
