"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (exact-SVD branch) in the build
container.  The fixtures travel to the GPU box; the reference does not.

    python tools/make_golden.py            # regenerate everything
    python tools/make_golden.py --check    # also compare oracle/ against the fresh reference output

What is stored
  geometry_N{2..5}.npz   node edges/angles/positions, sorted boundary edges, swallow orders
                         (ToMessage, ToCore, Full) for all six sides
  chain_D{D}_N{N}.npz    one ToMessage bubblecon call per side on a seeded unit cell with uniform
                         (UQ) messages: output MPS sites + (mantissa, exp10)
  bp_D{D}_N{N}[_damp].npz  belief_propagation(...) to convergence: per-iteration errors, iteration count,
                         final (hermitised) messages as dense vectors where tractable, else MPS sites
"""
from __future__ import annotations

import argparse
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


def _ref_modules():
    ref_env.setup()
    from libs import bmpslib
    _orig = bmpslib._perf_svd
    bmpslib._perf_svd = lambda m, svd_emthod="svd", check_result=False: _orig(m, "svd")  # pin to the numpy branch
    return bmpslib


def _ref_tn(D, N, seed=SEED):
    from containers import Config
    from tensor_networks.construction import kagome_tn_from_unit_cell
    uc = ref_env.seeded_unit_cell(D, seed)
    config = Config.derive_from_dimensions(D)
    config.dims.big_lattice_size = N
    config.bp.visuals.set_all_progress_bars(False)
    tn = kagome_tn_from_unit_cell(uc, config.dims)
    return uc, config, tn


def make_geometry():
    from algo.contraction_order import get_contraction_order
    from enums import ContractionDepth
    from lattices.directions import BlockSide
    from tensor_networks import tensor_network as tnmod
    for N in (2, 3, 4, 5):
        uc, config, tn = _ref_tn(2, N)
        tn.connect_uniform_messages()
        out = {}
        out["edges"] = np.array(["|".join(map(str, e)) for e in tn.edges_list])
        out["angles"] = np.array([np.pad(np.array(a, float), (0, 4 - len(a)), constant_values=-1) for a in tn.angles])
        out["positions"] = np.array(tn.positions, float)
        out["kets"] = np.array(tn.kets)
        for side in BlockSide.all_in_counter_clockwise_order():
            out[f"boundary_{side}"] = np.array(tn.lattice.sorted_boundary_edges(side))
            for depth in (ContractionDepth.ToMessage, ContractionDepth.ToCore, ContractionDepth.Full):
                # the reference reverses a cached list in place for Full (contraction_order.py:583-586 on the
                # list cached by tensor_network.py:1005-1013); record its clean-cache behaviour
                tnmod._kagome_lattice_derive_message_indices.cache_clear()
                out[f"order_{side}_{depth.name}"] = np.array(get_contraction_order(tn, side, depth))
            tnmod._kagome_lattice_derive_message_indices.cache_clear()
        np.savez_compressed(os.path.join(GOLD, f"geometry_N{N}.npz"), **out)
        print("geometry", N)


def make_chain(D, N, check, sides=None):
    from algo.contract_tensor_network import contract_tensor_network
    from enums import ContractionDepth
    from lattices.directions import BlockSide
    uc, config, tn = _ref_tn(D, N)
    tn.connect_uniform_messages()
    out = {"A": uc.A, "B": uc.B, "C": uc.C, "chi": config.bp.trunc_dim}
    for side in BlockSide.all_in_counter_clockwise_order():
        if sides is not None and str(side) not in sides:
            continue
        mps, _, _ = contract_tensor_network(tn, side, ContractionDepth.ToMessage, config.bp.trunc_dim, allow_progressbar=False)
        for k, a in enumerate(mps.A):
            out[f"{side}_site{k}"] = a
        out[f"{side}_nr"] = np.array([mps.nr_mantissa, mps.nr_exp], float)
        if check:
            from oracle import bp_np, mps_np
            mine = bp_np.outgoing_message(N, (uc.A, uc.B, uc.C), bp_np.uniform_messages(N, D), str(side), config.bp.trunc_dim)
            r = mps_np.MPS(mps.N); r.A = mps.A; r.nr_mantissa, r.nr_exp = mps.nr_mantissa, mps.nr_exp
            if N <= 3:
                a, b = mps_np.mps_to_dense(r), mps_np.mps_to_dense(mine)
                print(f"  chain D={D} N={N} {side}: rel diff {np.linalg.norm(a - b) / np.linalg.norm(a):.2e}")
            else:
                ov = mps_np.mps_inner_product(r, mine, True) / np.sqrt(abs(mps_np.mps_inner_product(r, r, True) * mps_np.mps_inner_product(mine, mine, True)))
                print(f"  chain D={D} N={N} {side}: 1-|overlap| {1 - abs(ov):.2e}")
    np.savez_compressed(os.path.join(GOLD, f"chain_D{D}_N{N}.npz"), **out)
    print("chain", D, N)


def make_bp(D, N, damping, check, terminate=1e-6):
    from algo import belief_propagation as ref_bp
    from enums import MessageModel
    uc, config, tn = _ref_tn(D, N)
    config.bp.damping = damping
    config.bp.msg_diff_terminate = terminate
    config.bp.init_msg = MessageModel("UQ")
    tn.connect_uniform_messages()
    msgs0 = tn.messages
    errors = []
    _step = ref_bp._belief_propagation_step

    def spy(*a, **k):
        r = _step(*a, **k)
        errors.append(float(r[2]))
        return r
    ref_bp._belief_propagation_step = spy
    try:
        msgs, stats = ref_bp.belief_propagation(tn, msgs0, config.bp)
    finally:
        ref_bp._belief_propagation_step = _step
    out = {"A": uc.A, "B": uc.B, "C": uc.C, "chi": config.bp.trunc_dim, "damping": -1.0 if damping is None else damping,
           "terminate": terminate, "errors": np.array(errors), "iterations": stats.iterations,
           "final_error": float(stats.final_error), "success": bool(stats.success)}
    for side, m in msgs.items():
        for k, a in enumerate(m.mps.A):
            out[f"{side}_site{k}"] = a
        out[f"{side}_nr"] = np.array([m.mps.nr_mantissa, m.mps.nr_exp], float)
    tag = f"bp_D{D}_N{N}" + ("_damp" if damping else "")
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    print(tag, "iterations", stats.iterations, "errors", errors)
    if check:
        from oracle import bp_np, mps_np
        cfg = bp_np.BPConfigNP(trunc_dim=config.bp.trunc_dim, msg_diff_terminate=terminate, damping=damping)
        mine, st = bp_np.belief_propagation(N, (uc.A, uc.B, uc.C), bp_np.uniform_messages(N, D), cfg)
        print("  oracle iterations", st["iterations"], "final_error", st["final_error"], "ref", float(stats.final_error))
        for side, m in msgs.items():
            r = mps_np.MPS(m.mps.N); r.A = m.mps.A; r.nr_mantissa, r.nr_exp = m.mps.nr_mantissa, m.mps.nr_exp
            a, b = mps_np.mps_to_dense(r), mps_np.mps_to_dense(mine[str(side)])
            print(f"  bp {side}: rel diff {np.linalg.norm(a - b) / np.linalg.norm(a):.2e}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    _ref_modules()
    if args.only in ("", "geometry"):
        make_geometry()
    if args.only in ("", "chain"):
        make_chain(2, 2, args.check)
        make_chain(2, 3, args.check)
        make_chain(3, 2, args.check)
        make_chain(3, 3, args.check)
    if args.only == "d4":       # the benchmarked bond dimension (CPU-heavy: minutes); D=4, N=3 keeps two sides to bound the fixture size
        make_chain(4, 2, args.check)
        make_bp(4, 2, 0.1, args.check)
        make_chain(4, 3, args.check, sides=("D", "UR"))
    if args.only == "d6":       # config C4 (chi_bp = 72): one side, tens of minutes of LAPACK on 2592 x 2592 matrices
        make_chain(6, 2, args.check, sides=("D",))
    if args.only in ("", "bp"):
        make_bp(2, 2, None, args.check)
        make_bp(2, 2, 0.1, args.check)
        make_bp(2, 3, 0.1, args.check)
        make_bp(3, 2, 0.1, args.check)
