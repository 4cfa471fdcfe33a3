"""Device time per opcode of the resident ITE backend (reduction to the edge + gate / ALS update) for consecutive ITE steps.
usage: python tools/ite_op_profile.py D N [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kagomeperiodicbp_b200 import edge_env, ite_flow
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell

D, N = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
chi = 2 * D * D + 10
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ", max_iterations=50)
cell, msgs = UnitCell.random(2, D, seed=0), None
names = {1: "permute", 2: "gemm", 3: "qr", 4: "svd", 5: "normalize", 6: "embed", 7: "zero", 10: "eye"}
eng = ite_flow.backend().eng
for it in range(steps):
    prof = it >= 2
    if prof:
        eng.profile_enable(True)
    t0 = time.perf_counter()
    cell, msgs, energy, st = ite_flow.ite_edge_update(cell, msgs, N, "A", edge_env.EDGES[it % 6], 1e-2, cfg, chi)
    dt = time.perf_counter() - t0
    line = f"step {it} edge {edge_env.EDGES[it % 6]}: {1e3 * dt:.0f} ms (bp {1e3 * st.t_bp:.0f}, reduce {1e3 * st.t_reduce:.0f}, update {1e3 * st.t_update:.0f})"
    if prof:
        ms, cnt = eng.profile_read()
        eng.profile_enable(False)
        line += "   resident-backend device ms per op: " + ", ".join(f"{names[k]} {ms[k]:.1f} ({int(cnt[k])})" for k in names if cnt[k])
    print(line, flush=True)
