"""Time ITE steps (one loop body of ite_per_mode) on the device.  usage: python tools/ite_bench.py D N steps"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kagomeperiodicbp_b200 import edge_env, ite_flow
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell

D, N, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
chi = 2 * D * D + 10
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ", max_iterations=50)
cell = UnitCell.random(2, D, seed=0)
msgs = None
k = 0
for it in range(steps):
    e = edge_env.EDGES[k % 6]
    k += 1
    t0 = time.perf_counter()
    cell, msgs, energy, st = ite_flow.ite_edge_update(cell, msgs, N, "A", e, 1e-2, cfg, chi)
    dt = time.perf_counter() - t0
    print(f"step {it} edge {e}: {dt*1e3:8.1f} ms  (bp {st.t_bp*1e3:.1f} ms / {st.bp_iterations} it, reduce {st.t_reduce*1e3:.1f}, update {st.t_update*1e3:.1f}; "
          f"ALS {st.als_iterations} it, dist {st.truncation_distance:.2e})  energy {energy:+.8f}  backend calls {ite_flow.backend().calls}", flush=True)
