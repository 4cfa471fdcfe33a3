"""Host-side profile (cProfile, cumulative) of ITE steps on the device: where the wall time of the reduction and the
gate / ALS update goes.  usage: python tools/ite_profile.py D N [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kagomeperiodicbp_b200 import edge_env, ite_flow
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell

D, N = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
chi = 2 * D * D + 10
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ", max_iterations=50)
cell, msgs = UnitCell.random(2, D, seed=0), None
pr = cProfile.Profile()
for it in range(steps):
    if it == 3:
        pr.enable()
    t0 = time.perf_counter()
    cell, msgs, energy, st = ite_flow.ite_edge_update(cell, msgs, N, "A", edge_env.EDGES[it % 6], 1e-2, cfg, chi)
    print(f"step {it}: {1e3 * (time.perf_counter() - t0):.1f} ms (bp {1e3 * st.t_bp:.1f}, reduce {1e3 * st.t_reduce:.1f}, update {1e3 * st.t_update:.1f}; "
          f"ALS {st.als_iterations} it) backend calls {ite_flow.backend().calls}", flush=True)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
print(s.getvalue()[:14000])
