"""Run BP iterations on the device and report which SVD path each truncation took (debugging aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
from kagomeperiodicbp_b200.lattice import BLOCK_SIDES_CCW
from kagomeperiodicbp_b200.runtime import get_engine

D, N, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
cell = UnitCell.random(2, D, seed=0)
msgs = bp.initial_messages(D, N, "UQ")
for it in range(iters):
    t0 = time.perf_counter()
    out, msgs, err, trunc = bp.bp_step_batch(N, [cell], [msgs], cfg)[0]
    dt = time.perf_counter() - t0
    tot = {}
    for s in BLOCK_SIDES_CCW:
        for k, v in get_engine(("side", s)).svd_counters().items():
            tot[k] = tot.get(k, 0) + v
    print(f"iter {it} err {err:.3e} trunc {trunc:.3e} {dt*1e3:.1f} ms cumulative {tot}", flush=True)
