"""Are the message tensors themselves (not just the states) continuous across BP iterations?  (debugging aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
D, N, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
cell = UnitCell.random(2, D, seed=0)
msgs = bp.initial_messages(D, N, "UQ")
prev = None
for it in range(iters):
    out, msgs, err, trunc = bp.bp_step_batch(N, [cell], [msgs], cfg)[0]
    if prev is not None:
        d = []
        for s in msgs:
            for a, b in zip(msgs[s].mps.A, prev[s].mps.A):
                d.append(np.max(np.abs(a - b)) if a.shape == b.shape else np.nan)
        print(f"iter {it} err {err:.2e} max elementwise change of next-message sites: {np.nanmax(d):.3e}  median {np.nanmedian(d):.3e}")
    prev = msgs
