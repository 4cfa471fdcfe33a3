import sys, os
sys.path.insert(0, '/root/repo')
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
D, N, seed = 4, 6, int(sys.argv[1])
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
cell = UnitCell.random(2, D, seed=seed)
msgs = bp.initial_messages(D, N, "UQ")
for it in range(3):
    print(f"--- iteration {it}", file=sys.stderr, flush=True)
    out, msgs, err, _ = bp.bp_step_batch(N, [cell], [msgs], cfg)[0]
    print(f"iteration {it}: error {err:.3e}", file=sys.stderr, flush=True)
