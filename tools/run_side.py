"""Debug tool: run two BP iterations of one configuration (python tools/run_side.py D N), printing per-iteration error."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
D, N = int(sys.argv[1]), int(sys.argv[2])
cell = UnitCell.random(2, D, seed=0)
msgs = bp.initial_messages(D, N, "UQ")
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    t = time.time()
    out, msgs, err, trunc = bp.bp_step_batch(N, [cell], [msgs], cfg)[0]
    print(f"iteration {it}: error {err:.3e} trunc {trunc:.3e} {time.time()-t:.2f} s", flush=True)
