"""Debug tool: contract growing prefixes of a swallow order on the GPU and with the oracle; report the first
swallow after which the (gauge-invariant) dense boundary state differs.  usage: prefix_check.py D N side"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from kagomeperiodicbp_b200 import belief_propagation as bp, block_tn, contraction_order  # noqa: E402
from kagomeperiodicbp_b200.bubblecon import bubblecon  # noqa: E402
from kagomeperiodicbp_b200.containers import UnitCell  # noqa: E402
from kagomeperiodicbp_b200.lattice import SIDE_ANGLE  # noqa: E402
from oracle import bubblecon_np, mps_np  # noqa: E402
from helpers import to_oracle_mps  # noqa: E402

D, N, side = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
cell = UnitCell.random(2, D, seed=1234)
tn = bp.KagomeTNRepeatedUnitCell(cell, N)
tn.connect_uniform_messages()
msgs = {s: m.mps.A for s, m in tn.messages.items()}
T, E, A, K, P = block_tn.assemble(N, cell.tensors(), msgs)
T, E, A = block_tn.connect_corner(N, T, E, A, P, side)
order = list(contraction_order.kagome_order(N, side, "ToMessage"))
chi = 2 * D * D
for k in range(2, len(order) + 1):
    mp = bubblecon(T, E, A, SIDE_ANGLE[side], order[:k], D_trunc=chi, ket_tensors=K)
    ref = bubblecon_np.bubblecon(T, E, A, SIDE_ANGLE[side], order[:k], D_trunc=chi, ket_tensors=K)
    if not hasattr(mp, "A"):
        continue
    a = mps_np.mps_to_dense(ref)
    b = mps_np.mps_to_dense(to_oracle_mps(mp))
    ph = np.vdot(a, b); ph /= abs(ph)
    d = np.linalg.norm(a * ph - b) / np.linalg.norm(a)
    print(k, order[k - 1], [x.shape for x in mp.A], f"{d:.2e}", flush=True)
    if d > 1e-9:
        break
