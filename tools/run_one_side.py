"""Debug tool: run one side program in the main thread.  python tools/run_one_side.py D N side seed"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.containers import UnitCell
from kagomeperiodicbp_b200.engine import Engine
D, N, side, seed = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
cell = UnitCell.random(2, D, seed=seed)
msgs = bp.initial_messages(D, N, "UQ")
comp = bp.compile_side_program(N, 2, D, side, 2 * D * D, bp._msg_shapes(msgs), 0.1)
eng = Engine(0)
t = time.time()
outs, slots, rc = comp.run(eng, [bp._side_inputs(cell, msgs, comp)], soft_errors=(-4,))
print("ok rc", rc, "time", time.time() - t, "slots", slots[0][:6], "launches", eng.launch_count())
