"""Mini counterpart of the reference's scripts/run_ite.py on the device: full-update imaginary time evolution of a random
unit cell (Kagome Heisenberg AFM), modes A, B, C x the six edges per sweep, decreasing time steps; prints energy per site.
usage: python tools/run_ite.py D N sweeps_per_dt [dt ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kagomeperiodicbp_b200 import edge_env, ite_flow
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell

D, N, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dts = [float(x) for x in sys.argv[4:]] or [0.1, 0.05, 0.02, 0.01]
chi = 2 * D * D + 10
cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ", max_iterations=50)
cell = UnitCell.random(2, D, seed=0)
msgs = None
t0 = time.perf_counter()
steps = 0
for dt in dts:
    for sw in range(sweeps):
        for mode in edge_env.MODES:
            order = [(e, dt) for e in edge_env.EDGES]
            cell, msgs, energies, stats = ite_flow.ite_per_mode(cell, msgs, N, mode, order, cfg, chi)
            steps += len(order)
        m = ite_flow.measure_energies(cell, msgs, N, chi, mode="A")
        print(f"dt {dt:g} sweep {sw}: energy/site {m.mean_energy:+.6f}  ({steps} edge updates, {time.perf_counter()-t0:.1f} s, "
              f"{1e3*(time.perf_counter()-t0)/steps:.1f} ms/step)", flush=True)
