"""Warm, in-stream time per ZGEMM launch for the shapes the subspace SVD uses.  usage: python tools/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kagomeperiodicbp_b200.engine import Engine
from kagomeperiodicbp_b200.program import Program
from kagomeperiodicbp_b200.runtime import Compiled

eng = Engine(0)
rng = np.random.default_rng(0)
shapes = [(512, 80, 512, 0, 0, 200), (512, 80, 512, 2, 0, 200), (80, 80, 512, 2, 0, 200), (512, 80, 80, 0, 0, 200),
          (32, 512, 80, 0, 2, 200), (512, 32, 512, 0, 2, 200), (512, 512, 32, 0, 0, 200), (8192, 512, 16, 0, 0, 50)]
if len(sys.argv) > 1 and sys.argv[1] == "d4":        # the products of one subspace iteration at D = 4 (block 64) and its Rayleigh-Ritz step
    shapes = [(512, 64, 512, 0, 0, 200), (512, 64, 512, 2, 0, 200), (64, 64, 512, 2, 0, 200), (256, 64, 512, 0, 0, 200), (512, 64, 256, 2, 0, 200),
              (32, 512, 64, 0, 2, 200), (512, 32, 512, 0, 2, 200), (32, 512, 512, 2, 0, 200), (512, 512, 32, 0, 0, 200), (64, 64, 64, 0, 0, 200)]
if len(sys.argv) > 1 and sys.argv[1] == "ksweep":
    shapes = [(512, 80, k, 0, 0, 200) for k in (16, 64, 128, 256, 512, 1024, 2048)] + [(2048, 80, k, 0, 0, 100) for k in (64, 512)]
for (m, n, k, oa, ob, reps) in shapes:
    p = Program(8)
    sa = (m, k) if oa in (0, 3) else (k, m)
    sb = (k, n) if ob in (0, 3) else (n, k)
    a, b = p.input("a", sa), p.input("b", sb)
    cs = [p.matmul(a, b, m, n, k, oa, ob) for _ in range(reps)]
    comp = Compiled(p, [("a", a), ("b", b)], [("c", cs[-1])])
    comp.load(eng, 1)
    eng.upload(0, comp.pack_inputs([{"a": rng.normal(size=sa) + 0j, "b": rng.normal(size=sb) + 0j}]))
    for _ in range(3):                        # first sight (host-driven), capture, first replay
        comp.run_resident(eng); eng.sync()
    eng.timer_start()
    comp.run_resident(eng)
    ms = eng.timer_stop_ms()
    print(f"gemm {m}x{n}x{k} op({oa},{ob}): {1e3*ms/reps:7.2f} us per launch  ({8*m*n*k/(ms/reps*1e-3)/1e12:.2f} TF/s)")
