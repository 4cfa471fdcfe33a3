"""Micro-benchmark of the batched truncated SVD op (KBP_OP_SVD) through the C ABI.
usage: python tools/svd_bench.py m n keep nb reps [spectrum: random|decay]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kagomeperiodicbp_b200.engine import Engine  # noqa: E402
from kagomeperiodicbp_b200.program import Program  # noqa: E402
from kagomeperiodicbp_b200.runtime import Compiled  # noqa: E402

m, n, keep, nb, reps = (int(x) for x in sys.argv[1:6])
spec = sys.argv[6] if len(sys.argv) > 6 else "decay"
rng = np.random.default_rng(0)
mats = []
for c in range(nb):
    a = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
    if spec == "decay":
        u, s, vh = np.linalg.svd(a, full_matrices=False)
        a = (u * np.exp(-0.12 * np.arange(len(s)))) @ vh      # ~ boundary-MPS like: sigma_32/sigma_0 ~ 2e-2
    mats.append(a)
p = Program(8)
t = p.input("a", (m, n))
us, vh = p.svd_trunc(t, keep, True, 0, 1)
comp = Compiled(p, [("a", t)], [("us", us), ("vh", vh)])
eng = Engine(0)
comp.load(eng, nb)
eng.upload(0, comp.pack_inputs([{"a": a} for a in mats]))
comp.run_resident(eng)
eng.sync()
s0, l0 = eng.svd_sweeps(), eng.launch_count()
eng.timer_start()
for _ in range(reps):
    comp.run_resident(eng)
ms = eng.timer_stop_ms() / reps
sw = (eng.svd_sweeps() - s0) / reps
flops = 4.0 * (14.0 * max(m, n) * min(m, n) ** 2 + 8.0 * min(m, n) ** 3) * nb
print(f"svd {m}x{n} keep {keep} nb {nb} [{spec}]: {ms:.2f} ms per batched SVD, {sw:.1f} sweeps, {(eng.launch_count()-l0)/reps:.0f} launches, "
      f"{flops / ms / 1e9:.3f} TFLOP/s algorithmic")
outs = eng.download(comp.out_block.off, comp.out_elems)
a = mats[0]
u, s, v = np.linalg.svd(a, full_matrices=False)
ref = (u[:, :keep] * s[:keep]) @ v[:keep]
us0 = outs[0, :m * keep].reshape(m, keep)
q = (m * keep + 7) // 8 * 8
vh0 = outs[0, q:q + keep * n].reshape(keep, n)
print("  accuracy vs numpy (chain 0):", np.linalg.norm(us0 @ vh0 * np.linalg.norm(s) - ref) / np.linalg.norm(s))
