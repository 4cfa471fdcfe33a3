"""Debug tool: run a compiled tensor program ONE OP AT A TIME on the GPU and check every op's output against
numpy applied to the GPU's own inputs of that op (gauge-free invariants for QR / SVD).  Localises the first
op whose result is off.   usage: python tools/lockstep.py D N side"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kagomeperiodicbp_b200 import belief_propagation as bp  # noqa: E402
from kagomeperiodicbp_b200.containers import UnitCell  # noqa: E402
from kagomeperiodicbp_b200.engine import Engine  # noqa: E402
from kagomeperiodicbp_b200.program import _prod  # noqa: E402


def split_ops(w):
    i, ops = 0, []
    fixed = {2: 9, 3: 7, 4: 11, 5: 4, 6: 12, 7: 3, 8: 4, 9: 4, 10: 4}
    while i < len(w):
        op = int(w[i])
        n = 5 + 2 * int(w[i + 4]) if op == 1 else fixed[op]
        ops.append(w[i:i + n])
        i += n
    return ops


def main():
    D, N, side = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    cell = UnitCell.random(2, D, seed=int(os.environ.get("LOCKSTEP_SEED", "1234")))
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    comp = bp.compile_side_program(N, 2, D, side, 2 * D * D, bp._msg_shapes(tn.messages), 0.1)
    eng = Engine(0)
    comp.load(eng, 1)
    eng.upload(0, comp.pack_inputs([bp._side_inputs(cell, tn.messages, comp)]))
    get = lambda off, n: eng.download(int(off), int(n), chain=0)
    worst = {}
    wpos = 0
    for k, o in enumerate(split_ops(comp.words)):
        op = int(o[0])
        wpos_here = wpos; wpos += len(o)
        pre = None
        if op == 5:
            pre = get(o[1], o[2])
        if wpos_here > int(os.environ.get('LOCKSTEP_MAX_WORD', '1000000000')):
            break
        if op == 4 and wpos_here == int(os.environ.get('LOCKSTEP_DUMP_WORD', '-1')):
            A_, m_, n_ = int(o[1]), int(o[5]), int(o[6])
            os.makedirs('gpurun_out', exist_ok=True)
            np.save('gpurun_out/svd_in.npy', get(A_, m_ * n_).reshape(m_, n_))
            print('dumped svd input at word', wpos_here, flush=True)
        eng.run(np.array(o, dtype=np.int64), soft_errors=(-4,))
        err, what = 0.0, ""
        if op == 1:
            nd = int(o[4]); dims = [int(x) for x in o[5:5 + nd]]; perm = [int(x) for x in o[5 + nd:5 + 2 * nd]]
            n = _prod(dims)
            x = get(o[2], n).reshape(dims).transpose(perm)
            x = np.conj(x) if o[3] else x
            err = np.abs(get(o[1], n) - x.ravel()).max(); what = f"permute {dims}"
        elif op == 2:
            C, A, B, m, n, kk, oa, ob = [int(x) for x in o[1:9]]
            def mat(off, r, c, opx):
                a = get(off, r * c)
                M = a.reshape(r, c) if opx in (0, 3) else a.reshape(c, r).T
                return np.conj(M) if opx in (2, 3) else M
            e = mat(A, m, kk, oa) @ mat(B, kk, n, ob)
            err = np.abs(get(C, m * n).reshape(m, n) - e).max() / max(1e-300, np.abs(e).max()); what = f"gemm {m}x{n}x{kk} ops {oa}{ob}"
        elif op == 3:
            A, Q, R, wk, m, n = [int(x) for x in o[1:7]]
            kk = min(m, n)
            a = get(A, m * n).reshape(m, n); q = get(Q, m * kk).reshape(m, kk); r = get(R, kk * n).reshape(kk, n)
            err = max(np.abs(q @ r - a).max() / max(1e-300, np.abs(a).max()), np.abs(q.conj().T @ q - np.eye(kk)).max()); what = f"qr {m}x{n}"
        elif op == 4:
            A, US, Vh, wk, m, n, keep, nrb = [int(x) for x in o[1:9]]
            a = get(A, m * n).reshape(m, n); us = get(US, m * keep).reshape(m, keep); vh = get(Vh, keep * n).reshape(keep, n)
            u, s, v = np.linalg.svd(a, full_matrices=False)
            ref = (u[:, :keep] * s[:keep]) @ v[:keep]
            fro = np.linalg.norm(s)
            got = us @ vh * (fro if nrb else 1.0)
            err = np.linalg.norm(got - ref) / fro
            gap = (s[keep - 1] - s[keep]) / s[0] if keep < len(s) else 1.0
            what = f"svd {m}x{n} keep {keep} relgap {gap:.1e} s_keep/s0 {s[keep-1]/s[0]:.1e}"
        elif op == 5:
            x = get(o[1], o[2])
            err = np.abs(x - pre / np.linalg.norm(pre)).max(); what = f"normalize {int(o[2])}"
        elif op == 10:
            r, c_ = int(o[2]), int(o[3])
            err = np.abs(get(o[1], r * c_).reshape(r, c_) - np.eye(r, c_)).max(); what = f"eye {r}x{c_}"
        else:
            continue
        if not np.isfinite(err):
            print(f"op #{k} (word {wpos_here}): {what}: NON-FINITE result"); break
        tag = what.split()[0]
        worst[tag] = max(worst.get(tag, 0.0), err)
        if err > 1e-11:
            print(f"op #{k} (word {wpos_here}): {what}: err {err:.3e}", flush=True)
    print("worst per op type:", {k: f"{v:.2e}" for k, v in worst.items()})


if __name__ == "__main__":
    main()
