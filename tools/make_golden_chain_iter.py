"""Fixture for bubblecon with ITERATIVE compression (compression={'type': 'iter', ...}: src/libs/bubblecon.py:2612-2624,
2793-2798, 3035-3038 -> bmpslib.mps.reduceDiter): the arguments of one ToMessage bubblecon call of the reference on a
seeded D=2, N=2 block with uniform messages are captured, and the UNMODIFIED reference bubblecon is run on them with the
iterative compressor.

    python tools/make_golden_chain_iter.py    ->  tests/golden/chain_iter_D2_N2.npz
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import ref_env  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
COMPRESSION = {"type": "iter", "max-iter": 20, "err": 1e-8}


def main():
    ref_env.setup()
    import algo.contract_tensor_network as ctn
    from algo.contract_tensor_network import contract_tensor_network
    from enums import ContractionDepth
    from lattices.directions import BlockSide
    from tensor_networks.construction import kagome_tn_from_unit_cell
    D, N = 2, 2
    uc = ref_env.seeded_unit_cell(D, 1234)
    cfg = ref_env.quiet_config(D, N)
    tn = kagome_tn_from_unit_cell(uc, cfg.dims)
    tn.connect_uniform_messages()
    captured = {}
    orig = ctn.bubblecon

    def spy(T_list, edges_list, angles_list, bubble_angle, swallow_order, **kw):
        captured.update(T_list=T_list, edges_list=edges_list, angles_list=angles_list, bubble_angle=bubble_angle,
                        swallow_order=swallow_order, kw=kw)
        return orig(T_list, edges_list, angles_list, bubble_angle, swallow_order, **kw)

    out = {}
    for side in (BlockSide.D, BlockSide.UL):
        ctn.bubblecon = spy
        try:
            contract_tensor_network(tn, side, ContractionDepth.ToMessage, cfg.bp.trunc_dim, allow_progressbar=False)
        finally:
            ctn.bubblecon = orig
        kw = dict(captured["kw"])
        kw["compression"] = dict(COMPRESSION)
        kw["progress_bar"] = False
        mp = orig(captured["T_list"], captured["edges_list"], captured["angles_list"], captured["bubble_angle"],
                  captured["swallow_order"], **kw)
        s = str(side)
        for i, t in enumerate(captured["T_list"]):
            out[f"{s}/T{i}"] = np.asarray(t)
        meta = dict(edges=[[str(e) for e in es] for es in captured["edges_list"]],
                    angles=[[float(a) for a in x] for x in captured["angles_list"]],
                    bubble_angle=float(captured["bubble_angle"]), order=[int(v) for v in captured["swallow_order"]],
                    kets=[bool(k) for k in kw.get("ket_tensors")], D_trunc=int(kw.get("D_trunc")), compression=COMPRESSION)
        out[f"{s}/meta"] = np.array(json.dumps(meta))
        for k, a in enumerate(mp.A):
            out[f"{s}/out{k}"] = np.asarray(a)
        out[f"{s}/nr"] = np.array([mp.nr_mantissa, mp.nr_exp], float)
        out[f"{s}/corder"] = np.array([c if c is not None else "" for c in mp.Corder])
        print(s, [a.shape for a in mp.A], mp.Corder, mp.nr_mantissa, mp.nr_exp)
    np.savez_compressed(os.path.join(GOLD, "chain_iter_D2_N2.npz"), **out)
    print("wrote chain_iter_D2_N2.npz")


if __name__ == "__main__":
    main()
