"""Drop-in for the reference's ``bubblecon`` entry point (src/libs/bubblecon.py:2465-2468): same
positional/keyword arguments, same return conventions, executed as one device program.

    bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None, D_trunc2=None,
              eps=None, opt='high', break_points=[], ket_tensors=None, separate_exp=False, bubbleket=False,
              compression=None, progress_bar=False)
      -> MPS                     open legs remain
      -> (value, exp10)          no legs remain and separate_exp=True   (reference :3084-3086)
      -> value                   no legs remain and separate_exp=False  (reference :3088)

Only the modes the Kagome path enables are implemented (src/algo/contract_tensor_network.py:193-205):
``opt='high'``, ``D_trunc2=None``, ``eps=None``, ``bubbleket=False``, SVD compression.  Anything else
raises ``NotImplementedError`` -- there is no silent fallback.
"""
from __future__ import annotations

import math

import numpy as np

from .dev_bubblecon import trace_bubblecon
from .dev_mps import SLOT_LOGNORM, SLOT_NONFINITE, SLOT_TRUNC
from .engine import E_SVD_NOCONV, BubbleConError
from .mps import MPS
from .program import Program
from .runtime import Compiled, get_engine

_cache: dict = {}
last_stats: dict = {}


def _signature(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors, alias):
    return (tuple(tuple(t.shape) for t in T_list), tuple(tuple(e) for e in edges_list),
            tuple(tuple(round(float(a), 9) for a in ang) for ang in angles_list), round(float(bubble_angle), 9),
            tuple(int(v) for v in swallow_order), D_trunc, tuple(bool(k) for k in ket_tensors), tuple(alias))


def compile_bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors):
    n = len(T_list)
    # tensors that are the same object (the repeated unit cell) share one input buffer
    ids, alias = {}, []
    for t in T_list:
        alias.append(ids.setdefault(id(t), len(ids)))
    key = _signature(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors, alias)
    if key in _cache:
        return _cache[key], alias
    p = Program(8)
    ins, dts = [], {}
    used = set(int(v) for v in swallow_order)
    for i, t in enumerate(T_list):
        a = alias[i]
        if a not in dts and i in used:
            dts[a] = p.input(f"t{a}", t.shape)
            ins.append((f"t{a}", dts[a]))
    TL = [dts.get(alias[i]) for i in range(n)]
    mp, edges = trace_bubblecon(p, TL, edges_list, angles_list, bubble_angle, list(swallow_order), D_trunc, ket_tensors)
    sites = mp.dense_sites()
    for t in sites:
        p.nonfinite(t, SLOT_NONFINITE)
    comp = Compiled(p, ins, [(f"o{k}", t) for k, t in enumerate(sites)], meta=dict(edges=edges, n_out=mp.N))
    _cache[key] = comp
    return comp, alias


def bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None, D_trunc2=None, eps=None,
              opt="high", break_points=(), ket_tensors=None, separate_exp=False, bubbleket=False, compression=None,
              progress_bar=False, engine_key="bubblecon"):
    if opt != "high":
        raise BubbleConError("bubblecon: opt parameter can only be set to 'high'")   # reference :2627-2629
    if D_trunc2 is not None or eps is not None or bubbleket or list(break_points):
        raise NotImplementedError("device bubblecon implements D_trunc2=None, eps=None, bubbleket=False, no break points")
    if compression is not None and compression.get("type", "SVD") != "SVD":
        raise NotImplementedError("device bubblecon implements SVD compression only (iterative compression is selected for D > 10)")
    n = len(T_list)
    if ket_tensors is None:
        ket_tensors = [False] * n
    T_list = [np.asarray(t) for t in T_list] if not all(isinstance(t, np.ndarray) for t in T_list) else T_list
    comp, alias = compile_bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors)
    names = {nm for nm, _, _ in comp.in_layout}
    inputs = {}
    for i, t in enumerate(T_list):
        nm = f"t{alias[i]}"
        if nm in names and nm not in inputs:
            inputs[nm] = np.ascontiguousarray(t, dtype=np.complex128)
    eng = get_engine(engine_key)
    outs, slots, rc = comp.run(eng, [inputs], soft_errors=(E_SVD_NOCONV,))
    if slots[0, SLOT_NONFINITE] > 0:
        raise BubbleConError("bubblecon: the boundary MPS contains nan/inf values")      # reference bmpslib.py:711-717
    last_stats.update(trunc_error=float(slots[0, SLOT_TRUNC]), svd_noconv=(rc == E_SVD_NOCONV), flops=comp.flops)
    n_out = comp.meta["n_out"]
    mp = MPS.from_sites([outs[0][f"o{k}"] for k in range(n_out)], ln_scale=float(slots[0, SLOT_LOGNORM]))
    if not comp.meta["edges"]:
        val = complex(mp.A[0][0, 0, 0])
        if separate_exp:
            return val * mp.nr_mantissa, mp.nr_exp
        return val * mp.overall_factor()
    return mp
