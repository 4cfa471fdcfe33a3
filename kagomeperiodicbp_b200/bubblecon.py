"""Drop-in for the reference's ``bubblecon`` entry point (src/libs/bubblecon.py:2465-2468): same
positional/keyword arguments, same return conventions, executed as one device program.

    bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None, D_trunc2=None,
              eps=None, opt='high', break_points=[], ket_tensors=None, separate_exp=False, bubbleket=False,
              compression=None, progress_bar=False)
      -> MPS                     open legs remain
      -> (value, exp10)          no legs remain and separate_exp=True   (reference :3084-3086)
      -> value                   no legs remain and separate_exp=False  (reference :3088)

Only the modes the Kagome path enables are implemented (src/algo/contract_tensor_network.py:193-205):
``opt='high'``, ``D_trunc2=None``, ``eps=None``, ``bubbleket=False``.  Anything else raises ``NotImplementedError`` -- there
is no silent fallback.  ``compression={'type': 'SVD'}`` (default) compiles the whole contraction into ONE device program;
``compression={'type': 'iter', 'max-iter': m, 'err': e}`` (what the reference selects for D > 10,
src/containers/contractions.py:18-35) runs one device program per swallow and the QR-only compressor
(`reduce_iter.reduceDiter`, device resident backend) between them, because its number of rounds is decided on the host from
data, exactly as in the reference (src/libs/bubblecon.py:2793-2798, 3035-3038).
"""
from __future__ import annotations

import math

import numpy as np

from .dev_bubblecon import fuse_tensor, swallow_ket_T, swallow_T, tensor_to_mps_id, trace_bubblecon
from .dev_mps import SLOT_LOGNORM, SLOT_NONFINITE, SLOT_TRUNC, DevMPS
from .engine import E_SVD_NOCONV, BubbleConError, raise_if_not_converged
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


# ---------------------------------------------------------------------------------------------------------------------
# iterative compression: one program per swallow, reduceDiter between them
_step_cache: dict = {}


def _plan(edges_list, angles_list, bubble_angle, swallow_order):
    """topology of the contraction (no tensor data): root leg order and, per swallow, which MPS legs the vertex takes and in
    which order its remaining legs enter the MPS (src/libs/bubblecon.py:2654-2720, 2900-2990; same rules as trace_bubblecon)."""
    n = len(edges_list)
    vertices = {}
    for i in range(n):
        for e in edges_list[i]:
            vertices[e] = (i, vertices[e][0]) if e in vertices else (i, i)
    root = swallow_order[0]
    r_ang = np.array(angles_list[root])
    r_edges = edges_list[root]
    rot = (bubble_angle + math.pi - r_ang) % (2 * math.pi)
    Ls = sorted((rot[i], i, r_edges[i]) for i in range(len(r_edges)))
    perm = [x[1] for x in Ls]
    mp_edges = [x[2] for x in Ls]
    steps = []
    for l in range(1, len(swallow_order)):
        v = swallow_order[l]
        v_edges = edges_list[v]
        v_ang = np.array(angles_list[v])
        k = len(v_edges)
        legs = [(i, e) for i, e in enumerate(mp_edges) if v in vertices[e]]
        if not legs:
            raise ValueError(f"bubblecon: no MPS leg connects to vertex {v}")
        i0, i1 = legs[0][0], legs[-1][0]
        in_legs = [v_edges.index(e) for _, e in legs]
        if len(in_legs) != i1 - i0 + 1:
            raise ValueError(f"bubblecon: legs of vertex {v} are not contiguous in the MPS")
        out1 = list(set(range(k)) - set(in_legs))
        if len(out1) > 1:
            rv = (v_ang[in_legs[0]] * np.ones(k) - v_ang + 2 * math.pi) % (2 * math.pi)
            out_legs = [i for _, i in sorted((rv[i], i) for i in out1)]
        else:
            out_legs = out1
        steps.append((v, i0, i1, in_legs, out_legs))
        mp_edges = mp_edges[:i0] + [v_edges[i] for i in out_legs] + mp_edges[i1 + 1:]
    return root, perm, steps, mp_edges


def _root_program(t_shape, ket, perm):
    key = ("root", tuple(t_shape), bool(ket), tuple(perm))
    if key not in _step_cache:
        p = Program(8)
        t = p.input("t", t_shape)
        T_root = fuse_tensor(p, t) if ket else t
        mp = tensor_to_mps_id(p, p.transpose(T_root, perm), (SLOT_LOGNORM, SLOT_TRUNC))
        sites = mp.dense_sites()
        for x in sites:
            p.nonfinite(x, SLOT_NONFINITE)
        _step_cache[key] = Compiled(p, [("t", t)], [(f"o{k}", x) for k, x in enumerate(sites)], meta=dict(corder=list(mp.Corder), n_out=mp.N))
    return _step_cache[key]


def _swallow_program(site_shapes, corder, t_shape, ket, i0, i1, in_legs, out_legs):
    key = ("step", tuple(tuple(s) for s in site_shapes), tuple(corder), tuple(t_shape), bool(ket), i0, i1, tuple(in_legs), tuple(out_legs))
    if key not in _step_cache:
        p = Program(8)
        ins = []
        mp = DevMPS(p, len(site_shapes))
        for k, sh in enumerate(site_shapes):
            dt = p.input(f"s{k}", sh)
            ins.append((f"s{k}", dt))
            mp.set_site(dt, k, corder[k])
        t = p.input("t", t_shape)
        ins.append(("t", t))
        mp = (swallow_ket_T if ket else swallow_T)(p, mp, t, i0, i1, in_legs, out_legs)
        sites = mp.dense_sites()
        for x in sites:
            p.nonfinite(x, SLOT_NONFINITE)
        _step_cache[key] = Compiled(p, ins, [(f"o{k}", x) for k, x in enumerate(sites)], meta=dict(corder=list(mp.Corder), n_out=mp.N))
    return _step_cache[key]


def _bubblecon_iterative(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors, compression,
                         separate_exp, engine_key):
    from . import reduce_iter
    max_iter, err = int(compression["max-iter"]), float(compression["err"])
    root, perm, steps, final_edges = _plan(edges_list, angles_list, bubble_angle, list(swallow_order))
    eng = get_engine(engine_key)
    B = reduce_iter.backend()
    rounds = []

    def run(comp, inputs):
        outs, slots, rc = comp.run(eng, [inputs], soft_errors=(E_SVD_NOCONV,))
        if slots[0, SLOT_NONFINITE] > 0:
            raise BubbleConError("bubblecon: the boundary MPS contains nan/inf values")
        return [outs[0][f"o{k}"] for k in range(comp.meta["n_out"])], list(comp.meta["corder"])

    def compress(mp):
        st = {}
        reduce_iter.reduceDiter(B, mp, D_trunc, nr_bulk=True, max_iter=max_iter, err=err, stats=st)
        rounds.append(st.get("rounds", 0))

    c128 = lambda t: np.ascontiguousarray(t, dtype=np.complex128)
    sites, cor = run(_root_program(T_list[root].shape, ket_tensors[root], perm), {"t": c128(T_list[root])})
    mp = MPS.from_sites(sites, Corder=cor)
    compress(mp)
    for v, i0, i1, in_legs, out_legs in steps:
        comp = _swallow_program([a.shape for a in mp.A], mp.Corder, T_list[v].shape, ket_tensors[v], i0, i1, in_legs, out_legs)
        inputs = {f"s{k}": c128(a) for k, a in enumerate(mp.A)}
        inputs["t"] = c128(T_list[v])
        sites, cor = run(comp, inputs)
        nxt = MPS.from_sites(sites, Corder=cor)
        nxt.nr_mantissa, nxt.nr_exp = mp.nr_mantissa, mp.nr_exp
        mp = nxt
        compress(mp)
    last_stats.update(trunc_error=None, svd_noconv=False, flops=None, reduce_iter_rounds=rounds)
    if not final_edges:
        val = complex(mp.A[0][0, 0, 0])
        if separate_exp:
            return val * mp.nr_mantissa, mp.nr_exp
        return val * mp.overall_factor()
    return mp


def bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None, D_trunc2=None, eps=None,
              opt="high", break_points=(), ket_tensors=None, separate_exp=False, bubbleket=False, compression=None,
              progress_bar=False, engine_key="bubblecon"):
    if opt != "high":
        raise BubbleConError("bubblecon: opt parameter can only be set to 'high'")   # reference :2627-2629
    if D_trunc2 is not None or eps is not None or bubbleket or list(break_points):
        raise NotImplementedError("device bubblecon implements D_trunc2=None, eps=None, bubbleket=False, no break points")
    ctype = "SVD" if compression is None else compression.get("type", "SVD")
    if ctype not in ("SVD", "iter"):
        raise NotImplementedError(f"bubblecon: unknown compression type {ctype!r}")
    n = len(T_list)
    if ket_tensors is None:
        ket_tensors = [False] * n
    T_list = [np.asarray(t) for t in T_list] if not all(isinstance(t, np.ndarray) for t in T_list) else T_list
    if ctype == "iter" and D_trunc is not None:
        return _bubblecon_iterative(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc, ket_tensors, compression,
                                    separate_exp, engine_key)
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
    raise_if_not_converged(rc, "bubblecon")
    n_out = comp.meta["n_out"]
    mp = MPS.from_sites([outs[0][f"o{k}"] for k in range(n_out)], ln_scale=float(slots[0, SLOT_LOGNORM]))
    if not comp.meta["edges"]:
        val = complex(mp.A[0][0, 0, 0])
        if separate_exp:
            return val * mp.nr_mantissa, mp.nr_exp
        return val * mp.overall_factor()
    return mp
