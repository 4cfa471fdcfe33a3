"""Device boundary-MPS contractor: records one whole ``bubblecon`` call (the reference's S1 seam,
src/libs/bubblecon.py:2465-3107) as a tensor program.

The leg bookkeeping (which MPS legs point at the next vertex, in/out leg order from the angles) is the
reference's and runs on the host at *compile* time; only shapes are needed.  The numerics recorded per
swallow are
    swallow_ket_T  (src/libs/bubblecon.py:1855-2172)  ket and bra layer contracted separately, D^2 legs never built
    swallow_T      (src/libs/bubblecon.py:2180-2453)  already-fused tensors (message sites)
    merge_T / tensor_to_MPS_ID (:994-1184, :390-483)  SVD-free re-expansion into sites
    mps.reduceD(D_trunc, nr_bulk=True)  (:3029-3038)   after every swallow
"""
from __future__ import annotations

import math

import numpy as np

from .dev_mps import DevMPS, Id, absorb_left, absorb_right
from .program import DT, Program


def _id_site(p: Program, DL, Dm, DR):
    return Id(DL, Dm, DR)          # stays symbolic; see dev_mps.py


def fuse_tensor(p: Program, T: DT) -> DT:
    n = T.ndim
    T2 = p.tensordot(T, T, ([0], [0]), conj_b=True)
    perm = [x for i in range(n - 1) for x in (i, i + n - 1)]
    return p.transpose(T2, perm).reshape([T.shape[i] ** 2 for i in range(1, n)])


def tensor_to_mps_id(p: Program, T: DT, slots) -> DevMPS:
    dims = T.shape
    n = len(dims)
    total = T.size
    k_mid = n // 2
    if n % 2 == 0 and int(np.prod(dims[:k_mid])) ** 2 > total:
        k_mid -= 1
    mp = DevMPS(p, n, *slots)
    DL = 1
    for i in range(k_mid):
        mp.set_site(_id_site(p, DL, dims[i], DL * dims[i]), i)
        DL *= dims[i]
    mp.set_site(T.reshape(DL, dims[k_mid], total // (DL * dims[k_mid])), k_mid)
    DR = 1
    for i in range(n - 1, k_mid, -1):
        mp.set_site(_id_site(p, DR * dims[i], dims[i], DR), i)
        DR *= dims[i]
    return mp


def merge_T(p: Program, mp: DevMPS, A: DT, i0: int, i1: int) -> DevMPS:
    n_out = A.ndim - 2
    As, Cs = mp.A, mp.Corder
    if n_out == 0:
        if i0 == 0 and i1 == mp.N - 1:
            mp.set_lists([None], [None])
            mp.set_site(A.reshape(1, 1, 1), 0)
            return mp
        if i0 == 0:
            mp.set_lists(As[i1 + 1:], Cs[i1 + 1:])
            mp.set_site(absorb_left(p, A, mp.A[0]), 0)
            return mp
        if i1 == mp.N - 1:
            mp.set_lists(As[:i0], Cs[:i0])
            mp.set_site(absorb_right(p, mp.A[i0 - 1], A), i0 - 1)
            return mp
        mp.set_lists(As[:i0] + As[i1 + 1:], Cs[:i0] + Cs[i1 + 1:])
        if A.shape[0] < A.shape[1]:
            mp.set_site(absorb_left(p, A, mp.A[i0]), i0)
        else:
            mp.set_site(absorb_right(p, mp.A[i0 - 1], A), i0 - 1)
        return mp
    sub = tensor_to_mps_id(p, A, (mp.slot_lognorm, mp.slot_trunc))
    # the end sites of `sub` are [1, DL, DL] and [DR, DR, 1] identities: absorbing them is a reshape
    mid_A = list(sub.A[1:sub.N - 1])
    mid_C = [None] * len(mid_A)
    mp.set_lists(As[:i0] + mid_A + As[i1 + 1:], Cs[:i0] + mid_C + Cs[i1 + 1:])
    return mp


def swallow_T(p: Program, mp: DevMPS, T: DT, i0, i1, in_legs, out_legs) -> DevMPS:
    T0 = p.transpose(T, list(in_legs) + list(out_legs))
    nin = len(in_legs)
    out_shape = list(T0.shape[nin:])
    seg = mp.site(i0)
    for i in range(i0 + 1, i1 + 1):
        seg = p.tensordot(seg, mp.site(i), ([seg.ndim - 1], [0]))
    A = p.tensordot(seg, T0, (list(range(1, 1 + nin)), list(range(nin))))      # [DL, DR, out...]
    if out_legs:
        A = p.transpose(A, [0] + list(range(2, 2 + len(out_shape))) + [1])
    return merge_T(p, mp, A, i0, i1)


def swallow_ket_T(p: Program, mp: DevMPS, ket_T: DT, i0, i1, in_legs, out_legs) -> DevMPS:
    nin, nout = len(in_legs), len(out_legs)
    n = ket_T.ndim
    # [in..., out..., p]  (physical leg is axis 0 of the ket tensor)
    T0 = p.transpose(ket_T, [1 + x for x in in_legs] + [1 + x for x in out_legs] + [0])
    out_shape = list(T0.shape[nin:nin + nout])
    seg = None
    for k, i in enumerate(range(i0, i1 + 1)):
        a = mp.site(i)
        dk = T0.shape[k]
        assert a.shape[1] == dk * dk, "MPS physical leg must be the fused (ket, bra) pair of the tensor leg"
        a = a.reshape(a.shape[0], dk, dk, a.shape[2])
        seg = a if seg is None else p.tensordot(seg, a, ([seg.ndim - 1], [0]))
    ket_axes = [1 + 2 * k for k in range(nin)]
    X = p.tensordot(seg, T0, (ket_axes, list(range(nin))))                      # [DL, b..., DR, out..., p]
    Xb = list(range(1, 1 + nin))
    X = p.tensordot(X, T0, (Xb + [X.ndim - 1], list(range(nin)) + [nin + nout]), conj_b=True)
    perm = [0] + [x for i in range(nout) for x in (2 + i, 2 + nout + i)] + [1]
    A = p.transpose(X, perm).reshape([X.shape[0]] + [s * s for s in out_shape] + [X.shape[1]])
    return merge_T(p, mp, A, i0, i1)


def trace_bubblecon(p: Program, T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc,
                    ket_tensors=None, slots=(0, 1)) -> tuple[DevMPS, list]:
    """records the whole contraction; returns the final DevMPS and the list of edges its legs carry."""
    n = len(T_list)
    if ket_tensors is None:
        ket_tensors = [False] * n
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
    T_root = fuse_tensor(p, T_list[root]) if ket_tensors[root] else T_list[root]
    mp = tensor_to_mps_id(p, p.transpose(T_root, perm), slots)
    if D_trunc is not None:
        mp.reduceD(D_trunc, nr_bulk=True)
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
        if ket_tensors[v]:
            mp = swallow_ket_T(p, mp, T_list[v], i0, i1, in_legs, out_legs)
        else:
            mp = swallow_T(p, mp, T_list[v], i0, i1, in_legs, out_legs)
        if D_trunc is not None:
            mp.reduceD(D_trunc, nr_bulk=True)
        mp_edges = mp_edges[:i0] + [v_edges[i] for i in out_legs] + mp_edges[i1 + 1:]
    return mp, mp_edges
