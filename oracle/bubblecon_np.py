"""ORACLE (test infrastructure, not product): numpy restatement of the reference's boundary-MPS
("bubblecon") contractor, double-layer mode only (``bubbleket=False``, ``opt='high'``,
``D_trunc2=None``, SVD compression) -- the only mode the Kagome path ever enables
(src/algo/contract_tensor_network.py:193-205).

Parity status: PINNED against the real reference by ``tools/make_golden.py`` (see mps_np.py).
"""
from __future__ import annotations

import math

import numpy as np

from .mps_np import MPS


def fuse_tensor(T):
    """ket [d, D1, D2, ...] -> double-layer [D1^2, D2^2, ...] (src/libs/bubblecon.py:303-337)."""
    n = T.ndim
    T2 = np.tensordot(T, np.conj(T), axes=([0], [0]))
    perm = [x for i in range(n - 1) for x in (i, i + n - 1)]
    return T2.transpose(perm).reshape([T.shape[i] ** 2 for i in range(1, n)])


def _id_site(DL, Dm, DR):
    """identity reshaped to an MPS site (src/libs/bubblecon.py:345-382)."""
    n = DL if DL == Dm * DR else DR
    return np.eye(n).reshape(DL, Dm, DR)


def tensor_to_MPS_ID(T) -> MPS:
    """SVD-free tensor -> MPS: the middle site carries T, the rest are reshaped identities
    (src/libs/bubblecon.py:390-483)."""
    dims = T.shape
    n = len(dims)
    total = T.size
    k_mid = n // 2
    if n % 2 == 0 and int(np.prod(dims[:k_mid])) ** 2 > total:
        k_mid -= 1
    mp = MPS(n)
    DL = 1
    for i in range(k_mid):
        mp.set_site(_id_site(DL, dims[i], DL * dims[i]), i)
        DL *= dims[i]
    mp.set_site(T.reshape(DL, dims[k_mid], total // (DL * dims[k_mid])), k_mid)
    DR = 1
    for i in range(n - 1, k_mid, -1):
        mp.set_site(_id_site(DR * dims[i], dims[i], DR), i)
        DR *= dims[i]
    return mp


def merge_T(mp: MPS, A, i0, i1) -> MPS:
    """replace sites [i0..i1] of ``mp`` by the MPS form of A[DL, out..., DR]
    (src/libs/bubblecon.py:994-1184)."""
    n_out = A.ndim - 2
    As, Cs = mp.A, mp.Corder
    if n_out == 0:
        if i0 == 0 and i1 == mp.N - 1:
            mp.set_lists([None], [None])
            mp.set_site(A.reshape(1, 1, 1), 0)
            return mp
        if i0 == 0:
            mp.set_lists(As[i1 + 1:], Cs[i1 + 1:])
            mp.set_site(np.tensordot(A, mp.A[0], axes=([1], [0])), 0)
            return mp
        if i1 == mp.N - 1:
            mp.set_lists(As[:i0], Cs[:i0])
            mp.set_site(np.tensordot(mp.A[i0 - 1], A, axes=([2], [0])), i0 - 1)
            return mp
        mp.set_lists(As[:i0] + As[i1 + 1:], Cs[:i0] + Cs[i1 + 1:])
        if A.shape[0] < A.shape[1]:
            mp.set_site(np.tensordot(A, mp.A[i0], axes=([1], [0])), i0)
        else:
            mp.set_site(np.tensordot(mp.A[i0 - 1], A, axes=([2], [0])), i0 - 1)
        return mp
    sub = tensor_to_MPS_ID(A)
    AL = sub.A[0].reshape(sub.A[0].shape[1], sub.A[0].shape[2])
    sub.set_site(np.tensordot(AL, sub.A[1], axes=([1], [0])), 1)
    AR = sub.A[sub.N - 1].reshape(sub.A[sub.N - 1].shape[0], sub.A[sub.N - 1].shape[1])
    sub.set_site(np.tensordot(sub.A[sub.N - 2], AR, axes=([2], [0])), sub.N - 2)
    mid_A, mid_C = sub.A[1:sub.N - 1], sub.Corder[1:sub.N - 1]
    mp.set_lists(As[:i0] + mid_A + As[i1 + 1:], Cs[:i0] + mid_C + Cs[i1 + 1:])
    return mp


def swallow_T(mp: MPS, T, i0, i1, in_legs, out_legs) -> MPS:
    """contract a double-layer tensor with MPS sites [i0..i1] (src/libs/bubblecon.py:2180-2453).
    The reference chains reshaped matmuls; contraction order does not change the exact result."""
    T0 = T.transpose(list(in_legs) + list(out_legs))
    nin = len(in_legs)
    out_shape = list(T0.shape[nin:])
    # fuse the MPS segment:  seg[DL, p_0, ..., p_{nin-1}, DR]
    seg = mp.A[i0]
    for i in range(i0 + 1, i1 + 1):
        seg = np.tensordot(seg, mp.A[i], axes=([seg.ndim - 1], [0]))
    A = np.tensordot(seg, T0, axes=(list(range(1, 1 + nin)), list(range(nin))))  # [DL, DR, out...]
    if out_legs:
        A = A.transpose([0] + list(range(2, 2 + len(out_shape))) + [1])
    return merge_T(mp, A, i0, i1)


def swallow_ket_T(mp: MPS, ket_T, i0, i1, in_legs, out_legs) -> MPS:
    """contract ket and bra layers of a PEPS tensor (physical leg first) with MPS sites [i0..i1]
    whose physical legs are fused (ket,bra) pairs, trace the physical leg, fuse the out pairs
    (src/libs/bubblecon.py:1855-2172)."""
    nin, nout = len(in_legs), len(out_legs)
    n = ket_T.ndim
    T = ket_T.transpose(list(range(1, n)) + [0])              # phys last
    T0 = T.transpose(list(in_legs) + list(out_legs) + [n - 1])  # [in..., out..., p]
    out_shape = list(T0.shape[nin:nin + nout])
    # segment with every physical leg split into (ket, bra)
    seg = None
    for k, i in enumerate(range(i0, i1 + 1)):
        a = mp.A[i]
        dk = T0.shape[k]
        a = a.reshape(a.shape[0], dk, dk, a.shape[2])
        seg = a if seg is None else np.tensordot(seg, a, axes=([seg.ndim - 1], [0]))
    # seg legs: DL, (k0,b0), (k1,b1), ..., DR
    ket_axes = [1 + 2 * k for k in range(nin)]
    bra_axes = [2 + 2 * k for k in range(nin)]
    X = np.tensordot(seg, T0, axes=(ket_axes, list(range(nin))))     # [DL, b..., DR, out..., p]
    Xb = list(range(1, 1 + nin))
    X = np.tensordot(X, np.conj(T0), axes=(Xb + [X.ndim - 1], list(range(nin)) + [nin + nout]))
    # X: [DL, DR, ket-out..., bra-out...]
    perm = [0] + [x for i in range(nout) for x in (2 + i, 2 + nout + i)] + [1]
    A = X.transpose(perm).reshape([X.shape[0]] + [s * s for s in out_shape] + [X.shape[1]])
    return merge_T(mp, A, i0, i1)


def bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None,
              eps=None, break_points=(), ket_tensors=None, separate_exp=False, compression=None):
    """main loop (src/libs/bubblecon.py:2465-3107): root tensor -> MPS, then for each vertex in
    ``swallow_order`` find the contiguous MPS legs that point at it, swallow it, and truncate the
    whole boundary MPS with ``reduceD(D_trunc, eps, nr_bulk=True)``."""
    n = len(T_list)
    if ket_tensors is None:
        ket_tensors = [False] * n
    # compression = {'type': 'SVD'} | {'type': 'iter', 'max-iter': ..., 'err': ...}      (:2612-2624, 2793-2798, 3035-3038)
    comp_type = "SVD" if compression is None else compression.get("type", "SVD")

    def compress(m):
        if comp_type == "SVD":
            m.reduceD(D_trunc, eps, nr_bulk=True)
        else:
            from .reduce_iter_np import reduceDiter
            reduceDiter(m, D_trunc, nr_bulk=True, max_iter=compression["max-iter"], err=compression["err"])
    # edge -> (i, j) vertex pair; open edges map to (i, i)          (:2654-2664)
    vertices = {}
    for i in range(n):
        for e in edges_list[i]:
            vertices[e] = (i, vertices[e][0]) if e in vertices else (i, i)

    root = swallow_order[0]
    r_ang = np.array(angles_list[root])
    r_edges = edges_list[root]
    rot = (bubble_angle + math.pi - r_ang) % (2 * math.pi)            # (:2700)
    L = sorted((rot[i], i, r_edges[i]) for i in range(len(r_edges)))
    perm = [x[1] for x in L]
    mp_edges = [x[2] for x in L]
    T_root = fuse_tensor(T_list[root]) if ket_tensors[root] else T_list[root]
    mp = tensor_to_MPS_ID(T_root.transpose(perm))
    if D_trunc is not None:
        compress(mp)

    snapshots = []
    for l in range(1, len(swallow_order)):
        if (l - 1) in break_points:
            snapshots.append(mp.copy())
        v = swallow_order[l]
        v_edges = edges_list[v]
        v_ang = np.array(angles_list[v])
        k = len(v_edges)
        legs = [(i, e) for i, e in enumerate(mp_edges) if v in vertices[e]]   # (:2918-2919)
        if not legs:
            raise RuntimeError(f"bubblecon: no MPS leg connects to vertex {v}")
        i0, i1 = legs[0][0], legs[-1][0]
        in_legs = [v_edges.index(e) for _, e in legs]
        if len(in_legs) != i1 - i0 + 1:
            raise RuntimeError(f"bubblecon: legs of vertex {v} are not contiguous in the MPS")
        out1 = list(set(range(k)) - set(in_legs))                              # (:2958)
        if len(out1) > 1:
            rv = (v_ang[in_legs[0]] * np.ones(k) - v_ang + 2 * math.pi) % (2 * math.pi)
            out_legs = [i for _, i in sorted((rv[i], i) for i in out1)]
        else:
            out_legs = out1
        if ket_tensors[v]:
            mp = swallow_ket_T(mp, T_list[v], i0, i1, in_legs, out_legs)
        else:
            mp = swallow_T(mp, T_list[v], i0, i1, in_legs, out_legs)
        if D_trunc is not None:
            compress(mp)
        mp_edges = mp_edges[:i0] + [v_edges[i] for i in out_legs] + mp_edges[i1 + 1:]

    if not mp_edges and not snapshots:                                          # (:3077-3088)
        val = mp.A[0][0, 0, 0]
        if separate_exp:
            return val * mp.nr_mantissa, mp.nr_exp
        return val * mp.overall_factor()
    if snapshots:
        snapshots.append(mp)
        return snapshots
    return mp
