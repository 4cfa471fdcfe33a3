"""ORACLE (test infrastructure, not product): numpy restatement of the block-BP loop
(src/algo/belief_propagation.py) on top of ``bubblecon_np``.

Geometry (sites, edges, angles, swallow order) is integer host logic shared with the product
(``kagomeperiodicbp_b200.lattice / block_tn / contraction_order``), itself pinned against the
reference by ``tests/test_geometry_golden.py``.  All floating-point work here is numpy.

Parity status: PINNED -- ``tools/make_golden.py`` runs the real reference and this module on the
same seeded unit cells and stores the reference's results in ``tests/golden/``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass

import numpy as np

from kagomeperiodicbp_b200 import block_tn, contraction_order
from kagomeperiodicbp_b200.lattice import BLOCK_SIDES_CCW, SIDE_ANGLE, SIDE_OPPOSITE

from .bubblecon_np import bubblecon
from .mps_np import (MPS, add_two_MPSs, hermitize_a_message, init_mps_quantum, mps_distance,
                     mps_inner_product)


@dataclass
class BPConfigNP:
    """the fields of the reference's BPConfig that affect numerics
    (src/containers/belief_propagation.py:30-45)."""
    max_iterations: int = 50
    trunc_dim: int = 9
    msg_diff_terminate: float = 1e-10
    msg_diff_good_enough: float = 1e-5
    msg_diff_squared: bool = True
    allowed_retries: int = 2
    times_to_deem_failure_when_diff_increases: int = 3
    damping: float | None = None
    hermitize_msgs_when_finished: bool = True
    fix_msg_each_step: bool = True


def uniform_messages(N: int, D: int) -> dict:
    """(src/tensor_networks/tensor_network.py:272-285)"""
    return {s: init_mps_quantum([D] * (2 * N - 1), random=False) for s in BLOCK_SIDES_CCW}


def random_messages(N: int, D: int, rng) -> dict:
    return {s: init_mps_quantum([D] * (2 * N - 1), random=True, rng=rng) for s in BLOCK_SIDES_CCW}


def outgoing_message(N, cell, messages: dict, side: str, chi: int, depth="ToMessage", break_points=()):
    """one ToMessage chain (src/algo/belief_propagation.py:90-106 ->
    src/algo/contract_tensor_network.py:146-213)."""
    T, E, A, K, P = block_tn.assemble(N, cell, {s: m.A for s, m in messages.items()})
    T, E, A = block_tn.connect_corner(N, T, E, A, P, side)
    order = list(contraction_order.kagome_order(N, side, depth))
    return bubblecon(T, E, A, SIDE_ANGLE[side], order, D_trunc=chi, ket_tensors=K,
                     separate_exp=True, break_points=break_points)


def fix_messages(messages: dict):
    """(src/algo/belief_propagation.py:113-117)"""
    for m in messages.values():
        m.right_canonical(nr_bulk=True)
        m.reset_nr()


def compute_error(prev: dict, out: dict, squared: bool) -> float:
    """(src/algo/belief_propagation.py:44-56)"""
    d = [mps_distance(prev[s], out[s]) for s in BLOCK_SIDES_CCW]
    return sum(d) / len(d) if squared else float(np.sqrt(sum(d)) / len(d))


def single_mps_damping(old: MPS, new: MPS, damping: float, trunc_dim: int) -> MPS:
    """(src/algo/belief_propagation.py:59-74)"""
    ip = mps_inner_product(new, old, conjB=True)
    sign = 1 if ip.real > 0 else -1
    c = add_two_MPSs(new, 1 - damping, old, sign * damping)
    c.left_canonical_QR()
    c.right_canonical(maxD=trunc_dim, nr_bulk=True)
    c.reset_nr()
    return c


def bp_step(N, cell, prev: dict, cfg: BPConfigNP):
    """(src/algo/belief_propagation.py:120-188)"""
    out = {}
    for side in BLOCK_SIDES_CCW:
        out[SIDE_OPPOSITE[side]] = outgoing_message(N, cell, prev, side, cfg.trunc_dim)
    if cfg.fix_msg_each_step:
        fix_messages(out)
    err = compute_error(prev, out, cfg.msg_diff_squared)
    if cfg.damping is None or cfg.damping == 0:
        nxt = out
    else:
        nxt = {s: single_mps_damping(prev[s], out[s], cfg.damping, cfg.trunc_dim) for s in out}
    return out, nxt, err


def belief_propagation(N, cell, messages: dict, cfg: BPConfigNP):
    """(src/algo/belief_propagation.py:192-281); ``messages`` must be given (the reference draws
    random ones from the global np.random when None)."""
    errors = []
    min_err, min_msgs = np.inf, messages
    nxt = messages
    err, success, it = None, False, 0
    out = messages
    for it in range(cfg.max_iterations):
        out, nxt, err = bp_step(N, cell, nxt, cfg)
        if err < cfg.msg_diff_terminate:
            success = True
            break
        if err < min_err:
            min_err, min_msgs = err, copy.deepcopy(out)
        errors.append(err)
        k = cfg.times_to_deem_failure_when_diff_increases
        if len(errors) > k and all(errors[-k:][i] <= errors[-k:][i + 1] for i in range(k - 1)):
            break
    if not success:
        out, err = min_msgs, min_err
    if cfg.hermitize_msgs_when_finished:
        out = {s: hermitize_a_message(m) for s, m in out.items()}
    return out, dict(iterations=it + 1, final_error=float(err), success=success)


def robust_belief_propagation(N, cell, messages: dict, cfg: BPConfigNP, fresh_messages=None):
    """retry ladder (src/algo/belief_propagation.py:285-350).  On a retry the reference restarts
    from random messages (``messages_in=None``); ``fresh_messages()`` supplies them here."""
    cfg = copy.deepcopy(cfg)
    msgs_in = copy.deepcopy(messages)
    min_msgs, min_err, total = msgs_in, np.inf, 0
    msgs_out, err_out = None, None
    attempt = 0
    for attempt in range(cfg.allowed_retries):
        msgs, st = belief_propagation(N, cell, msgs_in, cfg)
        total += st["iterations"]
        if st["final_error"] < cfg.msg_diff_terminate:
            msgs_out, err_out = msgs, st["final_error"]
            break
        if st["final_error"] < min_err:
            min_err, min_msgs = st["final_error"], copy.deepcopy(msgs)
        cfg.trunc_dim = int(1.5 * cfg.trunc_dim)
        cfg.max_iterations += 11
        msgs_in = fresh_messages() if fresh_messages is not None else copy.deepcopy(messages)
    else:
        msgs_out, err_out = min_msgs, min_err
    return msgs_out, dict(attempts=attempt + 1, iterations=total, final_error=float(err_out),
                          success=err_out < cfg.msg_diff_good_enough)
