"""Drop-in binding of the accelerated path into a RUNNING copy of the reference (NGBigField/KagomePeriodicBP).

The reference has no plugin interface; its seams are plain Python names.  ``install(...)`` rebinds them in the
already-imported reference modules, so that ``scripts/run_ite.py``, ``algo.belief_propagation.belief_propagation``,
``reduce_tn`` and the measurement functions keep their code and call the device through ``libkbp.so``:

  S1  ``algo.contract_tensor_network.bubblecon``            (src/algo/contract_tensor_network.py:27, called at :193-205)
        -> ``bubblecon.bubblecon`` of this package; the result is converted into a genuine ``libs.bmpslib.mps`` object
           (src/libs/bmpslib.py:214) -- the reference indexes ``.A[i]``, calls ``.copy(full=True)``, pickles it (dill) and
           reads ``nr_mantissa / nr_exp``.
  S2  ``algo.belief_propagation._belief_propagation_step``   (src/algo/belief_propagation.py:164-188)
        -> one batched device iteration (six side programs, epilogue on the device), messages converted both ways.
  S4  ``rho_ij`` / ``apply_2local_gate`` as imported by ``algo.imaginary_time_evolution._tn_update`` and ``tensor_networks``
        (src/libs/ITE.py:555, 1761)  -> ``ite.rho_ij`` / ``ite.apply_2local_gate`` on the device backend.

Nothing here imports the reference: the caller passes the modules it has already imported (the tests do so through
``tools/ref_env.py`` in the build container; a user does it in their own checkout).  ``uninstall()`` restores the originals.
"""
from __future__ import annotations

import numpy as np

from . import belief_propagation as kbp
from . import bubblecon as kbubblecon
from .containers import BPConfig, Message, MPSOrientation, UnitCell
from .lattice import BLOCK_SIDES_CCW
from .mps import MPS

_saved: list = []


# ---------------------------------------------------------------------------------------------------------------------
# conversions (a22): package objects <-> the reference's own classes
# ---------------------------------------------------------------------------------------------------------------------
def to_reference_mps(m: MPS, bmpslib):
    """a genuine ``bmpslib.mps`` with fresh host arrays (``set_site`` copies, src/libs/bmpslib.py:501-505)."""
    r = bmpslib.mps(m.N)
    for i, a in enumerate(m.A):
        r.set_site(np.array(a, dtype=np.complex128), i)
    r.Corder = list(m.Corder)
    r.nr_mantissa, r.nr_exp = m.nr_mantissa, m.nr_exp
    return r


def from_reference_mps(r) -> MPS:
    m = MPS.from_sites([np.asarray(a) for a in r.A], Corder=list(r.Corder))
    m.nr_mantissa, m.nr_exp = r.nr_mantissa, r.nr_exp
    return m


def _side_name(side) -> str:
    return side if isinstance(side, str) else str(side.name if hasattr(side, "name") else side)


def from_reference_messages(messages: dict) -> dict:
    """{BlockSide: Message(mps, orientation)} of the reference -> {side name: Message} of this package."""
    out = {}
    for side, msg in messages.items():
        s = _side_name(side)
        out[s] = Message(from_reference_mps(msg.mps), MPSOrientation.standard(kbp.SIDE_OPPOSITE[s]))
    return out


def to_reference_messages(messages: dict, ref_like: dict, bmpslib, ref_message_cls) -> dict:
    """back into the reference's container: same keys (BlockSide members) and orientations as ``ref_like``."""
    by_name = {_side_name(k): k for k in ref_like}
    return {by_name[s]: ref_message_cls(to_reference_mps(m.mps, bmpslib), ref_like[by_name[s]].orientation) for s, m in messages.items()}


def bp_config_from_reference(config) -> BPConfig:
    """the numerics-relevant fields of the reference's BPConfig (src/containers/belief_propagation.py:30-45)."""
    init = getattr(config, "init_msg", "UQ")
    init = getattr(init, "name", init)
    init = {"UNIFORM_QUANTUM": "UQ", "RANDOM_QUANTUM": "RQ"}.get(str(init), str(init))
    return BPConfig(max_iterations=config.max_iterations, trunc_dim=int(config.trunc_dim), msg_diff_terminate=config.msg_diff_terminate,
                    msg_diff_good_enough=config.msg_diff_good_enough, msg_diff_squared=config.msg_diff_squared,
                    allowed_retries=config.allowed_retries,
                    times_to_deem_failure_when_diff_increases=config.times_to_deem_failure_when_diff_increases,
                    damping=config.damping, hermitize_msgs_when_finished=config.hermitize_msgs_when_finished,
                    fix_msg_each_step=config.fix_msg_each_step, init_msg=init)


# ---------------------------------------------------------------------------------------------------------------------
# the replacement bodies
# ---------------------------------------------------------------------------------------------------------------------
def make_bubblecon(bmpslib):
    """S1: same signature and return conventions as libs.bubblecon.bubblecon (src/libs/bubblecon.py:2465-2468)."""

    def bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=None, D_trunc2=None, eps=None, opt="high",
                  break_points=[], ket_tensors=None, separate_exp=False, bubbleket=False, compression=None, progress_bar=True):
        r = kbubblecon.bubblecon(T_list, edges_list, angles_list, bubble_angle, swallow_order, D_trunc=D_trunc, D_trunc2=D_trunc2,
                                 eps=eps, opt=opt, break_points=break_points, ket_tensors=ket_tensors, separate_exp=separate_exp,
                                 bubbleket=bubbleket, compression=compression, progress_bar=False)
        return to_reference_mps(r, bmpslib) if isinstance(r, MPS) else r
    return bubblecon


def make_bp_step(bmpslib, ref_message_cls):
    """S2: body of _belief_propagation_step(tn, prev_messages, prev_error, config, prog_bar_obj)."""

    def _belief_propagation_step(tn, prev_messages, prev_error, config, prog_bar_obj=None):
        uc = tn.unit_cell
        cell = UnitCell(np.asarray(uc.A), np.asarray(uc.B), np.asarray(uc.C))
        msgs = from_reference_messages(prev_messages)
        N = tn.lattice.N
        out, nxt, err, _ = kbp.bp_step_batch(N, [cell], [msgs], bp_config_from_reference(config))[0]
        return (to_reference_messages(out, prev_messages, bmpslib, ref_message_cls),
                to_reference_messages(nxt, prev_messages, bmpslib, ref_message_cls), float(err))
    return _belief_propagation_step


def make_ite_functions():
    """S4: rho_ij(Ti, Tj, env_i=None, env_j=None, mps_env=None) and apply_2local_gate(g, Dmax, Ti, Tj, env_i=None, env_j=None,
    mps_env=None) with the reference's positional conventions (src/libs/ITE.py:555, 1761); only the mps_env branch exists on
    the Kagome path."""
    from . import ite, ite_flow

    def rho_ij(Ti, Tj, env_i=None, env_j=None, mps_env=None):
        if mps_env is None:
            raise NotImplementedError("rho_ij: only the mps_env branch is on the Kagome path")
        return np.asarray(ite.rho_ij(ite_flow.backend(), Ti, Tj, mps_env))

    def apply_2local_gate(g, Dmax, Ti, Tj, env_i=None, env_j=None, mps_env=None):
        if mps_env is None:
            raise NotImplementedError("apply_2local_gate: only the mps_env branch is on the Kagome path")
        ti, tj, ev = ite.apply_2local_gate(ite_flow.backend(), g, Dmax, Ti, Tj, mps_env)
        return np.asarray(ti), np.asarray(tj), ev
    return rho_ij, apply_2local_gate


def _rebind(module, name, new):
    _saved.append((module, name, getattr(module, name)))
    setattr(module, name, new)


def install(seams=("S1",)):
    """rebind the chosen seams in the reference modules that are ALREADY importable (``libs``, ``algo``, ``containers`` on
    sys.path, as in the reference's own scripts).  S1 and S2 are alternatives for BP (S2 short-circuits S1 inside BP; S1
    still serves reduce_tn and the measurements)."""
    import importlib
    bmpslib = importlib.import_module("libs.bmpslib")
    if "S1" in seams:
        ctn = importlib.import_module("algo.contract_tensor_network")
        _rebind(ctn, "bubblecon", make_bubblecon(bmpslib))
    if "S2" in seams:
        rbp = importlib.import_module("algo.belief_propagation")
        cont = importlib.import_module("containers")
        _rebind(rbp, "_belief_propagation_step", make_bp_step(bmpslib, cont.Message))
    if "S4" in seams:
        rho_ij, gate = make_ite_functions()
        upd = importlib.import_module("algo.imaginary_time_evolution._tn_update")
        _rebind(upd, "rho_ij", rho_ij)
        _rebind(upd, "apply_2local_gate", gate)
        for nm in ("tensor_networks.tensor_network", "algo.measurements"):        # the other importers of libs.ITE.rho_ij
            _rebind(importlib.import_module(nm), "rho_ij", rho_ij)
    return list(seams)


def uninstall():
    while _saved:
        module, name, old = _saved.pop()
        setattr(module, name, old)
