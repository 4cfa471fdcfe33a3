"""The ITE step on the device: one loop body of the reference's ``ite_per_mode``
(src/algo/imaginary_time_evolution/main.py:566-590) --

    robust_belief_propagation (warm messages)              -> belief_propagation.py (six device chains per iteration)
    reduce_tn(full_tn, ModeTN) / reduce_tn(mode_tn, EdgeTN) -> two device ToCore chains + edge_env.edge_environment
    ite_update_unit_cell (_tn_update.py:165-205)            -> ite.rho_ij / ite.apply_2local_gate on the device backend

and the energy measurement ``measure_energies_and_observables_together`` (src/algo/measurements.py:163-243: six edge RDMs of
one mode, energy per site = sum / 3).  Host code orchestrates and moves data; all tensor algebra runs through libkbp.so.
The core is reduced with bottom-up direction U unless told otherwise (the reference draws it at random from {U, DL, DR},
kagome_to_core.py:178-179; SURVEY 8d fixes U for reproducible runs; `reduce_to_core(..., direction=)` takes all three).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import belief_propagation as bp
from . import edge_env, ite
from .bubblecon import bubblecon as device_bubblecon
from .containers import BPConfig, UnitCell
from .lattice import BLOCK_SIDES_CCW
from .linalg import ResidentBackend
from .runtime import get_engine

_backend = None


def backend() -> ResidentBackend:
    global _backend
    if _backend is None:
        _backend = ResidentBackend("ite")
    return _backend


def _device_bubblecon_fn(T_list, edges, angles, bubble_angle, order, chi, kets):
    mp = device_bubblecon([np.ascontiguousarray(np.asarray(t), dtype=np.complex128) for t in T_list], edges, angles, bubble_angle, order,
                          D_trunc=chi, ket_tensors=kets, engine_key="ite-bubblecon")
    return mp.A


def reduce_to_core(unit_cell: UnitCell, messages: dict, N: int, chi: int, device: int = 0, direction: str = "U"):
    """the 12 ring tensors of the CoreTN (kagome_to_core.py:322-364): two truncated ToCore chains (bottom-up ``direction`` in
    {U, DL, DR} and its opposite) run concurrently on two streams, then the overlap zip on the device backend."""
    bu_side, td_side, _ = edge_env.CORE_DIRECTIONS[direction]
    d, D = unit_cell.A.shape[0], unit_cell.A.shape[1]
    shapes = bp._msg_shapes(messages)
    futs = []
    for side in (bu_side, td_side):
        comp = bp.compile_side_program(N, d, D, side, chi, shapes, None, depth="ToCore", epilogue=False, arbitrary=bp._is_arbitrary(unit_cell))
        batch = [bp._side_inputs(unit_cell, messages, comp)]
        futs.append(bp._pool.submit(bp._run_side, side, comp, batch, device))
    res = {}
    for f in futs:
        side, outs, slots, rc = f.result()
        bp.raise_if_not_converged(rc, f"ToCore contraction towards {side}")
        if slots[0, bp.SLOT_NONFINITE] > 0:
            raise bp.BubbleConError(f"non-finite values in the ToCore contraction towards {side}")
        o = outs[0]
        res[side] = [o[f"out{k}"] for k in range(len(o))]
    return edge_env.core_env_tensors(backend(), N, res[bu_side], res[td_side], direction)


def edge_tn(unit_cell: UnitCell, env12, N: int, mode: str, edge: str, chi: int):
    """(Ti, Tj, mps_env, info) of the EdgeTN in canonical order (reduce_core_to_mode + reduce_mode_to_edge)."""
    return edge_env.edge_environment(backend(), N, unit_cell.tensors(), env12, mode, edge, chi, _device_bubblecon_fn)


_PAULI = {"x": np.array([[0, 1], [1, 0]], dtype=np.complex128), "y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
          "z": np.array([[1, 0], [0, -1]], dtype=np.complex128)}


def expectation_values_with_rdm(rdm) -> dict:
    """(src/algo/measurements.py:522-545)  {pauli: (<sigma> on site i, <sigma> on site j)} from the one-site marginals of a
    two-site RDM rho[i, i*, j, j*] -- 2 x 2 host arithmetic, as in the reference."""
    rho_i = np.trace(rdm, axis1=2, axis2=3)
    rho_j = np.trace(rdm, axis1=0, axis2=1)
    return {k: (float(np.real(np.trace(s @ rho_i))), float(np.real(np.trace(s @ rho_j)))) for k, s in _PAULI.items()}


def negativity_of_rdm(rdm) -> float:
    """(src/algo/measurements.py:91-94, src/physics/metrics/_negativity.py:43-73)  sum of |negative eigenvalues| of the partial
    transpose (first site) of the 4 x 4 two-site density matrix."""
    pt = np.transpose(rdm, (1, 0, 2, 3))                               # i <-> i*
    mat = np.transpose(pt, (0, 2, 1, 3)).reshape(4, 4)
    w = np.linalg.eigvals(mat)
    return float(sum(abs(x) for x in w if x.real < 0))


@dataclass
class MeasurementsOnUnitCell:
    """(src/containers/results.py:8-30)"""
    energies: dict
    rdms: dict = field(default_factory=dict)
    expectations: dict = field(default_factory=dict)      # {A|B|C: {x|y|z: mean over the edges the site appears on}}
    entanglement: dict = field(default_factory=dict)      # negativity per edge

    @property
    def mean_energy(self) -> float:
        return sum(self.energies.values()) / 3

    @property
    def mean_expectation_values(self) -> dict:
        return {k: sum(self.expectations[f][k] for f in "ABC") / 3 for k in "xyz"}


def measure_energies(unit_cell: UnitCell, messages: dict, N: int, chi: int, h=None, mode: str = "A", env12=None) -> MeasurementsOnUnitCell:
    """(src/algo/measurements.py:163-243, energies part)"""
    h = ite.heisenberg_afm() if h is None else np.asarray(h)
    if env12 is None:
        env12 = reduce_to_core(unit_cell, messages, N, chi)
    B = backend()
    energies, rdms, ent = {}, {}, {}
    acc = {f: {k: [0.0, 0] for k in "xyz"} for f in "ABC"}
    for e in edge_env.EDGES:
        ti, tj, env, info = edge_tn(unit_cell, env12, N, mode, e, chi)
        rho = ite.rho_ij(B, ti, tj, env)
        rdms[e] = rho
        key = f"({e[0]}, {e[1]})"
        energies[key] = float(np.real(np.dot(np.asarray(rho).flatten(), h.flatten())))
        ent[key] = negativity_of_rdm(rho)
        for k, vals in expectation_values_with_rdm(rho).items():
            for v, f in zip(vals, info["flavors"]):
                acc[f][k][0] += v
                acc[f][k][1] += 1
    expectations = {f: {k: (acc[f][k][0] / acc[f][k][1] if acc[f][k][1] else 0.0) for k in "xyz"} for f in "ABC"}
    return MeasurementsOnUnitCell(energies, rdms, expectations, ent)


ENV_HERMICITY_THRESHOLD = 1e-4       # src/algo/imaginary_time_evolution/_constants.py
DEBUG_MODE = False                   # the reference's configuration.json "debug_mode": violations raise instead of warning


@dataclass
class MatrixMetrics:
    """(src/containers/density_matrices.py:5-13)"""
    eigenvalues: list
    negativity: float
    sum_eigenvalues: complex
    hermicity: float
    norm: float
    trace: complex
    other: dict = field(default_factory=dict)


def calc_metrics(rho: np.ndarray) -> MatrixMetrics:
    """(src/algo/density_matrices.py:31-42) of a d^2 x d^2 density matrix: 4 x 4 host arithmetic, as in the reference."""
    w = np.linalg.eigvals(rho)
    return MatrixMetrics(eigenvalues=list(w), negativity=float(sum(abs(np.real(v)) for v in w if np.real(v) < 0)),
                         sum_eigenvalues=complex(np.sum(w)), hermicity=float(np.linalg.norm(rho - np.conj(rho.T)) / np.linalg.norm(rho)),
                         norm=float(np.linalg.norm(rho)), trace=complex(np.trace(rho)))


def _raise_ite_error_or_print_warning(message: str):
    if DEBUG_MODE:
        raise ite.ITEError(message)
    import warnings
    warnings.warn(message, RuntimeWarning, stacklevel=3)


def check_rdms_metrics(rdm) -> MatrixMetrics:
    """the guard of every edge update (src/algo/imaginary_time_evolution/_tn_update.py:51-60): the two-site density matrix
    must be Hermitian to 1e-4, of unit trace (sum of eigenvalues) to 1e-4, and positive up to a total negativity of 0.1."""
    r = np.asarray(rdm)
    d = r.shape[0]
    m = calc_metrics(np.transpose(r, (0, 2, 1, 3)).reshape(d * d, d * d))          # rho_ij_to_rho (density_matrices.py:11-18)
    if m.hermicity > ENV_HERMICITY_THRESHOLD:
        _raise_ite_error_or_print_warning(f"env_hermicity={m.hermicity}")
    if abs(np.real(m.sum_eigenvalues) - 1) > ENV_HERMICITY_THRESHOLD:
        _raise_ite_error_or_print_warning(f"env is not psd. sum-eigenvalues={np.real(m.sum_eigenvalues)}")
    if m.negativity > 0.1:
        _raise_ite_error_or_print_warning(f"env is not psd. negativity={m.negativity}")
    return m


def _original_negativity_ratio(eigen_vals) -> float:
    """(_tn_update.py:63-73)"""
    if eigen_vals is None:
        return 0.0
    ev = np.real(np.asarray(eigen_vals))
    pos, neg = float(np.sum(ev[ev > 0])), float(abs(np.sum(ev[ev < 0])))
    return neg / pos if pos > 0 else 0.0


@dataclass
class ITEStepStats:
    bp_iterations: int = 0
    bp_error: float = 0.0
    als_iterations: int = 0
    truncation_distance: float = 0.0
    t_bp: float = 0.0
    t_reduce: float = 0.0
    t_update: float = 0.0
    env_metrics: MatrixMetrics | None = None      # of the RDM after the update (what ite_update_unit_cell returns, :183-205)
    saved_to: str | None = None


def ite_edge_update(unit_cell: UnitCell, messages: dict | None, N: int, mode: str, edge: str, delta_t: float, bp_config: BPConfig,
                    chi: int, h=None, normalize: bool = True, save: bool | str = False):
    """one loop body of ite_per_mode with bp_every_edge=True: -> (unit_cell, messages, energy_after, ITEStepStats).
    ``save``: keep a copy of the updated unit cell on disk as the reference does after every update (_tn_update.py:203);
    True = default folder, a string = that folder.  The write happens on the background saver thread (persistence.py)."""
    st = ITEStepStats()
    h = ite.heisenberg_afm() if h is None else np.asarray(h)
    g = ite.g_from_exp_h(h, delta_t)
    B = backend()
    t0 = time.perf_counter()
    tn = bp.KagomeTNRepeatedUnitCell(unit_cell, N)
    messages, bp_stats = bp.robust_belief_propagation(tn, messages, bp_config)
    st.bp_iterations, st.bp_error = bp_stats.iterations, bp_stats.final_error
    t1 = time.perf_counter()
    env12 = reduce_to_core(unit_cell, messages, N, chi)
    ti, tj, env, info = edge_tn(unit_cell, env12, N, mode, edge, chi)
    t2 = time.perf_counter()
    d_virtual = ti.shape[1]
    aux = {}
    ti_new, tj_new, origin_eigen_vals = ite.apply_2local_gate(B, g, d_virtual, ti, tj, env, aux=aux)
    # _measures_on_edge before and after the update (_tn_update.py:181, 190): both RDMs from the reduced environment the gate
    # application has built anyway (same contraction as rho_ij, two passes over the ring tensors less); a trivial / product
    # gate does not build it
    check_rdms_metrics(aux["rho_before"] if "rho_before" in aux else ite.rho_ij(B, ti, tj, env))
    last = getattr(ite.ALS_optimization, "last", None)
    if last:
        st.als_iterations, st.truncation_distance = last["iterations"], last["distance"]
    rho = aux["rho_after"] if "rho_after" in aux else ite.rho_ij(B, ti_new, tj_new, env)
    energy = float(np.real(np.dot(np.asarray(rho).flatten(), h.flatten())))
    st.env_metrics = check_rdms_metrics(rho)
    st.env_metrics.other["original_negativity_ratio"] = _original_negativity_ratio(origin_eigen_vals)
    if normalize:
        ti_new = B.scale(ti_new, 1.0 / B.norm(ti_new))
        tj_new = B.scale(tj_new, 1.0 / B.norm(tj_new))
    new_cell = edge_env.write_back(unit_cell.tensors(), info, ti_new, tj_new)
    t3 = time.perf_counter()
    st.t_bp, st.t_reduce, st.t_update = t1 - t0, t2 - t1, t3 - t2
    out_cell = UnitCell(*new_cell, unit_cell._file_name if isinstance(unit_cell, UnitCell) else None)
    if save:
        st.saved_to = out_cell.save(folder=save if isinstance(save, str) else None, asynchronous=True)
    return out_cell, messages, energy, st


def ite_per_mode(unit_cell: UnitCell, messages: dict | None, N: int, mode: str, edge_order, bp_config: BPConfig, chi: int, h=None):
    """(main.py:540-594)  edge_order: iterable of (edge, delta_t)  -> (unit_cell, messages, edge_energies, [stats])"""
    energies, stats = {}, []
    for e, dt in edge_order:
        unit_cell, messages, energy, st = ite_edge_update(unit_cell, messages, N, mode, e, dt, bp_config, chi, h)
        energies[f"({e[0]}, {e[1]})"] = energy
        stats.append(st)
    return unit_cell, messages, energies, stats
