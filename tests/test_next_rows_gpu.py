"""GPU tier: device runs of the SURVEY 8f rows -- bubblecon with iterative (reduceDiter) compression, the non-repeated block
(KagomeTNArbitrary), the shift-averaged measurement, the best D=3 unit cell -- and the run-to-run determinism of a BP step.
Their op streams are also validated on the CPU against the reference fixtures through the numpy interpreter
(tests/test_chain_iter_cpu.py, tests/test_arbitrary_tn_cpu.py, tests/test_shifting_cpu.py)."""
import os

import numpy as np
import pytest

from helpers import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("side", ["D", "UL"])
def test_iterative_compression_chain_on_device(side):
    from kagomeperiodicbp_b200 import bubblecon as bc
    from test_chain_iter_cpu import load
    T, meta, outs, cor, nr = load(side)
    mp = bc.bubblecon(T, meta["edges"], meta["angles"], meta["bubble_angle"], meta["order"], D_trunc=meta["D_trunc"],
                      ket_tensors=meta["kets"], compression=meta["compression"])
    assert [a.shape for a in mp.A] == [o.shape for o in outs] and list(mp.Corder) == cor
    assert abs(mp.nr_mantissa - nr[0]) <= 1e-9 * abs(nr[0]) and mp.nr_exp == int(nr[1])

    def dense(sites):
        t = np.asarray(sites[0])
        for a in sites[1:]:
            t = np.tensordot(t, np.asarray(a), ([t.ndim - 1], [0]))
        return t
    ref = dense(outs)
    assert np.linalg.norm(dense(mp.A) - ref) <= 1e-9 * np.linalg.norm(ref)      # QR gauge differs from LAPACK's: compare as states


def test_arbitrary_block_on_device():
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig
    from test_arbitrary_tn_cpu import load
    g, sites, chi_bp, chi, iters, term, damping, ref = load()
    tn = bp.KagomeTNArbitrary(sites)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=chi_bp, msg_diff_terminate=term, damping=damping, init_msg="UQ")
    msgs, stats = bp.belief_propagation(tn, tn.messages, cfg)
    assert stats.iterations == iters
    m = ite_flow.measure_energies(tn.unit_cell, msgs, 2, chi, g["h"], mode="A")
    for e, v in ref.items():
        assert abs(m.energies[f"({e[0]}, {e[1]})"] - v) < 1e-8, (e, m.energies)


def test_shift_averaged_measurement_on_device():
    from kagomeperiodicbp_b200 import shifting
    from kagomeperiodicbp_b200.containers import BPConfig
    g = golden("shifting.npz")
    sites = [g[f"site{i}"] for i in range(21)]
    chi_bp, chi, term, damping = g["cfg"].tolist()
    cfg = BPConfig(trunc_dim=int(chi_bp), msg_diff_terminate=term, damping=damping, init_msg="UQ")
    e = shifting.calc_measurement_non_unit_cell_kagome_tn(sites, cfg, int(chi))
    assert abs(e - float(g["measurement"][0])) < 1e-8


def test_bp_step_is_bitwise_reproducible():
    """run-to-run determinism on one GPU (SURVEY 8b conventions): no atomics on data, split-K partials and the cluster kernels'
    partial sums are added in a fixed order, so two executions of the same step agree bit for bit."""
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    from helpers import SIDES
    D, N = 4, 2
    cell = UnitCell.random(2, D, seed=9)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    runs = [bp.bp_step_batch(N, [cell], [tn.messages], cfg)[0] for _ in range(3)]
    for r in runs[1:]:
        assert r[2] == runs[0][2]
        for s in SIDES:
            for a, b in zip(r[0][s].mps.A, runs[0][0][s].mps.A):
                assert np.array_equal(a, b)


def test_best_D3_unit_cell_energy_on_device():
    """the D=3 case of tests/test_known_answer_gpu.py (the D=2 cases there were run on the B200; this fixture came later)"""
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    g = golden("best_D3.npz")
    N = 2
    chi_bp, chi, iters, term, damping = g[f"N{N}_cfg"].tolist()
    cell = UnitCell(g["A"], g["B"], g["C"])
    cfg = BPConfig(trunc_dim=int(chi_bp), msg_diff_terminate=term, damping=None if damping < 0 else damping, init_msg="UQ")
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    msgs, stats = bp.belief_propagation(tn, tn.messages, cfg)
    assert stats.iterations == int(iters)
    m = ite_flow.measure_energies(cell, msgs, N, int(chi), g["h"], mode="A")
    ref = dict(zip(g[f"N{N}_edges"].tolist(), g[f"N{N}_edge_energies"].tolist()))
    for e, v in ref.items():
        assert abs(m.energies[f"({e[0]}, {e[1]})"] - v) < 1e-8, (e, m.energies, v)
