"""The non-repeated Kagome block (KagomeTNArbitrary): 21 independent site tensors, block BP from uniform
messages and the six mode-A edge energies against the reference's own run (tests/golden/arbitrary_D2_N2.npz,
tools/make_golden_arbitrary.py) -- with the oracle, and with the product's programs on the numpy interpreter of the op stream."""
import numpy as np
import pytest

from helpers import golden


def load():
    g = golden("arbitrary_D2_N2.npz")
    sites = [g[f"site{i}"] for i in range(21)]
    chi_bp, chi, iters, term, damping, final_err = g["cfg"].tolist()
    ref = dict(zip(g["edges"].tolist(), g["edge_energies"].tolist()))
    return g, sites, int(chi_bp), int(chi), int(iters), term, damping, ref


def test_core_table_order_matches_block_indices():
    from kagomeperiodicbp_b200 import edge_env
    tb = edge_env.tables()["core"]
    idx = edge_env.core_site_indices(2)
    assert ["ABC"[i % 3] for i in idx] == [tb[k]["name"] for k in range(9)]


def test_oracle_arbitrary_block_matches_reference():
    from kagomeperiodicbp_b200 import edge_env
    from oracle import bp_np, ite_np
    from oracle.bubblecon_np import bubblecon as obub
    g, sites, chi_bp, chi, iters, term, damping, ref = load()
    N = 2
    cfg = bp_np.BPConfigNP(trunc_dim=chi_bp, msg_diff_terminate=term, damping=damping)
    msgs, st = bp_np.belief_propagation(N, sites, bp_np.uniform_messages(N, 2), cfg)
    assert st["iterations"] == iters
    bu = bp_np.outgoing_message(N, sites, msgs, "U", chi, depth="ToCore")
    td = bp_np.outgoing_message(N, sites, msgs, "D", chi, depth="ToCore")
    env12 = edge_env.core_env_tensors(ite_np.NP, N, bu.A, td.A)
    fn = lambda T, E, A, ang, order, c, kets: obub(T, E, A, ang, order, D_trunc=c, ket_tensors=kets).A
    for e, v in ref.items():
        ti, tj, env, info = edge_env.edge_environment(ite_np.NP, N, sites, env12, "A", e, chi, fn)
        rho = ite_np.rho_ij(ti, tj, env)
        assert np.max(np.abs(rho - g[f"rdm_{e}"])) < 1e-9, e
        assert abs(float(np.real(np.dot(rho.flatten(), g["h"].flatten()))) - v) < 1e-8, e


def test_product_arbitrary_block_on_the_interpreter(vm_engines, monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(ite_flow, "_backend", linalg.ResidentBackend("vm-ite", arena_elems=1 << 23))
    g, sites, chi_bp, chi, iters, term, damping, ref = load()
    tn = bp.KagomeTNArbitrary(sites)
    assert tn.N == 2
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=chi_bp, msg_diff_terminate=term, damping=damping, init_msg="UQ")
    msgs, stats = bp.belief_propagation(tn, tn.messages, cfg)
    assert stats.iterations == iters
    m = ite_flow.measure_energies(tn.unit_cell, msgs, 2, chi, g["h"], mode="A")
    for e, v in ref.items():
        assert abs(m.energies[f"({e[0]}, {e[1]})"] - v) < 1e-8, (e, m.energies)
