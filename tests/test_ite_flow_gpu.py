"""GPU tier: the device ITE step (ToCore chains, core -> mode -> edge reduction, RDM, gate + ALS) against the REFERENCE's
outputs, starting from the reference's converged messages (tests/golden/ite_D{2,3}_N2.npz), and one full loop body of
ite_per_mode (BP from uniform messages included) against the oracle.  Tolerances: energies 1e-8 (north star), RDM entries
1e-9, updated pair tensor 1e-7 relative."""
import numpy as np
import pytest

from helpers import SIDES, golden

pytestmark = pytest.mark.gpu


def device_messages(g, N):
    from kagomeperiodicbp_b200.containers import Message, MPSOrientation
    from kagomeperiodicbp_b200.lattice import SIDE_OPPOSITE
    from kagomeperiodicbp_b200.mps import MPS
    return {s: Message(MPS.from_sites([g[f"msg_{s}_{k}"] for k in range(2 * N - 1)]), MPSOrientation.standard(SIDE_OPPOSITE[s])) for s in SIDES}


@pytest.mark.parametrize("D", [2, 3])
def test_device_edge_rdm_energy_gate_match_reference(D):
    from kagomeperiodicbp_b200 import edge_env, ite, ite_flow
    from kagomeperiodicbp_b200.containers import UnitCell
    g = golden(f"ite_D{D}_N2.npz")
    cell = UnitCell(g["A"], g["B"], g["C"])
    msgs = device_messages(g, 2)
    chi = int(g["chi"])
    env12 = ite_flow.reduce_to_core(cell, msgs, 2, chi)
    B = ite_flow.backend()
    worst_e = 0.0
    for mode in edge_env.MODES:
        for e in edge_env.EDGES:
            key = f"{mode}_{e}"
            ti, tj, env, info = ite_flow.edge_tn(cell, env12, 2, mode, e, chi)
            rho = ite.rho_ij(B, ti, tj, env)
            assert np.max(np.abs(rho - g[f"rdm_{key}"])) < 1e-9, key
            energy = float(np.real(np.dot(rho.flatten(), g["h"].flatten())))
            worst_e = max(worst_e, abs(energy - g[f"energy_{key}"][0]))
            assert abs(energy - g[f"energy_{key}"][0]) < 1e-8, key
            if D == 2 or (mode == "A" and e in ("AB", "BC")):
                tin, tjn, w = ite.apply_2local_gate(B, g["g"], D, ti, tj, env)
                tin, tjn = tin / np.linalg.norm(tin), tjn / np.linalg.norm(tjn)
                pair = np.tensordot(tin, tjn, axes=([1], [1]))
                ref = g[f"pair_{key}"]
                ph = np.vdot(pair, ref)
                ph /= abs(ph)
                assert np.linalg.norm(pair * ph - ref) / np.linalg.norm(ref) < 1e-7, key
    m = ite_flow.measure_energies(cell, msgs, 2, chi, g["h"], mode="A", env12=env12)
    ref_mean = sum(g[f"energy_A_{e}"][0] for e in edge_env.EDGES) / 3
    assert abs(m.mean_energy - ref_mean) < 1e-8                      # energy per site
    print(f"D={D}: worst edge-energy deviation from the reference {worst_e:.2e}")


def test_full_ite_step_matches_oracle():
    """robust BP from uniform messages + reduction + gate on the device vs the same loop body with oracle numerics."""
    from kagomeperiodicbp_b200 import edge_env, ite, ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    from oracle import bp_np, ite_np
    from oracle.bubblecon_np import bubblecon as obub
    D, N, chi, dt = 2, 2, 18, 0.01
    cell = UnitCell.random(2, D, seed=11)
    cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-8, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ",
                   hermitize_msgs_when_finished=True)
    new_cell, msgs, energy, st = ite_flow.ite_edge_update(cell, None, N, "A", "AB", dt, cfg, chi)
    # oracle
    ocfg = bp_np.BPConfigNP(trunc_dim=2 * D * D, msg_diff_terminate=1e-8, damping=0.1)
    om, ost = bp_np.belief_propagation(N, cell.tensors(), bp_np.uniform_messages(N, D), ocfg)
    assert st.bp_iterations == ost["iterations"]
    bu = bp_np.outgoing_message(N, cell.tensors(), om, "U", chi, depth="ToCore")
    td = bp_np.outgoing_message(N, cell.tensors(), om, "D", chi, depth="ToCore")
    env12 = edge_env.core_env_tensors(ite_np.NP, N, bu.A, td.A)
    fn = lambda T, E, A, ang, order, c, kets: obub(T, E, A, ang, order, D_trunc=c, ket_tensors=kets).A
    ti, tj, env, info = edge_env.edge_environment(ite_np.NP, N, cell.tensors(), env12, "A", "AB", chi, fn)
    h = ite.heisenberg_afm()
    tin, tjn, _ = ite_np.apply_2local_gate(ite.g_from_exp_h(h, dt), D, ti, tj, env)
    rho = ite_np.rho_ij(tin, tjn, env)
    e_ref = float(np.real(np.dot(rho.flatten(), h.flatten())))
    assert abs(energy - e_ref) < 1e-8, (energy, e_ref)
    tin, tjn = tin / np.linalg.norm(tin), tjn / np.linalg.norm(tjn)
    ocell = edge_env.write_back(cell.tensors(), info, tin, tjn)
    # updated unit cell: the pair's bond gauge is free, compare the two updated tensors contracted over it
    fl = {"A": 0, "B": 1, "C": 2}
    a, b = fl[info["flavors"][0]], fl[info["flavors"][1]]
    pa = np.tensordot(new_cell.tensors()[a], new_cell.tensors()[b], axes=0)
    pb = np.tensordot(ocell[a], ocell[b], axes=0)
    # (outer products differ by the bond gauge; compare through the rank-1-invariant: both re-derived pair contractions)
    ti_d = np.transpose(new_cell.tensors()[a], [0] + [1 + p for p in info["perm_i"]])
    tj_d = np.transpose(new_cell.tensors()[b], [0] + [1 + p for p in info["perm_j"]])
    pair_d = np.tensordot(ti_d, tj_d, axes=([1], [1]))
    pair_o = np.tensordot(tin, tjn, axes=([1], [1]))
    ph = np.vdot(pair_d, pair_o)
    ph /= abs(ph)
    assert np.linalg.norm(pair_d * ph - pair_o) / np.linalg.norm(pair_o) < 1e-7
    # the untouched tensor is unchanged
    c = ({0, 1, 2} - {a, b}).pop()
    assert np.array_equal(new_cell.tensors()[c], cell.tensors()[c])


def test_ite_descends_into_the_known_energy_band():
    """known-answer sanity: full-update ITE of a random D=2 unit cell of the Kagome Heisenberg AFM must bring
    the energy per site below the simple-update value region and never below the best variational D=2 energy
    (reference: scripts/plot/afmh_benchmarking.py:34-42, data/unit_cells/best/D=2 energy=-0.4046...)."""
    from kagomeperiodicbp_b200 import edge_env, ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    D, N, chi = 2, 2, 18
    cfg = BPConfig(trunc_dim=8, msg_diff_terminate=1e-6, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ", max_iterations=50)
    cell, msgs = UnitCell.random(2, D, seed=0), None
    energies = []
    for dt, sweeps in ((0.1, 3), (0.05, 2)):
        for _ in range(sweeps):
            for mode in edge_env.MODES:
                cell, msgs, _, _ = ite_flow.ite_per_mode(cell, msgs, N, mode, [(e, dt) for e in edge_env.EDGES], cfg, chi)
            energies.append(ite_flow.measure_energies(cell, msgs, N, chi, mode="A").mean_energy)
    assert energies[-1] < energies[0]
    assert -0.4047 < energies[-1] < -0.375, energies


@pytest.mark.parametrize("direction", ["DL", "DR"])
def test_device_other_core_directions_match_reference(direction):
    from kagomeperiodicbp_b200 import edge_env, ite, ite_flow
    from kagomeperiodicbp_b200.containers import UnitCell
    g = golden("ite_D2_N2.npz")
    cell = UnitCell(g["A"], g["B"], g["C"])
    msgs = device_messages(g, 2)
    chi = int(g["chi"])
    env12 = ite_flow.reduce_to_core(cell, msgs, 2, chi, direction=direction)
    B = ite_flow.backend()
    for e in edge_env.EDGES:
        ti, tj, env, _ = ite_flow.edge_tn(cell, env12, 2, "A", e, chi)
        rho = ite.rho_ij(B, ti, tj, env)
        assert np.max(np.abs(rho - g[f"rdm_{direction}_A_{e}"])) < 1e-9, (direction, e)
        energy = float(np.real(np.dot(rho.flatten(), g["h"].flatten())))
        assert abs(energy - g[f"energy_{direction}_A_{e}"][0]) < 1e-8
