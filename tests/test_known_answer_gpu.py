"""End-to-end known answer on the device: the reference's best D=2 unit cell (tests/golden/best_D2.npz, measured by the
reference itself) through device BP from uniform messages, the ToCore chains, core -> mode -> edge and the two-site RDMs.
Edge energies to 1e-8 of the reference's, energy per site within the 3e-4 config dependence of the published value."""
import numpy as np
import pytest

from helpers import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [2, 3])
def test_best_unit_cell_energy_on_device(N):
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import ite_flow
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    g = golden("best_D2.npz")
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
    assert abs(m.mean_energy - float(g["file_energy"][0])) < 3e-4
