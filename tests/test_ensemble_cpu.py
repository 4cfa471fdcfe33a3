"""CPU tier: the ensemble launcher (SURVEY 8f-4 / BASELINE config C5) on the op-stream interpreter: cells batched into the same
launches converge individually and give what the one-cell path gives; the CSV carries the condor worker's columns."""
import csv

import numpy as np
import pytest


@pytest.fixture
def all_vm(vm_engines, monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import ite_flow
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(ite_flow, "_backend", linalg.ResidentBackend("vm-ite", arena_elems=1 << 23))


def test_batched_bp_matches_single_cell_bp(all_vm):
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import ensemble
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    D, N = 2, 2
    cells = [UnitCell.random(2, D, seed=s) for s in (0, 1, 2)]
    cfg = BPConfig(trunc_dim=8, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    res = ensemble.belief_propagation_batch(N, cells, cfg, batch=2)
    for cell, (msgs, its, err, ok) in zip(cells, res):
        tn = bp.KagomeTNRepeatedUnitCell(cell, N)
        tn.connect_uniform_messages()
        m1, st = bp.belief_propagation(tn, tn.messages, cfg)
        assert ok and st.success and its == st.iterations and err == st.final_error
        for s in m1:
            assert all(np.allclose(a, b, atol=1e-13) for a, b in zip(m1[s].mps.A, msgs[s].mps.A))


def test_run_ensemble_rows_and_csv(all_vm, tmp_path):
    from kagomeperiodicbp_b200 import ensemble
    rows0 = ensemble.run_ensemble(range(4), 2, 2, rank=0, world=2, batch=2)
    rows1 = ensemble.run_ensemble(range(4), 2, 2, rank=1, world=2, batch=2)
    assert [r["seed"] for r in rows0] == [0, 2] and [r["seed"] for r in rows1] == [1, 3]
    path = ensemble.write_csv(rows0 + rows1, str(tmp_path / "condor" / "results_ite_afm.csv"))
    with open(path) as f:
        got = list(csv.reader(f))
    assert got[0] == ensemble.RESULT_KEYS and [int(r[0]) for r in got[1:]] == [0, 1, 2, 3]
    for r in rows0 + rows1:
        assert -0.75 < r["energy"] < 0.75 and r["bp_error"] < 1e-5 and r["D"] == 2
