"""CPU tier: the product's whole ITE step (device BP programs, ToCore chains, resident-backend reduction, RDM, gate + ALS) with
every engine routed to the numpy interpreter of the op stream -- the bodies of the GPU-tier tests in test_ite_flow_gpu.py,
so the host logic of the step (program compilation, arena lifetimes, slot bookkeeping, message relabelling, write-back) is
checked against the reference fixtures and the oracle without a GPU."""
import pytest


@pytest.fixture
def all_vm(vm_engines, monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import ite_flow
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(ite_flow, "_backend", linalg.ResidentBackend("vm-ite", arena_elems=1 << 23))


def test_edge_rdm_energy_gate_match_reference_on_the_interpreter(all_vm):
    import test_ite_flow_gpu as t
    t.test_device_edge_rdm_energy_gate_match_reference(2)


def test_full_ite_step_matches_oracle_on_the_interpreter(all_vm):
    import test_ite_flow_gpu as t
    t.test_full_ite_step_matches_oracle()
