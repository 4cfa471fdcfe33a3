"""CPU tier, build container only (the reference cannot travel): the REAL reference's own code -- belief_propagation,
contract_tensor_network, its containers -- runs with this package bound at the seams of SURVEY 8b
(kagomeperiodicbp_b200/reference_adapter.py) and reproduces what the unmodified reference produced (tests/golden, written by
tools/make_golden*.py).  The device engine is replaced by the numpy interpreter of the op stream (tests/np_vm.py): what is
under test is the drop-in boundary -- argument conventions, returned types the reference can index / copy / pickle, the
iteration logic of the reference driving our step -- not the kernels (GPU tier)."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only mounted in the build container")

from helpers import dense_rel_diff, golden, golden_mps  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_env
    ref_env.setup()
    from libs import bmpslib
    orig = bmpslib._perf_svd
    bmpslib._perf_svd = lambda m, svd_emthod="svd", check_result=False: orig(m, "svd")      # the numpy branch
    yield ref_env
    bmpslib._perf_svd = orig


def _ref_problem(ref_env, D, N, damping):
    from containers import Config
    from enums import MessageModel
    from tensor_networks.construction import kagome_tn_from_unit_cell
    uc = ref_env.seeded_unit_cell(D, 1234)
    config = Config.derive_from_dimensions(D)
    config.dims.big_lattice_size = N
    config.bp.visuals.set_all_progress_bars(False)
    config.bp.damping = damping
    config.bp.msg_diff_terminate = 1e-6
    config.bp.init_msg = MessageModel("UQ")
    tn = kagome_tn_from_unit_cell(uc, config.dims)
    tn.connect_uniform_messages()
    return uc, config, tn


def _to_oracle(m):
    from oracle import mps_np
    r = mps_np.MPS(m.N)
    r.A = list(m.A)
    r.nr_mantissa, r.nr_exp = m.nr_mantissa, m.nr_exp
    return r


@pytest.mark.parametrize("seams", [("S1",), ("S2",)])
def test_reference_bp_runs_on_the_package(ref, vm_engines, seams):
    """reference belief_propagation(...) with bubblecon (S1) resp. the whole BP step (S2) rebound: same iteration count,
    errors and final hermitised messages as the unmodified reference (bp_D2_N2_damp.npz)."""
    from algo import belief_propagation as ref_bp
    from libs import bmpslib
    from kagomeperiodicbp_b200 import reference_adapter as ad
    g = golden("bp_D2_N2_damp.npz")
    uc, config, tn = _ref_problem(ref, 2, 2, 0.1)
    ad.install(seams)
    try:
        msgs, stats = ref_bp.belief_propagation(tn, tn.messages, config.bp)
    finally:
        ad.uninstall()
    assert stats.iterations == int(g["iterations"])
    assert abs(stats.final_error - float(g["final_error"])) < 1e-8
    for side, m in msgs.items():
        assert isinstance(m.mps, bmpslib.mps)                                   # a genuine reference object (a22)
        assert dense_rel_diff(golden_mps(g, str(side)), _to_oracle(m.mps)) < 1e-10
        back = pickle.loads(pickle.dumps(m.mps))                                # what _ite_tracker / unit_cell.save rely on
        assert isinstance(back, bmpslib.mps) and all(np.array_equal(a, b) for a, b in zip(back.A, m.mps.A))
        c = m.mps.copy(mode="full-copy")
        c.A[0][...] = 0                                                         # callers mutate returned sites in place
        assert np.any(m.mps.A[0] != 0)


def test_reference_chain_call_site(ref, vm_engines):
    """contract_tensor_network(tn, side, ToMessage, chi) of the reference with S1 bound == the reference's own output
    (chain_D2_N2.npz), for every side."""
    from algo.contract_tensor_network import contract_tensor_network
    from enums import ContractionDepth
    from lattices.directions import BlockSide
    from kagomeperiodicbp_b200 import reference_adapter as ad
    g = golden("chain_D2_N2.npz")
    uc, config, tn = _ref_problem(ref, 2, 2, None)
    ad.install(("S1",))
    try:
        for side in BlockSide.all_in_counter_clockwise_order():
            mps, order, orientation = contract_tensor_network(tn, side, ContractionDepth.ToMessage, config.bp.trunc_dim, allow_progressbar=False)
            assert dense_rel_diff(golden_mps(g, str(side)), _to_oracle(mps)) < 1e-10, side
    finally:
        ad.uninstall()


def test_reference_energy_measurement_on_the_package(ref, vm_engines):
    """the reference's measure_energies_and_observables_together (reduce_tn -> ModeTN -> six EdgeTN -> rho_ij) with S1 bound for
    every bubblecon it issues (the two ToCore chains and the truncated ModeTN -> EdgeTN chains): the unmodified reference's
    edge energies to 1e-8, returned in the reference's own MeasurementsOnUnitCell."""
    from algo import belief_propagation as ref_bp
    from algo.measurements import measure_energies_and_observables_together
    from enums import UpdateMode
    from physics import hamiltonians
    from kagomeperiodicbp_b200 import reference_adapter as ad
    h = hamiltonians.heisenberg_afm()

    def run(bound):
        uc, config, tn = _ref_problem(ref, 2, 2, 0.1)
        if bound:
            ad.install(("S1",))
        try:
            ref_bp.belief_propagation(tn, tn.messages, config.bp)
            np.random.seed(11)                       # reduce_full_kagome_to_core draws its direction (kagome_to_core.py:178-179)
            return measure_energies_and_observables_together(tn, h, config.contraction, mode=UpdateMode.A)
        finally:
            ad.uninstall()
    base, mine = run(False), run(True)
    assert type(mine) is type(base)
    assert set(mine.energies) == set(base.energies) and len(base.energies) == 6
    for k in base.energies:
        assert abs(base.energies[k] - mine.energies[k]) < 1e-8, (k, base.energies[k], mine.energies[k])
    assert abs(base.mean_energy - mine.mean_energy) < 1e-8


@pytest.fixture
def all_vm(vm_engines, monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import ite_flow
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(ite_flow, "_backend", linalg.ResidentBackend("vm-ite", arena_elems=1 << 23))


def test_reference_ite_edge_update_on_the_package(ref, all_vm):
    """the reference's own ITE update of one edge -- reduce_tn to the EdgeTN, edge_and_environment, _tn_update.rho_ij /
    apply_2local_gate -- with S1 (every bubblecon) and S4 (RDM, gate + ALS) bound: energy before/after the gate to 1e-8 and
    the updated pair tensor to 1e-7 of the unmodified reference (the ALS tolerance of the reference is 1e-6 on the distance)."""
    from algo import belief_propagation as ref_bp
    from algo.imaginary_time_evolution import _tn_update as upd
    from algo.tn_reduction import reduce_core_to_mode, reduce_full_kagome_to_core, reduce_mode_to_edge
    from containers import UpdateEdge
    from enums import UpdateMode
    from lattices.directions import BlockSide
    from physics.hamiltonians import heisenberg_afm
    from kagomeperiodicbp_b200 import reference_adapter as ad
    h = np.asarray(heisenberg_afm())

    def run(bound):
        uc, config, tn = _ref_problem(ref, 2, 2, 0.1)
        if bound:
            ad.install(("S1", "S4"))
        try:
            ref_bp.belief_propagation(tn, tn.messages, config.bp)
            core = reduce_full_kagome_to_core(tn, config.contraction, direction=BlockSide.U)
            mode_tn = reduce_core_to_mode(core, UpdateMode.A)
            res = []
            for e in list(UpdateEdge.all_options())[:3]:
                et = reduce_mode_to_edge(mode_tn, e, config.contraction, arange_legs=False)
                et.rearrange_tensors_and_legs_into_canonical_order()
                t1, t2, env = et.edge_and_environment()
                rdm = upd.rho_ij(t1, t2, mps_env=env)
                g = upd.g_from_exp_h(heisenberg_afm(), 1e-2)
                t1n, t2n, eig = upd.apply_2local_gate(g=g, Dmax=2, Ti=t1, Tj=t2, mps_env=env)
                rdm2 = upd.rho_ij(t1n, t2n, mps_env=env)
                pair = np.tensordot(t1n / np.linalg.norm(t1n), t2n / np.linalg.norm(t2n), axes=([1], [1]))
                res.append((float(np.real(np.dot(rdm.flatten(), h.flatten()))), float(np.real(np.dot(rdm2.flatten(), h.flatten()))), pair))
            return res
        finally:
            ad.uninstall()
    base, mine = run(False), run(True)
    for (e0, e1, p), (f0, f1, q) in zip(base, mine):
        assert abs(e0 - f0) < 1e-8 and abs(e1 - f1) < 1e-8, (e0, f0, e1, f1)
        assert np.linalg.norm(p - q) < 1e-7 * np.linalg.norm(p)
