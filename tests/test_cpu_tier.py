"""CPU tier (-m "not gpu"): oracle vs the reference's golden vectors, geometry/swallow orders vs the reference,
host-side program compilation through the numpy op-stream interpreter, and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, SIDES, dense_rel_diff, golden, golden_mps, to_oracle_mps
from kagomeperiodicbp_b200 import block_tn, contraction_order
from kagomeperiodicbp_b200.lattice import SIDE_ANGLE, get_block
from oracle import bp_np, mps_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ geometry vs reference
@pytest.mark.parametrize("N", [2, 3, 4, 5])
def test_geometry_and_orders_match_reference(N):
    g = golden(f"geometry_N{N}.npz")
    blk = get_block(N)
    D = 2
    msgs = {s: [m for m in bp_np.uniform_messages(N, D)[s].A] for s in SIDES}
    cell = [np.zeros((2, D, D, D, D))] * 3
    T, E, A, K, P = block_tn.assemble(N, cell, msgs)
    assert ["|".join(map(str, e)) for e in E] == list(g["edges"])
    for a, ga in zip(A, g["angles"]):
        assert np.allclose(a, ga[:len(a)], atol=1e-12)
    assert np.allclose(np.array(P, float), g["positions"], atol=1e-9)
    assert list(K) == list(g["kets"])
    for side in SIDES:
        assert blk.boundary_edges[side] == list(g[f"boundary_{side}"])
        for depth in ("ToMessage", "ToCore", "Full"):
            assert list(contraction_order.kagome_order(N, side, depth)) == list(g[f"order_{side}_{depth}"]), (N, side, depth)


def test_block_counts():
    for N in (2, 3, 6):
        blk = get_block(N)
        assert blk.n_sites == 3 * (3 * N * N - 3 * N + 1)
        for side in SIDES:
            assert len(blk.boundary_edges[side]) == 2 * N - 1
            assert len(contraction_order.kagome_order(N, side, "ToMessage")) == 5 * (2 * N - 1) + blk.n_sites


# ------------------------------------------------------------------ oracle vs reference
@pytest.mark.parametrize("D,N", [(2, 2), (2, 3), (3, 2), (4, 2)])
def test_oracle_chain_matches_reference(D, N):
    """(D=4 is the benchmarked bond dimension: chain_D4_N2.npz, and chain_D4_N3.npz on the GPU tier)"""
    g = golden(f"chain_D{D}_N{N}.npz")
    cell = (g["A"], g["B"], g["C"])
    um = bp_np.uniform_messages(N, D)
    for side in SIDES:
        if f"{side}_site0" not in g:
            continue
        mine = bp_np.outgoing_message(N, cell, um, side, int(g["chi"]))
        assert dense_rel_diff(golden_mps(g, side), mine) < 1e-12


@pytest.mark.parametrize("tag,D,N,damping", [("bp_D2_N2", 2, 2, None), ("bp_D2_N2_damp", 2, 2, 0.1)])
def test_oracle_bp_matches_reference(tag, D, N, damping):
    g = golden(tag + ".npz")
    cfg = bp_np.BPConfigNP(trunc_dim=int(g["chi"]), msg_diff_terminate=float(g["terminate"]), damping=damping)
    msgs, st = bp_np.belief_propagation(N, (g["A"], g["B"], g["C"]), bp_np.uniform_messages(N, D), cfg)
    assert st["iterations"] == int(g["iterations"])
    assert abs(st["final_error"] - float(g["final_error"])) < 1e-12
    for side in SIDES:
        assert dense_rel_diff(golden_mps(g, side), msgs[side]) < 1e-12


def test_oracle_mps_edge_cases():
    rng = np.random.default_rng(3)
    m = mps_np.MPS(2)                       # N < 3: reduceD is a no-op (reference bmpslib.py:875-876)
    m.set_site(rng.normal(size=(1, 4, 9)), 0)
    m.set_site(rng.normal(size=(9, 4, 1)), 1)
    assert m.reduceD(2) is None and m.A[0].shape == (1, 4, 9)
    m3 = mps_np.MPS(3)
    for i, sh in enumerate([(1, 2, 2), (2, 2, 2), (2, 2, 1)]):
        m3.set_site(rng.normal(size=sh) + 0j, i)
    before = mps_np.mps_to_dense(m3)
    assert m3.reduceD(8) == 0                # nothing oversized -> untouched
    m3.right_canonical(nr_bulk=True)
    assert np.allclose(mps_np.mps_to_dense(m3), before)
    with pytest.raises(FloatingPointError):
        m3.A[2][0, 0, 0] = np.nan
        m3.Corder[2] = None
        m3.right_canonical()


# ------------------------------------------------------------------ host logic through the op-stream interpreter
def test_bp_host_pipeline_via_interpreter(vm_engines):
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    g = golden("bp_D2_N2_damp.npz")
    cell = UnitCell(g["A"], g["B"], g["C"])
    tn = bp.KagomeTNRepeatedUnitCell(cell, 2)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=8, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    msgs, stats = bp.belief_propagation(tn, tn.messages, cfg)
    assert stats.iterations == int(g["iterations"])
    assert np.allclose(stats.errors, g["errors"], atol=1e-10)
    for side in SIDES:
        assert dense_rel_diff(golden_mps(g, side), to_oracle_mps(msgs[side].mps)) < 1e-10


def test_bubblecon_api_via_interpreter(vm_engines):
    from kagomeperiodicbp_b200.bubblecon import bubblecon
    from oracle.bubblecon_np import bubblecon as obub
    D, N = 2, 2
    g = golden("chain_D2_N2.npz")
    cell = (g["A"], g["B"], g["C"])
    msgs = {s: m.A for s, m in bp_np.uniform_messages(N, D).items()}
    T, E, A, K, P = block_tn.assemble(N, cell, msgs)
    T, E, A = block_tn.connect_corner(N, T, E, A, P, "D")
    mp = bubblecon(T, E, A, SIDE_ANGLE["D"], list(contraction_order.kagome_order(N, "D", "ToMessage")), D_trunc=8, ket_tensors=K)
    assert dense_rel_diff(golden_mps(g, "D"), to_oracle_mps(mp)) < 1e-10
    # scalar branch
    order = list(contraction_order.kagome_order(N, "D", "Full"))
    val, ex = bubblecon(T, E, A, SIDE_ANGLE["D"], order, D_trunc=18, ket_tensors=K, separate_exp=True)
    oval, oex = obub(T, E, A, SIDE_ANGLE["D"], order, D_trunc=18, ket_tensors=K, separate_exp=True)
    assert abs(val * 10.0 ** ex - oval * 10.0 ** oex) <= 1e-10 * abs(oval * 10.0 ** oex)
    with pytest.raises(NotImplementedError):
        bubblecon(T, E, A, SIDE_ANGLE["D"], order, D_trunc=8, D_trunc2=16, ket_tensors=K)


def test_program_allocator_reuses_and_never_overlaps():
    from kagomeperiodicbp_b200.program import Program
    p = Program()
    a = p.new((100,))
    b = p.new((50,))
    off_a = a.off
    del a
    c = p.new((60,))
    assert c.off == off_a                     # first fit into the freed block
    live = [(b.off, b.size), (c.off, c.size)]
    assert live[0][0] + live[0][1] <= live[1][0] or live[1][0] + live[1][1] <= live[0][0]


# ------------------------------------------------------------------ C ABI surface (no compute without a GPU)
def test_cabi_library_exports_every_declared_symbol():
    from kagomeperiodicbp_b200.build import build_library
    from kagomeperiodicbp_b200.engine import EXPORTED
    lib = ctypes.CDLL(build_library())
    header = open(os.path.join(ROOT, "include", "kbp.h")).read()
    declared = set(re.findall(r"\b(kbp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kbp.h but not exported"
    assert set(EXPORTED) <= declared
    lib.kbp_svd_work_elems.restype = ctypes.c_int64
    lib.kbp_svd_work_elems.argtypes = [ctypes.c_int64, ctypes.c_int64]
    from kagomeperiodicbp_b200.engine import svd_work_elems
    for m, n in [(512, 512), (256, 512), (162, 81), (5, 3)]:
        assert lib.kbp_svd_work_elems(m, n) == svd_work_elems(m, n)
    # the Python mirrors that size program buffers must follow the library's rules (small-kernel fit, block width)
    lib.kbp_svd_warm_elems.restype = ctypes.c_int64
    lib.kbp_svd_warm_elems.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
    lib.kbp_qr_work_elems.restype = ctypes.c_int64
    lib.kbp_qr_work_elems.argtypes = [ctypes.c_int64, ctypes.c_int64]
    from kagomeperiodicbp_b200.engine import qr_work_elems, svd_warm_elems
    for m, n, k in [(512, 512, 32), (256, 512, 32), (162, 162, 18), (81, 162, 18), (162, 81, 18), (84, 162, 18), (128, 110, 32), (129, 110, 32),
                    (48, 512, 32), (2592, 2592, 72), (16, 32, 8), (100, 37, 37), (1296, 648, 72)]:
        assert lib.kbp_svd_warm_elems(m, n, k) == svd_warm_elems(m, n, k), (m, n, k)
        assert lib.kbp_qr_work_elems(m, n) == qr_work_elems(m, n)


def test_no_cpu_fallback_without_device():
    """on a box without a GPU the engine must refuse to exist -- the product path never computes on the host."""
    from kagomeperiodicbp_b200.engine import Engine, EngineUnavailable, load_library
    if load_library().kbp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(EngineUnavailable):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "kagomeperiodicbp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_compression_selection_rule_and_chain_entry(vm_engines):
    """the reference's rule D <= 10 -> SVD, else iterative (src/containers/contractions.py:27-35), and the
    contract_tensor_network mirror that applies it (ToMessage chain on the interpreter == oracle)."""
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200.containers import BubbleConGlobalConfig, UnitCell
    gc = BubbleConGlobalConfig()
    assert gc.bubblecon_compression(10) == {"type": "SVD"}
    assert gc.bubblecon_compression(11) == {"type": "iter", "max-iter": 200, "err": 1e-8}
    D, N = 2, 2
    cell = UnitCell.random(2, D, seed=4)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    mps, order, orientation = bp.contract_tensor_network(tn, "UR", "ToMessage", 8)
    mine = bp_np.outgoing_message(N, cell.tensors(), bp_np.uniform_messages(N, D), "UR", 8)
    from helpers import to_oracle_mps
    assert dense_rel_diff(mine, to_oracle_mps(mps)) < 1e-12
    assert orientation.open_towards == "UR" and len(order) == 5 * (2 * N - 1) + 21


def test_unit_cell_persistence_and_rdm_guard(tmp_path):
    """per-step save of the unit cell (async writer, atomic replace) and the RDM guard of the ITE update
    (src/algo/imaginary_time_evolution/_tn_update.py:51-60, 203)."""
    import warnings
    from kagomeperiodicbp_b200 import ite_flow, persistence
    from kagomeperiodicbp_b200.containers import UnitCell
    uc = UnitCell.random(2, 3, seed=1)
    uc.set_filename("cell_a")
    p = uc.save(folder=str(tmp_path), asynchronous=True)
    persistence.saver().flush()
    back = UnitCell.load("cell_a", folder=str(tmp_path))
    assert p.endswith("cell_a.dat") and all(np.array_equal(a, b) for a, b in zip(back.tensors(), uc.tensors()))
    assert UnitCell.load("last", folder=str(tmp_path))._file_name == "cell_a"
    assert UnitCell.load("missing", folder=str(tmp_path)) is None
    rho = np.zeros((2, 2, 2, 2), complex)
    rho[0, 0, 0, 0] = rho[1, 1, 1, 1] = 0.5
    m = ite_flow.check_rdms_metrics(rho)
    assert m.hermicity == 0 and abs(m.sum_eigenvalues - 1) < 1e-15 and m.negativity == 0
    bad = rho.copy()
    bad[0, 1, 0, 0] = 0.3                                    # not Hermitian
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        ite_flow.check_rdms_metrics(bad)
    assert any("hermicity" in str(x.message) for x in w)
    ite_flow.DEBUG_MODE = True
    try:
        with pytest.raises(ite_flow.ite.ITEError):
            ite_flow.check_rdms_metrics(bad)
    finally:
        ite_flow.DEBUG_MODE = False
