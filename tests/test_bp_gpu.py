"""GPU tier: the CUDA path through the C ABI against the oracle and the committed golden vectors.
Tolerances (BASELINE.json north_star): messages 1e-10 relative (dense-contracted, where (D^2)^L is
tractable) and 1-|<ref|new>| <= 1e-10; BP iteration counts equal; per-iteration errors to 1e-8."""
import numpy as np
import pytest

from helpers import SIDES, dense_rel_diff, golden, golden_mps, overlap_defect, to_oracle_mps
from kagomeperiodicbp_b200 import belief_propagation as bp
from kagomeperiodicbp_b200.bubblecon import bubblecon
from kagomeperiodicbp_b200 import block_tn, contraction_order
from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
from kagomeperiodicbp_b200.lattice import SIDE_ANGLE
from oracle import bp_np, bubblecon_np, mps_np

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _cell_from(g):
    return UnitCell(g["A"], g["B"], g["C"])


@pytest.mark.parametrize("D,N", [(2, 2), (2, 3), (3, 2), (3, 3), (4, 2), (4, 3), (6, 2)])
def test_chain_vs_golden_reference(D, N):
    """one ToMessage bubblecon call per side, uniform messages: device result vs the REFERENCE's own output
    (D=4: the benchmarked bond dimension; the N=3 fixture holds two of the six sides to bound its size;
    D=6, chi_bp=72: BASELINE config C4 -- 2592 x 2592 truncations on the wide-block subspace path, one side)."""
    g = golden(f"chain_D{D}_N{N}.npz")
    cell = _cell_from(g)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    msgs = {s: m.mps.A for s, m in tn.messages.items()}
    for side in SIDES:
        if f"{side}_site0" not in g:
            continue
        T, E, A, K, P = block_tn.assemble(N, cell.tensors(), msgs)
        T, E, A = block_tn.connect_corner(N, T, E, A, P, side)
        order = list(contraction_order.kagome_order(N, side, "ToMessage"))
        mp = bubblecon(T, E, A, SIDE_ANGLE[side], order, D_trunc=int(g["chi"]), ket_tensors=K, separate_exp=True)
        ref = golden_mps(g, side)
        assert [a.shape for a in mp.A] == [a.shape for a in ref.A]
        assert dense_rel_diff(ref, to_oracle_mps(mp)) < TOL, (D, N, side)
        assert overlap_defect(ref, to_oracle_mps(mp)) < TOL


@pytest.mark.parametrize("tag,D,N,damping", [("bp_D2_N2", 2, 2, None), ("bp_D2_N2_damp", 2, 2, 0.1), ("bp_D2_N3_damp", 2, 3, 0.1),
                                             ("bp_D3_N2_damp", 3, 2, 0.1), ("bp_D4_N2_damp", 4, 2, 0.1)])
def test_bp_vs_golden_reference(tag, D, N, damping):
    """belief_propagation to convergence vs the REFERENCE's run: iteration count, error trace, final messages."""
    g = golden(tag + ".npz")
    cell = _cell_from(g)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=int(g["chi"]), msg_diff_terminate=float(g["terminate"]), damping=damping, init_msg="UQ")
    msgs, stats = bp.belief_propagation(tn, tn.messages, cfg)
    assert stats.iterations == int(g["iterations"])
    assert np.allclose(stats.errors, g["errors"], rtol=0, atol=1e-8)
    assert abs(stats.final_error - float(g["final_error"])) < 1e-8
    for side in SIDES:
        ref = golden_mps(g, side)
        assert dense_rel_diff(ref, to_oracle_mps(msgs[side].mps)) < TOL, (tag, side)


def test_bp_step_batch_ensemble_matches_oracle():
    """an ensemble of independent unit cells through ONE batched launch sequence vs the oracle, cell by cell."""
    D, N = 2, 2
    cells = [UnitCell.random(2, D, seed=s) for s in (0, 1, 2, 3)]
    tn = bp.KagomeTNRepeatedUnitCell(cells[0], N)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=8, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    res = bp.bp_step_batch(N, cells, [tn.messages] * len(cells), cfg)
    ocfg = bp_np.BPConfigNP(trunc_dim=8, msg_diff_terminate=1e-6, damping=0.1)
    for cell, (out, nxt, err, _) in zip(cells, res):
        o_out, o_nxt, o_err = bp_np.bp_step(N, cell.tensors(), bp_np.uniform_messages(N, D), ocfg)
        assert abs(err - o_err) < 1e-10
        for side in SIDES:
            assert dense_rel_diff(o_out[side], to_oracle_mps(out[side].mps)) < TOL
            assert dense_rel_diff(o_nxt[side], to_oracle_mps(nxt[side].mps)) < TOL


def test_full_contraction_scalar_matches_oracle():
    """ContractionDepth.Full: the scalar branch (mantissa, exp10) of bubblecon (reference :3077-3088)."""
    D, N = 2, 2
    cell = UnitCell.random(2, D, seed=11)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    msgs = {s: m.mps.A for s, m in tn.messages.items()}
    T, E, A, K, P = block_tn.assemble(N, cell.tensors(), msgs)
    T, E, A = block_tn.connect_corner(N, T, E, A, P, "U")
    order = list(contraction_order.kagome_order(N, "U", "Full"))
    val, ex = bubblecon(T, E, A, SIDE_ANGLE["U"], order, D_trunc=18, ket_tensors=K, separate_exp=True)
    oval, oex = bubblecon_np.bubblecon(T, E, A, SIDE_ANGLE["U"], order, D_trunc=18, ket_tensors=K, separate_exp=True)
    a, b = val * 10.0 ** ex, oval * 10.0 ** oex
    assert abs(a - b) <= 1e-10 * abs(b)


def test_d4_chain_properties():
    """D=4 (the benchmark size): too big for a dense check, so size-independent properties + oracle overlap."""
    D, N = 4, 2
    cell = UnitCell.random(2, D, seed=5)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    out, nxt, err, trunc = bp.bp_step_batch(N, [cell], [tn.messages], cfg)[0]
    ocfg = bp_np.BPConfigNP(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1)
    o_out, o_nxt, o_err = bp_np.bp_step(N, cell.tensors(), bp_np.uniform_messages(N, D), ocfg)
    assert abs(err - o_err) < 1e-8
    for side in SIDES:
        m = to_oracle_mps(out[side].mps)
        assert abs(mps_np.mps_inner_product(m, m, True) - 1) < 1e-10          # unit norm after _fix_messages
        for a in m.A[1:]:                                                          # right-canonical sites
            M = a.reshape(a.shape[0], -1)
            assert np.linalg.norm(M @ M.conj().T - np.eye(a.shape[0])) < 1e-9
        assert overlap_defect(o_out[side], m) < TOL
        assert overlap_defect(o_nxt[side], to_oracle_mps(nxt[side].mps)) < TOL
        assert dense_rel_diff(o_out[side], m) < TOL                                # (16^3 entries: tractable at N = 2)
    # second iteration: full-rank boundary-MPS spectra, i.e. the regime the subspace-iteration SVD is built for
    before = {k: sum(bp.get_engine(("side", s)).svd_counters()[k] for s in SIDES) for k in ("subspace", "subspace_fallback")}
    out2, nxt2, err2, _ = bp.bp_step_batch(N, [cell], [nxt], cfg)[0]
    after = {k: sum(bp.get_engine(("side", s)).svd_counters()[k] for s in SIDES) for k in ("subspace", "subspace_fallback")}
    assert after["subspace"] - before["subspace"] >= 80 and after["subspace_fallback"] - before["subspace_fallback"] <= 2
    o_out2, o_nxt2, o_err2 = bp_np.bp_step(N, cell.tensors(), o_nxt, ocfg)
    assert abs(err2 - o_err2) < 1e-8
    for side in SIDES:
        assert dense_rel_diff(o_out2[side], to_oracle_mps(out2[side].mps)) < TOL, side
        assert dense_rel_diff(o_nxt2[side], to_oracle_mps(nxt2[side].mps)) < TOL, side


def test_missing_device_path_is_loud():
    from kagomeperiodicbp_b200.engine import Engine, BubbleConError
    e = Engine(0)
    with pytest.raises(BubbleConError):
        e.run(np.array([99, 0, 0], dtype=np.int64))   # unknown opcode before reserve -> error, not silence
    e.close()


def test_robust_bp_retry_and_random_messages():
    """robust_belief_propagation's retry path (trunc_dim x 1.5, +11 iterations, fresh messages: reference :285-350) and random
    quantum initial messages run on the device and end converged."""
    D, N = 2, 2
    cell = UnitCell.random(2, D, seed=21)
    tn = bp.KagomeTNRepeatedUnitCell(cell, N)
    cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-7, msg_diff_good_enough=1e-5, damping=0.1, init_msg="UQ",
                   max_iterations=3, allowed_retries=3)
    tn.connect_uniform_messages()
    msgs, stats = bp.robust_belief_propagation(tn, tn.messages, cfg)
    assert stats.attempts >= 2                      # 3 iterations are not enough: the first attempt must fail
    assert stats.final_error < 1e-5 and stats.success
    rng = np.random.RandomState(4)
    tn2 = bp.KagomeTNRepeatedUnitCell(cell, N)
    tn2.connect_random_messages(rng)
    cfg2 = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-7, damping=None, init_msg="RQ", hermitize_msgs_when_finished=False)
    m2, st2 = bp.belief_propagation(tn2, tn2.messages, cfg2)
    assert st2.success and st2.final_error < 1e-7
    # both runs describe the same fixed point (the first one only to msg_diff_good_enough)
    for side in SIDES:
        assert overlap_defect(to_oracle_mps(msgs[side].mps), to_oracle_mps(m2[side].mps)) < 1e-3


def test_nonfinite_input_raises():
    """the reference prints and exits on nan/inf in the boundary MPS (src/libs/bmpslib.py:711-717): here BubbleConError."""
    from kagomeperiodicbp_b200.engine import BubbleConError
    D, N = 2, 2
    cell = UnitCell.random(2, D, seed=3)
    bad = UnitCell(cell.A.copy(), cell.B.copy(), cell.C.copy())
    bad.A[0, 0, 0, 0, 0] = np.nan
    tn = bp.KagomeTNRepeatedUnitCell(bad, N)
    tn.connect_uniform_messages()
    cfg = BPConfig(trunc_dim=8, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    with pytest.raises(BubbleConError):
        bp.bp_step_batch(N, [bad], [tn.messages], cfg)
