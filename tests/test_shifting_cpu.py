"""Periodic shifts of the non-repeated block and the shift-averaged measurement against the reference's own
permutations and result (tests/golden/shifting.npz, tools/make_golden_shifting.py).  The measurement runs the product's
programs on the numpy interpreter of the op stream."""
import numpy as np
import pytest

from helpers import golden


@pytest.mark.parametrize("N", [2, 3])
def test_shift_permutations_match_reference(N):
    from kagomeperiodicbp_b200 import shifting
    g = golden("shifting.npz")
    for d in ("R", "L", "DL", "DR", "UR", "UL"):
        assert list(shifting.shift_permutation(N, d)) == g[f"N{N}_dir_{d}"].tolist()
    assert shifting.all_shift_permutations(N) == g[f"N{N}_all"].tolist()          # same breadth-first order, identity first
    assert g[f"N{N}_triangles"].tolist() == [[3 * t, 3 * t + 1, 3 * t + 2] for t in range(len(g[f"N{N}_triangles"]))]


def test_shifted_blocks_hold_the_reference_tensors():
    from kagomeperiodicbp_b200 import belief_propagation as bp
    g = golden("shifting.npz")
    sites = [g[f"site{i}"] for i in range(21)]
    tn = bp.KagomeTNArbitrary(sites)
    for shifted, src in zip(tn.all_lattice_shifting_options(), g["shift_sources"].tolist()):
        for t, j in zip(shifted.tensors, src):
            assert np.array_equal(t, sites[j])
    one = tn.shift_periodically_in_direction("R")
    perm = g["N2_dir_R"].tolist()
    for prev, nxt in enumerate(perm):
        for c in range(3):
            assert np.array_equal(one.tensors[3 * nxt + c], sites[3 * prev + c])


def test_shift_averaged_measurement_on_the_interpreter(vm_engines, monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import ite_flow, shifting
    from kagomeperiodicbp_b200.containers import BPConfig
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(ite_flow, "_backend", linalg.ResidentBackend("vm-ite", arena_elems=1 << 23))
    g = golden("shifting.npz")
    sites = [g[f"site{i}"] for i in range(21)]
    chi_bp, chi, term, damping = g["cfg"].tolist()
    cfg = BPConfig(trunc_dim=int(chi_bp), msg_diff_terminate=term, damping=damping, init_msg="UQ")
    e = shifting.calc_measurement_non_unit_cell_kagome_tn(sites, cfg, int(chi))
    assert abs(e - float(g["measurement"][0])) < 1e-8, (e, float(g["measurement"][0]))
