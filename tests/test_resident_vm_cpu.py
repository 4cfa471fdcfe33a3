"""CPU tier: the product's RESIDENT backend (linalg.ResidentBackend: eager tensor programs, first-fit arena recycling,
RArr handles) executed on the numpy interpreter of the op stream (tests/np_vm.py) instead of libkbp.so.  What is under test
is host logic only -- op encoding, buffer lifetimes / reuse in the arena, slot bookkeeping, the backend-generic algorithms
(`reduce_iter.reduceDiter`, `ite.rho_ij`, `ite.apply_2local_gate`) driving it -- against the reference fixtures."""
import numpy as np
import pytest

from helpers import golden
from test_reduce_iter_cpu import CASES, load_case


@pytest.fixture
def resident_vm(monkeypatch):
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    engines = {}

    def fake(key="default", device=0):
        if key not in engines:
            engines[key] = NumpyEngine()
        return engines[key]

    monkeypatch.setattr(linalg, "get_engine", fake)
    return linalg.ResidentBackend("vm", arena_elems=1 << 22)


@pytest.mark.parametrize("name", [c for c in CASES if c != "kagome_like"])
def test_reduce_iter_on_resident_backend(resident_vm, name):
    from kagomeperiodicbp_b200 import reduce_iter
    from kagomeperiodicbp_b200.mps import MPS
    omp, maxD, nr_bulk, max_iter, err, outs, ocor, nr = load_case(name)
    mp = MPS.from_sites([a.copy() for a in omp.A], Corder=list(omp.Corder))
    mp.nr_mantissa, mp.nr_exp = 1.0, 0
    reduce_iter.reduceDiter(resident_vm, mp, maxD, nr_bulk=nr_bulk, max_iter=max_iter, err=err)
    assert mp.Corder == ocor
    for a, o in zip(mp.A, outs):
        assert a.shape == o.shape
        assert np.linalg.norm(a - o) <= 1e-9 * max(1.0, np.linalg.norm(o)), name     # the interpreter's QR is LAPACK's: same gauge
    assert abs(mp.nr_mantissa - nr[0]) <= 1e-9 * abs(nr[0]) and mp.nr_exp == int(nr[1])
    # the arena was recycled, not grown without bound: far fewer live elements than the sum of everything ever allocated
    assert resident_vm.p.peak < 1 << 22


def test_ite_algebra_on_resident_backend(resident_vm):
    from kagomeperiodicbp_b200 import ite
    g = golden("ite_D2_N2.npz")
    for key in ("A_AB", "B_CA"):
        ti, tj = g[f"in_ti_{key}"], g[f"in_tj_{key}"]
        env = [g[f"in_env{k}_{key}"] for k in range(6)]
        rho = ite.rho_ij(resident_vm, ti, tj, env)
        assert np.max(np.abs(np.asarray(rho) - g[f"rdm_{key}"])) < 1e-9
        tin, tjn, w = ite.apply_2local_gate(resident_vm, g["g"], 2, ti, tj, env)
        tin, tjn = np.asarray(tin), np.asarray(tjn)
        tin, tjn = tin / np.linalg.norm(tin), tjn / np.linalg.norm(tjn)
        pair = np.tensordot(tin, tjn, axes=([1], [1]))
        ref = g[f"pair_{key}"]
        ph = np.vdot(pair, ref)
        ph /= abs(ph)
        assert np.linalg.norm(pair * ph - ref) / np.linalg.norm(ref) < 1e-7
