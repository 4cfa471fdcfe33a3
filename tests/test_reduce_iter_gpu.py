"""reduceDiter on the device (ResidentBackend: ZGEMM + cluster Householder QR + normalise kernels) against the reference's
outputs (tests/golden/reduce_iter.npz) and the CPU oracle.  The QR gauge of the device kernel differs from LAPACK's by a
phase per bond index, so tensors are compared as states (dense contraction); shapes, canonical tags, the (mantissa, exp)
scale, the number of rounds and the stopping measure must agree exactly / to rounding."""
import numpy as np
import pytest

from kagomeperiodicbp_b200 import reduce_iter
from kagomeperiodicbp_b200.mps import MPS as DevMPS
from oracle.reduce_iter_np import reduceDiter as oracle_reduce
from test_reduce_iter_cpu import CASES, load_case

pytestmark = pytest.mark.gpu


def dense(sites):
    """contract the bonds, keep the physical legs and the two (possibly open) end bonds"""
    t = np.asarray(sites[0])
    for a in sites[1:]:
        t = np.tensordot(t, np.asarray(a), ([t.ndim - 1], [0]))
    return t


@pytest.mark.parametrize("name", CASES)
def test_device_matches_reference(name):
    omp, maxD, nr_bulk, max_iter, err, outs, ocor, nr = load_case(name)
    dmp = DevMPS.from_sites([a.copy() for a in omp.A], Corder=list(omp.Corder))
    dmp.nr_mantissa, dmp.nr_exp = 1.0, 0
    st_d, st_o = {}, {}
    reduce_iter.reduceDiter(reduce_iter.backend(), dmp, maxD, nr_bulk=nr_bulk, max_iter=max_iter, err=err, stats=st_d)
    oracle_reduce(omp, maxD, nr_bulk=nr_bulk, max_iter=max_iter, err=err, stats=st_o)
    assert dmp.Corder == ocor
    assert [a.shape for a in dmp.A] == [o.shape for o in outs]
    assert st_d["rounds"] == st_o["rounds"]
    if st_o["delta"] is not None:
        assert abs(st_d["delta"] - st_o["delta"]) <= 1e-8 * max(st_o["delta"], 1e-12) + 1e-14
    assert abs(dmp.nr_mantissa - nr[0]) <= 1e-10 * abs(nr[0]) and dmp.nr_exp == int(nr[1])
    ref = dense(outs)
    got = dense(dmp.A)
    assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref), name
    # canonical tags are honest (site 0 keeps its tag through update_A0_norm although it is rescaled to unit norm: as in the reference)
    for i, (a, c) in enumerate(zip(dmp.A, dmp.Corder)):
        if i == 0 and nr_bulk:
            continue
        if c == "L":
            m = a.reshape(-1, a.shape[2])
            assert np.linalg.norm(m.conj().T @ m - np.eye(m.shape[1])) <= 1e-12 * m.shape[1]
        if c == "R":
            m = a.reshape(a.shape[0], -1)
            assert np.linalg.norm(m @ m.conj().T - np.eye(m.shape[0])) <= 1e-12 * m.shape[0]


def test_method_on_the_container():
    omp, maxD, nr_bulk, max_iter, err, outs, ocor, nr = load_case("bulk")
    dmp = DevMPS.from_sites([a.copy() for a in omp.A], Corder=list(omp.Corder))
    dmp.nr_mantissa, dmp.nr_exp = 1.0, 0
    dmp.reduceDiter(maxD, nr_bulk=True, max_iter=max_iter, err=err)
    assert np.linalg.norm(dense(dmp.A) - dense(outs)) <= 1e-10 * np.linalg.norm(dense(outs))
