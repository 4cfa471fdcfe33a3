"""GPU tier: the device ITE algebra (every tensordot / QR / SVD / eigh a device program through the C ABI) against the
reference's outputs on the reference's own edge environments, and against the oracle."""
import numpy as np
import pytest

from helpers import golden

pytestmark = pytest.mark.gpu

EDGES = ("AB", "AC", "BA", "BC", "CA", "CB")


@pytest.fixture(scope="module", params=["per-call", "resident"])
def B(request):
    from kagomeperiodicbp_b200.linalg import DeviceBackend, ResidentBackend
    return DeviceBackend() if request.param == "per-call" else ResidentBackend("ite-test-resident", arena_elems=1 << 26)


def _inputs(g, key):
    return g[f"in_ti_{key}"], g[f"in_tj_{key}"], [g[f"in_env{k}_{key}"] for k in range(6)]


def pair_of(ti, tj):
    ti, tj = ti / np.linalg.norm(ti), tj / np.linalg.norm(tj)
    return np.tensordot(ti, tj, axes=([1], [1]))


def pair_rel_diff(a, b):
    ph = np.vdot(a, b)
    ph = ph / abs(ph)
    return np.linalg.norm(a * ph - b) / np.linalg.norm(b)


def test_backend_primitives(B):
    rng = np.random.default_rng(5)
    a = rng.normal(size=(7, 5, 6)) + 1j * rng.normal(size=(7, 5, 6))
    b = rng.normal(size=(6, 5, 3)) + 1j * rng.normal(size=(6, 5, 3))
    assert np.allclose(np.asarray(B.tensordot(a, b, ([1, 2], [1, 0]), conj_b=True)), np.tensordot(a, np.conj(b), axes=([1, 2], [1, 0])), atol=1e-13)
    assert np.allclose(np.asarray(B.tensordot(np.eye(2), a[:2, :2, 0], 0)), np.tensordot(np.eye(2), a[:2, :2, 0], 0), atol=1e-14)
    assert abs(B.norm(a) - np.linalg.norm(a)) < 1e-12
    assert np.allclose(np.asarray(B.scale(a, 0.3 - 2j)), a * (0.3 - 2j), atol=1e-13)
    h = rng.normal(size=(36, 36)) + 1j * rng.normal(size=(36, 36))
    h = h + h.conj().T
    h[:, 5] = 0
    h[5, :] = 0                       # a zero eigenvalue and indefinite spectrum
    w, u = B.eigh(h)
    assert np.allclose(w, np.linalg.eigvalsh(h), atol=1e-12 * np.linalg.norm(h))
    assert np.linalg.norm(u @ np.diag(w) @ u.conj().T - h) < 1e-12 * np.linalg.norm(h)
    assert np.linalg.norm(u.conj().T @ u - np.eye(36)) < 1e-12
    m = rng.normal(size=(12, 8)) + 1j * rng.normal(size=(12, 8))
    u, s, vh = B.svd(m)
    u, vh = np.asarray(u), np.asarray(vh)
    assert np.allclose(s, np.linalg.svd(m, compute_uv=False), atol=1e-13)
    assert np.linalg.norm((u * s) @ vh - m) < 1e-13 * np.linalg.norm(m)
    q, r = B.qr(m)
    assert np.linalg.norm(np.asarray(q) @ np.asarray(r) - m) < 1e-13 * np.linalg.norm(m)
    assert np.allclose(np.asarray(B.hermitize(m[:8])), 0.5 * (m[:8] + m[:8].conj().T), atol=1e-14)
    x = B.transpose(B.tensordot(a, b, ([2], [0])), (3, 1, 0, 2))
    assert np.allclose(np.asarray(B.reshape(x, (3 * 5, -1))), np.tensordot(a, b, axes=([2], [0])).transpose(3, 1, 0, 2).reshape(15, -1), atol=1e-13)


@pytest.mark.parametrize("mode", "ABC")
def test_device_rho_energy_gate_match_reference(B, mode):
    from kagomeperiodicbp_b200 import ite
    g = golden("ite_D2_N2.npz")
    for e in EDGES:
        key = f"{mode}_{e}"
        ti, tj, env = _inputs(g, key)
        rho = ite.rho_ij(B, ti, tj, env)
        assert np.max(np.abs(rho - g[f"rdm_{key}"])) < 1e-12                       # north star: energies to 1e-8
        energy = np.dot(rho.flatten(), g["h"].flatten())
        assert abs(energy.real - g[f"energy_{key}"][0]) < 1e-10
        tin, tjn, w = ite.apply_2local_gate(B, g["g"], 2, ti, tj, env)
        assert np.allclose(w, g[f"eig_{key}"], rtol=0, atol=1e-11 * np.max(np.abs(w)))
        assert pair_rel_diff(pair_of(tin, tjn), g[f"pair_{key}"]) < 1e-8, key         # truncation / ALS result to 1e-8


def test_device_rho_ij_and_reduced_rdms_against_independent_einsum(B):
    """D = 4 sized edge (ring bonds of mixed size as the Mode -> Edge reduction makes them, non-Hermitian environment): the device
    rho_ij and the RDMs the update loop takes from the reduced environment against the independent einsum of oracle/rho_np.py."""
    from kagomeperiodicbp_b200 import ite
    from oracle.rho_np import rho_ij_einsum
    rng = np.random.default_rng(21)
    rnd = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)
    d, D, bonds = 2, 4, [42, 96, 42, 16, 42, 96]
    Ti, Tj = rnd(d, D, D, D, D), rnd(d, D, D, D, D)
    env = [rnd(bonds[k], D, D, bonds[(k + 1) % 6]) / 8 for k in range(6)]
    ref0 = rho_ij_einsum(Ti, Tj, env)
    assert np.abs(np.asarray(ite.rho_ij(B, Ti, Tj, env)) - ref0).max() < 1e-11 * np.abs(ref0).max()
    aux = {}
    tin, tjn, _ = ite.apply_2local_gate(B, ite.g_from_exp_h(ite.heisenberg_afm(), 1e-2), D, Ti, Tj, env, aux=aux)
    assert np.abs(aux["rho_before"] - ref0).max() < 1e-10 * np.abs(ref0).max()
    ref1 = rho_ij_einsum(np.asarray(tin), np.asarray(tjn), env)
    assert np.abs(aux["rho_after"] - ref1).max() < 1e-9 * np.abs(ref1).max()
