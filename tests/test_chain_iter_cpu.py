"""bubblecon with iterative (QR-only, reduceDiter) compression: oracle vs the reference's output on the same call
(tests/golden/chain_iter_D2_N2.npz, tools/make_golden_chain_iter.py), and the product's stepwise path executed on the
numpy interpreter of the op stream (host logic: per-swallow programs, canonical tags and scale carried between steps)."""
import json

import numpy as np
import pytest

from helpers import golden

SIDES2 = ("D", "UL")


def load(side):
    g = golden("chain_iter_D2_N2.npz")
    meta = json.loads(str(g[f"{side}/meta"]))
    n = len(meta["edges"])
    T = [g[f"{side}/T{i}"] for i in range(n)]
    nout = len([k for k in g.files if k.startswith(f"{side}/out")])
    outs = [g[f"{side}/out{k}"] for k in range(nout)]
    cor = [c if c != "" else None for c in g[f"{side}/corder"].tolist()]
    return T, meta, outs, cor, g[f"{side}/nr"].tolist()


@pytest.mark.parametrize("side", SIDES2)
def test_oracle_iterative_chain_matches_reference(side):
    from oracle.bubblecon_np import bubblecon
    T, meta, outs, cor, nr = load(side)
    mp = bubblecon(T, meta["edges"], meta["angles"], meta["bubble_angle"], meta["order"], D_trunc=meta["D_trunc"],
                   ket_tensors=meta["kets"], compression=meta["compression"])
    assert [a.shape for a in mp.A] == [o.shape for o in outs] and list(mp.Corder) == cor
    for a, o in zip(mp.A, outs):
        assert np.linalg.norm(a - o) <= 1e-9 * max(1.0, np.linalg.norm(o))
    assert abs(mp.nr_mantissa - nr[0]) <= 1e-9 * abs(nr[0]) and mp.nr_exp == int(nr[1])


@pytest.mark.parametrize("side", SIDES2)
def test_stepwise_device_path_on_the_interpreter(side, vm_engines, monkeypatch):
    """the product's bubblecon(compression={'type': 'iter'}) with every engine routed to the numpy interpreter of the op stream"""
    from np_vm import NumpyEngine
    import kagomeperiodicbp_b200.linalg as linalg
    from kagomeperiodicbp_b200 import bubblecon as bc
    from kagomeperiodicbp_b200 import reduce_iter
    engines = {}
    monkeypatch.setattr(linalg, "get_engine", lambda key="default", device=0: engines.setdefault(key, NumpyEngine()))
    monkeypatch.setattr(reduce_iter, "_backend", linalg.ResidentBackend("vm-reduce", arena_elems=1 << 22))
    T, meta, outs, cor, nr = load(side)
    mp = bc.bubblecon(T, meta["edges"], meta["angles"], meta["bubble_angle"], meta["order"], D_trunc=meta["D_trunc"],
                      ket_tensors=meta["kets"], compression=meta["compression"])
    assert [a.shape for a in mp.A] == [o.shape for o in outs] and list(mp.Corder) == cor
    for a, o in zip(mp.A, outs):
        assert np.linalg.norm(a - o) <= 1e-9 * max(1.0, np.linalg.norm(o))      # interpreter QR = LAPACK: same gauge as the reference
    assert abs(mp.nr_mantissa - nr[0]) <= 1e-9 * abs(nr[0]) and mp.nr_exp == int(nr[1])
    assert len(bc.last_stats["reduce_iter_rounds"]) == len(meta["order"])
