"""Stand the UNMODIFIED reference up in THIS container (never on the GPU box).

Used only by the golden-vector generators under ``tools/`` and by ad-hoc
validation of ``oracle/`` against the real reference.  Nothing in the product,
``tests/``, ``bench.py`` or ``__graft_entry__`` imports this module.

Recipe:
  * scratch copy of ``src/ scripts/ configuration.json`` under ``/tmp/kbp_ref``
    (the reference writes ``data/`` and ``logs/`` next to ``src/``,
    ``src/project_paths.py:7-11``), with ``keep_logs`` switched off in the
    scratch copy of ``configuration.json``;
  * a stub ``quimb.linalg.rand_linalg.rsvd`` that is ``numpy.linalg.svd`` --
    i.e. the reference's own deterministic ``"svd"`` branch
    (``src/libs/bmpslib.py:2874-2875``) -- because quimb is not installed;
  * dummy ``matplotlib`` / ``mpl_toolkits`` / ``moviepy`` packages whose every
    attribute is a class (annotations are evaluated at import time,
    ``src/utils/visuals.py:123``).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import json
import os
import shutil
import sys
import types
from pathlib import Path

REFERENCE = Path("/root/reference")
SCRATCH = Path(os.environ.get("KBP_REF_SCRATCH", "/tmp/kbp_ref"))

_DUMMY_ROOTS = ("matplotlib", "mpl_toolkits", "moviepy")


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy(name)

    def __or__(cls, other):
        return cls

    def __ror__(cls, other):
        return cls

    def __call__(cls, *a, **k):
        return cls


def _make_dummy(name: str):
    return _DummyMeta(name, (), {})


class _DummyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy(name)


class _DummyFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _DUMMY_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _DummyModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        return None


def _install_quimb_stub():
    import numpy as np

    def rsvd(m, eps_or_k, *a, **k):
        return np.linalg.svd(m, full_matrices=False)

    quimb = types.ModuleType("quimb")
    quimb.__path__ = []
    linalg = types.ModuleType("quimb.linalg")
    linalg.__path__ = []
    rl = types.ModuleType("quimb.linalg.rand_linalg")
    rl.rsvd = rsvd
    quimb.linalg = linalg
    linalg.rand_linalg = rl
    sys.modules["quimb"] = quimb
    sys.modules["quimb.linalg"] = linalg
    sys.modules["quimb.linalg.rand_linalg"] = rl


def setup(fresh: bool = False) -> Path:
    """Create the scratch copy, install the shims, put ``src`` on sys.path."""
    if not REFERENCE.exists():
        raise RuntimeError("/root/reference is not mounted: the reference can only be run in the build container")
    if fresh and SCRATCH.exists():
        shutil.rmtree(SCRATCH)
    if not (SCRATCH / "src").exists():
        SCRATCH.mkdir(parents=True, exist_ok=True)
        shutil.copytree(REFERENCE / "src", SCRATCH / "src")
        shutil.copytree(REFERENCE / "scripts", SCRATCH / "scripts")
        cfg = json.loads((REFERENCE / "configuration.json").read_text())
        cfg["keep_logs"] = False
        (SCRATCH / "configuration.json").write_text(json.dumps(cfg, indent=4))
        (SCRATCH / "data").mkdir(exist_ok=True)
        (SCRATCH / "logs").mkdir(exist_ok=True)
    if not any(isinstance(f, _DummyFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _DummyFinder())
    _install_quimb_stub()
    for p in (str(SCRATCH / "scripts"), str(SCRATCH), str(SCRATCH / "src")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    return SCRATCH


def seeded_unit_cell(D: int, seed: int, d: int = 2):
    """The distribution of ``UnitCell.random`` (``src/unit_cell/definition.py:294-299``)
    but from a seeded generator: real part U(0,1), imaginary part N(0,1), Frobenius-normalised."""
    import numpy as np
    from unit_cell import UnitCell  # type: ignore

    rs = np.random.RandomState(seed)
    ts = []
    for _ in range(3):
        t = rs.rand(d, D, D, D, D) + 1j * rs.normal(size=(d, D, D, D, D))
        t = t / np.linalg.norm(t)
        ts.append(t)
    return UnitCell(A=ts[0], B=ts[1], C=ts[2])


def quiet_config(D: int, N: int):
    from containers import Config  # type: ignore

    config = Config.derive_from_dimensions(D)
    config.dims.big_lattice_size = N
    config.visuals.progress_bars = "all_disabled" if hasattr(config.visuals, "progress_bars") else None
    try:
        config.bp.visuals.set_all_progress_bars(False)
    except Exception:
        pass
    try:
        config.contraction.progress_bar = False
    except Exception:
        pass
    return config
