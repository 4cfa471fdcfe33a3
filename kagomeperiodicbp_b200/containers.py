"""Configuration / result containers mirroring the reference's dataclasses for the block-BP path
(src/containers/belief_propagation.py:30-79, src/containers/contractions.py:10-52,
src/containers/global_config.py:28-48, src/unit_cell/definition.py:25-43).  Same field names and
defaults, so a reference user's configuration code carries over.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass, field
from typing import NamedTuple

import numpy as np

from .lattice import SIDE_OPPOSITE, side_ortho_cw
from .mps import MPS


class BlockSide:
    """block sides are plain strings in this package; this namespace gives the reference's spelling."""
    U, UR, UL, D, DL, DR = "U", "UR", "UL", "D", "DL", "DR"

    @staticmethod
    def all_in_counter_clockwise_order():
        from .lattice import BLOCK_SIDES_CCW
        return iter(BLOCK_SIDES_CCW)

    @staticmethod
    def opposite(side: str) -> str:
        return SIDE_OPPOSITE[side]


@dataclass
class BubbleConGlobalConfig:
    """(src/containers/contractions.py:17-35) incl. the rule that selects the truncation algorithm of bubblecon from the bond
    dimension: exact-SVD truncation up to D = 10, QR-only iterative compression (reduceDiter) above."""
    override_progress_bar: bool | None = None
    separate_exp: bool = True
    iterative_compression_max_ier: int = 200
    iterative_compression_error: float = 1e-8
    d_threshold_for_compression: int = 10

    def bubblecon_compression(self, D: int) -> dict:
        if D <= self.d_threshold_for_compression:
            return {"type": "SVD"}
        return {"type": "iter", "max-iter": self.iterative_compression_max_ier, "err": self.iterative_compression_error}


@dataclass
class MPSOrientation:
    open_towards: str
    ordered: str

    @staticmethod
    def standard(main_direction: str) -> "MPSOrientation":
        return MPSOrientation(open_towards=main_direction, ordered=side_ortho_cw(main_direction))


class Message(NamedTuple):
    mps: MPS
    orientation: MPSOrientation

    def copy(self) -> "Message":
        return Message(mps=self.mps.copy(full=True), orientation=self.orientation)


@dataclass
class BPConfig:
    init_msg: str = "RQ"                      # MessageModel: 'UQ' uniform quantum, 'RQ' random quantum
    max_iterations: int | None = 50
    trunc_dim: int = 9
    msg_diff_terminate: float = 1e-10
    msg_diff_good_enough: float = 1e-5
    msg_diff_squared: bool = True
    allowed_retries: int = 2
    times_to_deem_failure_when_diff_increases: int = 3
    parallel_msgs: bool = True                # the six sides always run concurrently on the device
    damping: float | None = None
    hermitize_msgs_when_finished: bool = True
    fix_msg_each_step: bool = True

    def __post_init__(self):
        if self.msg_diff_terminate > self.msg_diff_good_enough:
            raise ValueError("msg_diff_terminate must not exceed msg_diff_good_enough")

    def copy(self) -> "BPConfig":
        return copy.deepcopy(self)


@dataclass
class BPStats:
    iterations: int = -1
    attempts: int = 1
    final_error: float = -1.0
    success: bool = False
    final_config: BPConfig = field(default_factory=BPConfig)
    execution_time: float | None = None
    errors: list = field(default_factory=list)
    truncation_errors: list = field(default_factory=list)


@dataclass
class TNDimensions:
    virtual_dim: int = 3
    physical_dim: int = 2
    big_lattice_size: int = 3


@dataclass
class BubbleconContractionConfig:
    trunc_dim: int = 20
    parallel: bool = False
    progress_bar: bool = False


@dataclass
class Config:
    bp: BPConfig = field(default_factory=BPConfig)
    dims: TNDimensions = field(default_factory=TNDimensions)
    contraction: BubbleconContractionConfig = field(default_factory=BubbleconContractionConfig)

    @staticmethod
    def derive_from_dimensions(D: int) -> "Config":
        c = Config()
        c.dims = TNDimensions(virtual_dim=D)
        c.bp = BPConfig(trunc_dim=2 * D ** 2)
        c.contraction = BubbleconContractionConfig(trunc_dim=2 * D ** 2 + 10)
        return c

    @property
    def chi(self) -> int:
        return self.contraction.trunc_dim

    @chi.setter
    def chi(self, v):
        self.contraction.trunc_dim = int(v)

    @property
    def chi_bp(self) -> int:
        return self.bp.trunc_dim

    @chi_bp.setter
    def chi_bp(self, v):
        self.bp.trunc_dim = int(v)


class LatticeTensors:
    """one independent site tensor [d, D, D, D, D] per lattice site of the block, in block-index order (what the reference's
    KagomeTNArbitrary holds, src/tensor_networks/tensor_network.py:400-431); leg orders by site kind as in UnitCell."""

    def __init__(self, site_tensors):
        self.site_tensors = [np.asarray(t) for t in site_tensors]

    @property
    def A(self):                      # shape queries (d, D) of the callers
        return self.site_tensors[0]

    def tensors(self):
        return tuple(self.site_tensors)

    def copy(self) -> "LatticeTensors":
        return LatticeTensors([t.copy() for t in self.site_tensors])


@dataclass
class UnitCell:
    """three site tensors [d, D, D, D, D] of one upper triangle; leg orders
    A: [p, UL, DL, DR, UR]   B: [p, L, DL, R, UR]   C: [p, UL, L, DR, R]."""
    A: np.ndarray
    B: np.ndarray
    C: np.ndarray
    _file_name: str | None = None

    # ---- persistence (src/unit_cell/definition.py:119-137): same call shapes; see persistence.py for the file format
    def save(self, file_name: str | None = None, folder: str | None = None, asynchronous: bool = False) -> str:
        from . import persistence
        name = file_name or self._file_name or persistence.time_stamp()
        fullpath = os.path.join(folder or persistence.DEFAULT_FOLDER, name + ".dat")
        payload = persistence.unit_cell_payload(self, name)
        if asynchronous:
            persistence.saver().submit(payload, fullpath)
            return fullpath
        return persistence.write_payload(payload, fullpath)

    @staticmethod
    def load(file_name: str, folder: str | None = None, none_if_not_exist: bool = True):
        from . import persistence
        folder = folder or persistence.DEFAULT_FOLDER
        try:
            if file_name == "last":
                files = sorted(os.listdir(folder), key=lambda f: os.path.getmtime(os.path.join(folder, f)))
                file_name = files[-1]
            if not file_name.endswith(".dat"):
                file_name += ".dat"
            a, b, c = persistence.read_unit_cell_arrays(os.path.join(folder, file_name))
        except (FileNotFoundError, IndexError):
            if none_if_not_exist:
                return None
            raise
        return UnitCell(a, b, c, file_name[:-4])

    def set_filename(self, name: str):
        self._file_name = name

    def __getitem__(self, key):
        return {"A": self.A, "B": self.B, "C": self.C}[key]

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def tensors(self):
        return (self.A, self.B, self.C)

    def copy(self) -> "UnitCell":
        return UnitCell(self.A.copy(), self.B.copy(), self.C.copy(), self._file_name)

    @staticmethod
    def random(d: int, D: int, seed=None) -> "UnitCell":
        """distribution of the reference's UnitCell.random (src/unit_cell/definition.py:294-299), seedable."""
        rs = np.random.RandomState(seed)
        ts = []
        for _ in range(3):
            t = rs.rand(d, D, D, D, D) + 1j * rs.normal(size=(d, D, D, D, D))
            ts.append(t / np.linalg.norm(t))
        return UnitCell(*ts)
