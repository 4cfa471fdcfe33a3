"""Geometry of the hexagonal Kagome block (host side, integer/graph work only).

The block is a hexagon of ``3N^2-3N+1`` upper triangles; triangle ``t`` holds Kagome sites
``3t`` (A, top), ``3t+1`` (B, left), ``3t+2`` (C, right).  Everything the contraction needs is
derived once per N and cached: site positions, per-leg edge labels and angles, boundary edges
per block side (in the order the message MPS sites attach to them) and the row orderings used by
the swallow-order generator.

Conventions follow the reference so that its users (and its golden vectors) carry over:
  * directions and angles:            src/lattices/directions.py:183-197, 226-272
  * triangle rows / widths / indices: src/lattices/triangle.py:62-75, 176-207
  * triangle neighbours:              src/lattices/triangle.py:77-110
  * site leg orders (= UnitCell):     src/lattices/kagome.py:122-127, src/unit_cell/definition.py:37-39
  * inter-triangle bonds:             src/lattices/kagome.py:242-290
  * boundary tagging / edge naming:   src/lattices/kagome.py:130-143, 200-239, 340-346
  * boundary node / edge order:       src/lattices/_common.py:94-106, src/lattices/kagome.py:471-499
  * row orderings:                    src/lattices/triangle.py:846-902, src/lattices/kagome.py:146-163, 411-428
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass, field

# --------------------------------------------------------------------------------------------
# directions
# --------------------------------------------------------------------------------------------
LATTICE_DIRS = ("R", "UR", "UL", "L", "DL", "DR")
LATTICE_ANGLE = {d: k * math.pi / 3 for k, d in enumerate(LATTICE_DIRS)}
LATTICE_OPPOSITE = {"R": "L", "L": "R", "UR": "DL", "DL": "UR", "UL": "DR", "DR": "UL"}
LATTICE_CCW = ("DL", "DR", "R", "UR", "UL", "L")

BLOCK_SIDES_CCW = ("D", "DR", "UR", "U", "UL", "DL")
SIDE_ANGLE = {
    "U": math.pi / 2, "UR": math.pi / 2 - math.pi / 3, "UL": math.pi / 2 + math.pi / 3,
    "D": 3 * math.pi / 2, "DL": 3 * math.pi / 2 - math.pi / 3, "DR": 3 * math.pi / 2 + math.pi / 3,
}
SIDE_OPPOSITE = {"U": "D", "D": "U", "UR": "DL", "DL": "UR", "UL": "DR", "DR": "UL"}
# lattice direction orthogonal (counter-clockwise) to each block side
SIDE_ORTHO_CCW = {"D": "R", "U": "L", "DR": "UR", "DL": "DR", "UR": "UL", "UL": "DL"}
# the two lattice directions that leave the block through each side
SIDE_MATCHING = {
    "D": ("DL", "DR"), "DR": ("DR", "R"), "UR": ("R", "UR"),
    "U": ("UR", "UL"), "UL": ("UL", "L"), "DL": ("L", "DL"),
}
_SORT_VEC = {  # integer stand-ins for unit vectors, used only for ordering
    "R": (1, 0), "L": (-1, 0), "UL": (-1, 1), "UR": (1, 1), "DL": (-1, -1), "DR": (1, -1),
}
_SORT_VEC_SIDE = {"U": (0, 1), "D": (0, -1), "UR": (1, 1), "UL": (-1, 1), "DL": (-1, -1), "DR": (1, -1)}

SITE_KINDS = ("up", "left", "right")
SITE_DIRS = {
    "up": ("UL", "DL", "DR", "UR"),
    "left": ("L", "DL", "R", "UR"),
    "right": ("UL", "L", "DR", "R"),
}
_SITE_DELTA = {"up": (0, 1), "left": (-1, -1), "right": (1, -1)}


def side_next_ccw(side: str) -> str:
    i = BLOCK_SIDES_CCW.index(side)
    return BLOCK_SIDES_CCW[(i + 1) % 6]


def side_next_cw(side: str) -> str:
    i = BLOCK_SIDES_CCW.index(side)
    return BLOCK_SIDES_CCW[(i - 1) % 6]


def side_ortho_cw(side: str) -> str:
    return LATTICE_OPPOSITE[SIDE_ORTHO_CCW[side]]


def _snap(x: float):
    r = int(round(x))
    return r if abs(r - x) < 1e-6 else x


def unit_vector(angle: float):
    return (_snap(math.cos(angle)), _snap(math.sin(angle)))


# --------------------------------------------------------------------------------------------
# triangular lattice of upper triangles
# --------------------------------------------------------------------------------------------
def num_triangles(N: int) -> int:
    return 3 * N * N - 3 * N + 1


def num_rows(N: int) -> int:
    return 2 * N - 1


def row_width(i: int, N: int) -> int:
    if i < 0 or i > 2 * N - 2:
        return 0
    return N + i if i < N else 3 * N - i - 2


def triangle_index(i: int, j: int, N: int) -> int:
    return sum(row_width(r, N) for r in range(i)) + j


def triangle_pos(i: int, j: int, N: int):
    return (N - row_width(i, N) + 2 * j, N - i)


def _triangle_neighbor(i: int, j: int, d: str, N: int):
    mid = num_rows(N) // 2
    if d == "L":
        i2, j2 = i, j - 1
    elif d == "R":
        i2, j2 = i, j + 1
    elif d == "UR":
        i2, j2 = (i - 1, j) if i <= mid else (i - 1, j + 1)
    elif d == "UL":
        i2, j2 = (i - 1, j - 1) if i <= mid else (i - 1, j)
    elif d == "DL":
        i2, j2 = (i + 1, j) if i < mid else (i + 1, j - 1)
    else:  # DR
        i2, j2 = (i + 1, j + 1) if i < mid else (i + 1, j)
    if i2 < 0 or i2 >= num_rows(N) or j2 < 0 or j2 >= row_width(i2, N):
        return None
    return i2, j2


def _triangle_sides(i: int, j: int, N: int):
    mid = num_rows(N) // 2
    w = row_width(i, N)
    out = []
    if i == 0:
        out.append("U")
    if i == num_rows(N) - 1:
        out.append("D")
    if j == 0:
        if i <= mid:
            out.append("UL")
        if i >= mid:
            out.append("DL")
    if j == w - 1:
        if i <= mid:
            out.append("UR")
        if i >= mid:
            out.append("DR")
    return out


_SIDE_TOUCH = {"U": ("up",), "DL": ("left",), "DR": ("right",), "D": ("left", "right"),
               "UR": ("up", "right"), "UL": ("up", "left")}
# which (site kind, site kind) a bond towards a neighbouring triangle joins
_BOND = {"L": ("left", "right"), "DL": ("left", "up"), "DR": ("right", "up"),
         "R": ("right", "left"), "UR": ("up", "left"), "UL": ("up", "right")}


@dataclass
class Site:
    index: int
    kind: str                    # 'up' | 'left' | 'right'
    pos: tuple
    dirs: tuple                  # lattice direction of each virtual leg (UnitCell leg order)
    edges: list                  # edge label of each virtual leg
    boundaries: set = field(default_factory=set)

    @property
    def angles(self):
        return [LATTICE_ANGLE[d] for d in self.dirs]

    def edge_in(self, d: str):
        return self.edges[self.dirs.index(d)]


class KagomeBlock:
    """All static geometry of the size-N block."""

    def __init__(self, N: int):
        assert N >= 2
        self.N = N
        self.L = 2 * N - 1                       # message length
        self.n_tri = num_triangles(N)
        self.n_sites = 3 * self.n_tri
        self.coords = [(i, j) for i in range(num_rows(N)) for j in range(row_width(i, N))]
        self.sites: list[Site] = []
        for t, (i, j) in enumerate(self.coords):
            x0, y0 = triangle_pos(i, j, N)
            for k, kind in enumerate(SITE_KINDS):
                dx, dy = _SITE_DELTA[kind]
                self.sites.append(Site(3 * t + k, kind, (2 * x0 + dx + 3, 4 * y0 + dy + 1),
                                       SITE_DIRS[kind], [None] * 4))
        self._wire_bulk()
        self._tag_and_name_boundaries()
        self.boundary_edges = {s: self._sorted_boundary_edges(s) for s in BLOCK_SIDES_CCW}
        # open edge label -> (site index, leg index)
        self.open_edge_owner = {}
        for s in self.sites:
            for l, e in enumerate(s.edges):
                if self._is_open(e):
                    self.open_edge_owner[e] = (s.index, l)
        self.center_triangle = triangle_index(num_rows(N) // 2, num_rows(N) // 2, N)

    # ---------------------------------------------------------------------------------------
    def site(self, t: int, kind: str) -> Site:
        return self.sites[3 * t + SITE_KINDS.index(kind)]

    @staticmethod
    def _name(a: int, b: int) -> str:
        return f"{min(a, b)}-{max(a, b)}"

    def _is_open(self, e: str) -> bool:
        return e.split("-")[0] in SIDE_ANGLE

    def _set(self, s: Site, d: str, name: str):
        s.edges[s.dirs.index(d)] = name

    def _wire_bulk(self):
        N = self.N
        for t in range(self.n_tri):
            up, left, right = (self.site(t, k) for k in SITE_KINDS)
            self._set(up, "DL", self._name(up.index, left.index)); self._set(left, "UR", self._name(up.index, left.index))
            self._set(up, "DR", self._name(up.index, right.index)); self._set(right, "UL", self._name(up.index, right.index))
            self._set(left, "R", self._name(left.index, right.index)); self._set(right, "L", self._name(left.index, right.index))
        for t, (i, j) in enumerate(self.coords):
            for d in LATTICE_CCW:
                nb = _triangle_neighbor(i, j, d, N)
                if nb is None:
                    continue
                t2 = triangle_index(nb[0], nb[1], N)
                k1, k2 = _BOND[d]
                a, b = self.site(t, k1), self.site(t2, k2)
                name = self._name(a.index, b.index)
                self._set(a, d, name)
                self._set(b, LATTICE_OPPOSITE[d], name)

    def sorted_boundary_sites(self, side: str) -> list[Site]:
        nodes = [s for s in self.sites if side in s.boundaries]
        key = {"U": lambda s: -s.pos[0], "UR": lambda s: s.pos[1], "DR": lambda s: s.pos[1],
               "UL": lambda s: -s.pos[1], "DL": lambda s: -s.pos[1], "D": lambda s: s.pos[0]}[side]
        return sorted(nodes, key=key)

    def _tag_and_name_boundaries(self):
        N = self.N
        for t, (i, j) in enumerate(self.coords):
            for side in _triangle_sides(i, j, N):
                for kind in _SIDE_TOUCH[side]:
                    self.site(t, kind).boundaries.add(side)
        # later sides overwrite the shared corner legs of earlier ones, exactly in this order
        for side in BLOCK_SIDES_CCW:
            for k, s in enumerate(self.sorted_boundary_sites(side)):
                nm = lambda q: f"{side}-{q}"
                if side == "D":
                    self._set(s, "DL" if s.kind == "left" else "DR", nm(k))
                elif side == "DR":
                    self._set(s, "DR", nm(2 * k)); self._set(s, "R", nm(2 * k + 1))
                elif side == "UR":
                    self._set(s, "R" if s.kind == "right" else "UR", nm(k))
                elif side == "U":
                    self._set(s, "UR", nm(2 * k)); self._set(s, "UL", nm(2 * k + 1))
                elif side == "UL":
                    self._set(s, "UL" if s.kind == "up" else "L", nm(k))
                else:  # DL
                    self._set(s, "L", nm(2 * k)); self._set(s, "DL", nm(2 * k + 1))
        self._set(self.sorted_boundary_sites("D")[0], "DL", "D-0")

    def _sorted_boundary_edges(self, side: str) -> list[str]:
        nodes = self.sorted_boundary_sites(side)
        dirs = SIDE_MATCHING[side]
        omit_last_edge = len(nodes) == self.N
        out = []
        for ni, s in enumerate(nodes):
            last_node = ni == len(nodes) - 1
            if (not omit_last_edge) and last_node:
                break
            for di, d in enumerate(dirs):
                if omit_last_edge and last_node and di == len(dirs) - 1:
                    break
                if d in s.dirs:
                    out.append(s.edge_in(d))
        assert len(out) == self.L
        return out

    # ---------------------------------------------------------------------------------------
    # row orderings for the swallow-order generator
    # ---------------------------------------------------------------------------------------
    def _sort_coords(self, items, vec):
        def key(ij):
            x, y = triangle_pos(ij[0], ij[1], self.N)
            return x * vec[0] + y * vec[1]
        return sorted(items, key=key)

    @functools.lru_cache(maxsize=None)
    def triangle_rows(self, major: str, minor: str):
        pool = self._sort_coords(self.coords, _SORT_VEC_SIDE[SIDE_OPPOSITE[major]])
        rows = []
        for i in range(num_rows(self.N)):
            row = [pool.pop() for _ in range(row_width(i, self.N))]
            row = self._sort_coords(row, _SORT_VEC[minor])
            rows.append([triangle_index(a, b, self.N) for a, b in row])
        return rows

    @staticmethod
    def triangle_vertex_groups(major: str, minor: str):
        if major == "U":
            return [["left", "right"], ["up"]] if minor == "R" else [["right", "left"], ["up"]]
        if major == "UR":
            return [["left"], ["up", "right"]] if minor == "DR" else [["left"], ["right", "up"]]
        if major == "UL":
            return [["right"], ["left", "up"]] if minor == "UR" else [["right"], ["up", "left"]]
        flipped = {"D": "U", "DL": "UR", "DR": "UL"}[major]
        return list(reversed(KagomeBlock.triangle_vertex_groups(flipped, minor)))

    @functools.lru_cache(maxsize=None)
    def site_rows(self, major: str, minor: str):
        groups = self.triangle_vertex_groups(major, minor)
        rows = []
        for trow in self.triangle_rows(major, minor):
            for g in groups:
                rows.append([self.site(t, k).index for t in trow for k in g])
        return rows

    # ---------------------------------------------------------------------------------------
    def message_indices(self, side: str) -> list[int]:
        k = BLOCK_SIDES_CCW.index(side)
        return [self.n_sites + self.L * k + q for q in range(self.L)]

    def neighbor_site(self, s: Site, d: str):
        """the lattice site on the other end of leg ``d`` (None for an open leg)."""
        e = s.edge_in(d)
        if self._is_open(e):
            return None
        a, b = (int(x) for x in e.split("-"))
        return self.sites[b if a == s.index else a]

    def core_indices(self) -> set[int]:
        """centre triangle + its nearest neighbours (src/tensor_networks/tensor_network.py:907-952)."""
        center = [self.site(self.center_triangle, k) for k in SITE_KINDS]
        out = {s.index for s in center}
        for s in center:
            for d in s.dirs:
                nb = self.neighbor_site(s, d)
                if nb is not None:
                    out.add(nb.index)
        return out


@functools.lru_cache(maxsize=None)
def get_block(N: int) -> KagomeBlock:
    return KagomeBlock(N)
