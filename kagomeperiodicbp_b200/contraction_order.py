"""Swallow orders for the boundary-MPS contraction of the Kagome block (host side, integers only).

``kagome_order(N, side, depth)`` reproduces the order the reference derives in
src/algo/contraction_order.py:535-601 (``derive_kagome_tn_contraction_order``): first the message
opposite to ``side``, then the lattice rows towards ``side`` in boustrophedon order (direction flips
every two rows, :168-203), each row preceded/followed by the lateral message sites that touch it
(:355-412), with special handling where one lateral message ends and the next begins ('Break').
``depth='ToCore'`` stops at the base of the centre triangle (:442-483).

Truncation happens after every swallow, so results only match the reference to 1e-10 if the order
is identical; ``tests/test_geometry_golden.py`` pins it against orders dumped from the reference.
"""
from __future__ import annotations

import functools

from .lattice import (BLOCK_SIDES_CCW, SIDE_OPPOSITE, KagomeBlock, get_block, side_next_ccw,
                      side_next_cw, side_ortho_cw)

TO_MESSAGE, TO_CORE, FULL = "ToMessage", "ToCore", "Full"


class _EdgeCursor:
    """walks the lateral boundary edges of one flank (left or right) of the contraction front."""

    def __init__(self, edges):
        self.edges = edges
        self.k = 0
        self.done = False

    def cur(self):
        if self.done:
            return "End"
        return self.edges[self.k]

    def advance(self):
        self.k += 1
        if self.k >= len(self.edges):
            self.done = True
            return None
        return self.edges[self.k]


class _Flip:
    """row direction tracker: starts reversed, flips after every second row (:168-203)."""

    def __init__(self):
        self.state, self.count = True, 0

    def take(self):
        r = self.state
        self.count += 1
        if self.count > 1:
            self.state, self.count = (not self.state), 0
        return r

    def force(self, v: bool):
        self.state, self.count = v, 0


def _flank_edges(blk: KagomeBlock, side: str):
    be = blk.boundary_edges
    r_last = side_next_cw(side)
    r_first = side_next_cw(r_last)
    l_last = side_next_ccw(side)
    l_first = side_next_ccw(l_last)
    right = be[r_first] + ["Break"] + be[r_last]
    left = list(reversed(be[l_last] + ["Break"] + be[l_first]))
    return left, right


def _edge_to_message_node(blk: KagomeBlock):
    m = {}
    for s in BLOCK_SIDES_CCW:
        for k, e in enumerate(blk.boundary_edges[s]):
            m[e] = blk.message_indices(s)[k]
    return m


def _collect(blk, node, cursor, e2m):
    out = []
    e = cursor.cur()
    while e in node.edges:
        out.append(e2m[e])
        e = cursor.advance()
    return e, out


@functools.lru_cache(maxsize=None)
def kagome_order(N: int, side: str, depth: str = TO_MESSAGE) -> tuple:
    blk = get_block(N)
    e2m = _edge_to_message_node(blk)
    minor = side_ortho_cw(side)
    rows = blk.site_rows(side, minor)
    left_edges, right_edges = _flank_edges(blk, side)
    cur = {"L": _EdgeCursor(left_edges), "R": _EdgeCursor(right_edges)}
    flip = _Flip()

    core = blk.core_indices() if depth == TO_CORE else set()
    terminal = set()
    if depth == TO_CORE:
        groups = blk.triangle_vertex_groups(side, minor)
        base = groups[0] if len(groups[0]) == 2 else groups[1]
        terminal = {blk.site(blk.center_triangle, k).index for k in base}
    stop = False
    seen_break_side = {"L": False, "R": False}

    order = list(blk.message_indices(SIDE_OPPOSITE[side]))
    for row in rows:
        rev_now = flip.state
        ends = {"L": blk.sites[row[0]], "R": blk.sites[row[-1]]}
        seq = ("R", "L") if flip.take() else ("L", "R")          # (first, last) flank
        nb = {}
        brk = {}
        for fl in seq:
            last_e, got = _collect(blk, ends[fl], cur[fl], e2m)
            nb[fl] = got
            brk[fl] = last_e == "Break"
            if brk[fl]:
                cur[fl].advance()
        annex = []
        if brk[seq[0]] or brk[seq[1]]:
            before = {fl: len(nb[fl]) for fl in seq}
            after = dict(before)
            for fl in seq:
                _, got = _collect(blk, ends[fl], cur[fl], e2m)
                after[fl] += len(got)
                annex += got
            # which flank continues first on the next row (:306-336, :270-303)
            if brk[seq[0]] and not brk[seq[1]]:
                nxt = seq[1]
            elif brk[seq[1]] and not brk[seq[0]]:
                nxt = seq[0]
            else:
                nxt = None
                for fl in seq:
                    if before[fl] == 0 or after[fl] == 0:
                        raise ValueError("unexpected lateral-message layout")
                    if before[fl] == 1 and after[fl] == 2:
                        nxt = fl
                        break
            if nxt == "L":
                flip.force(False)
            elif nxt == "R":
                flip.force(True)

        first, last = nb[seq[0]], nb[seq[1]]
        body = list(reversed(row)) if rev_now else list(row)
        if depth != TO_CORE:
            order += first + body + last + annex
            continue
        if stop:
            continue
        core_here = set(row) & core
        if not core_here:
            order += first + body + last + annex
            continue
        # split the row around the core sites
        left_part, right_part, seen = [], [], False
        for i in row:
            if i in core_here:
                seen = True
            elif seen:
                right_part.append(i)
            else:
                left_part.append(i)
        brk_side = {seq[0]: brk[seq[0]], seq[1]: brk[seq[1]]}
        if core_here.isdisjoint(terminal):
            parts = (right_part, left_part) if rev_now else (left_part, right_part)
            for s_ in ("L", "R"):
                if brk_side[s_]:
                    seen_break_side[s_] = True
            order += first + parts[0] + parts[1] + last + annex
            continue
        stop = True
        nb_side = {seq[0]: first, seq[1]: last}
        if (not seen_break_side["R"]) and brk_side["R"]:
            order += nb_side["L"] + left_part + nb_side["R"]
        else:
            order += nb_side["L"] + left_part

    if depth == FULL:
        last_rev = (not flip.state) if flip.count - 1 < 0 else flip.state
        final_rev = not last_rev
        last_msg = list(blk.message_indices(side))
        if not final_rev:
            last_msg.reverse()
        order += last_msg
    return tuple(order)
