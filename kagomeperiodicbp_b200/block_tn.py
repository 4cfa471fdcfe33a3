"""Assemble the planar tensor network of one Kagome block + its six incoming MPS messages in the
form the boundary-MPS contractor takes: ``(tensors, edges_list, angles_list, kets)``.

Mirrors (without sharing code) what the reference's containers produce:
  * lattice sites = repeated unit cell, A/B/C by ``index % 3``   (src/tensor_networks/tensor_network.py:907-931)
  * message sites attached to the sorted boundary edges of their side, legs
    [prev, to-lattice, next], angles from the message orientation  (src/tensor_networks/tensor_network.py:815-890)
  * node numbering: lattice first, then messages in counter-clockwise side order (:893-903)
  * the dim-1 'fake_leg' that glues the end of the first message to the start of the next one so
    the MPS segment stays contiguous                             (src/algo/contract_tensor_network.py:99-143)
"""
from __future__ import annotations

import math

import numpy as np

from .lattice import (BLOCK_SIDES_CCW, LATTICE_ANGLE, LATTICE_OPPOSITE, SIDE_ANGLE, SIDE_OPPOSITE,
                      KagomeBlock, get_block, side_next_ccw, side_ortho_cw, unit_vector)


def message_order_direction(side: str) -> str:
    """lattice direction along which the message sitting on ``side`` is ordered
    (= MPSOrientation.standard(side.opposite()).ordered, src/containers/contractions.py:43-52)."""
    return side_ortho_cw(SIDE_OPPOSITE[side])


def message_node_positions(blk: KagomeBlock, side: str):
    ux, uy = unit_vector(SIDE_ANGLE[side])
    out = []
    for e in blk.boundary_edges[side]:
        si, leg = blk.open_edge_owner[e]
        s = blk.sites[si]
        dx, dy = unit_vector(LATTICE_ANGLE[s.dirs[leg]])
        out.append((s.pos[0] + ux + dx, s.pos[1] + uy + dy))
    return out


def assemble(N: int, cell, messages: dict | None):
    """``cell`` = (A, B, C) arrays [d, D, D, D, D] (repeated unit cell) or one tensor per lattice site in block-index order (the
    non-repeated block, KagomeTNArbitrary: src/tensor_networks/tensor_network.py:400-431); ``messages`` = {side: [site arrays
    [DL, D^2, DR]]}.  Returns lists indexed by node: tensors, edges, angles, kets, positions."""
    blk = get_block(N)
    tensors, edges, angles, kets, pos = [], [], [], [], []
    for s in blk.sites:
        tensors.append(cell[s.index % 3] if len(cell) == 3 else cell[s.index])
        edges.append(list(s.edges))
        angles.append(list(s.angles))
        kets.append(True)
        pos.append(s.pos)
    if messages:
        for side in BLOCK_SIDES_CCW:
            if side not in messages:
                continue
            sites = messages[side]
            L = len(sites)
            assert L == blk.L
            o = message_order_direction(side)
            a_fwd, a_back = LATTICE_ANGLE[o], LATTICE_ANGLE[LATTICE_OPPOSITE[o]]
            a_in = SIDE_ANGLE[SIDE_OPPOSITE[side]]
            mpos = message_node_positions(blk, side)
            for k, t in enumerate(sites):
                e_lat = blk.boundary_edges[side][k]
                if k == 0:
                    tensors.append(t.reshape(t.shape[1], t.shape[2]))
                    edges.append([e_lat, f"M-{side}-0"])
                    angles.append([a_in, a_fwd])
                elif k == L - 1:
                    tensors.append(t.reshape(t.shape[0], t.shape[1]))
                    edges.append([f"M-{side}-{k - 1}", e_lat])
                    angles.append([a_back, a_in])
                else:
                    tensors.append(t)
                    edges.append([f"M-{side}-{k - 1}", e_lat, f"M-{side}-{k}"])
                    angles.append([a_back, a_in, a_fwd])
                kets.append(False)
                pos.append(mpos[k])
    return tensors, edges, angles, kets, pos


def connect_corner(N: int, tensors, edges, angles, pos, outgoing: str):
    """add the dim-1 leg between the last site of the first swallowed message (side opposite to
    ``outgoing``) and the first site of the next message counter-clockwise."""
    blk = get_block(N)
    first = SIDE_OPPOSITE[outgoing]
    i_a = blk.message_indices(first)[-1]
    i_b = blk.message_indices(side_next_ccw(first))[0]
    tensors, edges, angles = list(tensors), [list(e) for e in edges], [list(a) for a in angles]
    dx, dy = pos[i_b][0] - pos[i_a][0], pos[i_b][1] - pos[i_a][1]
    ang = math.atan2(dy, dx) % (2 * math.pi)
    for idx, a in ((i_a, ang), (i_b, (ang + math.pi) % (2 * math.pi))):
        t = tensors[idx]
        tensors[idx] = t.reshape(t.shape + (1,))
        edges[idx].append("fake_leg")
        angles[idx].append(a)
    return tensors, edges, angles
