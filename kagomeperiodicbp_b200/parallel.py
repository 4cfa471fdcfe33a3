"""Multi-GPU block BP: one process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).

Two shardings, both from SURVEY 8e:
  * independent unit cells (ensembles, field sweeps): every rank runs whole BP / ITE problems, no data-path collective;
    only scalars are gathered at the end  (``gather_scalars``).
  * block sides of ONE unit cell: the six outgoing messages of an iteration depend only on the previous iteration's six
    messages (src/algo/belief_propagation.py:149-155), so rank r computes the chains of the sides ``sides_of_rank(r)`` on its
    GPU and ONE all-gather per BP iteration publishes every rank's new messages (+ its overlap / truncation scalars) to all
    ranks (``bp_step_sharded``).  Payload: 6 x (2N-1) sites of [chi, D^2, chi] complex128, e.g. 17 MB at D=4, N=6.
Inside a chain the swallow steps are strictly sequential: nothing else shards.
"""
from __future__ import annotations

import numpy as np

from . import belief_propagation as bp
from .lattice import BLOCK_SIDES_CCW


def sides_of_rank(rank: int, world: int):
    """round-robin over the counter-clockwise side order: 2 ranks -> 3+3, 4 ranks -> 2+2+1+1, >= 6 ranks -> 1 each (rest idle)."""
    return [s for k, s in enumerate(BLOCK_SIDES_CCW) if k % world == rank]


def _layout(comp):
    return [(name, int(np.prod(shape)), tuple(shape)) for name, _, shape in comp.out_layout]


def pack_side(comp, outs, slots) -> np.ndarray:
    """one side's raw result (all cells) as a flat float64 vector: [slots | out tensors re/im interleaved]."""
    parts = [np.asarray(slots, dtype=np.float64).ravel()]
    for o in outs:
        for name, n, shape in _layout(comp):
            parts.append(np.ascontiguousarray(o[name], dtype=np.complex128).ravel().view(np.float64))
    return np.concatenate(parts)


def unpack_side(comp, flat: np.ndarray, n_cells: int, n_slots: int):
    pos = n_cells * n_slots
    slots = flat[:pos].reshape(n_cells, n_slots).copy()
    outs = []
    for _ in range(n_cells):
        d = {}
        for name, n, shape in _layout(comp):
            d[name] = flat[pos:pos + 2 * n].view(np.complex128).reshape(shape).copy()
            pos += 2 * n
        outs.append(d)
    return outs, slots


def bp_step_sharded(N: int, cells: list, messages_list: list, config, rank: int, world: int, device: int = 0, group=None,
                    torch_device=None):
    """one BP iteration with the six sides sharded over ``world`` ranks and one all-gather of the new messages.
    Every rank returns the same (out_messages, next_messages, error, trunc_error) per cell."""
    import torch
    import torch.distributed as dist
    d, D = cells[0].A.shape[0], cells[0].A.shape[1]
    shapes = bp._msg_shapes(messages_list[0])
    damping = config.damping if config.damping else None
    comps = {s: bp.compile_side_program(N, d, D, s, config.trunc_dim, shapes, damping) for s in BLOCK_SIDES_CCW}
    nb = len(cells)
    mine = sides_of_rank(rank, world)
    local = bp.run_sides(N, cells, messages_list, config, device, sides=mine) if mine else {}
    sizes = {s: nb * bp.N_SLOTS + 2 * nb * sum(n for _, n, _ in _layout(comps[s])) for s in BLOCK_SIDES_CCW}
    per_rank = [sum(sizes[s] for s in sides_of_rank(r, world)) for r in range(world)]
    width = max(per_rank)
    buf = np.zeros(width, dtype=np.float64)
    pos = 0
    for s in mine:
        outs, slots, rc = local[s]
        v = pack_side(comps[s], outs, slots)
        buf[pos:pos + v.size] = v
        pos += v.size
    dev = torch_device if torch_device is not None else "cpu"
    send = torch.from_numpy(buf).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)                      # THE collective of a BP iteration
    res = {}
    for r in range(world):
        flat = recv[r].cpu().numpy()
        pos = 0
        for s in sides_of_rank(r, world):
            outs, slots = unpack_side(comps[s], flat[pos:pos + sizes[s]], nb, bp.N_SLOTS)
            res[s] = (outs, slots, 0)
            pos += sizes[s]
    return bp.assemble_step(res, nb, config), width * 8


def gather_scalars(values, world: int, group=None, torch_device=None):
    """all ranks' scalar results (energies, errors) on every rank."""
    import torch
    import torch.distributed as dist
    dev = torch_device if torch_device is not None else "cpu"
    send = torch.tensor(list(values), dtype=torch.float64, device=dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return [r.cpu().numpy() for r in recv]
