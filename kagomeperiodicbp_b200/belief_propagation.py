"""Block belief propagation on the periodic Kagome block -- the accelerated counterpart of
src/algo/belief_propagation.py with the same entry points:

    belief_propagation(tn, messages, config)          (reference :192-281)
    robust_belief_propagation(tn, messages, config)   (reference :285-350)

One BP iteration computes the six outgoing block messages Jacobi-style from the same incoming set
(reference :120-161).  Here each of the six chains is ONE device program that runs, without leaving the
GPU: the boundary-MPS contraction towards its side (``contract_tensor_network`` -> ``bubblecon``), the
normalisation ``_fix_messages`` (:113-117), the overlap with the previous message that ``_compute_error``
needs (:44-56), and the damping ``_single_mps_damping`` (:59-74).  The six programs run concurrently on six
CUDA streams; the host only relabels sides (periodic boundary, :155), averages six numbers into the error
and applies the reference's convergence / failure / retry logic.
"""
from __future__ import annotations

import copy
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import block_tn, contraction_order
from .containers import BPConfig, BPStats, Message, MPSOrientation, UnitCell
from .dev_bubblecon import trace_bubblecon
from .dev_mps import SLOT_LOGNORM, SLOT_NONFINITE, SLOT_TRUNC, DevMPS, add_two_mps, inner_product
from .engine import E_SVD_NOCONV, BubbleConError, raise_if_not_converged
from .lattice import BLOCK_SIDES_CCW, SIDE_ANGLE, SIDE_OPPOSITE, get_block
from .mps import MPS
from .program import Program
from .runtime import Compiled, get_engine

SLOT_IP_RE, SLOT_IP_IM = 3, 4
SLOT_TRUNC_DAMP = 5
N_SLOTS = 8

import os as _os
_pool = ThreadPoolExecutor(max_workers=int(_os.environ.get("KBP_SIDE_THREADS", "6")), thread_name_prefix="kbp-side")
_cache: dict = {}


# ------------------------------------------------------------------------------------------------
# the tensor network container (API of KagomeTNRepeatedUnitCell, src/tensor_networks/tensor_network.py:335)
# ------------------------------------------------------------------------------------------------
class KagomeTNRepeatedUnitCell:
    def __init__(self, unit_cell: UnitCell, N: int, d: int | None = None, D: int | None = None):
        self.unit_cell = unit_cell
        self.N = N
        self.lattice = get_block(N)
        self.d = unit_cell.A.shape[0] if d is None else d
        self.D = unit_cell.A.shape[1] if D is None else D
        self.messages: dict = {}

    @property
    def num_message_connections(self) -> int:
        return self.lattice.L

    @property
    def has_messages(self) -> bool:
        return len(self.messages) == 6

    def connect_messages(self, messages: dict):
        for side, m in messages.items():
            self.messages[side] = m

    def connect_uniform_messages(self):
        self.connect_messages(initial_messages(self.D, self.N, "UQ"))

    def connect_random_messages(self, rng=None):
        self.connect_messages(initial_messages(self.D, self.N, "RQ", rng))

    def message_indices(self, side: str):
        return self.lattice.message_indices(side)


class KagomeTNArbitrary(KagomeTNRepeatedUnitCell):
    """the block with an independent tensor on every lattice site (src/tensor_networks/tensor_network.py:400-431): same BP and
    measurement entry points, `unit_cell` is a `LatticeTensors`.  The lattice size follows from the number of tensors."""

    def __init__(self, tensors):
        from .containers import LatticeTensors
        n = len(tensors)
        N = next((k for k in range(2, 64) if len(get_block(k).sites) == n), None)
        if N is None:
            raise ValueError(f"{n} tensors do not fill a Kagome block")
        lt = tensors if isinstance(tensors, LatticeTensors) else LatticeTensors(tensors)
        super().__init__(lt, N)

    @property
    def tensors(self):
        return list(self.unit_cell.site_tensors)

    def shift_periodically_in_direction(self, direction: str) -> "KagomeTNArbitrary":
        """(src/tensor_networks/tensor_network.py:491-496)"""
        from . import shifting
        return KagomeTNArbitrary(shifting.shift_tensors(self.tensors, shifting.shift_permutation(self.N, direction)))

    def all_lattice_shifting_options(self):
        """every periodic translation of the block, the identity first (src/tensor_networks/tensor_network.py:484-489)"""
        from . import shifting
        for perm in shifting.all_shift_permutations(self.N):
            yield KagomeTNArbitrary(shifting.shift_tensors(self.tensors, perm))


def kagome_tn_from_unit_cell(unit_cell: UnitCell, dims) -> KagomeTNRepeatedUnitCell:
    """(src/tensor_networks/construction.py:45-52)"""
    return KagomeTNRepeatedUnitCell(unit_cell, dims.big_lattice_size, dims.physical_dim, dims.virtual_dim)


# ------------------------------------------------------------------------------------------------
# initial messages (src/tensor_networks/mps.py:77-180): product of vectorised identities (UQ) or random
# |v><v| (RQ) embedded with bond D^2, left-canonicalised on the device
# ------------------------------------------------------------------------------------------------
def _raw_initial_sites(D: int, L: int, random: bool, rng):
    D2, D3 = D * D, D ** 3
    sites = []
    for i in range(L):
        if random:
            rs = np.random if rng is None else rng
            a = rs.normal(size=[D3, D3]) + 1j * rs.normal(size=[D3, D3])
            a /= np.linalg.norm(a)
            kb = a @ np.conj(a.T)
        else:
            kb = np.eye(D3) / np.sqrt(D3)
        kb = kb.reshape([D] * 6).transpose([0, 3, 1, 4, 2, 5]).reshape([D2, D2, D2])
        if i == 0:
            kb = kb[0, :, :].reshape([1, D2, D2])
        if i == L - 1:
            kb = kb[:, :, 0].reshape([kb.shape[0], D2, 1])
        sites.append(np.ascontiguousarray(kb, dtype=np.complex128))
    return sites


def initial_message(D: int, N: int, message_model: str = "RQ", rng=None) -> MPS:
    """``N`` = number of MPS sites (= 2*block_size - 1), as in the reference's signature.

    Setup data, not hot path: the message is a rank-1 product state embedded with bond D^2, so its
    left-canonical form contains an ARBITRARY orthonormal completion of the null space of every bond, and
    the first BP iteration depends on that choice (a partially swallowed message exposes the completion
    columns to the truncation).  To start from the reference's exact tensors the completion must be LAPACK's,
    i.e. numpy.linalg.qr on the host as in src/tensor_networks/mps.py:150-154 / src/libs/bmpslib.py:553-595.
    For 'UQ' the result is a constant table that depends on (D, N) only."""
    sites = _raw_initial_sites(D, N, message_model in ("RQ", "RANDOM_QUANTUM"), rng)
    corder = [None] * N
    for i in range(N - 1):
        D1, d, D2 = sites[i].shape
        Q, R = np.linalg.qr(sites[i].reshape(D1 * d, D2))
        sites[i] = Q.reshape(D1, d, Q.shape[1])
        corder[i] = "L"
        sites[i + 1] = np.tensordot(R, sites[i + 1], axes=([1], [0]))
    sites[N - 1] = sites[N - 1] / np.linalg.norm(sites[N - 1])
    return MPS.from_sites(sites, Corder=corder)


def initial_messages(D: int, block_N: int, model: str = "UQ", rng=None) -> dict:
    L = 2 * block_N - 1
    out = {}
    for side in BLOCK_SIDES_CCW:
        out[side] = Message(initial_message(D, L, model, rng), MPSOrientation.standard(SIDE_OPPOSITE[side]))
    return out


# ------------------------------------------------------------------------------------------------
# one chain = one program
# ------------------------------------------------------------------------------------------------
def _msg_shapes(messages: dict):
    return tuple(tuple(tuple(a.shape) for a in messages[s].mps.A) for s in BLOCK_SIDES_CCW)


def _node_tensor(k: int, L: int, t):
    """message site k as the 2- or 3-leg node the TN uses (src/tensor_networks/tensor_network.py:855-870)."""
    if k == 0:
        return t.reshape(t.shape[1], t.shape[2])
    if k == L - 1:
        return t.reshape(t.shape[0], t.shape[1])
    return t


def compile_side_program(N: int, d: int, D: int, side: str, chi: int, msg_shapes, damping, depth="ToMessage",
                         epilogue=True, arbitrary=False) -> Compiled:
    """`arbitrary`: one input tensor per lattice site ("site{i}", the non-repeated block) instead of the unit cell's three."""
    key = ("side", N, d, D, side, chi, msg_shapes, damping, depth, epilogue, arbitrary)
    if key in _cache:
        return _cache[key]
    blk = get_block(N)
    L = blk.L
    p = Program(N_SLOTS)
    ins = []
    cell_dt = []
    for nm in ([f"site{s_.index}" for s_ in blk.sites] if arbitrary else ["cellA", "cellB", "cellC"]):
        t = p.input(nm, (d, D, D, D, D))
        ins.append((nm, t))
        cell_dt.append(t)
    used = [s for s in BLOCK_SIDES_CCW if s != side] if depth == "ToMessage" else list(BLOCK_SIDES_CCW)
    msg_dt = {}
    for si, s in enumerate(BLOCK_SIDES_CCW):
        if s not in used:
            continue
        msg_dt[s] = []
        for k, sh in enumerate(msg_shapes[si]):
            t = p.input(f"m{s}{k}", sh)
            ins.append((f"m{s}{k}", t))
            msg_dt[s].append(t)
    # TN description with shape-only placeholders, then swap in device tensors
    dummy_cell = [np.empty((d, D, D, D, D))] * 3
    dummy_msgs = {s: [np.empty(sh) for sh in msg_shapes[si]] for si, s in enumerate(BLOCK_SIDES_CCW)}
    T, E, A, K, P = block_tn.assemble(N, dummy_cell, dummy_msgs)
    T, E, A = block_tn.connect_corner(N, T, E, A, P, side)
    TL = [None] * len(T)
    for s_ in blk.sites:
        TL[s_.index] = cell_dt[s_.index] if arbitrary else cell_dt[s_.index % 3]
    for s in BLOCK_SIDES_CCW:
        for k, idx in enumerate(blk.message_indices(s)):
            if s in msg_dt:
                TL[idx] = _node_tensor(k, L, msg_dt[s][k]).reshape(T[idx].shape)
    order = list(contraction_order.kagome_order(N, side, depth))
    mp, _ = trace_bubblecon(p, TL, E, A, SIDE_ANGLE[side], order, chi, ket_tensors=K, slots=(SLOT_LOGNORM, SLOT_TRUNC))
    outs = []
    if epilogue:
        opp = SIDE_OPPOSITE[side]
        # _fix_messages: right-canonical + unit norm (reference :113-117)
        mp.right_canonical(nr_bulk=True)
        for k, t in enumerate(mp.dense_sites()):
            outs.append((f"out{k}", t))
            p.nonfinite(t, SLOT_NONFINITE)
        prev = DevMPS(p, L)
        for k, t in enumerate(msg_dt[opp]):
            prev.set_site(t, k)
        ip = inner_product(p, prev, mp)                       # <prev|out>
        p.scalar_to_slot(ip, SLOT_IP_RE, SLOT_IP_IM)
        if damping:
            comb = add_two_mps(p, mp, 1.0 - damping, prev, damping, sign_slot_beta=SLOT_IP_RE)
            comb.slot_lognorm, comb.slot_trunc = SLOT_LOGNORM, SLOT_TRUNC_DAMP
            comb.left_canonical_QR()
            comb.right_canonical(maxD=chi, nr_bulk=True)
            for k, t in enumerate(comb.dense_sites()):
                outs.append((f"next{k}", t))
    else:
        for k, t in enumerate(mp.dense_sites()):
            outs.append((f"out{k}", t))
    comp = Compiled(p, ins, outs, meta=dict(n_out=mp.N, L=L, side=side, svd_shapes=list(p.svd_shapes),
                                            qr_shapes=list(p.qr_shapes), gemm_flops=p.gemm_flops))
    _cache[key] = comp
    return comp


def contract_tensor_network(tn: "KagomeTNRepeatedUnitCell", direction: str, depth: str, bubblecon_trunc_dim: int, allow_progressbar: bool = False):
    """(src/algo/contract_tensor_network.py:146-213)  one boundary-MPS contraction of the block with its messages attached:
    -> (mps | (mantissa, exp10), contraction_order, MPSOrientation).  ``depth`` in {"ToMessage", "ToCore", "Full"}.  The truncation
    algorithm follows the reference's rule (BubbleConGlobalConfig.bubblecon_compression: SVD up to D = 10, iterative above)."""
    from .bubblecon import bubblecon
    from .containers import BubbleConGlobalConfig
    gc = BubbleConGlobalConfig()
    cell = tn.unit_cell
    msgs = {s: m.mps.A for s, m in tn.messages.items()}
    T, E, A, K, P = block_tn.assemble(tn.N, cell.tensors(), msgs)
    T, E, A = block_tn.connect_corner(tn.N, T, E, A, P, direction)
    order = list(contraction_order.kagome_order(tn.N, direction, depth))
    mps = bubblecon(T, E, A, SIDE_ANGLE[direction], order, D_trunc=bubblecon_trunc_dim, ket_tensors=K, separate_exp=gc.separate_exp,
                    compression=gc.bubblecon_compression(tn.D))
    return mps, order, MPSOrientation.standard(direction)


def _is_arbitrary(cell) -> bool:
    return hasattr(cell, "site_tensors")


def _side_inputs(cell, messages: dict, comp: Compiled) -> dict:
    if _is_arbitrary(cell):
        d = {f"site{i}": t for i, t in enumerate(cell.site_tensors)}
    else:
        d = {"cellA": cell.A, "cellB": cell.B, "cellC": cell.C}
    names = {n for n, _, _ in comp.in_layout}
    for s in BLOCK_SIDES_CCW:
        for k, a in enumerate(messages[s].mps.A):
            nm = f"m{s}{k}"
            if nm in names:
                d[nm] = a
    return d


def _run_side(side: str, comp: Compiled, batch_inputs: list, device: int):
    eng = get_engine(("side", side), device)
    outs, slots, rc = comp.run(eng, batch_inputs, soft_errors=(E_SVD_NOCONV,))
    return side, outs, slots, rc


def run_sides(N: int, cells: list, messages_list: list, config: BPConfig, device: int = 0, sides=None) -> dict:
    """run the chain programs of ``sides`` (default: all six) concurrently, one CUDA stream each.
    -> {side: (outs per cell, slots[nb, n_slots], rc)}"""
    d, D = cells[0].A.shape[0], cells[0].A.shape[1]
    shapes = _msg_shapes(messages_list[0])
    for m in messages_list[1:]:
        assert _msg_shapes(m) == shapes, "batched cells must share message shapes"
    damping = config.damping if config.damping else None
    from .containers import BubbleConGlobalConfig
    if BubbleConGlobalConfig().bubblecon_compression(D)["type"] != "SVD":
        raise NotImplementedError(f"block BP at D = {D} > 10: the reference switches bubblecon to iterative compression there; the fused "
                                  "side programs truncate by SVD -- use contract_tensor_network(...) (stepwise device path) for such bonds")
    if config.fix_msg_each_step is False:
        raise NotImplementedError("BPConfig.fix_msg_each_step=False: the side programs always normalise the new message "
                                  "(the reference's default, src/algo/belief_propagation.py:158-159)")
    todo = []
    for side in (BLOCK_SIDES_CCW if sides is None else sides):
        comp = compile_side_program(N, d, D, side, config.trunc_dim, shapes, damping, arbitrary=_is_arbitrary(cells[0]))
        batch = [_side_inputs(c, m, comp) for c, m in zip(cells, messages_list)]
        todo.append((side, comp, batch, get_engine(("side", side), device)))
    res = {}
    if all(eng.graph_ready(comp.words) for _, comp, _, eng in todo):
        # steady state: every side program is one CUDA-graph launch on its own stream -- queue all of them from this thread,
        # then collect; the host takes no part in the iteration and no launch thread spins on a stream
        if max(len(comp.words) for _, comp, _, _ in todo) > BIG_GRAPH_WORDS:
            # the launch call of a very large graph (N = 6: ~23k kernel nodes) keeps the host busy for 20-30 ms: six short-lived
            # launch threads (they return as soon as the graph is queued) start the sides together instead of 150 ms apart
            futs = [_pool.submit(comp.launch, eng, batch, (E_SVD_NOCONV,)) for _, comp, batch, eng in todo]
            rcs = [f.result() for f in futs]
        else:
            rcs = [comp.launch(eng, batch, soft_errors=(E_SVD_NOCONV,)) for _, comp, batch, eng in todo]
        for (side, comp, batch, eng), rc in zip(todo, rcs):
            res[side] = comp.collect(eng, len(batch), rc)
        return res
    # first sight of a program (or graphs disabled / a profiler attached): host-driven loops, one launch thread per side
    futs = [_pool.submit(_run_side, side, comp, batch, device) for side, comp, batch, _ in todo]
    for f in futs:
        side, outs, slots, rc = f.result()
        res[side] = (outs, slots, rc)
    return res


BIG_GRAPH_WORDS = 20000      # op-stream length from which a program's graph launch is worth its own host thread


def assemble_step(res: dict, n_cells: int, config: BPConfig):
    """per cell (out_messages, next_messages, error, trunc_error) from the six sides' raw results: relabel to the
    opposite side (periodic block, reference :155), error = mean of 1 - |<prev|out>| (:44-56)."""
    damping = config.damping if config.damping else None
    results = []
    for ci in range(n_cells):
        out_msgs, next_msgs, dists, trunc = {}, {}, [], 0.0
        for side in BLOCK_SIDES_CCW:
            outs, slots, rc = res[side]
            if slots[ci, -1] > 0:                                  # engine status slot: a truncation of this chain did not converge
                raise_if_not_converged(E_SVD_NOCONV, f"block message towards {side}")
            if slots[ci, SLOT_NONFINITE] > 0:
                raise BubbleConError(f"non-finite values in the outgoing message towards {side}")
            opp = SIDE_OPPOSITE[side]
            o = outs[ci]
            n_out = sum(1 for k in o if k.startswith("out"))
            orient = MPSOrientation.standard(side)
            out_mps = MPS.from_sites([o[f"out{k}"] for k in range(n_out)], Corder=[None] + ["R"] * (n_out - 1))
            out_msgs[opp] = Message(out_mps, orient)
            if damping:
                nx = MPS.from_sites([o[f"next{k}"] for k in range(n_out)], Corder=[None] + ["R"] * (n_out - 1))
                next_msgs[opp] = Message(nx, orient)
            ip = complex(slots[ci, SLOT_IP_RE], slots[ci, SLOT_IP_IM])
            dist = 1.0 - abs(ip)
            dists.append(0.0 if dist < 0 else dist)
            trunc += float(slots[ci, SLOT_TRUNC])
        if config.msg_diff_squared:
            err = sum(dists) / len(dists)
        else:
            err = float(np.sqrt(sum(dists)) / len(dists))
        results.append((out_msgs, next_msgs if damping else out_msgs, float(err), trunc))
    return results


def bp_step_batch(N: int, cells: list, messages_list: list, config: BPConfig, device: int = 0):
    """one BP iteration for a batch of independent unit cells that share all shapes.  Returns per cell
    (out_messages, next_messages, error, trunc_error)."""
    return assemble_step(run_sides(N, cells, messages_list, config, device), len(cells), config)


def _belief_propagation_step(tn: KagomeTNRepeatedUnitCell, prev_messages: dict, prev_error, config: BPConfig, prog_bar_obj=None):
    """(reference :164-188) -> (out_messages, next_messages, next_error)"""
    out, nxt, err, _ = bp_step_batch(tn.N, [tn.unit_cell], [prev_messages], config)[0]
    return out, nxt, err


# ------------------------------------------------------------------------------------------------
# hermitisation at the end of BP (src/libs/ITE.py:116-185)
# ------------------------------------------------------------------------------------------------
def _hermitize_program(shapes) -> Compiled:
    key = ("herm", tuple(shapes))
    if key in _cache:
        return _cache[key]
    p = Program(N_SLOTS)
    ins = [(f"s{k}", p.input(f"s{k}", sh)) for k, sh in enumerate(shapes)]
    N = len(shapes)
    mpA, mpB = DevMPS(p, N), DevMPS(p, N)
    Dmax = 0
    for k, (_, t) in enumerate(ins):
        DL, d2, DR = t.shape
        dd = int(round(np.sqrt(d2)))
        mpA.set_site(t, k)
        tb = p.transpose(t.reshape(DL, dd, dd, DR), (0, 2, 1, 3), conj=True).reshape(DL, d2, DR)
        mpB.set_site(tb, k)
        Dmax = max(Dmax, DL)
    mpC = add_two_mps(p, mpA, 0.5, mpB, 0.5)
    mpC.reduceD(Dmax)
    comp = Compiled(p, ins, [(f"o{k}", t) for k, t in enumerate(mpC.dense_sites())])
    _cache[key] = comp
    return comp


def _hermitize_messages(messages: dict) -> dict:
    def one(side):
        m = messages[side]
        comp = _hermitize_program([a.shape for a in m.mps.A])
        outs, _, rc = comp.run(get_engine(("side", side)), [{f"s{k}": a for k, a in enumerate(m.mps.A)}],
                               soft_errors=(E_SVD_NOCONV,))
        raise_if_not_converged(rc, f"hermitisation of the message of side {side}")
        o = outs[0]
        return side, Message(MPS.from_sites([o[f"o{k}"] for k in range(m.mps.N)]), m.orientation)
    return dict(_pool.map(one, list(messages.keys())))


# ------------------------------------------------------------------------------------------------
def belief_propagation(tn: KagomeTNRepeatedUnitCell, messages: dict | None = None, config: BPConfig = None):
    """(reference :192-281)  -> (messages, BPStats)"""
    config = BPConfig() if config is None else config
    t0 = time.perf_counter()
    if messages is None:
        tn.connect_random_messages() if config.init_msg in ("RQ", "RANDOM_QUANTUM") else tn.connect_uniform_messages()
    else:
        tn.connect_messages(messages)
    messages = tn.messages
    errors, trunc_errs = [], []
    min_error, min_messages = np.inf, messages
    next_messages = out_messages = messages
    error, success, i = None, False, 0
    it = 0
    while config.max_iterations is None or it < config.max_iterations:
        i = it
        out_messages, next_messages, error, trunc = bp_step_batch(tn.N, [tn.unit_cell], [next_messages], config)[0]
        trunc_errs.append(trunc)
        it += 1
        if error < config.msg_diff_terminate:
            success = True
            errors.append(error)
            break
        tn.connect_messages(next_messages)
        if error < min_error:
            min_error, min_messages = error, copy.deepcopy(out_messages)
        errors.append(error)
        k = config.times_to_deem_failure_when_diff_increases
        if len(errors) > k and all(a <= b for a, b in zip(errors[-k:], errors[-k:][1:])):
            break
    assert isinstance(error, float)
    if not success:
        out_messages, error = min_messages, min_error
    if config.hermitize_msgs_when_finished:
        out_messages = _hermitize_messages(out_messages)
    tn.connect_messages(out_messages)
    stats = BPStats(iterations=i + 1, final_error=float(error), final_config=config, success=success,
                    execution_time=time.perf_counter() - t0, errors=errors, truncation_errors=trunc_errs)
    return out_messages, stats


def robust_belief_propagation(tn: KagomeTNRepeatedUnitCell, messages: dict | None = None, config: BPConfig = None):
    """(reference :285-350)"""
    config = (BPConfig() if config is None else config).copy()
    t0 = time.perf_counter()
    messages_in = copy.deepcopy(messages)
    min_messages, min_error, total_iterations = messages_in, np.inf, 0
    messages_out = error_out = None
    attempt_ind, stats = 0, None
    for attempt_ind in range(config.allowed_retries):
        msgs, stats = belief_propagation(tn, messages_in, config)
        total_iterations += stats.iterations
        if stats.final_error < config.msg_diff_terminate:
            messages_out, error_out = msgs, stats.final_error
            break
        if stats.final_error < min_error:
            min_error, min_messages = stats.final_error, copy.deepcopy(msgs)
        config.trunc_dim = int(1.5 * config.trunc_dim)
        if isinstance(config.max_iterations, int):
            config.max_iterations += 11
        messages_in = None
    else:
        messages_out, error_out = min_messages, min_error
    tn.connect_messages(messages_out)
    return messages_out, BPStats(attempts=attempt_ind + 1, iterations=total_iterations, final_error=float(error_out),
                                 final_config=stats.final_config, success=error_out < config.msg_diff_good_enough,
                                 execution_time=time.perf_counter() - t0)
