"""CPU tier: the N > 1 path with world_size 2 over gloo.  The chain programs run on the numpy interpreter of the op
stream (tests/np_vm.py) in each spawned process; what is under test is the host logic: side sharding, packing, the one
all-gather per iteration, and that every rank assembles the same iteration as the unsharded step."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from np_vm import NumpyEngine
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200 import parallel
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    engines = {}
    bp.get_engine = lambda key="default", device=0: engines.setdefault(key, NumpyEngine())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, N = 2, 2
    cell = UnitCell.random(2, D, seed=3)
    cfg = BPConfig(trunc_dim=2 * D * D, msg_diff_terminate=1e-6, damping=0.1, init_msg="UQ")
    msgs = bp.initial_messages(D, N, "UQ")
    errs = []
    for it in range(3):
        (res,), nbytes = parallel.bp_step_sharded(N, [cell], [msgs], cfg, rank, world)
        out, msgs, err, trunc = res
        errs.append(err)
    ref_msgs = bp.initial_messages(D, N, "UQ")
    ref_errs = []
    for it in range(3):
        out_r, ref_msgs, err_r, _ = bp.bp_step_batch(N, [cell], [ref_msgs], cfg)[0]
        ref_errs.append(err_r)
    same = all(np.array_equal(a, b) for s in msgs for a, b in zip(msgs[s].mps.A, ref_msgs[s].mps.A))
    # the persistent form: messages stay in the engines' arenas between iterations, one all-gather + one indexed copy per step
    sh = parallel.ShardedSides(N, cell, msgs, cfg, rank, world, engine_key="sharded")
    m2 = ref_msgs
    for it in range(3):
        e_sh, _, ok = sh.step()
        _, m2, e_ref, _ = bp.bp_step_batch(N, [cell], [m2], cfg)[0]
        same = same and ok and e_sh == e_ref
    _, nxt = sh.messages()
    same = same and all(np.array_equal(a, b) for s in nxt for a, b in zip(nxt[s].mps.A, m2[s].mps.A))
    gathered = parallel.gather_scalars([errs[-1], float(rank)], world)
    dist.destroy_process_group()
    q.put((rank, errs, ref_errs, same, nbytes, [g.tolist() for g in gathered]))


def test_sides_of_rank_partition():
    from kagomeperiodicbp_b200.parallel import sides_of_rank
    from kagomeperiodicbp_b200.lattice import BLOCK_SIDES_CCW
    for world in (1, 2, 4, 8):
        got = [s for r in range(world) for s in sides_of_rank(r, world)]
        assert sorted(got) == sorted(BLOCK_SIDES_CCW)
    assert [len(sides_of_rank(r, 4)) for r in range(4)] == [2, 2, 1, 1]


def test_bp_step_sharded_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs, ref_errs, same, nbytes, gathered in results:
        assert errs == ref_errs                      # bit-identical to the unsharded iteration on every rank
        assert same
        assert nbytes > 0
        assert gathered[0][1] == 0.0 and gathered[1][1] == 1.0 and gathered[0][0] == gathered[1][0]
