#!/usr/bin/env python
"""Headline benchmark: block-BP message updates per second on the periodic Kagome block, D=4
(BASELINE.json metric), one process per GPU.

A "step" is ONE BP iteration of one unit cell per chain slot: the six outgoing block messages (six
boundary-MPS chains with per-swallow truncation + normalise + overlap + damping), i.e. 6 message
updates per unit cell.  With --gpus N (torchrun), every rank runs the same step on its own, differently
seeded unit cells (independent states: no data-path collective, weak scaling); rank 0 prints one JSON line.

  value     message updates/s, inputs resident in HBM, six side programs on six CUDA streams
  e2e       same metric through the public API call bp_step_batch(...) with HOST buffers
            (pack + H2D + run + D2H inside the timed region)
  roofline  FP64 tensor (DMMA) roofline of the dominant kernel family, the block-Jacobi truncated SVD:
            algorithmic flops (SURVEY 8d convention 4*(14 m n^2 + 8 n^3) per SVD) / CUDA-event time of the
            SVD ops in an instrumented repetition of the same step; peak = cuBLAS DGEMM measured here
  cpu_baseline  the numpy oracle (port of the reference path, exact-SVD branch) on this box's host cores,
            bounded sample (one chain or a prefix of it)
  ite       second headline metric (BASELINE.json: "ms per ITE step"): one loop body of ite_per_mode = warm BP to
            tolerance + two ToCore chains + core->mode->edge reduction + RDM + gate/ALS + energy, wall clock on rank 0
`--impl reference` times that CPU port as the reference arm (the reference is pure Python + numpy; it cannot
travel to the GPU box, the oracle is its pinned restatement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "bp_message_updates_per_s"
UNIT = "msg/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--D", type=int, default=4)
    ap.add_argument("--N", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1, help="unit cells per GPU batched into every launch")
    ap.add_argument("--damping", type=float, default=0.1)
    ap.add_argument("--shard", default="cells", choices=("cells", "sides"),
                    help="N > 1: independent unit cells per rank (weak scaling, no collective) or the six sides of ONE cell "
                         "sharded over the ranks with one all-gather of the new messages per iteration (strong scaling)")
    ap.add_argument("--ensemble", type=int, default=4, help="unit cells per launch for the extra ensemble measurement at N=1 (0 = skip)")
    ap.add_argument("--ite-steps", type=int, default=3, help="ITE steps (loop bodies of ite_per_mode) timed on rank 0 at N=1; 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def workload_name(a):
    n_s = 3 * (3 * a.N * a.N - 3 * a.N + 1)
    return (f"Kagome Heisenberg-PEPS block BP, D={a.D}, block N={a.N} ({n_s} sites, 6 messages x {2 * a.N - 1} MPS sites), "
            f"chi_bp={2 * a.D * a.D}, damping={a.damping}, complex128")


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.rows, self.stop_flag = dev, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_sample(a, cell, messages, budget_s):
    """time the oracle on one chain (or a prefix of its swallow order) of the SAME workload."""
    from helpers import to_oracle_mps
    from kagomeperiodicbp_b200 import block_tn, contraction_order
    from kagomeperiodicbp_b200.lattice import SIDE_ANGLE
    from oracle.bubblecon_np import bubblecon as obub
    side = "D"
    msgs = {s: m.mps.A for s, m in messages.items()}
    T, E, A, K, P = block_tn.assemble(a.N, cell.tensors(), msgs)
    T, E, A = block_tn.connect_corner(a.N, T, E, A, P, side)
    order = list(contraction_order.kagome_order(a.N, side, "ToMessage"))
    chi = 2 * a.D * a.D
    # pick a prefix whose cost fits the budget: probe a short prefix first
    k0 = min(len(order), 8 + 2 * (2 * a.N - 1))
    t0 = time.perf_counter()
    obub(T, E, A, SIDE_ANGLE[side], order[:k0], D_trunc=chi, ket_tensors=K)
    t_probe = time.perf_counter() - t0
    per_swallow = t_probe / k0
    k = len(order) if per_swallow * len(order) * 1.5 < budget_s else max(k0, int(budget_s / (1.5 * per_swallow)))
    k = min(k, len(order))
    t0 = time.perf_counter()
    obub(T, E, A, SIDE_ANGLE[side], order[:k], D_trunc=chi, ket_tensors=K)
    t = time.perf_counter() - t0
    # extrapolate a prefix by the steady-state cost of the swallows it did after the probe region
    if k < len(order):
        t_full = t + (t - t_probe) / max(1, k - k0) * (len(order) - k)
    else:
        t_full = t
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle (numpy port of the reference path, exact-SVD branch) on {k}/{len(order)} swallows of one ToMessage chain "
                      f"(side D) of the same workload, {t:.1f} s measured" + ("" if k == len(order) else f", extrapolated to {t_full:.1f} s per message"),
            "seconds_per_message": t_full}


def ite_metric(a, cell, messages, cfg, cpu_bp):
    """ms per ITE step: loop bodies of ite_per_mode (BASELINE.json's second metric) on the device, wall clock, after one
    untimed pass over the same edges (program compilation + arena allocation), messages warm as in a running ITE."""
    from kagomeperiodicbp_b200 import edge_env, ite_flow
    chi = 2 * a.D * a.D + 10
    edges = [edge_env.EDGES[k % 6] for k in range(a.ite_steps)]
    uc, msgs = cell, messages
    for e in edges:                                   # warm-up pass
        uc, msgs, _, _ = ite_flow.ite_edge_update(uc, msgs, a.N, "A", e, 1e-2, cfg, chi)
    ts, parts, its, energies = [], [0.0, 0.0, 0.0], [], []
    for e in edges:
        t0 = time.perf_counter()
        uc, msgs, energy, st = ite_flow.ite_edge_update(uc, msgs, a.N, "A", e, 1e-2, cfg, chi)
        ts.append(time.perf_counter() - t0)
        parts = [parts[0] + st.t_bp, parts[1] + st.t_reduce, parts[2] + st.t_update]
        its.append(st.bp_iterations)
        energies.append(energy)
    out = {"metric": "ms_per_ite_step", "value": 1e3 * float(np.mean(ts)), "unit": "ms", "higher_is_better": False, "steps": len(ts),
           "edges": edges, "mode": "A", "delta_t": 1e-2, "chi": chi, "bp_iterations_per_step": its,
           "breakdown_ms": {"bp": 1e3 * parts[0] / len(ts), "reduce_to_edge": 1e3 * parts[1] / len(ts), "rdm_gate_als": 1e3 * parts[2] / len(ts)},
           "edge_energies_after": energies, "device_linalg_calls": ite_flow.backend().calls}
    if cpu_bp is not None:
        # CPU side of the same step, composed from timed parts: BP iterations and the two ToCore chains at the oracle's
        # seconds per chain (measured above), core -> edge reduction + RDM + gate/ALS timed here with numpy/LAPACK
        from oracle import ite_np
        from oracle.bubblecon_np import bubblecon as obub
        from kagomeperiodicbp_b200 import ite
        env12 = ite_flow.reduce_to_core(uc, msgs, a.N, chi)             # inputs for the host-timed part
        t0 = time.perf_counter()
        fn = lambda T, E, A, ang, order, c, kets: obub(T, E, A, ang, order, D_trunc=c, ket_tensors=kets).A
        ti, tj, env, _ = edge_env.edge_environment(ite_np.NP, a.N, uc.tensors(), env12, "A", edges[0], chi, fn)
        ite_np.rho_ij(ti, tj, env)
        tin, tjn, _ = ite_np.apply_2local_gate(ite.g_from_exp_h(ite.heisenberg_afm(), 1e-2), a.D, ti, tj, env)
        ite_np.rho_ij(tin, tjn, env)
        t_edge = time.perf_counter() - t0
        spm = cpu_bp["seconds_per_message"]
        cpu_ms = 1e3 * ((6 * float(np.mean(its)) + 2) * spm + t_edge)
        out["cpu_baseline"] = {"value": cpu_ms, "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"composed: ({float(np.mean(its)):.1f} BP iterations x 6 + 2 ToCore) chains at {spm:.2f} s per chain (oracle, "
                                         f"measured above) + core->edge reduction, RDM, gate/ALS for edge {edges[0]} timed with numpy ({t_edge:.2f} s)"}
    return out


def run_reference_arm(a, rank):
    if rank != 0:
        return
    try:                      # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread back
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    from kagomeperiodicbp_b200.containers import UnitCell
    from oracle import bp_np
    cell = UnitCell.random(2, a.D, seed=0)

    class _M:   # minimal Message-like wrapper around oracle MPS objects
        def __init__(self, m):
            self.mps = m
    msgs = {s: _M(m) for s, m in bp_np.uniform_messages(a.N, a.D).items()}
    vals, times = [], []
    per = a.cpu_budget_s * 4.0 / max(1, a.steps + a.warmup)   # whole arm stays within a few minutes
    for i in range(a.warmup + a.steps):
        r = cpu_sample(a, cell, msgs, per)
        if i >= a.warmup:
            vals.append(r["value"])
            times.append(r["seconds_per_message"])
    v = float(np.mean(vals))
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * 6 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": r["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank)
        return

    import torch
    import torch.distributed as dist
    from kagomeperiodicbp_b200 import belief_propagation as bp
    from kagomeperiodicbp_b200.containers import BPConfig, UnitCell
    from kagomeperiodicbp_b200.lattice import BLOCK_SIDES_CCW
    from kagomeperiodicbp_b200.runtime import get_engine

    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner when NCCL_DEBUG is
    # set on the box) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    D, N, B = a.D, a.N, a.batch
    chi = 2 * D * D
    cfg = BPConfig(trunc_dim=chi, msg_diff_terminate=1e-6, damping=a.damping, init_msg="UQ")
    cells = [UnitCell.random(2, D, seed=rank * B + i) for i in range(B)]
    uq = bp.initial_messages(D, N, "UQ")
    msgs_list = [uq] * B
    # two untimed iterations: brings the messages (and therefore every program shape) to the steady state
    for _ in range(2):
        res = bp.bp_step_batch(N, cells, msgs_list, cfg, device=dev)
        msgs_list = [r[1] for r in res]
    shapes = bp._msg_shapes(msgs_list[0])
    comps = {s: bp.compile_side_program(N, 2, D, s, chi, shapes, a.damping) for s in BLOCK_SIDES_CCW}
    engs = {s: get_engine(("side", s), dev) for s in BLOCK_SIDES_CCW}
    h2d = d2h = 0
    for s in BLOCK_SIDES_CCW:
        batch = [bp._side_inputs(c, m, comps[s]) for c, m in zip(cells, msgs_list)]
        comps[s].load(engs[s], B)
        engs[s].upload(0, comps[s].pack_inputs(batch))
        h2d += comps[s].in_elems * 16 * B
        d2h += comps[s].out_elems * 16 * B + 8 * bp.N_SLOTS * B
        engs[s].sync()

    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        futs = [bp._pool.submit(comps[s].run_resident, engs[s], (bp.E_SVD_NOCONV,)) for s in BLOCK_SIDES_CCW]
        for f in futs:
            f.result()
        for s in BLOCK_SIDES_CCW:
            engs[s].sync()

    def launches():
        return sum(engs[s].launch_count() for s in BLOCK_SIDES_CCW)

    for _ in range(a.warmup):
        resident_step()
    sampler = ClockSampler(dev)
    sampler.start()
    barrier()
    l0 = launches()
    ms_total = 0.0
    for _ in range(a.steps):
        flush.zero_()                      # L2 flush between timed iterations (512 MiB > 126 MB L2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        resident_step()
        e1.record()
        torch.cuda.synchronize()
        ms_total += e0.elapsed_time(e1)
    n_launch = launches() - l0
    barrier()
    # ---- end to end through the public API with host buffers
    t_e2e = 0.0
    for i in range(a.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        bp.bp_step_batch(N, cells, msgs_list, cfg, device=dev)
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
    # ---- sides sharded over the ranks: one all-gather of the new messages per iteration (config C3)
    ms_sharded, gather_bytes = None, 0
    if a.shard == "sides" and world > 1:
        from kagomeperiodicbp_b200 import parallel
        tdev = torch.device("cuda", local)
        for _ in range(a.warmup):
            parallel.bp_step_sharded(N, cells[:1], msgs_list[:1], cfg, rank, world, dev, torch_device=tdev)
        tot = 0.0
        for i in range(a.steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            _, gather_bytes = parallel.bp_step_sharded(N, cells[:1], msgs_list[:1], cfg, rank, world, dev, torch_device=tdev)
            torch.cuda.synchronize()
            tot += time.perf_counter() - t0
        ms_sharded = 1e3 * tot / a.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)

    ms_step = ms_total / a.steps
    ms_e2e = 1e3 * t_e2e / a.steps
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, ms_sharded or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e = float(t[0]), float(t[1])
        ms_sharded = float(t[2]) if ms_sharded is not None else None
        tl = torch.tensor([n_launch], device="cuda", dtype=torch.int64)
        dist.all_reduce(tl)
        n_launch = int(tl[0])
    units = 6 * B * world
    value = units / (ms_step * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)
    scaling, sharding = "weak", "independent unit cells per rank, no data-path collective"
    if ms_sharded is not None:
        # strong scaling: the six messages of ONE cell; the all-gather needs host-visible results, so value == e2e here
        units, scaling = 6, "strong"
        value = e2e_value = units / (ms_sharded * 1e-3)
        ms_step = ms_e2e = ms_sharded
        sharding = f"six block sides of one unit cell round-robin over {world} ranks, one NCCL all-gather of {gather_bytes} B per iteration"

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "unit_cells_per_gpu": B, "sharding": sharding,
                       "l2": "512 MiB buffer rewritten between timed iterations (L2 flush)",
                       "swallows_per_message": len(bp.contraction_order.kagome_order(N, "D", "ToMessage")) - 1},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e},
            "gpu_launches": int(n_launch), "clocks": sampler.summary()}

    if rank == 0:
        # ---- roofline of the dominant kernel family (instrumented repetition of the same step)
        for s in BLOCK_SIDES_CCW:
            engs[s].profile_enable(True)
        resident_step()
        ms = np.zeros(16)
        cnt = np.zeros(16, dtype=np.int64)
        for s in BLOCK_SIDES_CCW:
            m_, c_ = engs[s].profile_read()
            ms += m_
            cnt += c_
            engs[s].profile_enable(False)
        svd_flops = sum(4.0 * (14.0 * max(m, n) * min(m, n) ** 2 + 8.0 * min(m, n) ** 3) for s in BLOCK_SIDES_CCW
                        for (m, n, k) in comps[s].meta["svd_shapes"]) * B
        total_flops = sum(comps[s].flops for s in BLOCK_SIDES_CCW) * B
        # measured FP64 peak: cuBLAS DGEMM 4096^3 via torch (library call, used only as the denominator)
        x = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
        torch.matmul(x, x)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, x)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peak = 2 * 4096 ** 3 / (best * 1e-3) / 1e12
        svd_ms = float(ms[4])
        # the six streams overlap, so per-op event times are inflated by concurrency; the chip-level achieved
        # rate of the SVD family is its flops over the wall time of the step scaled by its share of op time
        share = svd_ms / max(1e-9, float(ms.sum()))
        achieved = svd_flops / (ms_step * 1e-3 * share) / 1e12
        line["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            # dram__bytes_read.sum + dram__bytes_write.sum per launch, from `ncu --set full` captures summarised under
                            # profiles/ (r01_ncu_full_chol_kernel_v2.csv for the current Cholesky kernel, r01_ncu_full_tsvd_kernels.csv
                            # for the others; cold L2 -- in the running step the operands are L2 resident: the family is not HBM bound)
                            "traffic": 479744, "traffic_detail_bytes_per_launch": {"chol_kernel": 479744, "svd_small_kernel (one CTA)": 212480,
                                                                                   "zgemm_dmma_kernel<32,32>": 780544},
                            "kernel": "truncated-SVD family: subspace iteration = zgemm_dmma_kernel + chol_kernel + trsm_kernel per iteration, "
                                      "svd_cluster_kernel / svd_small_kernel Rayleigh-Ritz (svd_round_kernel: exact fallback)",
                            "peak_source": "cuBLAS DGEMM 4096^3 FP64 measured in this run (MEASURED_PEAKS.json carries no FP64 figure)",
                            "algorithmic_flops_per_step": svd_flops, "svd_share_of_op_time": share,
                            "all_ops_algorithmic_flops_per_step": total_flops,
                            "all_ops_achieved_tflops": total_flops / (ms_step * 1e-3) / 1e12,
                            "op_time_ms": {k: float(ms[i]) for k, i in (("permute", 1), ("gemm", 2), ("qr", 3), ("svd", 4), ("normalize", 5), ("embed", 6), ("zero", 7), ("eye", 10))},
                            "op_counts": {k: int(cnt[i]) for k, i in (("permute", 1), ("gemm", 2), ("qr", 3), ("svd", 4))},
                            "jacobi_sweeps_total": sum(engs[s].svd_sweeps() for s in BLOCK_SIDES_CCW),
                            "svd_paths": {k: sum(engs[s].svd_counters()[k] for s in BLOCK_SIDES_CCW) for k in engs[BLOCK_SIDES_CCW[0]].svd_counters()}}
        if world == 1 and a.ensemble > 1 and B == 1:
            # BASELINE config C5: an ensemble of independent unit cells batched into every launch (same programs, nb chains)
            E = a.ensemble
            ecells = [UnitCell.random(2, D, seed=100 + i) for i in range(E)]
            emsgs = [msgs_list[0]] * E
            for s in BLOCK_SIDES_CCW:
                batch = [bp._side_inputs(c, m, comps[s]) for c, m in zip(ecells, emsgs)]
                comps[s].load(engs[s], E)
                engs[s].upload(0, comps[s].pack_inputs(batch))
                engs[s].sync()
            resident_step()
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            resident_step()
            resident_step()
            e1.record()
            torch.cuda.synchronize()
            ems = e0.elapsed_time(e1) / 2
            line["ensemble"] = {"unit_cells_per_gpu": E, "ms_per_step": ems, "value": 6 * E / (ems * 1e-3), "unit": UNIT,
                                "note": "independent unit cells as extra chains of the same launches (kernels take a chain index)"}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(a, cells[0], msgs_list[0], a.cpu_budget_s)
        if world == 1 and a.ite_steps > 0:
            line["ite"] = ite_metric(a, cells[0], msgs_list[0], cfg, line.get("cpu_baseline"))
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
