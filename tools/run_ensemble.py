"""Ensemble / sweep launcher (BASELINE config C5): `python tools/run_ensemble.py --cells 64 --D 4 --N 2 --out data/condor/results_ite_afm.csv`,
or under torchrun (one rank per GPU) to shard the seeds over the GPUs.  Replaces scripts/condor/main_sender.py + worker.py of the
reference: same CSV columns, no cluster scheduler."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=8)
ap.add_argument("--first-seed", type=int, default=0)
ap.add_argument("--D", type=int, default=2)
ap.add_argument("--N", type=int, default=2)
ap.add_argument("--chi", type=float, default=1.0)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--ite-steps", type=int, default=0)
ap.add_argument("--out", default="data/condor/results_ite_afm.csv")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from kagomeperiodicbp_b200 import ensemble
rows = ensemble.run_ensemble(range(a.first_seed, a.first_seed + a.cells), a.D, a.N, a.chi, rank, world, device=local, batch=a.batch, ite_steps=a.ite_steps)
rows = ensemble.gather_rows(rows, world)
if rank == 0:
    print(ensemble.write_csv(rows, a.out), len(rows), "rows; mean energy per site", sum(r["energy"] for r in rows) / len(rows))
if world > 1:
    dist.destroy_process_group()
