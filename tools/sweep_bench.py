#!/usr/bin/env python
"""BASELINE config 5: compare_all_models-style sweep, 13 objects x 4 variants x n hypotheses,
hypothesis axis sharded across the ranks of one box, one NCCL all-reduce of the counts.
  python tools/sweep_bench.py [n_per_block] [n_points]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_bench.py ...
Prints one JSON line on rank 0."""
import importlib, json, os, sys, time
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("6d-pose-estimation_b200")

n_per_block = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 500
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
pts, dia = pkg.workloads.sweep_meshes(n_points)
pkg.evaluate_sweep(pts, dia, dev, 4096, rank=rank, world=world)          # warm-up (kernels, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
acc, launches, _ = pkg.evaluate_sweep(pts, dia, dev, n_per_block, rank=rank, world=world)
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
total = 13 * 4 * n_per_block
if rank == 0:
    tab = acc.table(pkg.sweep.VARIANTS)
    print(json.dumps({"workload": f"config 5: 13 objects x 4 variants x {n_per_block} hypotheses, {n_points}-point meshes",
                      "n_gpus": world, "seconds": float(dt.item()), "poses_per_s": total / float(dt.item()),
                      "launches_per_rank": launches, "valid_total": int(acc.valid.sum().item()),
                      "add_01d_acc_by_variant": {v: round(tab[v]["all"]["add_01d_acc"], 4) for v in pkg.sweep.VARIANTS},
                      "hits_total": int(acc.hits.sum().item()),
                      "note": "includes on-device generation of the synthetic hypotheses (p6d_synth_poses) and the translation kernels"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
