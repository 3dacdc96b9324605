#!/usr/bin/env python
"""Dev tool (GPU box): the same config-5 sweep several times in one process -- how often is a run slow?
  python tools/sweep_repeat.py [runs] [n_points] [n_per_block]"""
import importlib, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n_per_block = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
dev = torch.device("cuda", 0)
pts, dia = pkg.workloads.sweep_meshes(n_points)
ev = pkg.PoseEvaluator(pts, dia, dev, n_rows=len(pkg.sweep.VARIANTS))
pkg.evaluate_sweep(pts, dia, dev, 4096, evaluator=ev)
torch.cuda.synchronize()
secs, hits = [], set()
for _ in range(runs):
    ev.acc.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc, _ = pkg.evaluate_sweep(pts, dia, dev, n_per_block, evaluator=ev)[:2]
    torch.cuda.synchronize()
    secs.append(round(time.perf_counter() - t0, 4))
    hits.add(int(acc.hits.sum().item()))
print(json.dumps({"n_points": n_points, "n_per_block": n_per_block, "seconds": secs, "hit_totals": sorted(hits)}))
