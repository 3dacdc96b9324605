#!/usr/bin/env python
"""Dev tool (GPU box): config-5 sweep time as a function of the chunk size of p6d_sweep_run."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("6d-pose-estimation_b200")
dev = torch.device("cuda", 0)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 500
per_block = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
pts, dia = pkg.workloads.sweep_meshes(npts)
ev = pkg.PoseEvaluator(pts, dia, dev, n_rows=4)
pkg.evaluate_sweep(pts, dia, dev, 8192, evaluator=ev)
for chunk in (65536, 262144, 1 << 20):
    ev.acc.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc, launches, _ = pkg.evaluate_sweep(pts, dia, dev, per_block, chunk=chunk, evaluator=ev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"n_points": npts, "chunk": chunk, "seconds": dt, "Mposes_s": 52 * per_block / dt / 1e6,
                      "launches": launches, "hits": int(acc.hits.sum())}), flush=True)
