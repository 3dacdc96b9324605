#!/usr/bin/env python
"""One launch of kernel (a) at N = 1000 / 1 M poses and of the streamed pose-loss kernel at 4 M rows (for ncu)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
W, core = pkg.workloads, pkg.core
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
m = 1 << 20
pts = {0: W.sphere_mesh(1000, 0.102, 100)}
table = core.MeshTable(pts, {0: 0.102}, pkg.SYMMETRIC_OBJECT_IDS, dev)
obj = torch.zeros(m, dtype=torch.int64, device=dev)
qa, ta = torch.nn.functional.normalize(rnd(m, 4), dim=1), rnd(m, 3)
qb, tb = torch.nn.functional.normalize(qa + 0.05 * rnd(m, 4), dim=1), ta + 0.005 * rnd(m, 3)
for _ in range(2):
    table.evaluate(qb, tb, qa, ta, obj, want_adds=False)
n = 1 << 22
pq, gq, pt, gt = rnd(n, 4), rnd(n, 4), rnd(n, 3), rnd(n, 3)
x = pq.requires_grad_(True); y = pt.requires_grad_(True)
for _ in range(2):
    pkg.PoseLoss(1.0, 10.0, "geodesic")(x, y, gq, gt).backward()
torch.cuda.synchronize()
print("done")
