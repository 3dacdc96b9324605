import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("6d-pose-estimation_b200")
core, W = pkg.core, pkg.workloads
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
Bn = 1 << 18
pts = {9: W.box_mesh(500, (0.1, 0.12, 0.05), 700)}
t = core.MeshTable(pts, {9: 0.1646}, pkg.SYMMETRIC_OBJECT_IDS, dev)
ob = torch.full((Bn,), 9, dtype=torch.int64, device=dev)
qa, ta = torch.nn.functional.normalize(rnd(Bn, 4), dim=1), rnd(Bn, 3)
qb, tb = torch.nn.functional.normalize(qa + 0.05 * rnd(Bn, 4), dim=1), ta + 0.005 * rnd(Bn, 3)
for _ in range(3):
    t.evaluate(qb, tb, qa, ta, ob, want_adds=True, prune=bool(int(os.environ.get("PRUNE", "0"))))
torch.cuda.synchronize()
print("done")
