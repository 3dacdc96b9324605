import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("6d-pose-estimation_b200")
core, W = pkg.core, pkg.workloads
dev = torch.device("cuda", 0)
pts, dia = W.config2_meshes(2048)
table = core.MeshTable(pts, dia, pkg.SYMMETRIC_OBJECT_IDS, dev).set_pruning(True)
host = W.config2(65536, base_seed=2000)
d_in = [torch.from_numpy(x).to(dev) for x in host]
order = torch.argsort(d_in[4], stable=True).to(torch.int32)
for _ in range(3):
    table.evaluate_packed(*d_in, want_adds=True, order=order)
torch.cuda.synchronize()
print("done")
