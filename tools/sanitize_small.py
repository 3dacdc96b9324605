#!/usr/bin/env python
"""Every kernel once at small shapes -- run under compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
W = pkg.workloads
dev = torch.device("cuda", 0)
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
pts = {0: W.sphere_mesh(37, 0.1, 1), 3: W.sphere_mesh(300, 0.2, 2), 9: W.box_mesh(129, (0.1, 0.12, 0.05), 3), 10: W.box_mesh(5, (0.04, 0.17, 0.04), 4)}
dia = {0: 0.1, 3: 0.2, 9: 0.16, 10: 0.17}
crit = pkg.ADDLoss(tempfile.mkdtemp(), dev)
for k, v in pts.items():
    crit.points[k] = T(v)
crit.diameters.update(dia)
pq, pt, gq, gt = W.random_poses(300, 5)
obj = np.array([0, 3, 9, 10, 7], np.int64)[np.arange(300) % 5]
a = [T(x) for x in (pq, pt, gq, gt, obj)]
print(crit.eval_metrics(*a))                                   # adds kernel, sorted order, TMA re-staging
print(crit.eval_metrics(*(x[:7] for x in a)))                  # unsorted small batch
table = crit._mesh_table(dev)
print(table.evaluate(*a, want_adds=False)[0].sum().item())     # add_warp kernel
print(table.evaluate_host(pq, pt, gq, gt, obj)["obj_hits"])    # host entry
x = a[0].clone().requires_grad_(True); y = a[1].clone().requires_grad_(True)
crit(x, y, a[2], a[3], a[4]).backward(); print(x.grad.abs().sum().item())     # backward kernel
c = W.config3(32, 3)
r = T(c["rot_raw"]).requires_grad_(True); z = T(c["z_pred"]).requires_grad_(True)
tr = pkg.pinhole_translation(z, T(c["bbox_center"]), T(c["K"]))
for mode in ("geodesic", "l1"):
    l = pkg.PoseLoss(1.0, 10.0, mode)(r, tr, T(c["gt_rot"]), T(c["gt_trans"])); l.backward(retain_graph=True); print(l.item())
pq2, pt2, gq2, gt2 = (T(v) for v in W.random_poses(5000, 6))
print(pkg.PoseLoss()(pq2, pt2, gq2, gt2).item())               # multi-CTA loss kernel
depth, uv, K = W.config4(64, 7)
print(pkg.depth_backproject(T(depth), T(uv), T(K)).sum().item())
print(crit._quat_to_mat(a[0][:9]).sum().item())
torch.cuda.synchronize(); print("sanitize_small done")
