#!/usr/bin/env python
"""Dev tool (GPU box): where the ~120 us of an eager PoseLoss step and the ~125 us of an eval_metrics call go."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("6d-pose-estimation_b200")
core, W = pkg.core, pkg.workloads
dev = torch.device("cuda", 0)
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)

def per_call(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

c = W.config3(32, 6)
crit = pkg.PoseLoss(1.0, 10.0, "geodesic")
rot, tr = T(c["rot_raw"]).requires_grad_(True), T(c["gt_trans"] + 0.01).requires_grad_(True)
gr, gtr = T(c["gt_rot"]), T(c["gt_trans"])
L = core.lib()
out = torch.empty(3, device=dev); g1 = torch.empty(32, 4, device=dev); g2 = torch.empty(32, 3, device=dev)
ws = torch.zeros(64, dtype=torch.uint8, device=dev)
st = core.stream_ptr(dev)
res = {}
res["ctypes_call_only"] = per_call(lambda: L.p6d_pose_loss_fwd_bwd(rot.data_ptr(), tr.data_ptr(), gr.data_ptr(), gtr.data_ptr(), 32, 1.0, 10.0, 0,
                                                                    out.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), 0, st))
res["stream_ptr"] = per_call(lambda: core.stream_ptr(dev))
res["torch_empty"] = per_call(lambda: torch.empty(228, dtype=torch.float32, device=dev))
with torch.no_grad():
    res["forward_no_grad"] = per_call(lambda: crit(rot, tr, gr, gtr))
res["forward_with_grad"] = per_call(lambda: crit(rot, tr, gr, gtr))
def step():
    rot.grad = None; tr.grad = None
    crit(rot, tr, gr, gtr).backward()
res["forward_backward"] = per_call(step, 1000)
x = torch.ones(32, 4, device=dev, requires_grad=True)
def trivial():
    x.grad = None
    (x * 2.0).sum().backward()
res["torch_trivial_mul_sum_backward"] = per_call(trivial, 1000)
# eval_metrics, config 1
pts, dia, (pq, pt, gq, gt, obj) = W.config1(seed=5, mixed=True)
ec = pkg.ADDLoss(os.path.join(os.path.dirname(__file__), "..", "tests", "golden"), dev)
for k, v in pts.items(): ec.points[k] = T(v)
ec.diameters.update(dia)
d = [T(a) for a in (pq, pt, gq, gt, obj)]
res["eval_metrics_cfg1"] = per_call(lambda: ec.eval_metrics(*d), 500)
res["eval_poses_cfg1"] = per_call(lambda: ec.eval_poses(*d), 500)
tb = ec._mesh_table(dev)
res["mesh_table_key"] = per_call(lambda: ec._mesh_table(dev))
res["prepare"] = per_call(lambda: ec._prepare(*d))
res["table_evaluate"] = per_call(lambda: tb.evaluate(*d, True, None), 500)
def ev_and_copy():
    p = tb.evaluate(*d, True, None)[4]
    return p.cpu()
res["table_evaluate_plus_d2h"] = per_call(ev_and_copy, 500)
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
