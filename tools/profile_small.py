#!/usr/bin/env python
"""One launch of every secondary kernel at the sizes quoted in DESIGN.md (for ncu):
   ncu --set full --clock-control none -k regex:'add_pose|pose_loss|pinhole|depth|detection|add_backward|quat|tf32|synth|adds_cta|adds_pruned' -o gpurun_out/secondary python tools/profile_small.py"""
import importlib, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("6d-pose-estimation_b200")
W, core = pkg.workloads, pkg.core
dev = torch.device("cuda", 0)
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
# (a) ADD only: 1 M poses, N = 1000
B = 1 << 20
pts = {0: W.sphere_mesh(1000, 0.102, 100)}
table = core.MeshTable(pts, {0: 0.102}, pkg.SYMMETRIC_OBJECT_IDS, dev)
pq, pt, gq, gt = (T(x) for x in W.random_poses(B, 9))
obj = torch.zeros(B, dtype=torch.int64, device=dev)
table.evaluate(pq, pt, gq, gt, obj, want_adds=False)
# (c) PoseLoss fwd+bwd: B = 32 (config 3) and B = 4 M
for n in (32, 1 << 22):
    a, b, c, d = (T(x) for x in W.random_poses(n, 3, rot_sigma=0.2, trans_sigma=0.02))
    x = a.requires_grad_(True); y = b.requires_grad_(True)
    pkg.PoseLoss(1.0, 10.0, "geodesic")(x, y, c, d).backward()
# (d1) pinhole fwd/bwd 4 M rows, batched K; fused geometric step at B = 32
n = 1 << 22
z = (torch.rand(n, 1, device=dev) + 0.4).requires_grad_(True); uv = torch.rand(n, 2, device=dev) * 400
K = torch.tensor(pkg.DEFAULT_K, dtype=torch.float32, device=dev).expand(n, 3, 3).contiguous()
pkg.pinhole_translation(z, uv, K).sum().backward()
c3 = W.config3(32, 3)
l, _ = pkg.PoseLoss(1.0, 10.0).forward_geometric(T(c3["rot_raw"]).requires_grad_(True), T(c3["z_pred"]).requires_grad_(True),
                                                 T(c3["bbox_center"]), T(c3["K"]), T(c3["gt_rot"]), T(c3["gt_trans"]))
l.backward()
# (d2) depth back-projection: config 4 API level (256 x 224 x 224) and 1 M small crops; N1 fused crop kernel
depth, uvc, Kc = (T(x) for x in W.config4(256, 4))
pkg.depth_backproject(depth, uvc, Kc)
n = 1 << 20
pkg.depth_backproject(torch.rand(n, 8, 8, device=dev) * 1.5, torch.rand(n, 2, device=dev) * 8, Kc[:1].expand(n, 3, 3).contiguous(), clamp_hi=7.0)
frame, boxes = W.config4_frame(40, 256)
pkg.depth_crop_backproject(T(frame), T(boxes), T(pkg.DEFAULT_K.astype(np.float32)))
boxes_big = np.tile(boxes, (4096, 1))
pkg.depth_crop_backproject(T(frame), T(boxes_big), T(pkg.DEFAULT_K.astype(np.float32)))
# N3 backward of ADDLoss.forward: 32 poses, N = 500, mixed symmetric / asymmetric
ptsb, diab = W.sweep_meshes(500)
crit = pkg.ADDLoss(tempfile.mkdtemp(), dev)
for k, v in ptsb.items():
    crit.points[k] = T(v)
crit.diameters.update(diab)
a, b, c, d = (T(x) for x in W.random_poses(32, 5))
o = T(np.array(W.LINEMOD_IDS, np.int64)[np.arange(32) % 13])
x = a.requires_grad_(True); y = b.requires_grad_(True)
crit(x, y, c, d, o).backward()
crit._quat_to_mat(a)
# loss form: ADDLoss.forward value in one launch (p6d_add_forward), B = 32 and B = 4096
crit(a, b, c, d, o)
a2, b2, c2, d2 = (T(x) for x in W.random_poses(4096, 6))
crit(a2, b2, c2, d2, T(np.array(W.LINEMOD_IDS, np.int64)[np.arange(4096) % 13]))
# config-5 hypothesis generator (p6d_synth_poses), one chunk of each kind
for vi, var in enumerate(pkg.sweep.VARIANTS):
    pkg.sweep.synth_block(1 << 18, 0, 5000, 0, vi, var, 0, dev)
# the tcgen05 evidence kernel (opt-in, never on the product path): 2,048 poses of config 2
pts2, dia2 = W.config2_meshes(2048)
t2 = core.MeshTable(pts2, dia2, pkg.SYMMETRIC_OBJECT_IDS, dev)
c2 = [T(x) for x in W.config2(2048)]
out = torch.empty(2048, dtype=torch.float32, device=dev)
core.check(core.lib().p6d_adds_tf32_eval(t2.handle, *(core.ptr(x) for x in c2), 2048, 3, core.ptr(out), 0, core.stream_ptr(dev)))
# N1, inference form (xyxy detector boxes): 256 boxes and 1 M boxes of one frame
xyxy = boxes.copy(); xyxy[:, 2:] += xyxy[:, :2]
pkg.detection_backproject(T(frame), T(xyxy))
pkg.detection_backproject(T(frame), T(np.tile(xyxy, (4096, 1))))
# (b') the opt-in exact-pruned ADD-S kernel on 16,384 config-2 poses
c16 = [T(x) for x in W.config2(16384)]
t2.evaluate(*c16, prune=True)
torch.cuda.synchronize()
print("profile_small done")
