#!/usr/bin/env python
"""Latency at the reference batch and GB/s at a scaled batch for kernels (a), (c), (d1), (d2)
(SURVEY.md section 8d configs 3 and 4).  Prints one JSON object; run on a B200."""
import importlib, json, os, sys, time
import numpy as np, torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("6d-pose-estimation_b200")
W, core = pkg.workloads, pkg.core
dev = torch.device("cuda", 0)
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
HBM = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.min(ts)), float(np.median(ts))


def eager_pose_loss(pq, pt, gq, gt):
    """The reference's op chain (torch eager on the GPU) -- comparison only."""
    q1 = torch.nn.functional.normalize(pq, p=2, dim=1); q2 = torch.nn.functional.normalize(gq, p=2, dim=1)
    dot = torch.sum(q1 * q2, dim=1, keepdim=True); q2 = torch.where(dot < 0, -q2, q2)
    ang = 2 * torch.atan2(torch.norm(q1 - q2, dim=1), torch.norm(q1 + q2, dim=1))
    return 1.0 * ang.mean() + 10.0 * torch.nn.functional.l1_loss(pt, gt)


out = {"hbm_peak_gbs_measured": HBM}
# ---- (c) PoseLoss fwd+bwd
for B in (32, 1 << 20, 1 << 22):
    pq, pt, gq, gt = (T(x) for x in W.random_poses(B, 3, rot_sigma=0.2, trans_sigma=0.02))
    crit = pkg.PoseLoss(1.0, 10.0, "geodesic")
    def ours():
        a = pq.requires_grad_(True); b = pt.requires_grad_(True)
        a.grad = None; b.grad = None
        crit(a, b, gq, gt).backward()
    def ref():
        a = pq.detach().requires_grad_(True); b = pt.detach().requires_grad_(True)
        eager_pose_loss(a, b, gq, gt).backward()
    reps = 200 if B == 32 else 20
    mn, md = timed(ours, reps)
    rmn, rmd = timed(ref, reps)
    # kernel-only (no autograd wrapper): one C-ABI call
    o = torch.empty(3, device=dev); g1 = torch.empty_like(pq); g2 = torch.empty_like(pt)
    ws = torch.zeros(64, dtype=torch.uint8, device=dev)
    def kern():
        core.check(core.lib().p6d_pose_loss_fwd_bwd(pq.data_ptr(), pt.data_ptr(), gq.data_ptr(), gt.data_ptr(), B, 1.0,
                                                    10.0, 0, o.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(),
                                                    0, core.stream_ptr(dev)))
    kmn, kmd = timed(kern, reps)
    bytes_ = B * (56 + 28)
    out[f"pose_loss_B{B}"] = {"ours_autograd_us": [mn, md], "kernel_us": [kmn, kmd], "torch_eager_gpu_us": [rmn, rmd],
                              "algorithmic_bytes": bytes_, "kernel_gbs": bytes_ / (kmn * 1e-6) / 1e9,
                              "frac_of_measured_hbm": bytes_ / (kmn * 1e-6) / 1e9 / HBM}
# ---- (d1) pinhole
for B in (32, 1 << 22):
    z = torch.rand(B, 1, device=dev) + 0.4; uv = torch.rand(B, 2, device=dev) * 400
    K = torch.tensor(pkg.DEFAULT_K, dtype=torch.float32, device=dev).expand(B, 3, 3).contiguous()
    mn, md = timed(lambda: pkg.pinhole_translation(z, uv, K), 100 if B == 32 else 20)
    bytes_ = B * (4 + 8 + 16 + 12)          # 4 of the 9 K entries are read (sectors: 36)
    out[f"pinhole_B{B}"] = {"us": [mn, md], "algorithmic_bytes": bytes_, "gbs": bytes_ / (mn * 1e-6) / 1e9,
                            "frac_of_measured_hbm": bytes_ / (mn * 1e-6) / 1e9 / HBM}
# ---- (d2) depth back-projection: 256 boxes at the API level, and a scaled batch of small crops
depth, uv, K = (T(x) for x in W.config4(256, 4))
mn, md = timed(lambda: pkg.depth_backproject(depth, uv, K), 100)
out["depth_backproject_B256_224x224"] = {"us": [mn, md], "algorithmic_bytes": 256 * 60}
B = 1 << 20
d8 = torch.rand(B, 8, 8, device=dev) * 1.5; uv8 = torch.rand(B, 2, device=dev) * 8
K8 = K[:1].expand(B, 3, 3).contiguous()
mn, md = timed(lambda: pkg.depth_backproject(d8, uv8, K8, clamp_hi=7.0), 20)
bytes_ = B * (32 + 8 + 16 + 12)             # one 32-byte sector of depth per row
out[f"depth_backproject_B{B}_8x8"] = {"us": [mn, md], "algorithmic_bytes": bytes_, "gbs": bytes_ / (mn * 1e-6) / 1e9,
                                      "frac_of_measured_hbm": bytes_ / (mn * 1e-6) / 1e9 / HBM}
# ---- (a) ADD-only kernel: config 1 (32 poses) latency and 1M poses throughput at N = 500 / 1000
for N, B in ((1000, 32), (1000, 1 << 20), (500, 1 << 20)):
    pts = {0: W.sphere_mesh(N, 0.102, 100)}
    table = core.MeshTable(pts, {0: 0.102}, pkg.SYMMETRIC_OBJECT_IDS, dev)
    pq, pt, gq, gt = (T(x) for x in W.random_poses(B, 9))
    obj = torch.zeros(B, dtype=torch.int64, device=dev)
    mn, md = timed(lambda: table.evaluate(pq, pt, gq, gt, obj, want_adds=False), 100 if B == 32 else 10)
    out[f"add_only_N{N}_B{B}"] = {"us": [mn, md], "poses_per_s": B / (mn * 1e-6), "gflops_46N": 46 * N * B / (mn * 1e-6) / 1e9,
                                  "hbm_gbs": B * 73 / (mn * 1e-6) / 1e9}
# ---- eval_metrics end to end at the reference batch sizes (one launch + one D2H)
pts, dia, (pq, pt, gq, gt, obj) = W.config1(seed=1, mixed=True)
crit = pkg.ADDLoss(os.path.join(REPO, "tests", "golden"), dev)
for k, v in pts.items():
    crit.points[k] = T(v)
crit.diameters.update(dia)
args = [T(x) for x in (pq, pt, gq, gt, obj)]
crit.eval_metrics(*args)
t0 = time.perf_counter()
for _ in range(200):
    crit.eval_metrics(*args)
out["eval_metrics_cfg1_32poses_N1000_wall_us"] = (time.perf_counter() - t0) / 200 * 1e6
print(json.dumps(out, indent=1))
