#!/usr/bin/env python
"""Broader randomized check of the oracle against the REAL reference (build container only;
test infrastructure).  python oracle/validate_against_reference.py  -> oracle/VALIDATION.txt

tests/golden/ pins ~300 poses; this run compares a few thousand more (per-pose ADD, ADD-S
bit patterns and decisions through ADDLoss.eval_metrics at batch size 1, PoseLoss values and
autograd gradients, both translation methods) and records the outcome."""
import importlib.util, os, sys, tempfile, time
import numpy as np, torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("P6D_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
import oracle as O
spec = importlib.util.spec_from_file_location("p6d_workloads", os.path.join(REPO, "6d-pose-estimation_b200", "workloads.py"))
W = importlib.util.module_from_spec(spec); spec.loader.exec_module(W)
sys.path.insert(0, REF)
from models.add_loss import ADDLoss
from models.pose_loss import PoseLoss
from models.pose_net_rgb_geometric import PoseNetRGBGeometric
from models.pose_net_rgbd_geometric import PoseNetRGBDGeometric
T = torch.from_numpy
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
lines, t0 = [], time.time()

# ---- eval_metrics per pose
tot = mism = 0
for si, (n, B) in enumerate([(1, 40), (6, 40), (10, 40), (11, 40), (17, 60), (100, 120), (333, 150), (500, 200), (1000, 120), (1536, 60), (2048, 60)]):
    pts = {0: W.sphere_mesh(n, 0.11, 4000 + si), 10: W.box_mesh(n, (0.05, 0.16, 0.04), 4100 + si)}
    dia = {0: 0.11, 10: 0.1759}
    crit = ADDLoss(tempfile.mkdtemp(), "cpu")
    for k, v in pts.items():
        crit.points[k] = T(v)
    crit.diameters.update(dia)
    r = np.random.RandomState(4200 + si)
    pq, pt, gq, gt = W.random_poses(B, 4300 + si, rot_sigma=np.exp(r.uniform(np.log(2e-3), np.log(0.4), B)), trans_sigma=0.01)
    pq[::9] *= 1.3
    obj = np.where(r.rand(B) < 0.5, 0, 10).astype(np.int64)
    add, adds, hit, valid = O.add_eval(O.MeshTable(pts, dia), pq, pt, gq, gt, obj, n_threads=8)
    for i in range(B):
        m = crit.eval_metrics(T(pq[i:i + 1]), T(pt[i:i + 1]), T(gq[i:i + 1]), T(gt[i:i + 1]), T(obj[i:i + 1]))
        ok = (bits(np.float32(m["add_mean"] / 1000.0)) == bits(add[i]) and bits(np.float32(m["add_s_mean"] / 1000.0)) == bits(adds[i])
              and (m["add_01d_acc"] == 100.0) == bool(hit[i]))
        tot += 1; mism += (not ok)
lines.append(f"eval_metrics per pose (11 mesh sizes 1..2048, symmetric and asymmetric ids): {tot} poses, {mism} mismatches (bit-exact ADD, ADD-S, decision)")

# ---- PoseLoss value + grads
worst_l = worst_g = 0.0; n_l = 0
for seed in range(40):
    B = [1, 2, 7, 32, 100][seed % 5]
    pq, pt, gq, gt = W.random_poses(B, 5000 + seed, rot_sigma=[1e-3, 0.05, 0.5, 2.0][seed % 4], trans_sigma=0.03)
    pq = (pq * np.random.RandomState(seed).uniform(0.2, 4.0, (B, 1))).astype(np.float32)
    for mode in ("geodesic", "l1"):
        a = T(pq.copy()).requires_grad_(True); b = T(pt.copy()).requires_grad_(True)
        loss = PoseLoss(1.0, 10.0, mode)(a, b, T(gq), T(gt)); loss.backward()
        o = O.pose_loss(pq, pt, gq, gt, 1.0, 10.0, mode)
        worst_l = max(worst_l, abs(float(o["loss"]) - loss.item()) / abs(loss.item()))
        sc = np.maximum(np.abs(a.grad.numpy()).max(1, keepdims=True), 1e-30)
        worst_g = max(worst_g, float((np.abs(o["grad_q"] - a.grad.numpy()) / sc).max()))
        assert np.array_equal(bits(o["grad_t"]), bits(b.grad.numpy()))
        n_l += 1
lines.append(f"PoseLoss fwd+bwd ({n_l} batches, B in 1..100, unnormalised inputs): worst relative loss error {worst_l:.2e}, worst gradient error / row max {worst_g:.2e}, translation gradients bit-exact")

# ---- pinhole / depth back-projection
net = PoseNetRGBGeometric.__new__(PoseNetRGBGeometric); netd = PoseNetRGBDGeometric.__new__(PoseNetRGBDGeometric)
bad = 0
for seed in range(20):
    c = W.config3(64, 6000 + seed)
    ref = PoseNetRGBGeometric._compute_pinhole_translation(net, T(c["z_pred"]), T(c["bbox_center"]), T(c["K"])).numpy()
    bad += int((bits(O.pinhole(c["z_pred"], c["bbox_center"], c["K"])[0]) != bits(ref)).sum())
    depth, uv, K = W.config4(64, 6100 + seed)
    ref = PoseNetRGBDGeometric._compute_pinhole_translation(netd, T(depth), T(uv), T(K)).numpy()
    bad += int((bits(O.depth_backproject(depth, uv, K)) != bits(ref)).sum())
lines.append(f"pinhole + depth back-projection (20 x 64 rows each): {bad} differing float32 values")
lines.append(f"torch {torch.__version__} ({torch.backends.cpu.get_cpu_capability()}), numpy {np.__version__}, {time.time() - t0:.0f} s")
open(os.path.join(REPO, "oracle", "VALIDATION.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
