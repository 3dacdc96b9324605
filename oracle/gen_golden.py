#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (test infrastructure).

Run in the build container only:   python oracle/gen_golden.py
It imports the reference's own modules read-only from /root/reference (CPU eager
PyTorch), feeds them the seeded synthetic inputs of ``workloads.py`` and stores inputs
and outputs.  The GPU box has no /root/reference, so the vectors are committed; the
tests never import the reference.

Per-pose values come from ``ADDLoss.eval_metrics`` at batch size 1:
``add_mean/1000`` rounded back to float32 is exactly the float32 ``.item()`` the
reference appended (the float64 round trip through *1000 is far below float32 spacing).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("P6D_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")


def _load_workloads():
    p = os.path.join(REPO, "6d-pose-estimation_b200", "workloads.py")
    spec = importlib.util.spec_from_file_location("p6d_workloads", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


W = _load_workloads()
sys.path.insert(0, REF)
from models.add_loss import ADDLoss  # noqa: E402  (the reference)
from models.pose_loss import PoseLoss  # noqa: E402
from models.pose_net_rgb_geometric import PoseNetRGBGeometric  # noqa: E402
from models.pose_net_rgbd_geometric import PoseNetRGBDGeometric  # noqa: E402
from utils.camera import DEFAULT_K  # noqa: E402

T = torch.from_numpy


def make_crit(points, diameters):
    crit = ADDLoss(tempfile.mkdtemp(), "cpu")  # empty dir -> no meshes loaded
    for k, v in points.items():
        crit.points[k] = T(np.ascontiguousarray(v))
    for k, v in diameters.items():
        crit.diameters[k] = v
    return crit


def per_pose(crit, pq, pt, gq, gt, obj):
    B = len(obj)
    add = np.zeros(B, np.float32); adds = np.zeros(B, np.float32)
    hit = np.zeros(B, np.uint8); valid = np.zeros(B, np.uint8)
    for i in range(B):
        m = crit.eval_metrics(T(pq[i:i + 1]), T(pt[i:i + 1]), T(gq[i:i + 1]), T(gt[i:i + 1]), T(obj[i:i + 1]))
        if int(obj[i]) in crit.points:
            valid[i] = 1
            add[i] = np.float32(m["add_mean"] / 1000.0)
            adds[i] = np.float32(m["add_s_mean"] / 1000.0)
            hit[i] = 1 if m["add_01d_acc"] == 100.0 else 0
    return add, adds, hit, valid


def pack_points(points):
    ids = np.array(sorted(points), np.int64)
    return {"mesh_ids": ids, **{f"mesh_{i}": points[i] for i in ids}}


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def eval_case(name, points, diameters, poses, batch_dict=True):
    crit = make_crit(points, diameters)
    pq, pt, gq, gt, obj = poses
    add, adds, hit, valid = per_pose(crit, pq, pt, gq, gt, obj)
    extra = {}
    if batch_dict:
        m = crit.eval_metrics(T(pq), T(pt), T(gq), T(gt), T(obj))
        extra = {"agg": np.array([m["add_mean"], m["add_s_mean"], m["add_01d_acc"]], np.float64)}
    ids = np.array(sorted(diameters), np.int64)
    save(name, pq=pq, pt=pt, gq=gq, gt=gt, obj=obj, add=add, adds=adds, hit=hit, valid=valid,
         dia_ids=ids, dia=np.array([diameters[i] for i in ids], np.float64), **pack_points(points), **extra)


def gen_eval():
    # config 1 and its mixed/symmetric/unknown-id variant
    p, d, poses = W.config1(seed=1, mixed=False)
    eval_case("eval_cfg1", p, d, poses)
    p, d, poses = W.config1(seed=11, mixed=True)
    eval_case("eval_cfg1_mixed", p, d, poses)
    # config 2 subset: 2,048-point meshes, first 48 poses of chunk 0 and 16 of chunk 7
    p, d = W.config2_meshes(2048)
    c0, c7 = W.config2_chunk(0), W.config2_chunk(7)
    sel0 = np.linspace(0, 4095, 48).astype(int); sel7 = np.linspace(1, 4095, 16).astype(int)
    poses = tuple(np.concatenate([c0[i][sel0], c7[i][sel7]], 0) for i in range(5))
    eval_case("eval_cfg2_subset", p, d, poses)
    # ragged mesh sizes: every code path of the summation order (tails, <8, cascade >=512)
    sizes = [1, 2, 3, 7, 8, 9, 15, 31, 32, 33, 63, 100, 127, 255, 500, 513, 640]
    pts, dia = {}, {}
    for k, n in enumerate(sizes):
        pts[k] = W.sphere_mesh(n, 0.12, 300 + k); dia[k] = 0.12
    pts[9] = W.box_mesh(77, (0.1, 0.12, 0.05), 399); dia[9] = 0.164627
    pq, pt, gq, gt = W.random_poses(3 * len(sizes), 31, rot_sigma=0.04, trans_sigma=0.004)
    obj = np.repeat(np.arange(len(sizes)), 3).astype(np.int64)
    eval_case("eval_ragged", pts, dia, (pq, pt, gq, gt, obj))
    # all skipped -> int zeros; missing diameter -> default 0.1
    pts = {2: W.sphere_mesh(64, 0.1, 41)}
    crit = make_crit(pts, {})
    pq, pt, gq, gt = W.random_poses(4, 42)
    m = crit.eval_metrics(T(pq), T(pt), T(gq), T(gt), T(np.array([5, 5, 7, 1], np.int64)))
    assert m == {"add_mean": 0, "add_s_mean": 0, "add_01d_acc": 0} and all(isinstance(v, int) for v in m.values())
    obj = np.array([2, 2, 5, 2], np.int64)
    m = crit.eval_metrics(T(pq), T(pt), T(gq), T(gt), T(obj))
    add, adds, hit, valid = per_pose(crit, pq, pt, gq, gt, obj)
    save("eval_default_diameter", pq=pq, pt=pt, gq=gq, gt=gt, obj=obj, add=add, adds=adds, hit=hit,
         valid=valid, agg=np.array([m["add_mean"], m["add_s_mean"], m["add_01d_acc"]], np.float64),
         **pack_points(pts))
    # non-unit / degenerate quaternions, NaN translation, large offsets
    pts = {0: W.sphere_mesh(96, 0.1, 51)}
    pq, pt, gq, gt = W.random_poses(8, 52)
    pq[0] *= 1.7; gq[1] *= 0.3; pq[2] = 0.0; pt[3, 1] = np.nan; gt[4] += 50.0; pt[4] += 50.0
    pq[5] = gq[5]; pt[5] = gt[5]  # exact match -> distances exactly 0
    eval_case("eval_degenerate", pts, {0: 0.1}, (pq, pt, gq, gt, np.zeros(8, np.int64)))


def gen_quat():
    r = np.random.RandomState(61)
    q = r.standard_normal((64, 4)).astype(np.float32)
    q[:32] /= np.linalg.norm(q[:32], axis=1, keepdims=True)
    crit = make_crit({}, {})
    save("quat_to_mat", q=q, R=crit._quat_to_mat(T(q)).numpy())


def gen_forward():
    """ADDLoss.forward value + autograd grads w.r.t. the prediction (a5 / N3)."""
    pts = {0: W.sphere_mesh(200, 0.10, 71), 4: W.sphere_mesh(150, 0.2, 72),
           9: W.box_mesh(180, (0.10, 0.12, 0.05), 73), 10: W.box_mesh(64, (0.04, 0.17, 0.04), 74)}
    dia = {k: W.LINEMOD_DIAMETERS[k] for k in pts}
    crit = make_crit(pts, dia)
    pq, pt, gq, gt = W.random_poses(24, 75, rot_sigma=0.08, trans_sigma=0.01)
    obj = np.array([0, 9, 4, 10, 0, 9, 6, 4] * 3, np.int64)
    a, b = T(pq).requires_grad_(True), T(pt).requires_grad_(True)
    loss = crit(a, b, T(gq), T(gt), T(obj))
    loss.backward()
    empty = crit(T(pq[:2]), T(pt[:2]), T(gq[:2]), T(gt[:2]), T(np.array([6, 6], np.int64)))
    assert empty.item() == 0.0 and empty.requires_grad
    save("add_forward", pq=pq, pt=pt, gq=gq, gt=gt, obj=obj, loss=np.float32(loss.item()),
         grad_q=a.grad.numpy(), grad_t=b.grad.numpy(), dia_ids=np.array(sorted(dia), np.int64),
         dia=np.array([dia[i] for i in sorted(dia)], np.float64), **pack_points(pts))


def gen_forward_more():
    """Further ADDLoss.forward values (no grads): the reference's default mesh size, tiny meshes
    (torch.matmul's naive kernel rounds differently from MKL for n <= 44) and a mixed batch."""
    def one(name, pts, dia, B, seed, ids):
        crit = make_crit(pts, dia)
        pq, pt, gq, gt = W.random_poses(B, seed, rot_sigma=np.geomspace(0.01, 0.2, B), trans_sigma=0.01)
        obj = np.asarray(ids, np.int64)[np.random.RandomState(seed + 1).randint(0, len(ids), B)]
        with torch.no_grad():
            loss = crit(T(pq), T(pt), T(gq), T(gt), T(obj))
        save(name, pq=pq, pt=pt, gq=gq, gt=gt, obj=obj, loss=np.float32(loss.item()),
             dia_ids=np.array(sorted(dia), np.int64), dia=np.array([dia[i] for i in sorted(dia)], np.float64),
             **pack_points(pts))
    pts, dia = W.sweep_meshes(500)
    one("add_forward_n500", pts, dia, 32, 81, W.LINEMOD_IDS)
    small = {0: W.sphere_mesh(1, 0.1, 82), 1: W.sphere_mesh(2, 0.1, 83), 3: W.sphere_mesh(10, 0.1, 84),
             4: W.sphere_mesh(11, 0.1, 85), 9: W.box_mesh(44, (0.1, 0.12, 0.05), 86), 10: W.box_mesh(33, (0.04, 0.17, 0.04), 87)}
    one("add_forward_small", small, {k: 0.1 for k in small}, 48, 88, sorted(small))
    mixed = {0: W.sphere_mesh(44, 0.1, 89), 1: W.sphere_mesh(45, 0.1, 90), 9: W.box_mesh(300, (0.1, 0.12, 0.05), 91),
             10: W.box_mesh(1000, (0.04, 0.17, 0.04), 92), 12: W.sphere_mesh(640, 0.2, 93)}
    one("add_forward_mixed", mixed, {k: 0.15 for k in mixed}, 40, 94, sorted(mixed) + [6, 7])


def gen_pose_loss():
    c = W.config3(32, 3)
    net = PoseNetRGBGeometric.__new__(PoseNetRGBGeometric)  # stateless method, no ResNet build
    out = {k: v for k, v in c.items()}
    for mode in ("geodesic", "l1"):
        for tag, B in (("b32", 32), ("b5", 5)):
            rot_raw = T(c["rot_raw"][:B].copy()).requires_grad_(True)
            z = T(c["z_pred"][:B].copy()).requires_grad_(True)
            # model-side normalisation stays in autograd (pose_net_rgb_geometric.py:75)
            rot = rot_raw / (torch.norm(rot_raw, dim=1, keepdim=True) + 1e-8)
            rot.retain_grad()
            trans = PoseNetRGBGeometric._compute_pinhole_translation(net, z, T(c["bbox_center"][:B]), T(c["K"][:B]))
            trans.retain_grad()
            crit = PoseLoss(1.0, 10.0, mode)
            loss = crit(rot, trans, T(c["gt_rot"][:B]), T(c["gt_trans"][:B]))
            loss.backward()
            rl = (crit._geodesic_distance if mode == "geodesic" else crit._quaternion_l1)(rot.detach(), T(c["gt_rot"][:B]))
            tl = torch.nn.functional.l1_loss(trans.detach(), T(c["gt_trans"][:B]))
            out.update({f"{mode}_{tag}_loss": np.float32(loss.item()), f"{mode}_{tag}_rot": np.float32(rl.item()),
                        f"{mode}_{tag}_trans": np.float32(tl.item()),
                        f"{mode}_{tag}_rot_in": rot.detach().numpy(), f"{mode}_{tag}_trans_in": trans.detach().numpy(),
                        f"{mode}_{tag}_grad_rot": rot.grad.numpy(), f"{mode}_{tag}_grad_trans": trans.grad.numpy(),
                        f"{mode}_{tag}_grad_z": z.grad.numpy(), f"{mode}_{tag}_grad_rot_raw": rot_raw.grad.numpy()})
    # other weights, direct translation (RGB / RGBD heads)
    crit = PoseLoss(0.5, 2.0, "geodesic")
    a = T(c["rot_raw"].copy()).requires_grad_(True); b = T(c["pred_trans_direct"].copy()).requires_grad_(True)
    loss = crit(a, b, T(c["gt_rot"]), T(c["gt_trans"]))
    loss.backward()
    out.update({"w_loss": np.float32(loss.item()), "w_grad_rot": a.grad.numpy(), "w_grad_trans": b.grad.numpy()})
    save("pose_loss_cfg3", **out)


def gen_pinhole():
    c = W.config3(32, 3)
    net = PoseNetRGBGeometric.__new__(PoseNetRGBGeometric)
    z = T(c["z_pred"].copy()).requires_grad_(True)
    o = PoseNetRGBGeometric._compute_pinhole_translation(net, z, T(c["bbox_center"]), T(c["K"]))
    r = np.random.RandomState(81)
    go = r.standard_normal((32, 3)).astype(np.float32)
    o.backward(T(go))
    K1 = DEFAULT_K.astype(np.float32)
    o1 = PoseNetRGBGeometric._compute_pinhole_translation(net, T(c["z_pred"]), T(c["bbox_center"]), T(K1))
    save("pinhole", z=c["z_pred"], uv=c["bbox_center"], K=c["K"], out=o.detach().numpy(), grad_out=go,
         grad_z=z.grad.numpy(), K_shared=K1, out_shared=o1.numpy(), default_K=DEFAULT_K)


def gen_depth():
    depth, uv, K = W.config4(256, 4)
    net = PoseNetRGBDGeometric.__new__(PoseNetRGBDGeometric)
    o = PoseNetRGBDGeometric._compute_pinhole_translation(net, T(depth), T(uv), T(K))
    o1 = PoseNetRGBDGeometric._compute_pinhole_translation(net, T(depth), T(uv), T(K[0]))
    # only the sampled pixel matters; store the full crops of 16 rows + the gathered pixel of all
    u = np.clip(np.clip(uv[:, 0], 0, 223).astype(np.int64), 0, 223)
    v = np.clip(np.clip(uv[:, 1], 0, 223).astype(np.int64), 0, 223)
    save("depth_backproject", uv=uv, K=K, out=o.numpy(), out_shared=o1.numpy(), depth2=depth[:2],
         sampled=depth[np.arange(256), v, u], seed=np.int64(4))


def gen_loader():
    """ADDLoss.__init__ on a synthetic model dir: ASCII PLY incl. face lines (ingested as
    bogus vertices, add_loss.py:90-97), models_info.yml, diameter fallbacks, 500-point cap."""
    import yaml
    r = np.random.RandomState(91)
    d = tempfile.mkdtemp()
    files = {}

    def ply(name, verts, faces=0, extra_cols=0):
        lines = ["ply", "format ascii 1.0", f"element vertex {len(verts)}", "property float x",
                 "property float y", "property float z"]
        lines += [f"property float c{i}" for i in range(extra_cols)]
        if faces:
            lines += [f"element face {faces}", "property list uchar int vertex_indices"]
        lines.append("end_header")
        for vtx in verts:
            lines.append(" ".join(f"{x:.6f}" for x in vtx) + "".join(" 0.5" for _ in range(extra_cols)))
        for _ in range(faces):
            i, j, k = r.randint(0, len(verts), 3)
            lines.append(f"3 {i} {j} {k}")
        text = "\n".join(lines) + "\n"
        with open(os.path.join(d, name), "w") as f:
            f.write(text)
        files[name] = text

    ply("obj_01.ply", r.uniform(-60, 60, (800, 3)), faces=40, extra_cols=3)   # >500 -> downsample
    ply("obj_02.ply", r.uniform(-90, 90, (300, 3)))                            # official diameter
    ply("obj_10.ply", np.vstack([r.uniform(-50, 50, (120, 3)), [[900, 0, 0]]]))  # outlier dropped, fallback diameter
    ply("obj_15.ply", r.uniform(-10, 10, (6, 3)))                              # <=10 pts -> 0.1
    ply("notes.ply", r.uniform(-1, 1, (3, 3)))                                 # malformed name skipped
    info = {1: {"diameter": 102.09865663, "min_x": -1.0}, 2: {"diameter": 247.50624233}, 9: {"diameter": 5.0}}
    with open(os.path.join(d, "models_info.yml"), "w") as f:
        yaml.safe_dump(info, f)
    files["models_info.yml"] = open(os.path.join(d, "models_info.yml")).read()
    np.random.seed(1234)
    crit = ADDLoss(d, "cpu")
    arrs = {"file_names": np.array(list(files)), "file_texts": np.array(list(files.values())),
            "ids": np.array(sorted(crit.points), np.int64),
            "dia": np.array([crit.diameters[i] for i in sorted(crit.points)], np.float64)}
    for i in crit.points:
        arrs[f"pts_{i}"] = crit.points[i].numpy()
    save("loader", **arrs)


def gen_crop():
    """N1: the reference's own crop pipeline (LineMODDatasetRGBD.__getitem__,
    data/dataset_rgbd.py:85-206, cv2 pad/crop/resize on uint16) followed by its depth
    back-projection, on one synthetic 480x640 frame with 256 boxes.  cv2's optimised
    (IPP) bilinear differs from its generic C++ path by +-1 LSB in ~0.1 % of pixels, so
    both are recorded."""
    import cv2, yaml
    from data.dataset_rgbd import LineMODDatasetRGBD
    depth, boxes = W.config4_frame(40, 256)
    root = tempfile.mkdtemp()
    d = os.path.join(root, "01")
    os.makedirs(os.path.join(d, "rgb")); os.makedirs(os.path.join(d, "depth"))
    cv2.imwrite(os.path.join(d, "rgb", "0000.png"), np.zeros((480, 640, 3), np.uint8))
    cv2.imwrite(os.path.join(d, "depth", "0000.png"), depth)
    K = [float(v) for v in DEFAULT_K.reshape(-1)]
    eye = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]
    yaml.safe_dump({0: [{"obj_id": 1, "obj_bb": [int(v) for v in b], "cam_R_m2c": eye, "cam_t_m2c": [0.0, 0.0, 800.0]}
                        for b in boxes]}, open(os.path.join(d, "gt.yml"), "w"))
    yaml.safe_dump({0: {"cam_K": K, "depth_scale": 1.0}}, open(os.path.join(d, "info.yml"), "w"))
    net = PoseNetRGBDGeometric.__new__(PoseNetRGBDGeometric)
    out = {"seed": np.int64(40), "boxes": boxes, "K": DEFAULT_K.astype(np.float32)}
    for tag, opt in (("generic", False), ("optimized", True)):
        cv2.setUseOptimized(opt)
        ds = LineMODDatasetRGBD(root, mode="train", augment_bbox=False)
        assert len(ds) == len(boxes)
        centers, Ks, raws, zs = [], [], [], []
        for i in range(len(ds)):
            _, _, depth_raw, _, _, _, c, Kc = ds[i]
            centers.append(c.numpy()); Ks.append(Kc.numpy()); raws.append(depth_raw)
        depth_raw = torch.stack(raws); c = T(np.stack(centers)); Kc = T(np.stack(Ks))
        xyz = PoseNetRGBDGeometric._compute_pinhole_translation(net, depth_raw, c, Kc).numpy()
        u = np.clip(np.clip(c.numpy()[:, 0], 0, 223).astype(np.int64), 0, 223)
        v = np.clip(np.clip(c.numpy()[:, 1], 0, 223).astype(np.int64), 0, 223)
        z_mm = np.rint(depth_raw.numpy()[np.arange(len(ds)), v, u] * 1000).astype(np.uint16)
        out.update({f"{tag}_xyz": xyz, f"{tag}_z_mm": z_mm, f"{tag}_center": c.numpy(), f"{tag}_Kcrop": Kc.numpy()})
    cv2.setUseOptimized(True)
    out["cv2_version"] = np.array(cv2.__version__)
    save("crop_backproject", **out)


def gen_crop_ipp():
    """N1, the case that matters: boxes whose centre pixel comes out DIFFERENT from cv2's default
    (IPP) and from its generic C++ bilinear.  Candidates are found with the NumPy restatement
    (oracle.crop_depth_backproject in both modes); the values stored are those of the reference's own
    pipeline (LineMODDatasetRGBD.__getitem__ + PoseNetRGBDGeometric._compute_pinhole_translation) run
    twice: with cv2 as the reference runs it (default: IPP on) and with cv2.ipp.setUseIPP(False)."""
    import cv2, yaml
    if REPO not in sys.path:
        sys.path.append(REPO)
    import oracle as O
    from data.dataset_rgbd import LineMODDatasetRGBD
    depth, _ = W.config4_frame(44, 8)
    r = np.random.RandomState(45)
    n_cand = 48000        # ~0.25 % of the boxes land on a pixel where the two arithmetics disagree
    w = r.randint(30, 420, n_cand); h = r.randint(30, 420, n_cand)
    x = (r.rand(n_cand) * 640 - 40).astype(np.int64); y = (r.rand(n_cand) * 480 - 40).astype(np.int64)
    cand = np.stack([x, y, w, h], 1).astype(np.int32)
    Kf = DEFAULT_K.astype(np.float32)
    a = O.crop_depth_backproject(depth, cand, Kf, bilinear="cv2")["z_mm"].astype(np.int64)
    b = O.crop_depth_backproject(depth, cand, Kf, bilinear="generic")["z_mm"].astype(np.int64)
    differ = np.nonzero(a != b)[0]
    same = np.nonzero(a == b)[0]
    assert len(differ) >= 64, len(differ)
    special = np.array([(300, 200, 374, 100),      # crop 448: exact 2x (generic path switches to INTER_AREA)
                        (200, 150, 187, 100),      # crop 224: identity
                        (10, 10, 20, 30),          # crop 36: x6.2 up-sampling
                        (0, 0, 640, 480),          # whole frame, padded on all sides
                        (-30, -20, 200, 180),      # starts outside the frame
                        (500, 380, 300, 260)], np.int32)
    boxes = np.concatenate([cand[differ[:64]], special, cand[same[:256 - 64 - len(special)]]]).astype(np.int32)
    root = tempfile.mkdtemp()
    d = os.path.join(root, "01")
    os.makedirs(os.path.join(d, "rgb")); os.makedirs(os.path.join(d, "depth"))
    cv2.imwrite(os.path.join(d, "rgb", "0000.png"), np.zeros((480, 640, 3), np.uint8))
    cv2.imwrite(os.path.join(d, "depth", "0000.png"), depth)
    K = [float(v) for v in DEFAULT_K.reshape(-1)]
    eye = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]
    yaml.safe_dump({0: [{"obj_id": 1, "obj_bb": [int(v) for v in bx], "cam_R_m2c": eye, "cam_t_m2c": [0.0, 0.0, 800.0]}
                        for bx in boxes]}, open(os.path.join(d, "gt.yml"), "w"))
    yaml.safe_dump({0: {"cam_K": K, "depth_scale": 1.0}}, open(os.path.join(d, "info.yml"), "w"))
    net = PoseNetRGBDGeometric.__new__(PoseNetRGBDGeometric)
    out = {"seed": np.int64(44), "boxes": boxes, "K": Kf}      # the frame is W.config4_frame(seed, 8)[0]
    for tag, ipp in (("cv2", True), ("generic", False)):
        cv2.setUseOptimized(True)
        cv2.ipp.setUseIPP(ipp)
        ds = LineMODDatasetRGBD(root, mode="train", augment_bbox=False)
        assert len(ds) == len(boxes)
        centers, Ks, raws = [], [], []
        for i in range(len(ds)):
            _, _, depth_raw, _, _, _, c, Kc = ds[i]
            centers.append(c.numpy()); Ks.append(Kc.numpy()); raws.append(depth_raw)
        depth_raw = torch.stack(raws); c = T(np.stack(centers)); Kc = T(np.stack(Ks))
        xyz = PoseNetRGBDGeometric._compute_pinhole_translation(net, depth_raw, c, Kc).numpy()
        u = np.clip(np.clip(c.numpy()[:, 0], 0, 223).astype(np.int64), 0, 223)
        v = np.clip(np.clip(c.numpy()[:, 1], 0, 223).astype(np.int64), 0, 223)
        z_mm = np.rint(depth_raw.numpy()[np.arange(len(ds)), v, u] * 1000).astype(np.uint16)
        out.update({f"{tag}_xyz": xyz, f"{tag}_z_mm": z_mm, f"{tag}_center": c.numpy(), f"{tag}_Kcrop": Kc.numpy()})
    cv2.ipp.setUseIPP(True)
    out["cv2_version"] = np.array(cv2.__version__)
    n_diff = int((out["cv2_z_mm"] != out["generic_z_mm"]).sum())
    assert n_diff >= 64, n_diff
    print(f"crop_backproject_ipp: {n_diff} of {len(boxes)} boxes differ between cv2's default and its generic bilinear")
    save("crop_backproject_ipp", **out)


def gen_crop_xyxy():
    """N1, inference form: the per-detection crop code of the reference's inference script
    (scripts/inference/inference_rgbd_geometric.py, the body of `for box in results[0].boxes:` up to the model
    call) EXECUTED FROM THE REFERENCE'S OWN SOURCE FILE -- the script is not importable (YOLO, weights,
    matplotlib), so its lines are read from /root/reference at generation time, dedented and exec'ed per box with
    a stand-in detection object; nothing of it is stored here -- followed by the reference's
    PoseNetRGBDGeometric._compute_pinhole_translation on what those lines produced."""
    import textwrap, types
    import cv2
    src = open(os.path.join(REF, "scripts", "inference", "inference_rgbd_geometric.py")).read().splitlines()
    lo = next(i for i, l in enumerate(src) if "map(int, box.xyxy[0])" in l)
    hi = next(i for i, l in enumerate(src) if "# Pose inference" in l)
    body = compile(textwrap.dedent("\n".join(src[lo:hi])), "inference_rgbd_geometric.py[per-detection]", "exec")
    depth, boxes = W.config4_frame(41, 256)
    xyxy = np.stack([boxes[:, 0], boxes[:, 1], boxes[:, 0] + boxes[:, 2], boxes[:, 1] + boxes[:, 3]], 1).astype(np.int32)
    xyxy[:6] = [(300, 200, 674, 300),      # crop 448 = 2 x 224, crosses the right border
                (200, 150, 387, 250),      # crop 224: identity
                (10, 10, 30, 40),          # crop 36: x6.2 up-sampling
                (-40, -30, 120, 90),       # negative corner: padding on the left and top
                (500, 380, 700, 520),      # crosses the bottom-right corner
                (100, 100, 101, 103)]      # crop of 3 pixels
    h_img, w_img = depth.shape
    net = PoseNetRGBDGeometric.__new__(PoseNetRGBDGeometric)
    centers, Ks, zs, xyzs = [], [], [], []
    for b in xyxy:
        ns = dict(box=types.SimpleNamespace(xyxy=[torch.tensor([float(v) for v in b])], cls=[0], conf=[0.9]),
                  CLASS_ID_TO_OBJ_NAME={0: "01"}, rgb_img=np.zeros((h_img, w_img, 3), np.uint8), depth_img=depth,
                  h_img=h_img, w_img=w_img, img_size=224, K=DEFAULT_K.copy(), cv2=cv2, np=np, torch=torch,
                  transform=lambda a: torch.zeros(3, 224, 224), device="cpu")
        exec(body, ns)
        xyz = PoseNetRGBDGeometric._compute_pinhole_translation(net, ns["input_depth_raw"], ns["bbox_center"], ns["cam_matrix"])
        c = ns["bbox_center"][0].numpy()
        u, v = (int(np.clip(np.clip(c[k], 0, 223).astype(np.int64), 0, 223)) for k in (0, 1))
        centers.append(c); Ks.append(ns["cam_matrix"][0].numpy()); xyzs.append(xyz[0].numpy())
        zs.append(ns["input_depth_raw"][0, v, u].numpy())
    save("crop_backproject_xyxy", seed=np.int64(41), boxes_xyxy=xyxy, K=DEFAULT_K.astype(np.float64), center=np.stack(centers),
         Kcrop=np.stack(Ks), z_m=np.array(zs, np.float32), xyz=np.stack(xyzs), cv2_version=np.array(cv2.__version__))


def gen_projection():
    """N4: utils/mesh_utils.load_mesh_corners (PLY -> 1/99 percentile box corners) and
    utils/visualization.project_points (scipy quaternion -> R, pinhole projection, int
    truncation) on synthetic inputs."""
    from utils.mesh_utils import load_mesh_corners
    from utils.visualization import project_points
    r = np.random.RandomState(111)
    d = tempfile.mkdtemp()
    verts = np.vstack([r.uniform(-80, 80, (400, 3)) * np.array([1.0, 0.6, 0.3]), [[500.0, 0, 0]]])
    text = "\n".join(["ply", "format ascii 1.0", f"element vertex {len(verts)}", "property float x", "property float y",
                      "property float z", "element face 2", "property list uchar int vertex_indices", "end_header"] +
                     [" ".join(f"{x:.5f}" for x in v) for v in verts] + ["3 0 1 2", "3 5 6 7"]) + "\n"
    open(os.path.join(d, "obj_07.ply"), "w").write(text)
    corners = load_mesh_corners(d, "07")
    assert load_mesh_corners(d, "08") is None
    B = 64
    pq, pt, gq, gt = W.random_poses(B, 112, rot_sigma=0.3, trans_sigma=0.05)
    gq64 = gq.astype(np.float64) * r.uniform(0.5, 2.0, (B, 1))        # scipy normalises
    gt64 = gt.astype(np.float64)
    gt64[0, 2] = -0.2                                                  # behind the camera -> z clipped to 0.001
    uv = np.stack([project_points(corners, gq64[b], gt64[b], DEFAULT_K) for b in range(B)])
    from scipy.spatial.transform import Rotation as R
    Rm = np.stack([R.from_quat(gq64[b]).as_matrix() for b in range(B)])
    uv_mat = np.stack([project_points(corners, Rm[b], gt64[b], DEFAULT_K) for b in range(B)])
    assert np.array_equal(uv, uv_mat)
    save("projection", ply_text=np.array(text), corners=corners, quat=gq64, trans=gt64, K=DEFAULT_K, uv=uv.astype(np.int64),
         Rmat=Rm)


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_quat(); gen_eval(); gen_forward(); gen_forward_more(); gen_pose_loss(); gen_pinhole(); gen_depth(); gen_loader(); gen_crop(); gen_crop_ipp(); gen_crop_xyxy(); gen_projection()
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by oracle/gen_golden.py from {REF}\n"
                f"torch {torch.__version__} cpu_capability {torch.backends.cpu.get_cpu_capability()} "
                f"numpy {np.__version__} threads {torch.get_num_threads()}\n")
