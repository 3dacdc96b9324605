"""Seeded synthetic inputs for the BASELINE.json configs.

NumPy only, no device code and no oracle: the same generators feed the golden-vector
script, the parity tests and bench.py, so every side sees identical bytes.
There is no dataset in the image (LineMOD is not shipped with the reference,
/root/reference/.gitignore:6); meshes are point clouds on simple surfaces with the public
LineMOD diameters, poses are drawn the way SURVEY.md section 8d specifies.
"""
from __future__ import annotations

import numpy as np

# 0-based LineMOD object ids used by the reference (scripts/inference/inference_rgb.py:28-31)
LINEMOD_IDS = (0, 1, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14)
# models_info.yml diameters in metres (public dataset metadata; only has to be identical
# on both sides of a parity check)
LINEMOD_DIAMETERS = {
    0: 0.10209865663, 1: 0.24750624233, 3: 0.17249224865, 4: 0.20140358597,
    5: 0.15454551808, 7: 0.26166403443, 8: 0.10899920102, 9: 0.16462758848,
    10: 0.17588933422, 11: 0.14554287471, 12: 0.27807811733, 13: 0.28260129399,
    14: 0.21235825148,
}
DEFAULT_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]])


def sphere_mesh(n: int, diameter: float, seed: int) -> np.ndarray:
    """n float32 points uniform on a sphere of the given diameter, centred at 0."""
    r = np.random.RandomState(seed)
    v = r.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return (v * (diameter / 2.0)).astype(np.float32)


def box_mesh(n: int, extent, seed: int) -> np.ndarray:
    """n float32 points uniform on the surface of an axis-aligned box (eggbox/glue-like)."""
    r = np.random.RandomState(seed)
    e = np.asarray(extent, np.float64)
    p = (r.rand(n, 3) - 0.5) * e
    face = r.randint(0, 3, n)
    sign = r.randint(0, 2, n) * 2 - 1
    p[np.arange(n), face] = sign * e[face] / 2.0
    return p.astype(np.float32)


def random_poses(batch: int, seed: int, rot_sigma=0.05, trans_sigma=0.005):
    """GT quaternion/translation and a perturbed prediction (float32, scalar-last quats).

    gt_q = normalize(randn); pred_q = normalize(gt_q + rot_sigma*randn);
    gt_t = (U(-.2,.2), U(-.2,.2), U(.4,1.2)); pred_t = gt_t + trans_sigma*randn.
    rot_sigma may be an array [batch] to mix difficulty levels.
    """
    r = np.random.RandomState(seed)
    gq = r.standard_normal((batch, 4))
    gq /= np.linalg.norm(gq, axis=1, keepdims=True)
    sig = np.broadcast_to(np.asarray(rot_sigma, np.float64).reshape(-1, 1), (batch, 1))
    pq = gq + sig * r.standard_normal((batch, 4))
    pq /= np.linalg.norm(pq, axis=1, keepdims=True)
    gt = np.stack([r.uniform(-0.2, 0.2, batch), r.uniform(-0.2, 0.2, batch),
                   r.uniform(0.4, 1.2, batch)], 1)
    pt = gt + trans_sigma * r.standard_normal((batch, 3))
    f = np.float32
    return pq.astype(f), pt.astype(f), gq.astype(f), gt.astype(f)


def config1(seed: int = 1, mixed: bool = False):
    """BASELINE config 1: 32 poses, 1,000-point ape-sized mesh (ids 0, and 9 when mixed)."""
    pts = {0: sphere_mesh(1000, LINEMOD_DIAMETERS[0], 100)}
    dia = {0: LINEMOD_DIAMETERS[0]}
    pq, pt, gq, gt = random_poses(32, seed)
    obj = np.zeros(32, np.int64)
    if mixed:
        pts[9] = box_mesh(1000, (0.10, 0.12, 0.05), 109)
        dia[9] = LINEMOD_DIAMETERS[9]
        obj[1::2] = 9
        obj[5] = 6  # id absent from points -> skipped (add_loss.py:171-172)
    return pts, dia, (pq, pt, gq, gt, obj)


def config2_meshes(n_points: int = 2048):
    """BASELINE config 2: two symmetric-object meshes (eggbox 9, glue 10)."""
    pts = {9: box_mesh(n_points, (0.10, 0.12, 0.05), 209),
           10: box_mesh(n_points, (0.04, 0.17, 0.04), 210)}
    dia = {9: LINEMOD_DIAMETERS[9], 10: LINEMOD_DIAMETERS[10]}
    return pts, dia


def config2_chunk(chunk: int, chunk_size: int = 4096, base_seed: int = 2000):
    """Poses of one seeded chunk of config 2 (any chunk can be regenerated on its own).
    Rotation noise sweeps 0.01..0.2 so nearest-neighbour distances span ~0.5-20 mm;
    object ids alternate 9, 10."""
    sig = np.geomspace(0.01, 0.2, chunk_size)
    pq, pt, gq, gt = random_poses(chunk_size, base_seed + chunk, rot_sigma=sig)
    obj = np.where(np.arange(chunk_size) % 2 == 0, 9, 10).astype(np.int64)
    return pq, pt, gq, gt, obj


def config2(n_poses: int = 65536, chunk_size: int = 4096, base_seed: int = 2000):
    chunks = [config2_chunk(c, chunk_size, base_seed) for c in range((n_poses + chunk_size - 1) // chunk_size)]
    out = [np.concatenate([c[i] for c in chunks], 0)[:n_poses] for i in range(5)]
    return tuple(out)


def config3(batch: int = 32, seed: int = 3, edge_rows: bool = True):
    """BASELINE config 3: RGB-Geometric head outputs for PoseLoss + pinhole.
    Returns dict of float32 arrays: rot_raw (unnormalised head output), z_pred [B,1],
    bbox_center [B,2] (full-image px), K [B,3,3], gt_rot, gt_trans."""
    pq, pt, gq, gt = random_poses(batch, seed, rot_sigma=0.2, trans_sigma=0.02)
    rot_raw = (3.0 * pq).astype(np.float32)
    if edge_rows and batch >= 8:
        rot_raw[0] = gq[0]                    # identical quaternions -> zero rot grad
        rot_raw[1] = -gq[1]                   # antipodal
        rot_raw[2] = np.array([1, 0, 0, 0]); gq[2] = np.array([0, 1, 0, 0])  # dot == 0
        rot_raw[3] = 0.0                      # zero quaternion
        pt[4] = gt[4]                         # sign(0) = 0 in the L1 term
    K = np.broadcast_to(DEFAULT_K.astype(np.float32), (batch, 3, 3)).copy()
    uv = np.stack([gt[:, 0] / gt[:, 2] * K[:, 0, 0] + K[:, 0, 2],
                   gt[:, 1] / gt[:, 2] * K[:, 1, 1] + K[:, 1, 2]], 1).astype(np.float32)
    return {"rot_raw": rot_raw, "z_pred": pt[:, 2:3].copy(), "bbox_center": uv, "K": K,
            "gt_rot": gq, "gt_trans": gt, "pred_trans_direct": pt}


def config4(batch: int = 256, seed: int = 4, hw=(224, 224), clamp_hi: int = 223):
    """BASELINE config 4 at the reference API: depth crops [B,224,224] in metres with 10 %
    zeros and a few NaN / out-of-range pixels, crop-space centres partly outside the
    crop, per-row K_crop."""
    r = np.random.RandomState(seed)
    H, W = hw
    depth = r.uniform(0.0, 1.5, (batch, H, W)).astype(np.float32)
    depth[r.rand(batch, H, W) < 0.10] = 0.0
    uv = r.uniform(-3.0, 227.0, (batch, 2)).astype(np.float32)
    uv[0] = (112.0, 112.0)
    uv[1] = (223.0, 0.0)
    uv[2] = (222.99998, 223.5)
    uv[3] = (-0.5, 1e6)
    scale = r.uniform(0.8, 3.0, batch)
    K = np.zeros((batch, 3, 3), np.float32)
    K[:, 0, 0] = DEFAULT_K[0, 0] * scale
    K[:, 1, 1] = DEFAULT_K[1, 1] * scale
    K[:, 0, 2] = r.uniform(60, 160, batch)
    K[:, 1, 2] = r.uniform(60, 160, batch)
    K[:, 2, 2] = 1.0
    # force special depth values at the sampled pixel of a few rows
    def px(b):
        u = int(np.clip(uv[b, 0], 0, clamp_hi)); v = int(np.clip(uv[b, 1], 0, clamp_hi))
        return v, u
    for b, val in ((4, np.nan), (5, 0.0), (6, 0.01), (7, 0.010001), (8, 0.05), (9, 2.5), (10, -1.0),
                   (11, np.inf), (12, 0.1), (13, 2.0)):
        v, u = px(b)
        depth[b, v, u] = val
    return depth, uv, K


def sweep_meshes(n_points: int = 500):
    """Config 5 meshes: the 13 LineMOD ids with n_points each."""
    pts, dia = {}, {}
    for oid in LINEMOD_IDS:
        d = LINEMOD_DIAMETERS[oid]
        if oid in (9, 10):
            pts[oid] = box_mesh(n_points, (0.6 * d, 0.7 * d, 0.3 * d), 500 + oid)
        else:
            pts[oid] = sphere_mesh(n_points, d, 500 + oid)
        dia[oid] = d
    return pts, dia


def config4_frame(seed: int = 40, n_boxes: int = 256, hw=(480, 640)):
    """BASELINE config 4 at frame level: one uint16 depth frame in millimetres
    (U(400,1500), 10 % zeros) and `n_boxes` integer (x, y, w, h) boxes, w,h in [40,200],
    a third of them pushed against the image border so that the 1.2x square crop of the
    reference (data/dataset_rgbd.py:122-145) needs zero padding."""
    r = np.random.RandomState(seed)
    H, W = hw
    depth = r.randint(400, 1500, (H, W)).astype(np.uint16)
    depth[r.rand(H, W) < 0.10] = 0
    w = r.randint(40, 201, n_boxes)
    h = r.randint(40, 201, n_boxes)
    x = (r.rand(n_boxes) * (W - w)).astype(np.int64)
    y = (r.rand(n_boxes) * (H - h)).astype(np.int64)
    edge = np.arange(n_boxes) % 3 == 0
    side = r.randint(0, 4, n_boxes)
    x = np.where(edge & (side == 0), 0, x); x = np.where(edge & (side == 1), W - w, x)
    y = np.where(edge & (side == 2), 0, y); y = np.where(edge & (side == 3), H - h, y)
    boxes = np.stack([x, y, w, h], 1).astype(np.int32)
    # a few hand-picked shapes: exact 2x / 1x resize factors, tiny and huge boxes
    boxes[0] = (100, 100, 373, 200)      # size 447.6 -> crop 447
    boxes[1] = (200, 150, 187, 100)      # size 224.4 -> crop 224 (identity resize)
    boxes[2] = (10, 10, 20, 30)          # crop 36: up-sampling x6.2
    boxes[3] = (0, 0, 640, 480)          # whole frame: crop 768, padded on all sides
    boxes[4] = (300, 200, 374, 100)      # size 448.8 -> crop 448 (exact 2x)
    return depth, boxes
