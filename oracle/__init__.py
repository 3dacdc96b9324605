"""CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes/numpy front end of ``oracle/pose_oracle.c``, the plain-C restatement of the
reference's pose-geometry hot path (models/add_loss.py:156-215, models/pose_loss.py:19-61,
models/pose_net_rgb_geometric.py:93-109, models/pose_net_rgbd_geometric.py:56-85 of
SFR-Vision/6d-pose-estimation).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package, and only as the checker or the timed CPU
baseline.  Nothing under ``6d-pose-estimation_b200/`` imports it.

Parity status: pinned against the reference itself through ``tests/golden/`` (see
``oracle/gen_golden.py`` and ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpose_oracle.so")
_lib = None

SYMMETRIC_OBJECT_IDS = frozenset({9, 10})  # models/add_loss.py:10


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc, no reference sources)."""
    src = os.path.join(_HERE, "pose_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "CC=gcc"], check=True, capture_output=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, ip, dp, bp, lp = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                              C.POINTER(C.c_uint8), C.POINTER(C.c_int64))
        L.p6o_aten_sum_f32.restype = C.c_float
        L.p6o_aten_sum_f32.argtypes = [fp, C.c_int64]
        L.p6o_aten_mean_f32.restype = C.c_float
        L.p6o_aten_mean_f32.argtypes = [fp, C.c_int64]
        L.p6o_quat_to_mat.restype = None
        L.p6o_quat_to_mat.argtypes = [fp, fp]
        L.p6o_transform.restype = None
        L.p6o_transform.argtypes = [fp, C.c_int64, fp, fp, fp]
        L.p6o_add_eval.restype = C.c_int
        L.p6o_add_eval.argtypes = [fp, ip, ip, dp, bp, C.c_int, fp, fp, fp, fp, lp, C.c_int64,
                                   C.c_int, fp, fp, bp, bp, C.c_int]
        L.p6o_add_backward.restype = C.c_int
        L.p6o_add_backward.argtypes = [fp, ip, ip, bp, C.c_int, fp, fp, fp, fp, lp, C.c_int64, C.c_double, fp, fp]
        L.p6o_pose_loss.restype = C.c_int
        L.p6o_pose_loss.argtypes = [fp, fp, fp, fp, C.c_int64, C.c_float, C.c_float, C.c_int,
                                    fp, fp, fp, fp, fp, fp]
        L.p6o_pinhole.restype = None
        L.p6o_pinhole.argtypes = [fp, fp, fp, C.c_int, C.c_int64, fp, fp, fp]
        L.p6o_depth_backproject.restype = None
        L.p6o_depth_backproject.argtypes = [fp, C.c_int, C.c_int, fp, fp, C.c_int, C.c_int64,
                                            C.c_float, fp]
        L.p6o_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def max_threads() -> int:
    return int(lib().p6o_max_threads())


def aten_sum(x) -> np.float32:
    x = _f32(x).ravel()
    return np.float32(lib().p6o_aten_sum_f32(_ptr(x, C.c_float), x.size))


def aten_mean(x) -> np.float32:
    x = _f32(x).ravel()
    return np.float32(lib().p6o_aten_mean_f32(_ptr(x, C.c_float), x.size))


def quat_to_mat(q) -> np.ndarray:
    """[B,4] scalar-last quaternions -> [B,3,3] (models/add_loss.py:203-215)."""
    q = _f32(q).reshape(-1, 4)
    out = np.empty((q.shape[0], 3, 3), np.float32)
    for b in range(q.shape[0]):
        lib().p6o_quat_to_mat(_ptr(q[b], C.c_float), _ptr(out[b], C.c_float))
    return out


def transform(mesh, q, t) -> np.ndarray:
    mesh = _f32(mesh).reshape(-1, 3)
    q, t = _f32(q).ravel(), _f32(t).ravel()
    out = np.empty_like(mesh)
    lib().p6o_transform(_ptr(mesh, C.c_float), mesh.shape[0], _ptr(q, C.c_float),
                        _ptr(t, C.c_float), _ptr(out, C.c_float))
    return out


class MeshTable:
    """Host mirror of ``ADDLoss.points`` / ``ADDLoss.diameters`` (models/add_loss.py:21-22)
    flattened for the C entry points: slot = object id."""

    def __init__(self, points: dict, diameters: dict | None = None, symmetric_ids=SYMMETRIC_OBJECT_IDS):
        diameters = diameters or {}
        ids = [int(k) for k in points.keys()]
        if any(i < 0 for i in ids):
            raise ValueError("negative object id")
        self.n_slots = (max(ids) + 1) if ids else 1
        self.offsets = np.zeros(self.n_slots, np.int32)
        self.counts = np.zeros(self.n_slots, np.int32)
        self.diameters = np.full(self.n_slots, 0.1, np.float64)  # .get(oid, 0.1), add_loss.py:175
        self.symmetric = np.zeros(self.n_slots, np.uint8)
        chunks, off = [], 0
        for oid in sorted(ids):
            m = _f32(points[oid]).reshape(-1, 3)
            self.offsets[oid], self.counts[oid] = off, m.shape[0]
            off += m.shape[0]
            chunks.append(m)
            if oid in diameters:
                self.diameters[oid] = float(diameters[oid])
            self.symmetric[oid] = 1 if oid in symmetric_ids else 0
        self.xyz = np.concatenate(chunks, 0) if chunks else np.zeros((1, 3), np.float32)


def add_eval(table: MeshTable, pred_q, pred_t, gt_q, gt_t, obj_ids, want_adds=True, n_threads=1, bmm=False):
    """Per-pose ADD, ADD-S, ADD-0.1d hit and validity (models/add_loss.py:168-195).
    bmm=True: clouds transformed with the rounding of the batched torch.matmul of
    ``ADDLoss.forward`` (models/add_loss.py:132-133) instead of torch.mm's."""
    pq, pt, gq, gt = (_f32(pred_q).reshape(-1, 4), _f32(pred_t).reshape(-1, 3),
                      _f32(gt_q).reshape(-1, 4), _f32(gt_t).reshape(-1, 3))
    obj = np.ascontiguousarray(obj_ids, dtype=np.int64).ravel()
    B = obj.shape[0]
    add = np.zeros(B, np.float32)
    adds = np.zeros(B, np.float32)
    hit = np.zeros(B, np.uint8)
    valid = np.zeros(B, np.uint8)
    rc = lib().p6o_add_eval(
        _ptr(table.xyz, C.c_float), _ptr(table.offsets, C.c_int32), _ptr(table.counts, C.c_int32),
        _ptr(table.diameters, C.c_double), _ptr(table.symmetric, C.c_uint8), table.n_slots,
        _ptr(pq, C.c_float), _ptr(pt, C.c_float), _ptr(gq, C.c_float), _ptr(gt, C.c_float),
        _ptr(obj, C.c_int64), B, (1 if want_adds else 0) | (2 if bmm else 0), _ptr(add, C.c_float), _ptr(adds, C.c_float),
        _ptr(hit, C.c_uint8), _ptr(valid, C.c_uint8), int(n_threads))
    if rc != 0:
        raise MemoryError("oracle add_eval failed")
    return add, adds, hit, valid


def eval_metrics(table: MeshTable, pred_q, pred_t, gt_q, gt_t, obj_ids, n_threads=1) -> dict:
    """The dict ``ADDLoss.eval_metrics`` returns (models/add_loss.py:197-201): float64
    host means over the valid poses, int 0 when none is valid."""
    add, adds, hit, valid = add_eval(table, pred_q, pred_t, gt_q, gt_t, obj_ids, True, n_threads)
    v = valid.astype(bool)
    if not v.any():
        return {"add_mean": 0, "add_s_mean": 0, "add_01d_acc": 0}
    return {
        "add_mean": np.mean([float(x) for x in add[v]]) * 1000,
        "add_s_mean": np.mean([float(x) for x in adds[v]]) * 1000,
        "add_01d_acc": np.mean([float(x) for x in hit[v]]) * 100,
    }


def add_forward(table: MeshTable, pred_q, pred_t, gt_q, gt_t, obj_ids) -> np.float32:
    """Value of ``ADDLoss.forward`` (models/add_loss.py:101-150): per object group, the
    float32 ATen sum of the per-sample ADD (or ADD-S for symmetric ids), accumulated in
    first-appearance order, divided by the number of valid samples."""
    add, adds, _, valid = add_eval(table, pred_q, pred_t, gt_q, gt_t, obj_ids, True, bmm=True)
    obj = np.asarray(obj_ids, np.int64).ravel()
    order, groups = [], {}
    for i, o in enumerate(obj):
        if valid[i]:
            if int(o) not in groups:
                groups[int(o)] = []
                order.append(int(o))
            groups[int(o)].append(i)
    total, count = np.float32(0.0), 0
    for o in order:
        idx = groups[o]
        vals = (adds if table.symmetric[o] else add)[idx]
        total = np.float32(total + aten_sum(vals))
        count += len(idx)
    if count == 0:
        return np.float32(0.0)
    return np.float32(total / np.float32(count))


def add_backward(table: MeshTable, pred_q, pred_t, gt_q, gt_t, obj_ids, grad_out=1.0):
    """d ADDLoss.forward / d (pred_r, pred_t)  (autograd of models/add_loss.py:101-150)."""
    pq, pt, gq, gt = (_f32(pred_q).reshape(-1, 4), _f32(pred_t).reshape(-1, 3),
                      _f32(gt_q).reshape(-1, 4), _f32(gt_t).reshape(-1, 3))
    obj = np.ascontiguousarray(obj_ids, dtype=np.int64).ravel()
    gq_ = np.zeros_like(pq); gt_ = np.zeros_like(pt)
    rc = lib().p6o_add_backward(_ptr(table.xyz, C.c_float), _ptr(table.offsets, C.c_int32),
                                _ptr(table.counts, C.c_int32), _ptr(table.symmetric, C.c_uint8), table.n_slots,
                                _ptr(pq, C.c_float), _ptr(pt, C.c_float), _ptr(gq, C.c_float), _ptr(gt, C.c_float),
                                _ptr(obj, C.c_int64), obj.shape[0], float(grad_out), _ptr(gq_, C.c_float),
                                _ptr(gt_, C.c_float))
    if rc != 0:
        raise MemoryError("oracle add_backward failed")
    return gq_, gt_


def pose_loss(pred_q, pred_t, gt_q, gt_t, rot_weight=1.0, trans_weight=1.0, mode="geodesic",
              want_grads=True):
    """PoseLoss.forward (+ analytic grads) -- models/pose_loss.py:19-61."""
    pq, pt, gq, gt = (_f32(pred_q).reshape(-1, 4), _f32(pred_t).reshape(-1, 3),
                      _f32(gt_q).reshape(-1, 4), _f32(gt_t).reshape(-1, 3))
    B = pq.shape[0]
    loss, rot, tr = (np.zeros(1, np.float32) for _ in range(3))
    rows = np.zeros(B, np.float32)
    gq_ = np.zeros((B, 4), np.float32) if want_grads else None
    gt_ = np.zeros((B, 3), np.float32) if want_grads else None
    rc = lib().p6o_pose_loss(_ptr(pq, C.c_float), _ptr(pt, C.c_float), _ptr(gq, C.c_float),
                             _ptr(gt, C.c_float), B, float(rot_weight), float(trans_weight),
                             0 if mode == "geodesic" else 1, _ptr(loss, C.c_float),
                             _ptr(rot, C.c_float), _ptr(tr, C.c_float), _ptr(rows, C.c_float),
                             _ptr(gq_, C.c_float), _ptr(gt_, C.c_float))
    if rc != 0:
        raise ValueError("oracle pose_loss failed (B must be > 0)")
    return {"loss": loss[0], "rot": rot[0], "trans": tr[0], "rows": rows, "grad_q": gq_, "grad_t": gt_}


def _k_args(K, B):
    K = _f32(K)
    if K.ndim == 2:
        return K.reshape(9), 0
    return K.reshape(B, 9), 1


def pinhole(z_pred, bbox_center, K, grad_out=None):
    """models/pose_net_rgb_geometric.py:93-109; returns (out[B,3], grad_z[B,1] or None)."""
    z = _f32(z_pred).reshape(-1)
    B = z.shape[0]
    uv = _f32(bbox_center).reshape(B, 2)
    Kf, kb = _k_args(K, B)
    out = np.zeros((B, 3), np.float32)
    go = _f32(grad_out).reshape(B, 3) if grad_out is not None else None
    gz = np.zeros(B, np.float32) if grad_out is not None else None
    lib().p6o_pinhole(_ptr(z, C.c_float), _ptr(uv, C.c_float), _ptr(Kf, C.c_float), kb, B,
                      _ptr(out, C.c_float), _ptr(go, C.c_float), _ptr(gz, C.c_float))
    return out, (gz.reshape(B, 1) if gz is not None else None)


def depth_backproject(depth_raw, bbox_center, K, clamp_hi=223.0):
    """models/pose_net_rgbd_geometric.py:56-85."""
    d = _f32(depth_raw)
    B, H, W = d.shape
    uv = _f32(bbox_center).reshape(B, 2)
    Kf, kb = _k_args(K, B)
    out = np.zeros((B, 3), np.float32)
    lib().p6o_depth_backproject(_ptr(d, C.c_float), H, W, _ptr(uv, C.c_float), _ptr(Kf, C.c_float),
                                kb, B, float(clamp_hi), _ptr(out, C.c_float))
    return out


def _round_f32(fr):
    """Exact rational -> nearest float32, ties to even (no double rounding)."""
    c = np.float32(float(fr))
    best = c
    for cand in (np.nextafter(c, np.float32(-np.inf)), np.nextafter(c, np.float32(np.inf))):
        dc, db = abs(Fraction(float(cand)) - fr), abs(Fraction(float(best)) - fr)
        if dc < db or (dc == db and (int(cand.view(np.uint32)) & 1) == 0 and (int(best.view(np.uint32)) & 1) == 1):
            best = cand
    return best


def fma32(x, y, z):
    """float32 fused multiply-add, exactly rounded."""
    return _round_f32(Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z)))


def resize_linear_u16(img, size=224, bilinear="cv2"):
    """cv2.resize(img_u16, (size, size)) with INTER_LINEAR, restated for a whole square image (NumPy,
    vectorised; used by the live pin in tests/test_live_pins.py).
      bilinear="cv2"      what the pip wheel of OpenCV 4.x does for CV_16U by default -- the call the
                          reference makes (data/dataset_rgbd.py:173): hal::resize hands 16-bit linear
                          resizes to IPP (ippiResizeLinear_16u), whose arithmetic was identified by
                          probing (oracle/VALIDATION.txt): source coordinate (d + 0.5) * (src / dst) - 0.5
                          in float64, weight = float32(fraction), horizontal pass then vertical pass, each
                          lerp as ONE float32 fma(b - a, w, a), round half to even, saturate.
      bilinear="generic"  OpenCV's own C++ path (cv2.ipp.setUseIPP(False), cv2.setUseOptimized(False),
                          or a build without IPP): coordinate rounded to float32, weights (1 - w, w),
                          a * w0 + b * w1 with separate roundings; exact 2x down-scaling switches to
                          INTER_AREA ((a + b + c + d + 2) >> 2)."""
    f, d64 = np.float32, np.float64
    img = np.asarray(img)
    cs = img.shape[0]
    assert img.shape == (cs, cs)
    I = img.astype(f)
    d = np.arange(size)
    if bilinear == "cv2":
        c = (d + 0.5) * (cs / float(size)) - 0.5
        s = np.floor(c).astype(np.int64)
        w = np.where(s < 0, 0.0, c - s).astype(f)
        s0, s1 = np.clip(s, 0, cs - 1), np.clip(s + 1, 0, cs - 1)
        lerp = lambda a, b, ww: (b.astype(d64) - a.astype(d64)).astype(f).astype(d64) * ww.astype(d64) + a.astype(d64)
        hor = lerp(I[:, s0], I[:, s1], np.broadcast_to(w, (cs, size))).astype(f)
        out = lerp(hor[s0], hor[s1], np.broadcast_to(w[:, None], (size, size))).astype(f)
        return np.clip(np.rint(out), 0, 65535).astype(np.uint16)
    if bilinear != "generic":
        raise ValueError("bilinear must be 'cv2' or 'generic'")
    if cs == 2 * size:
        q = img.astype(np.int64)
        return ((q[0::2, 0::2] + q[0::2, 1::2] + q[1::2, 0::2] + q[1::2, 1::2] + 2) >> 2).astype(np.uint16)
    scale = 1.0 / (float(size) / float(cs))
    v = ((d + 0.5) * scale - 0.5).astype(f)
    s = np.floor(v).astype(np.int64)
    w = (v - s.astype(f)).astype(f)
    w = np.where((s < 0) | (s >= cs - 1), f(0), w).astype(f)
    s = np.clip(s, 0, cs - 1)
    s1 = np.minimum(s + 1, cs - 1)
    w0 = (f(1) - w).astype(f)
    hor = ((I[:, s] * w0).astype(f) + (I[:, s1] * w).astype(f)).astype(f)
    out = ((hor[s] * w0[:, None]).astype(f) + (hor[s1] * w[:, None]).astype(f)).astype(f)
    return np.clip(np.rint(out), 0, 65535).astype(np.uint16)


def resize_linear_f32(img, size=224):
    """cv2.resize(img_f32, (size, size)) with INTER_LINEAR for a square float32 image, as the OpenCV pip wheel
    computes it by default (IPP, ippiResizeLinear_32f) -- the call of the reference's inference script
    (scripts/inference/inference_rgbd_geometric.py:140): the arithmetic of resize_linear_u16(bilinear="cv2")
    without the final rounding.  Identified by probing (0 mismatching bit patterns for crops of 48..576 pixels,
    tests/test_live_pins.py re-checks it against the local cv2).  OpenCV's own C++ path for float32
    (cv2.ipp.setUseIPP(False)) is NOT restated: it differs from  a*(1-w) + b*w  in a few pixels when up-sampling."""
    f, d64 = np.float32, np.float64
    I = np.asarray(img, f)
    cs = I.shape[0]
    assert I.shape == (cs, cs)
    d = np.arange(size)
    c = (d + 0.5) * (cs / float(size)) - 0.5
    s = np.floor(c).astype(np.int64)
    w = np.where(s < 0, 0.0, c - s).astype(f)
    s0, s1 = np.clip(s, 0, cs - 1), np.clip(s + 1, 0, cs - 1)
    lerp = lambda a, b, ww: ((b.astype(d64) - a.astype(d64)).astype(f).astype(d64) * ww.astype(d64) + a.astype(d64)).astype(f)
    hor = lerp(I[:, s0], I[:, s1], np.broadcast_to(w, (cs, size)))
    return lerp(hor[s0], hor[s1], np.broadcast_to(w[:, None], (size, size)))


def detection_depth_backproject(depth_u16, boxes_xyxy, K, img_size=224):
    """N1, inference form (NumPy, per box): the per-detection crop code of the reference's inference script
    (scripts/inference/inference_rgbd_geometric.py:109-170) -- integer (x1, y1, x2, y2) detector boxes, square
    crop of 1.2 x max(w, h) cut from the zero-padded uint16 frame, cv2.resize of the crop CAST TO FLOAT32
    (resize_linear_f32), centre and K_crop computed in float64 (Python floats, float64 DEFAULT_K) and only then
    stored as float32 -- followed by the depth back-projection of models/pose_net_rgbd_geometric.py:56-85 at the
    one pixel the network reads.  Differs from the dataset form (crop_depth_backproject) in the box format, in
    where float32 rounding happens and in the un-rounded float32 depth."""
    f = np.float32
    depth = np.asarray(depth_u16)
    H, Wd = depth.shape
    K = np.asarray(K, np.float64)
    B = len(boxes_xyxy)
    center = np.zeros((B, 2), f); Kc = np.zeros((B, 3, 3), f); z_m = np.zeros(B, f); xyz = np.zeros((B, 3), f)
    for b in range(B):
        x1, y1, x2, y2 = (int(v) for v in boxes_xyxy[b])
        c_x, c_y = (x1 + x2) / 2, (y1 + y2) / 2
        w, h = x2 - x1, y2 - y1
        size = max(w, h) * 1.2
        crop_x1, crop_y1 = int(c_x - size / 2), int(c_y - size / 2)
        cs = int(size)
        pad_l, pad_t = max(0, -crop_x1), max(0, -crop_y1)
        adj_x1, adj_y1 = crop_x1 + pad_l, crop_y1 + pad_t
        scale = img_size / cs                                              # float64
        cr = np.clip(np.array([(c_x + pad_l - adj_x1) * scale, (c_y + pad_t - adj_y1) * scale], dtype=f), 0, img_size - 1)
        center[b] = cr
        fx_, fy_, cx_, cy_ = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        Kc[b] = np.array([[fx_ * scale, 0, (cx_ + pad_l - crop_x1) * scale],
                          [0, fy_ * scale, (cy_ + pad_t - crop_y1) * scale], [0, 0, 1]], dtype=f)
        u, v = cr[0], cr[1]
        ui, vi = min(max(int(u), 0), img_size - 1), min(max(int(v), 0), img_size - 1)

        def tx(yy, xx):
            fy, fx = adj_y1 + yy - pad_t, adj_x1 + xx - pad_l
            return f(depth[fy, fx]) if (0 <= fy < H and 0 <= fx < Wd) else f(0)

        def axis(d):
            c = (d + 0.5) * (cs / float(img_size)) - 0.5
            s = int(np.floor(c))
            return min(max(s, 0), cs - 1), min(max(s + 1, 0), cs - 1), (f(0) if s < 0 else f(c - s))

        sx0, sx1, wx = axis(ui)
        sy0, sy1, wy = axis(vi)
        h0 = fma32(f(tx(sy0, sx1) - tx(sy0, sx0)), wx, tx(sy0, sx0))
        h1 = fma32(f(tx(sy1, sx1) - tx(sy1, sx0)), wx, tx(sy1, sx0))
        val = fma32(f(h1 - h0), wy, h0)
        z = f(val / f(1000.0))
        z_m[b] = z
        z = z if z > f(0.01) else f(0.5)
        z = min(max(z, f(0.1)), f(2.0))
        xyz[b] = (f(f(f(u - Kc[b, 0, 2]) * z) / Kc[b, 0, 0]), f(f(f(v - Kc[b, 1, 2]) * z) / Kc[b, 1, 1]), z)
    return {"center": center, "Kcrop": Kc, "z_m": z_m, "xyz": xyz}


def crop_depth_backproject(depth_u16, boxes, K, img_size=224, bilinear="cv2"):
    """N1 restatement (NumPy, per box): the crop geometry of LineMODDatasetRGBD.__getitem__
    (data/dataset_rgbd.py:104-179, no augmentation), cv2.resize's INTER_LINEAR for uint16
    evaluated at the single pixel the network reads (`bilinear`: see resize_linear_u16 -- "cv2" is
    the reference's default call, "generic" OpenCV without IPP), and the depth back-projection of
    models/pose_net_rgbd_geometric.py:56-85.
    Python-int / float64 arithmetic where the reference uses Python scalars, float32 where
    it uses NumPy float32 (NumPy 2 promotion rules)."""
    if bilinear not in ("cv2", "generic"):
        raise ValueError("bilinear must be 'cv2' or 'generic'")
    f = np.float32
    depth = np.asarray(depth_u16)
    H, Wd = depth.shape
    K = np.asarray(K, np.float32)
    B = len(boxes)
    center = np.zeros((B, 2), f); Kc = np.zeros((B, 3, 3), f); z_mm = np.zeros(B, np.uint16); xyz = np.zeros((B, 3), f)

    def texel(yy, xx, x1, y1, pad_l, pad_t):
        # crop pixel (yy, xx) -> padded frame (y1 + yy, x1 + xx) -> frame (.. - pad); zero outside
        fy, fx_ = y1 + yy - pad_t, x1 + xx - pad_l
        if 0 <= fy < H and 0 <= fx_ < Wd:
            return f(depth[fy, fx_])
        return f(0)

    def axis(d, cs):
        scale = 1.0 / (float(img_size) / float(cs))
        v = f((d + 0.5) * scale - 0.5)
        s = int(np.floor(v))
        w = f(v - f(s))
        if s < 0:
            s, w = 0, f(0)
        if s >= cs - 1:
            s, w = cs - 1, f(0)
        return s, min(s + 1, cs - 1), f(f(1.0) - w), w

    def axis_ipp(d, cs):
        c = (d + 0.5) * (cs / float(img_size)) - 0.5          # float64
        s = int(np.floor(c))
        w = f(0) if s < 0 else f(c - s)
        return min(max(s, 0), cs - 1), min(max(s + 1, 0), cs - 1), w

    for b in range(B):
        x, y, w, h = (int(v) for v in boxes[b])
        cgt = np.array([x + w / 2, y + h / 2], dtype=f)
        c_x, c_y = x + w / 2, y + h / 2
        size = max(w, h) * 1.2
        x1, y1 = int(c_x - size / 2), int(c_y - size / 2)
        cs = int(size)
        pad_l, pad_t = max(0, -x1), max(0, -y1)
        x1 += pad_l; y1 += pad_t
        scale32 = f(img_size / cs)
        cc = np.array([f(f(cgt[0] + f(pad_l)) - f(x1)), f(f(cgt[1] + f(pad_t)) - f(y1))], f)
        cr = np.clip((cc * scale32).astype(f), 0, img_size - 1).astype(f)
        center[b] = cr
        fx_, fy_, cx_, cy_ = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        Kc[b] = np.array([[fx_ * scale32, 0, f(f(cx_ + f(pad_l)) - f(x1)) * scale32],
                          [0, fy_ * scale32, f(f(cy_ + f(pad_t)) - f(y1)) * scale32], [0, 0, 1]], f)
        u = min(max(cr[0], f(0)), f(img_size - 1)); v = min(max(cr[1], f(0)), f(img_size - 1))
        ui, vi = min(max(int(u), 0), img_size - 1), min(max(int(v), 0), img_size - 1)
        tx = lambda yy, xx: texel(yy, xx, x1, y1, pad_l, pad_t)
        if bilinear == "cv2":
            sx0, sx1, wx = axis_ipp(ui, cs)
            sy0, sy1, wy = axis_ipp(vi, cs)
            h0 = fma32(f(tx(sy0, sx1) - tx(sy0, sx0)), wx, tx(sy0, sx0))
            h1 = fma32(f(tx(sy1, sx1) - tx(sy1, sx0)), wx, tx(sy1, sx0))
            val = fma32(f(h1 - h0), wy, h0)
        elif cs == 2 * img_size:          # generic path, exact 2x: INTER_AREA
            q = [int(tx(2 * vi + dy, 2 * ui + dx)) for dy in (0, 1) for dx in (0, 1)]
            val = f((sum(q) + 2) >> 2)
        else:
            sx0, sx1, a0, a1 = axis(ui, cs)
            sy0, sy1, b0, b1 = axis(vi, cs)
            h0 = f(f(tx(sy0, sx0) * a0) + f(tx(sy0, sx1) * a1))
            h1 = f(f(tx(sy1, sx0) * a0) + f(tx(sy1, sx1) * a1))
            val = f(f(h0 * b0) + f(h1 * b1))
        zi = int(np.clip(np.rint(val), 0, 65535))
        z_mm[b] = zi
        z = f(f(zi) / f(1000.0))
        z = z if z > f(0.01) else f(0.5)
        z = min(max(z, f(0.1)), f(2.0))
        xyz[b] = (f(f(f(u - Kc[b, 0, 2]) * z) / Kc[b, 0, 0]), f(f(f(v - Kc[b, 1, 2]) * z) / Kc[b, 1, 1]), z)
    return {"center": center, "Kcrop": Kc, "z_mm": z_mm, "xyz": xyz}
