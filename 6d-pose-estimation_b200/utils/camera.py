"""Camera constants, the GT/intrinsics reader and the geometric-translation kernels.

``DEFAULT_K`` and ``get_gt_and_K`` keep the reference's surface (``utils/camera.py:8-56``
of SFR-Vision/6d-pose-estimation).  ``pinhole_translation`` and ``depth_backproject`` are
additions: the arithmetic the north-star assigns to this module lives, in the reference,
in two model methods (``models/pose_net_rgb_geometric.py:93-109`` and
``models/pose_net_rgbd_geometric.py:56-85``); those methods can delegate here one-to-one
(INTEGRATION.md).  CUDA only, no CPU path.
"""
import os

import numpy as np
import torch
import yaml


try:                                    # imported as part of the package ...
    from .._p6d_bootstrap import core as _core
except ImportError:                     # ... or as top-level `models` / `utils` (drop-in layout: this
    from _p6d_bootstrap import core as _core   # directory is at the front of sys.path, see dropin.py)


# LineMOD intrinsics used when a frame has none (reference utils/camera.py:8-12)
DEFAULT_K = np.array([
    [572.4114, 0.0, 325.2611],
    [0.0, 573.57043, 242.04899],
    [0.0, 0.0, 1.0],
])


def _read_yaml(path):
    if not os.path.exists(path):
        return None
    with open(path, "r") as fh:
        return yaml.safe_load(fh)


def get_gt_and_K(data_dir, obj_id_str, frame_id):
    """(R [3,3] or None, t [3] in metres or None, K [3,3]) for one frame of one object.

    K: the frame's ``cam_K`` from ``info.yml``; else the first entry of that file; else
    ``DEFAULT_K.copy()``.  Pose: the ``gt.yml`` annotation of the frame whose zero-padded
    ``obj_id`` equals ``obj_id_str`` (reference utils/camera.py:15-56).
    """
    folder = os.path.join(data_dir, obj_id_str)
    K = None
    infos = _read_yaml(os.path.join(folder, "info.yml"))
    if infos is not None:
        if frame_id in infos:
            K = np.array(infos[frame_id]['cam_K']).reshape(3, 3)
        elif infos:
            K = np.array(next(iter(infos.values()))['cam_K']).reshape(3, 3)
    if K is None:
        K = DEFAULT_K.copy()

    rotation = translation = None
    gts = _read_yaml(os.path.join(folder, "gt.yml"))
    if gts is not None and frame_id in gts:
        for entry in gts[frame_id]:
            if str(int(entry['obj_id'])).zfill(2) == obj_id_str:
                translation = np.array(entry['cam_t_m2c']) / 1000.0
                rotation = np.array(entry['cam_R_m2c']).reshape(3, 3)
                break
    return rotation, translation, K


def _k_arg(core, camera_matrix, dev, B):
    K = core.as_cuda_f32(camera_matrix, dev, ())
    if camera_matrix.dim() == 2:
        if tuple(camera_matrix.shape) != (3, 3):
            raise ValueError("camera_matrix must be [3,3] or [B,3,3]")
        return K, 0
    if tuple(camera_matrix.shape) != (B, 3, 3):
        raise ValueError("camera_matrix must be [3,3] or [B,3,3]")
    return K, 1


class _Pinhole(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_pred, bbox_center, camera_matrix):
        core = _core()
        dev = core.require_cuda(z_pred.device)
        B = z_pred.shape[0]
        z = core.as_cuda_f32(z_pred, dev, ())
        uv = core.as_cuda_f32(bbox_center, dev, (2,))
        if z.shape[0] != B or uv.shape[0] != B:
            raise ValueError("z_pred must be [B,1] (or [B]) and bbox_center [B,2]")
        K, kb = _k_arg(core, camera_matrix, dev, B)
        out = torch.empty(B, 3, dtype=torch.float32, device=dev)
        core.check(core.lib().p6d_pinhole_fwd(core.ptr(z), core.ptr(uv), core.ptr(K), kb, B, core.ptr(out),
                                              dev.index, core.stream_ptr(dev)))
        ctx.saved = (uv, K, kb, dev, z_pred.shape, z_pred.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        core = _core()
        uv, K, kb, dev, zshape, zdtype = ctx.saved
        go = core.as_cuda_f32(grad_out, dev, (3,))
        B = go.shape[0]
        gz = torch.empty(B, dtype=torch.float32, device=dev)
        core.check(core.lib().p6d_pinhole_bwd(core.ptr(go), core.ptr(uv), core.ptr(K), kb, B, core.ptr(gz),
                                              dev.index, core.stream_ptr(dev)))
        return gz.reshape(zshape).to(zdtype), None, None


def pinhole_translation(z_pred, bbox_center, camera_matrix):
    """[x, y, z] with x = ((u-cx)*z)/fx, y = ((v-cy)*z)/fy; differentiable in ``z_pred``.
    Same semantics as PoseNetRGBGeometric._compute_pinhole_translation
    (reference models/pose_net_rgb_geometric.py:93-109)."""
    return _Pinhole.apply(z_pred, bbox_center, camera_matrix)


@torch.no_grad()
def depth_backproject(depth_raw, bbox_center, camera_matrix, clamp_hi=223.0):
    """XYZ from the depth pixel under the (clamped, truncated) crop-space bbox centre.
    Same semantics as PoseNetRGBDGeometric._compute_pinhole_translation (reference
    models/pose_net_rgbd_geometric.py:56-85), including its hard-coded 223 clamp."""
    core = _core()
    dev = core.require_cuda(depth_raw.device)
    if depth_raw.dim() != 3:
        raise ValueError("depth_raw must be [B,H,W]")
    B, H, W = depth_raw.shape
    d = core.as_cuda_f32(depth_raw, dev, (H, W))
    uv = core.as_cuda_f32(bbox_center, dev, (2,))
    K, kb = _k_arg(core, camera_matrix, dev, B)
    out = torch.empty(B, 3, dtype=torch.float32, device=dev)
    core.check(core.lib().p6d_depth_backproject(core.ptr(d), H, W, core.ptr(uv), core.ptr(K), kb, B,
                                                float(clamp_hi), core.ptr(out), dev.index,
                                                core.stream_ptr(dev)))
    return out


@torch.no_grad()
def depth_crop_backproject(depth_frame, boxes, camera_matrix, img_size=224, return_aux=False, bilinear="cv2"):
    """XYZ for every (x, y, w, h) box of one uint16 depth frame [H,W] (millimetres) without
    materialising any crop: the reference's pad / square-crop / cv2.resize / K remap
    (data/dataset_rgbd.py:104-179) fused with its depth back-projection
    (models/pose_net_rgbd_geometric.py:56-85).  `camera_matrix` is the frame's [3,3] K.
    With return_aux also returns (centre [B,2], K_crop [B,3,3], z_mm [B] uint16).
    `bilinear`: "cv2" reproduces what ``cv2.resize`` of the OpenCV pip wheel computes for uint16 (its
    IPP path -- the reference's own call), "generic" OpenCV's C++ path (a build without IPP or
    ``cv2.ipp.setUseIPP(False)``); the two differ by 1 mm in 0.1-10 % of the pixels."""
    if bilinear not in ("cv2", "generic"):
        raise ValueError("bilinear must be 'cv2' or 'generic'")
    core = _core()
    if not isinstance(depth_frame, torch.Tensor):
        depth_frame = torch.from_numpy(np.ascontiguousarray(depth_frame, dtype=np.uint16))
    if not isinstance(boxes, torch.Tensor):
        boxes = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32))
    dev = core.require_cuda(depth_frame.device if depth_frame.is_cuda else boxes.device if boxes.is_cuda
                            else camera_matrix.device)
    if depth_frame.dim() != 2 or depth_frame.dtype not in (torch.uint16, torch.int16):
        raise ValueError("depth_frame must be a [H,W] uint16 tensor (millimetres)")
    d = depth_frame.to(dev).contiguous()
    bx = boxes.to(dev, torch.int32).reshape(-1, 4).contiguous()
    K = core.as_cuda_f32(camera_matrix, dev, ())
    if K.numel() != 9:
        raise ValueError("camera_matrix must be the frame's [3,3] intrinsics")
    B = bx.shape[0]
    H, W = d.shape
    xyz = torch.empty(B, 3, dtype=torch.float32, device=dev)
    center = torch.empty(B, 2, dtype=torch.float32, device=dev) if return_aux else None
    kcrop = torch.empty(B, 3, 3, dtype=torch.float32, device=dev) if return_aux else None
    zmm = torch.empty(B, dtype=torch.int16, device=dev) if return_aux else None
    core.check(core.lib().p6d_depth_crop_backproject(core.ptr(d), H, W, core.ptr(bx), B, core.ptr(K), int(img_size),
                                                     0 if bilinear == "cv2" else 1, core.ptr(xyz), core.ptr(center), core.ptr(kcrop), core.ptr(zmm),
                                                     dev.index, core.stream_ptr(dev)))
    if return_aux:
        return xyz, center, kcrop, zmm.view(torch.uint16)
    return xyz


def detection_backproject(depth_frame, boxes_xyxy, camera_matrix=None, img_size=224, return_aux=False):
    """XYZ for every integer (x1, y1, x2, y2) detector box of one uint16 depth frame [H,W] (millimetres)
    without materialising any crop: the per-detection crop code of the reference's inference script
    (scripts/inference/inference_rgbd_geometric.py:109-170: pad, square crop of 1.2 x max(w, h),
    ``cv2.resize`` of the crop cast to float32, centre and K_crop in float64) fused with the depth
    back-projection of models/pose_net_rgbd_geometric.py:56-85.  `camera_matrix` is the frame's [3,3] K
    in float64 (default: DEFAULT_K, as in the script).  With return_aux also returns (centre [B,2],
    K_crop [B,3,3], z_m [B] float32 = ``depth_meters`` under the centre)."""
    core = _core()
    if not isinstance(depth_frame, torch.Tensor):
        depth_frame = torch.from_numpy(np.ascontiguousarray(depth_frame, dtype=np.uint16))
    if not isinstance(boxes_xyxy, torch.Tensor):
        boxes_xyxy = torch.from_numpy(np.ascontiguousarray(boxes_xyxy, dtype=np.int32))
    K = DEFAULT_K if camera_matrix is None else camera_matrix
    if not isinstance(K, torch.Tensor):
        K = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64))
    dev = core.require_cuda(depth_frame.device if depth_frame.is_cuda else boxes_xyxy.device if boxes_xyxy.is_cuda
                            else K.device)
    if depth_frame.dim() != 2 or depth_frame.dtype not in (torch.uint16, torch.int16):
        raise ValueError("depth_frame must be a [H,W] uint16 tensor (millimetres)")
    if K.numel() != 9:
        raise ValueError("camera_matrix must be the frame's [3,3] intrinsics")
    d = depth_frame.to(dev).contiguous()
    bx = boxes_xyxy.to(dev, torch.int32).reshape(-1, 4).contiguous()
    K = K.to(dev, torch.float64).contiguous()
    B = bx.shape[0]
    H, W = d.shape
    xyz = torch.empty(B, 3, dtype=torch.float32, device=dev)
    center = torch.empty(B, 2, dtype=torch.float32, device=dev) if return_aux else None
    kcrop = torch.empty(B, 3, 3, dtype=torch.float32, device=dev) if return_aux else None
    zm = torch.empty(B, dtype=torch.float32, device=dev) if return_aux else None
    core.check(core.lib().p6d_detection_backproject(core.ptr(d), H, W, core.ptr(bx), B, core.ptr(K), int(img_size),
                                                    core.ptr(xyz), core.ptr(center), core.ptr(kcrop), core.ptr(zm),
                                                    dev.index, core.stream_ptr(dev)))
    if return_aux:
        return xyz, center, kcrop, zm
    return xyz
