"""3-D -> 2-D projection and box drawing -- same surface as the reference's
``utils/visualization.py`` (SFR-Vision/6d-pose-estimation), SURVEY.md N4.

``project_points`` / ``draw_3d_box`` / ``draw_axes`` are host-side (one image at a time,
float64 like the reference).  ``project_points_batch`` is an addition: the same projection
for B poses at once on the GPU (float64 kernel ``p6d_project_points``), e.g. the 8 box
corners of every hypothesis of a sweep.
"""
import numpy as np

_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))


try:                                    # imported as part of the package ...
    from .._p6d_bootstrap import core as _core
except ImportError:                     # ... or as top-level `utils` (drop-in layout, see dropin.py)
    from _p6d_bootstrap import core as _core


def _rotation_matrix(rotation):
    """Quaternion [x,y,z,w] (normalised first, like scipy's Rotation.from_quat, which the
    reference uses at utils/visualization.py:21-22) or a ready 3x3 matrix."""
    rotation = np.asarray(rotation, dtype=np.float64)
    if rotation.shape != (4,):
        return rotation
    x, y, z, w = rotation / np.linalg.norm(rotation)
    return np.array([
        [x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w],
    ])


def project_points(points_3d, rotation, translation, K):
    """[N,3] model points -> [N,2] integer pixels: R p + t, z clipped to >= 1 mm,
    u = x fx / z + cx, v = y fy / z + cy, truncated (reference :8-32)."""
    pts = np.asarray(points_3d, dtype=np.float64)
    cam = (_rotation_matrix(rotation) @ pts.T).T + np.asarray(translation, dtype=np.float64)
    depth = np.clip(cam[:, 2], 0.001, None)
    uv = np.empty((pts.shape[0], 2))
    uv[:, 0] = cam[:, 0] * K[0, 0] / depth + K[0, 2]
    uv[:, 1] = cam[:, 1] * K[1, 1] / depth + K[1, 2]
    return uv.astype(int)


def project_points_batch(points_3d, rotation, translation, K, device="cuda"):
    """B poses at once on the GPU: points [N,3], rotation [B,4] quaternions or [B,3,3],
    translation [B,3], K [3,3] -> int64 tensor [B,N,2] (same arithmetic as project_points)."""
    import torch
    core = _core()
    dev = core.require_cuda(device)
    f64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64) if not isinstance(a, torch.Tensor) else a,
                                    dtype=torch.float64).to(dev).contiguous()
    pts, rot, tr, Kd = f64(points_3d).reshape(-1, 3), f64(rotation), f64(translation).reshape(-1, 3), f64(K).reshape(9)
    B = tr.shape[0]
    is_quat = 1 if rot.dim() == 2 and rot.shape[1] == 4 else 0
    if (is_quat and rot.shape[0] != B) or (not is_quat and tuple(rot.shape) != (B, 3, 3)):
        raise ValueError("rotation must be [B,4] or [B,3,3] with the batch size of translation")
    out = torch.empty(B, pts.shape[0], 2, dtype=torch.int64, device=dev)
    core.check(core.lib().p6d_project_points(core.ptr(pts), pts.shape[0], core.ptr(rot), is_quat, core.ptr(tr),
                                             core.ptr(Kd), B, core.ptr(out), dev.index, core.stream_ptr(dev)))
    return out


def draw_3d_box(img, pts_2d, color=(0, 255, 0), thickness=2):
    """Draw the 12 edges of a projected box (corner order of load_mesh_corners) in place."""
    import cv2
    for a, b in _EDGES:
        cv2.line(img, (int(pts_2d[a][0]), int(pts_2d[a][1])), (int(pts_2d[b][0]), int(pts_2d[b][1])), color, thickness)


def draw_axes(img, rotation, translation, K, scale=0.1):
    """Draw the object's X (red), Y (green), Z (blue) axes, `scale` metres long, in place."""
    import cv2
    tips = project_points(np.array([[0, 0, 0], [scale, 0, 0], [0, scale, 0], [0, 0, scale]], dtype=np.float64),
                          rotation, translation, K)
    for tip, bgr in zip(tips[1:], ((0, 0, 255), (0, 255, 0), (255, 0, 0))):
        cv2.line(img, tuple(int(v) for v in tips[0]), tuple(int(v) for v in tip), bgr, 3)
