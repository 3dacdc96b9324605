"""3-D bounding-box corners of a LineMOD mesh -- same surface as the reference's
``utils/mesh_utils.py`` (SFR-Vision/6d-pose-estimation), host-side NumPy.

``load_mesh_corners(mesh_dir, obj_id_str)`` -> float64 [8,3] corners in metres or None:
ASCII PLY ``obj_<id>.ply`` parsed with the reference's permissive rule (every post-header
line with >= 3 tokens is a vertex, face lines included, utils/mesh_utils.py:20-33), mm -> m,
points farther than 0.3 m from the origin dropped (:37-39), box = 1st / 99th percentile per
axis (:44-45), corners in the reference's order (:47-52).
"""
import os

import numpy as np

# corner k takes min (0) or max (1) per axis -- the order draw_3d_box's edge list expects
_CORNER_BITS = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))


def _ply_rows(path):
    rows, body = [], False
    with open(path, "r") as fh:
        for raw in fh:
            if not body:
                body = "end_header" in raw
                continue
            tok = raw.split()
            if len(tok) >= 3:
                rows.append((float(tok[0]), float(tok[1]), float(tok[2])))
    return np.array(rows)


def load_mesh_corners(mesh_dir, obj_id_str):
    path = os.path.join(mesh_dir, f"obj_{obj_id_str}.ply")
    if not os.path.exists(path):
        return None
    cloud = _ply_rows(path) / 1000.0
    cloud = cloud[np.linalg.norm(cloud, axis=1) < 0.3]
    if len(cloud) == 0:
        return None
    lo_hi = np.stack([np.percentile(cloud, 1, axis=0), np.percentile(cloud, 99, axis=0)])   # [2,3]
    return np.array([[lo_hi[b][axis] for axis, b in enumerate(bits)] for bits in _CORNER_BITS])
