"""Reference-shaped module path ``utils``: ``utils.camera`` (hot path), ``utils.mesh_utils`` and
``utils.visualization`` (N4: host-side helpers with the reference's names and semantics).
``extend_path`` keeps any further module of the reference's own ``utils/`` reachable in
the drop-in layout."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .camera import (DEFAULT_K, depth_backproject, depth_crop_backproject, detection_backproject,  # noqa: E402
                     get_gt_and_K, pinhole_translation)
from .mesh_utils import load_mesh_corners  # noqa: E402
from .visualization import draw_3d_box, draw_axes, project_points, project_points_batch  # noqa: E402

__all__ = ["load_mesh_corners", "project_points", "draw_3d_box", "draw_axes", "get_gt_and_K", "DEFAULT_K",
           "pinhole_translation", "depth_backproject", "depth_crop_backproject", "detection_backproject", "project_points_batch"]
