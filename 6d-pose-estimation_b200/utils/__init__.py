"""Reference-shaped module path ``utils.camera``.

Only the camera module is on the hot path.  The reference's other ``utils`` modules
(``mesh_utils``, ``visualization``: cv2 drawing helpers, out of scope here) are picked up
from the reference's own ``utils/`` directory when it is on ``sys.path`` (``extend_path``),
so ``from utils import project_points`` keeps working in the drop-in layout."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from .camera import (DEFAULT_K, depth_backproject, depth_crop_backproject, get_gt_and_K,  # noqa: E402
                     pinhole_translation)

__all__ = ["get_gt_and_K", "DEFAULT_K", "pinhole_translation", "depth_backproject", "depth_crop_backproject"]

try:  # the reference's visualisation helpers, if its utils/ directory is reachable
    from .mesh_utils import load_mesh_corners  # noqa: E402,F401
    from .visualization import draw_3d_box, draw_axes, project_points  # noqa: E402,F401
    __all__ += ["load_mesh_corners", "project_points", "draw_3d_box", "draw_axes"]
except ImportError:
    pass
