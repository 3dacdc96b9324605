"""Run the reference's own scripts on the B200 path without editing them.

    python 6d-pose-estimation_b200/dropin.py /path/to/6d-pose-estimation scripts/visualization/compare_all_models.py [args]

The reference's scripts put their project root at ``sys.path[0]`` themselves
(e.g. scripts/training/train_rgb.py: ``sys.path.insert(0, PROJECT_ROOT)``), which would
shadow any replacement found later on the path.  ``install()`` therefore imports THIS
directory's ``models`` and ``utils`` packages first (they stay in ``sys.modules``); both
extend their ``__path__`` with the reference's directories, so

    models.add_loss, models.pose_loss, utils.camera      -> this repo (CUDA kernels)
    models.pose_net_*, utils.mesh_utils, utils.visualization, data.*  -> the reference

and the two geometric networks' ``_compute_pinhole_translation`` methods
(models/pose_net_rgb_geometric.py:93-109, models/pose_net_rgbd_geometric.py:56-85) are
re-pointed at ``utils.camera.pinhole_translation`` / ``depth_backproject``.
"""
import importlib
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_root: str, patch_networks: bool = True) -> None:
    reference_root = os.path.abspath(reference_root)
    for p in (reference_root, HERE):          # HERE ends up first, the reference second
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("models", "utils"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(HERE):
            raise RuntimeError(f"{name!r} was already imported from {mod.__file__}; call install() first")
        importlib.import_module(name)
    if patch_networks:
        cam = importlib.import_module("utils.camera")
        try:
            g = importlib.import_module("models.pose_net_rgb_geometric")
            g.PoseNetRGBGeometric._compute_pinhole_translation = (
                lambda self, z_pred, bbox_center, camera_matrix: cam.pinhole_translation(z_pred, bbox_center, camera_matrix))
            d = importlib.import_module("models.pose_net_rgbd_geometric")
            d.PoseNetRGBDGeometric._compute_pinhole_translation = (
                lambda self, depth_raw, bbox_center, camera_matrix: cam.depth_backproject(depth_raw, bbox_center, camera_matrix))
        except ImportError:
            pass                               # network files (torchvision) not importable: nothing to patch


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    root, script = argv[1], argv[2]
    install(root)
    path = script if os.path.isabs(script) else os.path.join(root, script)
    sys.argv = [path] + argv[3:]
    runpy.run_path(path, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
