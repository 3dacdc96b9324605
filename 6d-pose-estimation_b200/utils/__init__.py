"""Reference-shaped module path ``utils.camera`` (only the camera module is on the hot
path; the cv2 drawing helpers of the reference's ``utils`` are out of scope)."""
from .camera import DEFAULT_K, depth_backproject, depth_crop_backproject, get_gt_and_K, pinhole_translation

__all__ = ["get_gt_and_K", "DEFAULT_K", "pinhole_translation", "depth_backproject", "depth_crop_backproject"]
