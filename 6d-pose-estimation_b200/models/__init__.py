"""Reference-shaped module path: ``models.add_loss``, ``models.pose_loss``."""
