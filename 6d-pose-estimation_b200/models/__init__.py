"""Reference-shaped module path: ``models.add_loss``, ``models.pose_loss``.

``extend_path`` merges every other ``models/`` directory found on ``sys.path`` into this
package, so with the reference's root on ``sys.path`` (behind this directory)
``models.pose_net_rgb`` etc. still resolve to the reference's network files while
``models.add_loss`` / ``models.pose_loss`` resolve here (see dropin.py)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
