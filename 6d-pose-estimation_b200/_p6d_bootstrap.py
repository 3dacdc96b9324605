"""Locate the native core from any import layout.

The reference-shaped modules live in ``models/`` and ``utils/`` so that adding this
directory to ``sys.path`` makes ``from models.add_loss import ADDLoss`` resolve here
(drop-in).  In that layout they are top-level packages and cannot use relative imports to
reach ``_lib.py``; this helper (importable as ``.._p6d_bootstrap`` inside the package and as
top-level ``_p6d_bootstrap`` in the drop-in layout) loads it by path once and caches it in
``sys.modules``.
"""
import importlib.util
import os
import sys

_NAME = "p6d_b200_core"


def core():
    mod = sys.modules.get(_NAME)
    if mod is not None:
        return mod
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib.py")
    spec = importlib.util.spec_from_file_location(_NAME, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
