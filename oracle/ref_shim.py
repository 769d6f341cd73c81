"""
Import shim for the real vp-suite reference (test infrastructure).

The unmodified reference is looked for in ``baseline/_ref`` first (``pip install --no-deps --target baseline/_ref`` of
the checkout; git-ignored, but it travels to the GPU box with the repo snapshot) and then in the
authoring container's read-only mount ``/root/reference``.  Recipe from SURVEY.md App. A.1: register a bare
``vp_suite`` namespace (skips ``vp_suite/__init__.py`` and its dataset imports, which need packages this image lacks),
polyfill ``torch._utils._accumulate`` and stub ``piqa``.  Users: ``oracle/make_golden.py``, the tests that compare
against the real reference, and ``bench.py --impl reference`` -- never the product path.
"""
import importlib
import itertools
import os
import sys
import types
from unittest.mock import MagicMock

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get("VPK_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "vp_suite", "models")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vp_suite"))


def load_reference():
    """Returns the reference's ``MODEL_CLASSES`` dict; raises if the checkout is absent."""
    if not available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
    import torch
    import torch._utils
    if not hasattr(torch._utils, "_accumulate"):
        torch._utils._accumulate = itertools.accumulate
    if "vp_suite" not in sys.modules:
        pkg = types.ModuleType("vp_suite")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "vp_suite")]
        sys.modules["vp_suite"] = pkg
    for m in ("piqa", "piqa.lpips", "piqa.ssim"):
        try:
            importlib.import_module(m)
        except Exception:
            sys.modules[m] = MagicMock()
    from vp_suite.models import MODEL_CLASSES
    return MODEL_CLASSES
