"""
Import shim for the real vp-suite reference (test infrastructure; authoring container only).

``/root/reference`` is a read-only mount that exists only where the golden vectors are generated; it does not
exist on the GPU box, so nothing at test / bench time may call this.  Recipe from SURVEY.md App. A.1:
register a bare ``vp_suite`` namespace (skips ``vp_suite/__init__.py`` and its dataset imports), polyfill
``torch._utils._accumulate`` and stub ``piqa``.
"""
import importlib
import itertools
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("VPK_REFERENCE_ROOT", "/root/reference")


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
