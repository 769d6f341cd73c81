import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# When the unmodified reference is present (baseline/_ref, or the authoring container's mount) it is registered BEFORE
# vp_suite_b200 is imported, so that the drop-ins subclass the real vp_suite.base.VPModel -- the deployment situation
# (reference installed, drop-ins registered into it).  VPK_NO_REFERENCE=1 forces the mirror base instead.
from oracle import ref_shim          # noqa: E402
if ref_shim.available() and not os.environ.get("VPK_NO_REFERENCE"):
    ref_shim.load_reference()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Every `gpu`-marked test is skipped in ONE place when there is no CUDA device (plain `pytest tests` on a CPU box
    then reports skips, not 'Found no NVIDIA driver' failures from an early `.cuda()`)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}
