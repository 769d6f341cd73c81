"""
vp_suite_b200 -- B200-native (sm_100a) implementation of vp-suite's recurrent video-prediction hot path:
the conv-recurrent cell step (ConvLSTM, ST-LSTM, PhyCell) and the rollouts that call it, behind the reference's
VPModel / VPModelBlock API.  All frames are computed by hand-written CUDA kernels in libvpk.so (C ABI in
include/vpk.h); this package is the thin host side.  (The directory is ``vp_suite_b200`` because a hyphen is not
importable.)
"""
from .base import VPModel, VPModelBlock          # noqa: F401
from .models import MODEL_CLASSES                # noqa: F401
from . import models, _native                    # noqa: F401


def register_into(model_classes: dict, suffix: str = ""):
    """Adds/overrides entries of ``vp_suite.models.MODEL_CLASSES`` with the drop-in classes."""
    for key, cls in MODEL_CLASSES.items():
        model_classes[key + suffix] = cls
    return model_classes
