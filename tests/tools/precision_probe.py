"""Developer aid (CPU): emulates operand rounding of the PhyDNet path in the oracle to choose operand formats.
Cells (ConvLSTM / PhyCell convs) get bf16-rounded conv inputs and weights; the DCGAN encoder/decoder convs get the
format under test (fp32 / bf16 / fp16).  Accumulation stays fp32.  Prints per-frame max-abs error against the golden
vectors of the reference.      python tests/tools/precision_probe.py [phy_3x64 ...]"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import blocks as OB, models as OM            # noqa: E402
from oracle.weights import synth_state_dict, synth_frames  # noqa: E402

MODE = {"cells": torch.bfloat16, "dcgan": None}
_in_dcgan = [0]
_conv2d, _convT = F.conv2d, F.conv_transpose2d


def _r(t):
    dt = MODE["dcgan"] if _in_dcgan[0] else MODE["cells"]
    return t if (dt is None or t is None) else t.to(dt).to(torch.float32)


class _Fn:
    """torch.nn.functional stand-in for oracle.blocks / oracle.models: rounds conv operands."""
    def __getattr__(self, k):
        return getattr(F, k)

    @staticmethod
    def conv2d(x, w, b=None, **kw):
        return _conv2d(_r(x), _r(w), b, **kw)

    @staticmethod
    def conv_transpose2d(x, w, b=None, **kw):
        return _convT(_r(x), _r(w), b, **kw)


def _wrap(fn):
    def g(*a, **k):
        _in_dcgan[0] += 1
        try:
            return fn(*a, **k)
        finally:
            _in_dcgan[0] -= 1
    return g


def main():
    OB.F = _Fn()
    OM.F = _Fn()
    for name in ("dcgan_conv", "dcgan_conv_transpose", "dcgan_decoder"):
        setattr(OB, name, _wrap(getattr(OB, name)))
    man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
    names = sys.argv[1:] or ["phy_3x64", "phy_1x64", "branch_1x64"]
    for name in names:
        meta = man["models"][name]
        gold = np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))["pred"]
        x = synth_frames(meta["batch"], meta["context"], *meta["img_shape"], seed=meta["xseed"])
        sd = synth_state_dict(meta["shapes"], meta["wseed"], meta["gain"])
        for label, dc in (("fp32", None), ("bf16", torch.bfloat16), ("fp16", torch.float16)):
            MODE["dcgan"] = dc
            with torch.no_grad():
                pred, _ = OM.FORWARDS[meta["key"]](sd, x, meta["pred"])
            d = np.abs(pred.numpy() - gold)
            print(f"{name:12s} cells bf16, dcgan {label}: per-frame max abs err",
                  ["%.2e" % d[:, i].max() for i in range(d.shape[1])], flush=True)


if __name__ == "__main__":
    main()
