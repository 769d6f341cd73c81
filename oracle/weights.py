"""
Deterministic synthetic weights and inputs shared by the golden-vector generator and the tests
(test infrastructure).  The values depend only on (seed, key name, shape), never on torch's module
init order, so a fixture stores seeds instead of megabytes of weights.
"""
import math
import zlib

import torch


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def synth_tensor(key, shape, seed=0, gain=1.5):
    """Value for one state-dict entry.  Conv weights: U(-a, a) with a = gain / sqrt(fan_in) ("heated" so that
    gates leave the linear regime and outputs are not near-constant); biases U(-0.1, 0.1);
    peepholes N(0, 0.1); GroupNorm/LayerNorm affine: weight 1 + U(-0.2, 0.2), bias U(-0.1, 0.1)."""
    g = _gen(seed, key)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf in ("Wci", "Wcf", "Wco"):
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        if "deconv" in key or "upc" in key:            # ConvTranspose2d: [Cin, Cout, kh, kw]
            fan_in = shape[0] * shape[2] * shape[3]
        a = gain / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    if leaf == "weight" and len(shape) in (1, 3):      # norm scale (GroupNorm [C]; LayerNorm [C, H, W])
        return 1.0 + (torch.rand(shape, generator=g) * 2 - 1) * 0.2
    return (torch.rand(shape, generator=g) * 2 - 1) * 0.1


def synth_state_dict(shapes, seed=0, gain=1.5):
    """``shapes``: mapping key -> shape (e.g. ``{k: v.shape for k, v in model.state_dict().items()}``)."""
    return {k: synth_tensor(k, s, seed, gain) for k, s in shapes.items()}


def synth_frames(b, t, c, h, w, seed=1234):
    """Uniform [0, 1) frames, the reference datasets' default value range (base/base_dataset.py:74-75).
    Smooth-ish content: a low-resolution random field upsampled, plus noise, so that convs see structure."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    coarse = torch.rand(b * t, c, max(h // 8, 1), max(w // 8, 1), generator=g)
    up = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    x = 0.8 * up + 0.2 * torch.rand(b * t, c, h, w, generator=g)
    return x.reshape(b, t, c, h, w).clamp_(0.0, 1.0 - 1e-6).contiguous()


def measure_inputs(shape, seed):
    """Deterministic (pred, target) pair in [0, 1] for the measure fixtures (tests regenerate them from the seed)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    target = torch.rand(shape, generator=g)
    pred = (target + 0.15 * torch.randn(shape, generator=g)).clamp(0.0, 1.0)
    return pred, target


def synth_actions(b, t, a, seed=77):
    """Deterministic action vectors [b, t, a] in (-1, 1) (the reference's datasets emit float actions per frame pair)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((b, t, a), generator=g) * 2 - 1
