"""
CPU restatement of the SSIM measure on the evaluation path (test infrastructure; only tests/ may import it).

PARITY UNPINNED: the reference computes SSIM with ``piqa.ssim.SSIM()`` (vp_suite/measure/image_wise.py:100-121;
requirements.txt pins piqa 1.1.7).  piqa is a third-party package that is neither vendored in /root/reference nor
installed in this image, so there is nothing to run it against here.  What follows restates piqa's published algorithm
with its constructor defaults (window_size=11, sigma=1.5, n_channels=3, reduction='mean'; ssim(): value_range=1,
k1=0.01, k2=0.03, no padding, channel average) -- written as plain loops over window taps, independently of the torch
expression in vp_suite_b200/evaluation.py -- and is anchored on the reference's own tests for this measure
(tests/test_measure.py:26-50: best value when pred == target, symmetry) and on its call site's pre-processing
(base/base_measure.py:59-75 reshape_clamp: (v + 1) / 2 clamped to [0, 1]).
"""
import numpy as np


def gaussian_window(size=11, sigma=1.5):
    """piqa.utils.functional.gaussian_kernel: exp(-(i - (size - 1) / 2)^2 / (2 sigma^2)), normalised to sum 1 (fp32)."""
    k = np.arange(size, dtype=np.float32) - np.float32((size - 1) / 2)
    k = np.exp(-(k ** 2) / np.float32(2 * sigma ** 2)).astype(np.float32)
    return k / k.sum(dtype=np.float32)


def _filter_valid(v, win):
    """Separable Gaussian filter over the last two axes, valid region only (fp64 accumulation: this is the checker)."""
    n = len(win)
    h, w = v.shape[-2:]
    rows = sum(np.float64(win[k]) * v[..., k:h - n + 1 + k, :] for k in range(n))
    return sum(np.float64(win[k]) * rows[..., :, k:w - n + 1 + k] for k in range(n))


def ssim_images(pred, target, k1=0.01, k2=0.03, value_range=1.0):
    """Per-image SSIM of [b, t, c, h, w] arrays -> [b, t] (the reference's measure is 1 - mean of these)."""
    x = np.clip((np.asarray(pred, dtype=np.float64) + 1) / 2, 0.0, 1.0)     # base_measure.py:71-74
    y = np.clip((np.asarray(target, dtype=np.float64) + 1) / 2, 0.0, 1.0)
    win = gaussian_window()
    c1, c2 = (k1 * value_range) ** 2, (k2 * value_range) ** 2
    mx, my = _filter_valid(x, win), _filter_valid(y, win)
    mxx, myy, mxy = mx * mx, my * my, mx * my
    sxx = _filter_valid(x * x, win) - mxx
    syy = _filter_valid(y * y, win) - myy
    sxy = _filter_valid(x * y, win) - mxy
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    ss = (2 * mxy + c1) / (mxx + myy + c1) * cs
    return ss.mean(axis=(-1, -2, -3))


def ssim_measure(pred, target):
    """SSIM.forward (image_wise.py:111-116): 1 - mean over all (b, t) images; 3-channel frames only."""
    if np.shape(pred)[2] != 3 or np.shape(target)[2] != 3:
        raise ValueError("Structural Similarity (SSIM) needs 3-channel images with the channels at dim 2")
    return 1.0 - float(ssim_images(pred, target).mean())


# ---- MSE / PSNR (fully in the reference; pinned by tests/golden/measures.npz, made by the reference's own classes) ----
def mse_per_horizon(pred, target):
    """vp_suite/measure/image_wise.py:19-31 + base/base_measure.py:54-57: squared error summed over (c, h, w), mean over the
    first k frames, mean over the batch -- for every horizon k = 1..P (metric_provider.py:56-71).  fp64 numpy."""
    se = (np.asarray(pred, dtype=np.float64) - np.asarray(target, dtype=np.float64)) ** 2
    per_frame = se.sum(axis=(2, 3, 4))                                     # [b, P]
    return np.asarray([per_frame[:, :k].mean(axis=1).mean(axis=0) for k in range(1, per_frame.shape[1] + 1)])


def psnr_per_horizon(pred, target):
    """vp_suite/measure/image_wise.py:65-75: 10 * log10(mean_chw squared error) per frame, mean over frames then batch,
    negated for display (to_display)."""
    se = (np.asarray(pred, dtype=np.float64) - np.asarray(target, dtype=np.float64)) ** 2
    per_frame = 10.0 * np.log10(se.mean(axis=(2, 3, 4)))
    return np.asarray([-per_frame[:, :k].mean(axis=1).mean(axis=0) for k in range(1, per_frame.shape[1] + 1)])
