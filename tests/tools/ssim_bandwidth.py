"""Achieved HBM GB/s of the SSIM partial-sum kernels (algorithmic bytes = both frame tensors read once)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from vp_suite_b200 import evaluation as E

for shape in [(512, 20, 3, 128, 128), (256, 10, 3, 64, 64)]:
    x = torch.rand(*shape, device="cuda") * 2 - 1
    y = torch.rand(*shape, device="cuda") * 2 - 1
    for _ in range(3):
        E.ssim_partial_sums(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        E.ssim_partial_sums(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = 2 * x.numel() * 4 / 1e9
    print(f"ssim {shape}: {ms:.3f} ms, {gb / (ms * 1e-3):.0f} GB/s algorithmic ({gb:.2f} GB)")
    for _ in range(3):
        E.metric_partial_sums(x, y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        E.metric_partial_sums(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"mse/psnr {shape}: {ms:.3f} ms, {gb / (ms * 1e-3):.0f} GB/s algorithmic")
