"""
Batch sharding and metric reduction for multi-GPU evaluation of the rollout (host-side plumbing).

The rollout shards by independent sequences: rank r of G runs ``x[shard_bounds(B, r, G)]`` with a full weight
replica and no inter-GPU traffic.  The only collective is one all-reduce of a small fp64 vector of per-horizon
partial sums, divided after the reduction -- exact (up to fp reassociation) because the reference's measures are
means over the batch of per-sample values:

    MSE   vp_suite/measure/image_wise.py:19-31 + base/base_measure.py:39-57: sum over (c,h,w) of squared error,
          mean over frames, mean over batch
    PSNR  vp_suite/measure/image_wise.py:53-75: forward() = mean_{b,t} 10*log10(mean_chw (p-y)^2)  (lower is
          better); to_display() negates it
    SSIM  vp_suite/measure/image_wise.py:100-121: forward() = 1 - piqa.ssim.SSIM()(reshape_clamp(pred, target)), the
          mean over all (b, t) images of the per-image SSIM; 3-channel images only.  piqa is absent from the reference
          checkout and this image: its algorithm is restated (value parity unpinned, see include/vpk.h)
The per-horizon listing follows PredictionMetricProvider.get_metrics(all_frame_cnts=True)
(vp_suite/measure/metric_provider.py:56-71): horizon k uses the first k predicted frames.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous, balanced split: the first ``batch % world`` ranks get one extra sequence."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def metric_partial_sums(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """fp64 vector [P mse sums | P psnr sums | count] for this rank's sequences; entry t is the sum over the rank's
    batch of the per-frame value at predicted frame t (horizon means are prefix means of these, see finalize).
    CUDA tensors go through libvpk's two reduction kernels (``vpk_metric_partial_sums``: no intermediate tensors, no
    host sync, deterministic); the torch expression below is the definition and serves CPU tensors (gloo tests)."""
    if pred.is_cuda:
        return _native_partial_sums(pred, target)
    se = (pred - target).pow(2).flatten(2)                              # [b, P, chw]
    per_frame = se.sum(-1, dtype=torch.float64)                         # [b, P] sum_chw, fp64 accumulation
    mse_t = per_frame.sum(0)                                            # sum_b sum_chw
    psnr_t = (10.0 * torch.log10(per_frame / se.shape[-1])).sum(0)      # sum_b 10 log10(mean_chw)
    n = torch.tensor([float(pred.shape[0])], dtype=torch.float64, device=pred.device)
    return torch.cat([mse_t, psnr_t, n])


def _native_partial_sums(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    from . import _native as N
    if pred.shape != target.shape or pred.dim() < 3:
        raise ValueError(f"metric_partial_sums: shapes {tuple(pred.shape)} / {tuple(target.shape)}")
    b, p = pred.shape[:2]
    chw = pred[0, 0].numel()
    pred = pred.contiguous().float()
    target = target.to(pred.device).contiguous().float()
    scratch = torch.empty(b * p, dtype=torch.float64, device=pred.device)
    out = torch.empty(2 * p + 1, dtype=torch.float64, device=pred.device)
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    with torch.cuda.device(pred.device):
        N.check(N.lib().vpk_metric_partial_sums(N.ptr(pred), N.ptr(target), b, p, chw, N.ptr(scratch), N.ptr(out),
                                                stream))
    return out


def _ssim_images_torch(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Per-image SSIM [b, P] (definition for CPU tensors / gloo tests): piqa's defaults restated -- 11-tap Gaussian
    window (sigma 1.5), valid region, K1 = 0.01, K2 = 0.03, value range 1, after reshape_clamp's (v + 1) / 2 clamp."""
    import torch.nn.functional as F
    b, p, c = pred.shape[:3]
    x = ((pred.float() + 1) / 2).clamp(0.0, 1.0).flatten(0, 1)
    y = ((target.float() + 1) / 2).clamp(0.0, 1.0).flatten(0, 1)
    k = torch.arange(11, dtype=torch.float32, device=pred.device) - 5.0
    k = torch.exp(-(k ** 2) / (2 * 1.5 ** 2))
    k = k / k.sum()

    def filt(v):
        v = F.conv2d(v, k.view(1, 1, 11, 1).expand(c, 1, 11, 1), groups=c)
        return F.conv2d(v, k.view(1, 1, 1, 11).expand(c, 1, 1, 11), groups=c)

    mx, my = filt(x), filt(y)
    mxx, myy, mxy = mx * mx, my * my, mx * my
    sxx, syy, sxy = filt(x * x) - mxx, filt(y * y) - myy, filt(x * y) - mxy
    cs = (2 * sxy + 0.03 ** 2) / (sxx + syy + 0.03 ** 2)
    ss = (2 * mxy + 0.01 ** 2) / (mxx + myy + 0.01 ** 2) * cs
    return ss.flatten(1).mean(-1, dtype=torch.float64).view(b, p)


def ssim_partial_sums(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """fp64 vector [P]: entry t = sum over this rank's sequences of SSIM(pred[b, t], target[b, t]).  Like the
    reference's measure (image_wise.py:112-113) it takes 3-channel frames only.  CUDA tensors go through libvpk
    (``vpk_metric_ssim_sums``); CPU tensors use the torch definition above (gloo tests)."""
    if pred.shape != target.shape or pred.dim() != 5:
        raise ValueError(f"ssim_partial_sums: shapes {tuple(pred.shape)} / {tuple(target.shape)}")
    if pred.shape[2] != 3:
        raise ValueError("Structural Similarity (SSIM) needs 3-channel images with the channels at dim 2")
    if not pred.is_cuda:
        return _ssim_images_torch(pred, target).sum(0)
    from . import _native as N
    b, p, c, h, w = pred.shape
    n = N.lib().vpk_metric_ssim_scratch_elems(b, p, c, h, w)
    if n < 0:
        raise ValueError(f"ssim_partial_sums: unsupported image size {h} x {w} (the window is 11 x 11)")
    pred = pred.contiguous().float()
    target = target.to(pred.device).contiguous().float()
    scratch = torch.empty(n, dtype=torch.float64, device=pred.device)
    out = torch.empty(p, dtype=torch.float64, device=pred.device)
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    with torch.cuda.device(pred.device):
        N.check(N.lib().vpk_metric_ssim_sums(N.ptr(pred), N.ptr(target), b, p, c, h, w, N.ptr(scratch), N.ptr(out),
                                             stream))
    return out


def finalize_ssim(ssim_vec: torch.Tensor, sequences: float) -> list:
    """Per-horizon SSIM as the reference displays it (to_display(1 - mean) = mean over the first k frames and the
    batch of the per-image SSIM; image_wise.py:119-121)."""
    P = ssim_vec.numel()
    frames = torch.arange(1, P + 1, dtype=torch.float64, device=ssim_vec.device)
    return (torch.cumsum(ssim_vec, 0) / (frames * float(sequences))).tolist()


def all_reduce_sums(vec: torch.Tensor, group=None) -> torch.Tensor:
    """One SUM all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests); no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def finalize_metrics(vec: torch.Tensor) -> dict:
    """Per-horizon means as the reference reports them: horizon k = mean over the first k frames and the batch."""
    P = (vec.numel() - 1) // 2
    n = float(vec[-1])
    frames = torch.arange(1, P + 1, dtype=torch.float64, device=vec.device)
    mse = torch.cumsum(vec[:P], 0) / (frames * n)
    psnr = -torch.cumsum(vec[P:2 * P], 0) / (frames * n)                # to_display(): higher is better
    return {"mse": mse.tolist(), "psnr": psnr.tolist(), "sequences": n}
