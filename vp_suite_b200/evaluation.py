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
