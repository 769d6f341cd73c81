"""world_size-2 gloo test of the multi-GPU plumbing: batch sharding by independent sequences + one all-reduce of the
metric partial sums reproduces the single-process metrics of the reference's formulas."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import measure as OM
from vp_suite_b200 import evaluation as E


def _reference_metrics(pred, target):
    """vp_suite/measure/image_wise.py:19-75 + base/base_measure.py:39-57 restated for one horizon (all frames)."""
    se = (pred - target).double().pow(2)
    mse = se.sum(dim=(-1, -2, -3)).mean(dim=1).mean(dim=0)
    psnr = -(10 * torch.log10(se.mean(dim=(-1, -2, -3)))).mean(dim=1).mean(dim=0)
    return float(mse), float(psnr)


def _worker(rank, world, port, pred, target, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = E.shard_bounds(pred.shape[0], rank, world)
    vec = E.metric_partial_sums(pred[lo:hi], target[lo:hi])
    ssim = E.ssim_partial_sums(pred[lo:hi], target[lo:hi])   # travels in the same all-reduce as the MSE / PSNR sums
    vec = E.all_reduce_sums(torch.cat([vec, ssim]))
    if rank == 0:
        n_frames = pred.shape[1]
        res = E.finalize_metrics(vec[:-n_frames])
        res["ssim"] = E.finalize_ssim(vec[-n_frames:], res["sequences"])
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_metric_reduction_matches_single_process():
    g = torch.Generator().manual_seed(0)
    pred = torch.rand((5, 4, 3, 12, 14), generator=g)        # 5 sequences over 2 ranks: ragged shards (3 + 2)
    target = torch.rand((5, 4, 3, 12, 14), generator=g)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, pred, target, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got["sequences"] == 5
    for k in range(1, 5):                                    # horizon k = first k predicted frames
        mse, psnr = _reference_metrics(pred[:, :k], target[:, :k])
        assert abs(got["mse"][k - 1] - mse) <= 1e-6 * max(1.0, abs(mse))      # fp32 elementwise, fp64 accumulation
        assert abs(got["psnr"][k - 1] - psnr) <= 1e-6 * max(1.0, abs(psnr))
        ssim = float(OM.ssim_images(pred[:, :k].numpy(), target[:, :k].numpy()).mean())   # to_display(1 - mean)
        assert abs(got["ssim"][k - 1] - ssim) <= 1e-5


def test_shard_bounds_cover_batch_without_overlap():
    for batch in (1, 5, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [E.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_needs_no_process_group():
    v = torch.arange(5, dtype=torch.float64)
    assert torch.equal(E.all_reduce_sums(v.clone()), v)
