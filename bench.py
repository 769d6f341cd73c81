#!/usr/bin/env python
"""
bench.py -- predicted frames/s of the recurrent-rollout hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--seqs-per-gpu S] [--impl ours|reference]

One "step" = one rollout (VPModel.forward) over this rank's shard of synthetic sequences.  Default workload is
BASELINE config 5 (convlstm-shi, 3x128x128, 10 context + 20 predicted frames), sharded by independent sequences:
every rank processes `--seqs-per-gpu` sequences (weak scaling; 512/GPU = the named global batch 4096 at 8 GPUs).
The rollout has no inter-GPU traffic; NCCL only sums the evaluation metrics (MSE / PSNR / SSIM partial sums).

Printed JSON (rank 0, one line): metric/value/unit, ms_per_step, e2e (host buffers through the C ABI, H2D/D2H
inside the timed region), roofline (gate-GEMM kernels: algorithmic FLOPs / CUDA-event time vs the measured bf16
peak), cpu_baseline (the oracle's CPU port on a bounded sample), clocks, gpu_launches.

`--impl reference` times the reference's CPU implementation of the same path -- the oracle port (oracle/), since
the Python reference checkout does not travel to the GPU box -- on the host cores, on the same workload/metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model key, img_shape, context, pred, default seqs/GPU, description)
    "cfg1": ("convlstm-shi", (1, 64, 64), 10, 10, 8, "convlstm-shi 1x64x64 10+10"),
    "cfg2": ("convlstm-branch", (1, 64, 64), 10, 10, 256, "ConvLSTM composition 1x64x64 10+10"),
    "cfg3": ("predrnn-pp", (1, 64, 64), 10, 10, 256, "predrnn-pp 1x64x64 10+10"),
    "cfg4": ("phy", (3, 64, 64), 2, 10, 256, "phy 3x64x64 2+10"),
    "cfg5": ("convlstm-shi", (3, 128, 128), 10, 20, 512, "convlstm-shi 3x128x128 10+20"),
    # ST-LSTM with layer_norm=True, same shape as cfg3
    "cfg3ln": ("predrnn-pp", (1, 64, 64), 10, 10, 256, "predrnn-pp layer_norm=True 1x64x64 10+10"),
}
WORKLOAD_KW = {"cfg3ln": {"layer_norm": True}}          # extra model kwargs of a workload
# "required" GFLOP per sequence of the whole rollout; informational
REQUIRED_GFLOP_PER_SEQ = {"cfg1": 81.03, "cfg2": 26.319, "cfg3": 168.787, "cfg4": 17.74, "cfg5": 524.31,
                          "cfg3ln": 168.787}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    p.add_argument("--seqs-per-gpu", type=int, default=0)
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--microbatch", type=int, default=0)
    p.add_argument("--graph", type=int, default=-1, help="CUDA-graph replay (default: on for small batches)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_throughput(workload, seconds_budget=20.0, threads=None):
    """Times the oracle's CPU port of the rollout on a bounded sample; returns (frames/s, sample description, cores)."""
    import torch
    from oracle import models as OM
    from oracle.weights import synth_state_dict, synth_frames
    from oracle.shapes import SHAPES
    key, img, ctx, pred, _, _ = WORKLOADS[workload]
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    kw = WORKLOAD_KW.get(workload)
    sd = synth_state_dict(SHAPES[key](img, kw) if kw else SHAPES[key](img), 0, 1.0)
    b = 2 if workload == "cfg5" else 8
    t_in = ctx + (pred if key == "predrnn-pp" else 0)
    x = synth_frames(b, t_in, *img, seed=1234)
    fwd = OM.FORWARDS[key]
    with torch.no_grad():
        t0 = time.perf_counter()
        fwd(sd, x, pred)
        first = time.perf_counter() - t0
        reps = max(1, min(5, int(seconds_budget / max(first, 1e-3)) - 1))
        best = first
        for _ in range(reps):
            t0 = time.perf_counter()
            fwd(sd, x, pred)
            best = min(best, time.perf_counter() - t0)
    sample = f"{WORKLOADS[workload][5]} at batch {b}, best of {reps + 1} forwards, torch CPU fp32"
    return b * pred / best, sample, threads, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    key, img, ctx, pred, _, desc = WORKLOADS[args.workload]
    # each "step" is one bounded-sample forward; warm-up + K steps stay within a few minutes
    fps, sample, cores, sec = cpu_reference_throughput(args.workload, seconds_budget=8.0 * max(1, args.steps))
    line = {
        "impl": "reference", "metric": "predicted frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "note": "reference CPU path = oracle port (the Python reference checkout does "
                                             "not travel to the GPU box); throughput is batch-linear on CPU"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import vp_suite_b200 as V
    from vp_suite_b200 import evaluation as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (vp_suite_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to STDOUT when the communicator is created; stdout must carry exactly one JSON
        # line, so file descriptor 1 points at stderr while the process group comes up (init + first collective)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    key, img, ctx, pred, default_b, desc = WORKLOADS[args.workload]
    B = args.seqs_per_gpu or default_b
    t_in = ctx + (pred if key == "predrnn-pp" else 0)
    use_graph = args.graph if args.graph >= 0 else int(B <= 32)
    torch.manual_seed(0)      # random-init weights of the named architecture (torch default init, as the reference)
    model = V.MODEL_CLASSES[key](f"cuda:{local}", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0],
                                 precision=args.precision, max_microbatch=args.microbatch,
                                 use_cuda_graph=bool(use_graph), **WORKLOAD_KW.get(args.workload, {})).eval()

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x_host = torch.rand((B, t_in, *img), generator=g, dtype=torch.float32).pin_memory()
    tgt_host = torch.rand((B, pred, *img), generator=g, dtype=torch.float32)
    x_dev = x_host.to(dev)
    tgt_dev = tgt_host.to(dev)
    # bytes the host entry really copies: predrnn-pp reads only the context frames of its context + target input
    in_bytes = B * ctx * img[0] * img[1] * img[2] * 4 if key == "predrnn-pp" else x_host.numel() * 4
    out_bytes = B * pred * img[0] * img[1] * img[2] * 4

    def metrics_reduce(pred_frames):
        """Per-horizon MSE / PSNR partial sums, summed over ranks with one NCCL all-reduce (vp_suite_b200.evaluation)."""
        return E.all_reduce_sums(E.metric_partial_sums(pred_frames, tgt_dev))

    def step_device():
        with torch.no_grad():
            out, _ = model(x_dev, pred_frames=pred)
        return metrics_reduce(out)

    def step_host():
        out, _ = model.forward_host(x_host, pred_frames=pred)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        dev_ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([dev_ms, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), r

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, _, metric_vec = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = (model.last_launch_count() + 2) * args.steps       # rollout kernels + the two metric-reduction kernels
    ms_per_step = dev_ms / args.steps
    frames = B * pred * world
    value = frames / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (gate GEMMs): events around every launch, same steps ----
    model.set_timing(True)
    with torch.no_grad():
        model(x_dev, pred_frames=pred)
        torch.cuda.synchronize()
        gs = model.last_gemm_stats()
    model.set_timing(False)
    peaks, peak_kind = load_peaks()
    peak = peaks.get("bf16_tflops_sustained" if ms_per_step > 50 else "bf16_tflops", 1590.0)
    achieved = gs["flops"] / max(gs["ms"], 1e-9) * 1e-9 if gs["launches"] else 0.0     # TFLOP/s
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and args.workload == "cfg5":
        with open(tpath) as f:
            tj = json.load(f)["cfg5"]
        mb = model.microbatch_size(B)
        traffic = tj["dram_bytes_in_capture"] / tj["sequences_in_capture"] * mb
        traffic_note = (f"DRAM bytes per launch of {tj['kernel']} at this run's microbatch of {mb} sequences, scaled from the "
                        f"ncu capture at {tj['sequences_in_capture']} sequences ({tj.get('profile', 'profiles/')}); "
                        f"algorithmic bytes of that launch = {tj['algorithmic_bytes_per_position'] * tj['positions_per_sequence'] * mb:.3e}")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "conv_halo_kernel / conv_tc2_kernel <EPI_LSTM> (tcgen05 ConvLSTM gate GEMMs + fused state update)"
                if key != "predrnn-pp" else "conv_tc2_kernel / conv_halo_kernel <EPI_ST_*> (tcgen05 ST-LSTM gate GEMMs + fused update)",
                "gemm_launches_per_step": gs["launches"], "gemm_ms_per_step": gs["ms"],
                "gemm_share_of_step": gs["ms"] / ms_per_step if ms_per_step else None,
                "algorithmic_gflop_per_step": gs["flops"] * 1e-9,
                "peak_source": f"{peak_kind} MEASURED_PEAKS.json "
                               f"({'sustained' if ms_per_step > 50 else 'burst'} bf16 cuBLAS)"}

    # ---- SSIM sums of one (untimed) rollout, 3-channel workloads: the third metric the evaluation all-reduces ----
    ssim_disp = None
    if img[0] == 3:
        with torch.no_grad():
            out, _ = model(x_dev, pred_frames=pred)
            ssim_vec = E.all_reduce_sums(E.ssim_partial_sums(out, tgt_dev))
        ssim_disp = E.finalize_ssim(ssim_vec, B * world)
        del out

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 2)):
            step_host()
        _, wall_ms, _ = timed(step_host, args.steps)
        e2e = {"value": frames / (wall_ms / args.steps * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": wall_ms / args.steps,
               "path": "vpk_model_forward_host: pinned host x -> H2D -> rollout -> D2H host frames, per microbatch"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, sample, cores, _ = cpu_reference_throughput(args.workload)
        cpu_baseline = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        ev = E.finalize_metrics(metric_vec)
        line = {
            "metric": "predicted frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": desc, "seqs_per_gpu": B, "global_batch": B * world, "context": ctx, "pred": pred,
                       "sharding": f"independent sequences, {world} rank(s), no data-path collective",
                       "l2": "inputs/activations per step far exceed the 126 MB L2 (no flush needed)"
                       if B * t_in * img[0] * img[1] * img[2] * 4 > 2.6e8 else "small working set: latency-bound case",
                       "cuda_graph": bool(use_graph),
                       "required_gflop_per_seq": REQUIRED_GFLOP_PER_SEQ.get(args.workload)},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "eval_metrics": {"mse_h1": ev["mse"][0], "psnr_h1": ev["psnr"][0], "mse_hP": ev["mse"][-1],
                             "psnr_hP": ev["psnr"][-1], "sequences": ev["sequences"],
                             "ssim_h1": ssim_disp[0] if ssim_disp else None,
                             "ssim_hP": ssim_disp[-1] if ssim_disp else None,
                             "note": "synthetic random targets; exercises the NCCL metric reduction only (MSE / PSNR sums inside "
                                     "the timed step; SSIM sums, 3-channel workloads, once outside it)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
