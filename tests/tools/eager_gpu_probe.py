"""
The oracle's torch restatement of the rollouts run in PyTorch EAGER mode on the same B200 (cuDNN / cuBLAS kernels), in
fp32 (TF32 off), TF32 and bf16 autocast, timed with CUDA events -- SURVEY.md sec. 8(d): "the real bar on the box".  It is
what a vp-suite user gets today by calling ``model.to("cuda")``: the same algorithm, one library kernel per torch op.

Test infrastructure (imports ``oracle/``); not part of the product path and not a bench arm.  Prints one JSON line per
(workload, mode).  Usage:  python tests/tools/eager_gpu_probe.py [cfg1 cfg2 ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

import torch  # noqa: E402

import bench  # noqa: E402  (workload table only)
from oracle import models as OM  # noqa: E402
from oracle.shapes import SHAPES  # noqa: E402
from oracle.weights import synth_frames, synth_state_dict  # noqa: E402

# sequences per forward: the bench batch where eager's materialised intermediates fit comfortably, else a bounded sample
BATCH = {"cfg1": 8, "cfg2": 256, "cfg3": 256, "cfg3ln": 256, "cfg4": 256, "cfg5": 64}


def run(workload, mode, reps=3):
    key, img, ctx, pred, _, desc = bench.WORKLOADS[workload]
    kw = bench.WORKLOAD_KW.get(workload)
    sd = synth_state_dict(SHAPES[key](img, kw) if kw else SHAPES[key](img), 0, 1.0)
    sd = {k: v.cuda() for k, v in sd.items()}
    b = BATCH[workload]
    x = synth_frames(b, ctx + (pred if key == "predrnn-pp" else 0), *img, seed=1234).cuda()
    tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    fwd = OM.FORWARDS[key]
    times = []
    with torch.no_grad(), torch.device("cuda"), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
        for i in range(2 + reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fwd(sd, x, pred)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
    ms = min(times)
    return {"workload": workload, "desc": desc, "mode": mode, "batch": b, "ms_per_forward": round(ms, 3),
            "frames_per_s": round(b * pred / (ms * 1e-3), 1), "impl": "oracle port, torch eager on cuda:0 (cuDNN/cuBLAS)"}


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg3ln", "cfg4", "cfg5"]
    for w in names:
        for mode in ("fp32", "tf32", "bf16_autocast"):
            try:
                print(json.dumps(run(w, mode)), flush=True)
            except Exception as e:  # noqa: BLE001  (a probe: report and go on)
                print(json.dumps({"workload": w, "mode": mode, "error": repr(e)[:300]}), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
