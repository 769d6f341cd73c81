"""Developer aid for ncu captures: builds one model and runs `reps` forwards of one workload (no timing, no oracle).
    python tools/run_once.py cfg5 32 [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vp_suite_b200 as V          # noqa: E402
from bench import WORKLOADS, WORKLOAD_KW        # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    key, img, ctx, pred, default_b, desc = WORKLOADS[wl]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else default_b
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    t_in = ctx + (pred if key in ("predrnn-pp", "predrnn-pp-causal") else 0)
    torch.manual_seed(0)
    m = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0],
                             **WORKLOAD_KW.get(wl, {})).eval()
    x = torch.rand(B, t_in, *img, device="cuda")
    with torch.no_grad():
        for _ in range(reps):
            m(x, pred_frames=pred)
    torch.cuda.synchronize()
    print(f"{desc}: {reps} forward(s) of {B} sequences, {m.last_launch_count()} launches each")


if __name__ == "__main__":
    main()
