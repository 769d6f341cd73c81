"""Developer aid: per-layer device times (CUDA events around every kernel) of one rollout.
    python tools/layer_profile.py cfg5 64        # workload, sequences"""
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
    t_in = ctx + (pred if key in ("predrnn-pp", "predrnn-pp-causal") else 0)
    torch.manual_seed(0)
    m = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0],
                             **WORKLOAD_KW.get(wl, {})).eval()
    if os.environ.get("PROFILE_NO_PEEPHOLES"):      # experiment: ConvLSTM without the Wci/Wcf/Wco terms
        sd = {k: v for k, v in m.state_dict().items() if k.rsplit(".", 1)[-1] not in ("Wci", "Wcf", "Wco")}
        m.load_state_dict(sd)
    x = torch.rand(B, t_in, *img, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            m(x, pred_frames=pred)
        torch.cuda.synchronize()
        m.set_timing(2)
        m(x, pred_frames=pred)
        torch.cuda.synchronize()
    rows = m.layer_profile()
    total = sum(r[2] for r in rows)
    print(f"# {desc}, {B} sequences: {total:.2f} ms in kernels")
    print(f"{'layer':44s} {'n':>4s} {'ms':>9s} {'share':>6s} {'us/launch':>10s} {'TFLOP/s':>8s}")
    for name, n, ms, gf in rows:
        tf = gf / ms if ms > 0 else 0.0
        print(f"{name:44s} {n:4d} {ms:9.3f} {100 * ms / total:5.1f}% {1e3 * ms / n:10.1f} {tf:8.1f}")


if __name__ == "__main__":
    main()
