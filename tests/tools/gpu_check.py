"""Developer aid for GPU sessions: prints per-frame max-abs errors of every precision/backend against the golden
vectors (more telling than a pytest assertion when a kernel is wrong)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.weights import synth_state_dict, synth_frames   # noqa: E402
import vp_suite_b200 as V                                     # noqa: E402


def main():
    man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
    names = sys.argv[1:] or list(man["models"])
    for name in names:
        meta = man["models"][name]
        if meta["key"] not in V.MODEL_CLASSES:
            continue
        gold = np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))
        t = meta["context"] + (meta["pred"] if meta["key"] == "predrnn-pp" else 0)
        x = synth_frames(meta["batch"], t, *meta["img_shape"], seed=meta["xseed"]).cuda()
        sd = synth_state_dict(meta["shapes"], meta["wseed"], meta["gain"])
        for precision, backend in (("fp32", "auto"), ("bf16", "simt"), ("bf16", "auto")):
            try:
                m = V.MODEL_CLASSES[meta["key"]]("cuda:0", img_shape=tuple(meta["img_shape"]), action_size=0,
                                                 tensor_value_range=[0.0, 1.0], precision=precision, backend=backend,
                                                 **(meta.get("model_kwargs") or {}))
                m.load_state_dict(sd)
                with torch.no_grad():
                    pred, aux = m(x, pred_frames=meta["pred"])
                torch.cuda.synchronize()
                d = np.abs(pred.cpu().numpy() - gold["pred"])
                errs = [float(d[:, i].max()) for i in range(d.shape[1])]
                extra = ""
                if aux is not None and "loss" in gold.files:
                    (k, v), = aux.items()
                    extra = f" loss {float(v):.6f} (ref {float(gold['loss']):.6f})"
                print(f"{name:14s} {precision} {backend:5s} errs {['%.2e' % e for e in errs]} "
                      f"std {gold['pred'].std():.3f} launches {m.last_launch_count()}{extra}", flush=True)
            except Exception as e:   # keep going: one broken mode should not hide the others
                print(f"{name:14s} {precision} {backend:5s} FAILED: {type(e).__name__}: {e}", flush=True)


if __name__ == "__main__":
    main()
