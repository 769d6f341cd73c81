"""Developer aid (GPU): does sequence i's result depend on its batch position / on the run?  Repeats 3 sequences to a
batch, compares every copy with the first, and the same call twice.   python tests/tools/repl_check.py phy 48 [frames]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vp_suite_b200 as V                                     # noqa: E402
from oracle.shapes import SHAPES                               # noqa: E402
from oracle.weights import synth_state_dict, synth_frames     # noqa: E402

key = sys.argv[1]
B = int(sys.argv[2])
pred = int(sys.argv[3]) if len(sys.argv) > 3 else 10
img = (3, 64, 64) if key == "phy" else (1, 64, 64)
ctx = 2 if key == "phy" else 10
t_in = ctx + (pred if key == "predrnn-pp" else 0)
sd = synth_state_dict(SHAPES[key](img), seed=11, gain=1.5)
x3 = synth_frames(3, t_in, *img, seed=321)
m = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0], precision="bf16").eval()
m.load_state_dict(sd)
idx = (torch.arange(B) % 3).cuda()
x = x3.cuda()[idx]
with torch.no_grad():
    a = m(x, pred_frames=pred)[0].clone()
    b = m(x, pred_frames=pred)[0].clone()
    s = m(x3.cuda(), pred_frames=pred)[0].clone()
fmt = lambda u, v: ["%.1e" % float((u[:, t] - v[:, t]).abs().max()) for t in range(pred)]
print("run-to-run max diff per frame:", fmt(a, b))
ref = a[:3][idx]
print("copy vs first copy per frame: ", fmt(a, ref))
print("batch vs 3-sequence run:      ", fmt(a, s[idx]))
worst = (a - ref).abs().flatten(1).max(1).values
top = torch.topk(worst, min(5, B))
print("worst copies:", top.indices.tolist(), ["%.1e" % v for v in top.values.tolist()])
