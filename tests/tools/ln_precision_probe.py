"""Developer aid (CPU): which conv operand's 16-bit rounding drives the LayerNorm PredRNN-V2 rollout error?

Emulates operand rounding inside the oracle's ST-LSTM step (accumulation stays fp32) at cfg 3's full shape and
rollout length (3 sequences) and prints the per-frame max-abs error against the unrounded oracle.
    python tests/tools/ln_precision_probe.py [variant ...]
A variant is a comma list of <conv>:<act format>/<weight format>, conv in x,h,m,o,l (conv_x, conv_h, conv_m, conv_o,
conv_last) or '*', format in f32, f16, bf16, f16x2 (hi + lo fp16 split = ~22 bits)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import blocks as OB, models as OM            # noqa: E402
from oracle.shapes import SHAPES                         # noqa: E402
from oracle.weights import synth_state_dict, synth_frames  # noqa: E402


def rnd(t, fmt):
    if fmt == "f32":
        return t
    if fmt == "f16":
        return t.to(torch.float16).float()
    if fmt == "bf16":
        return t.to(torch.bfloat16).float()
    if fmt == "f16x2":
        hi = t.to(torch.float16).float()
        return hi + (t - hi).to(torch.float16).float()
    raise ValueError(fmt)


class Fn:
    """torch.nn.functional stand-in: the ST-LSTM step's convs are told apart by their weight shape."""
    def __init__(self, table, C, head_fmt):
        self.table, self.C, self.head = table, C, head_fmt

    def __getattr__(self, k):
        return getattr(F, k)

    def conv2d(self, x, w, b=None, **kw):
        C = self.C
        o, i, kh, _ = w.shape
        if kh == 1:
            name = "l" if (o == C and i == 2 * C) else "head"
        else:
            name = {7 * C: "x", 4 * C: "h", 3 * C: "m", C: "o"}[o]
        fa, fw = self.table.get(name, self.table.get("*", ("f32", "f32")))
        return F.conv2d(rnd(x, fa), rnd(w, fw), b, **kw)


def main():
    variants = sys.argv[1:] or [
        "*:f16/f16",
        "*:f16x2/f16x2",
        "*:f16x2/f16",
        "*:f16/f16x2",
        "x:f16x2/f16x2,h:f16x2/f16x2,m:f16x2/f16x2,o:f16/f16,l:f16/f16,head:f16/f16",
        "x:f16/f16,h:f16/f16,m:f16/f16,o:f16x2/f16x2,l:f16x2/f16x2,head:f16/f16",
    ]
    img, ctx, pred = (1, 64, 64), 10, 10
    sd = synth_state_dict(SHAPES["predrnn-pp"](img, {"layer_norm": True}), seed=11, gain=1.5)
    x = synth_frames(3, ctx + pred, *img, seed=321)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref, _ = OM.predrnn_v2_forward(sd, x, pred)
    print("reference std %.3f" % ref.std().item(), flush=True)
    for v in variants:
        table = {}
        for item in v.split(","):
            name, fm = item.split(":")
            table[name] = tuple(fm.split("/"))
        fn = Fn(table, 128, "f16")
        OB.F, OM.F = fn, fn
        with torch.no_grad():
            out, _ = OM.predrnn_v2_forward(sd, x, pred)
        OB.F, OM.F = F, F
        d = (out - ref).abs()
        print(v, ["%.1e" % d[:, t].max().item() for t in range(pred)], flush=True)


if __name__ == "__main__":
    main()
