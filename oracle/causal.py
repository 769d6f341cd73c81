"""
PredRNN++ (Causal LSTM + gradient highway unit) -- CPU restatement, TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  BASELINE.json's north star names "PredRNN++ CausalLSTMCell plus GHU", but the vp-suite checkout at
/root/reference contains neither (its ``predrnn-pp`` key maps to PredRNN-V2's ST-LSTM): there is no
reference module, test or golden vector to pin this file against.  It restates the published equations of

    Wang, Gao, Long, Wang, Yu: "PredRNN++: Towards A Resolution of the Deep-in-Time Dilemma in Spatiotemporal
    Predictive Learning", ICML 2018 -- Causal LSTM: eq. (1) / fig. 2; GHU: eq. (2); network: fig. 3

in the bias-free parameterisation used by the public PyTorch re-implementations of the authors' TensorFlow code (one
conv per source tensor, gates split along channels): conv_x -> 7C (i, f, g, i', f', g', o), conv_h -> 4C (i, f, g, o),
conv_c -> 3C (i, f, g), conv_m -> 3C (i', f', m_m), conv_c2m -> 4C (i', g', f', o), conv_om -> C, conv_last 1x1 over
cat(c', m'); forget bias 1.  The rollout contract (patches, context + target frames in ``x``, x_gen fed back after the
context, 1x1 head) is the reference's PredRNN_V2 eval path (models/predrnn_v2.py:131-230) with the cell swapped.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this module.
"""
import torch
import torch.nn.functional as F

from .models import reshape_patch, reshape_patch_back

FORGET_BIAS = 1.0


def _conv(x, w, pad=0, q=None):
    """conv2d; ``q`` (optional) rounds both operands first -- the 16-bit-operand emulation used by the GPU tests to separate
    operand rounding (bf16 tensor-core operands, fp32 accumulation and state) from kernel bugs."""
    if q is not None:
        x, w = q(x), q(w)
    return F.conv2d(x, w, padding=pad)


def bf16_operands(t):
    return t.to(torch.bfloat16).to(torch.float32)


def causal_lstm_step(x, h, c, m, w, q=None):
    """One Causal LSTM step.  ``w``: dict with conv_x / conv_h / conv_c / conv_m / conv_c2m / conv_om / conv_last weights.
    Returns (h', c', m')."""
    C = h.shape[1]
    pad = w["conv_x"].shape[-1] // 2
    xs = torch.split(_conv(x, w["conv_x"], pad, q), C, dim=1)          # i, f, g, i', f', g', o
    hs = torch.split(_conv(h, w["conv_h"], pad, q), C, dim=1)          # i, f, g, o
    cs = torch.split(_conv(c, w["conv_c"], pad, q), C, dim=1)          # i, f, g
    ms = torch.split(_conv(m, w["conv_m"], pad, q), C, dim=1)          # i', f', m_m
    # temporal memory (eq. 1, first block): all three gates see x_t, h_{t-1} AND c_{t-1}
    i = torch.sigmoid(xs[0] + hs[0] + cs[0])
    f = torch.sigmoid(xs[1] + hs[1] + cs[1] + FORGET_BIAS)
    g = torch.tanh(xs[2] + hs[2] + cs[2])
    c_new = f * c + i * g
    # spatial memory, cascaded behind c' (eq. 1, second block)
    c2m = torch.split(_conv(c_new, w["conv_c2m"], pad, q), C, dim=1)   # i', g', f', o
    i2 = torch.sigmoid(xs[3] + ms[0] + c2m[0])
    f2 = torch.sigmoid(xs[4] + ms[1] + c2m[2] + FORGET_BIAS)
    g2 = torch.tanh(xs[5] + c2m[1])
    m_new = f2 * torch.tanh(ms[2]) + i2 * g2
    # output (eq. 1, third block): tanh gate over x, h, c', m'; 1x1 conv over the two memories
    o = torch.tanh(xs[6] + hs[3] + c2m[3] + _conv(m_new, w["conv_om"], pad, q))
    h_new = o * torch.tanh(_conv(torch.cat([c_new, m_new], dim=1), w["conv_last"], 0, q))
    return h_new, c_new, m_new


def ghu_step(x, z, w_x, w_z, q=None):
    """Gradient highway unit (eq. 2): p = tanh(W_px x + W_pz z), s = sigmoid(W_sx x + W_sz z), z' = s z + (1 - s) p."""
    C = x.shape[1]
    pad = w_x.shape[-1] // 2
    p, u = torch.split(_conv(x, w_x, pad, q) + _conv(z, w_z, pad, q), C, dim=1)
    u = torch.sigmoid(u)
    return u * z + (1.0 - u) * torch.tanh(p)


def cell_weights(sd, prefix):
    return {"conv_x": sd[prefix + "conv_x.0.weight"], "conv_h": sd[prefix + "conv_h.0.weight"],
            "conv_c": sd[prefix + "conv_c.0.weight"], "conv_m": sd[prefix + "conv_m.0.weight"],
            "conv_c2m": sd[prefix + "conv_c2m.0.weight"], "conv_om": sd[prefix + "conv_om.0.weight"],
            "conv_last": sd[prefix + "conv_last.weight"]}


def state_dict_shapes(img_c, num_layers, num_hidden, patch_size=4, filter_size=5):
    """Key -> shape of the drop-in's state_dict (vp_suite_b200.models.PredRNNpp).  ``num_hidden``: one width for all layers
    or a list (the paper's stacks are 128-64-64-64): conv_m reads the memory written by the previous layer (the top layer for
    layer 0), conv_x the previous layer's h (the GHU's z for layer 1, the patch frame for layer 0)."""
    cp, k, L = patch_size * patch_size * img_c, filter_size, num_layers
    hid = [num_hidden] * L if isinstance(num_hidden, int) else list(num_hidden)[:L]
    shapes = {}
    for i in range(L):
        pre, C, cin, cm = f"cell_list.{i}.", hid[i], (cp if i == 0 else hid[i - 1]), hid[(i - 1) % L]
        shapes[pre + "conv_x.0.weight"] = (7 * C, cin, k, k)
        shapes[pre + "conv_h.0.weight"] = (4 * C, C, k, k)
        shapes[pre + "conv_c.0.weight"] = (3 * C, C, k, k)
        shapes[pre + "conv_m.0.weight"] = (3 * C, cm, k, k)
        shapes[pre + "conv_c2m.0.weight"] = (4 * C, C, k, k)
        shapes[pre + "conv_om.0.weight"] = (C, C, k, k)
        shapes[pre + "conv_last.weight"] = (C, 2 * C, 1, 1)
    shapes["gradient_highway.x_concat.0.weight"] = (2 * hid[0], hid[0], k, k)
    shapes["gradient_highway.z_concat.0.weight"] = (2 * hid[0], hid[0], k, k)
    shapes["conv_last.weight"] = (cp, hid[L - 1], 1, 1)
    return shapes


def predrnnpp_forward(sd, x, pred_frames, cfg=None, q=None):
    """Eval-mode rollout (fig. 3): layer 1 Causal LSTM -> GHU -> layers 2..L, spatial memory zig-zag, 1x1 head.
    ``q``: optional operand rounding of every conv (see _conv); state, gate math and the output frames stay fp32."""
    cfg = {"patch_size": 4, "num_layers": 4, **(cfg or {})}
    p, L = cfg["patch_size"], cfg["num_layers"]
    b, total = x.shape[:2]
    ctx = total - pred_frames
    if ctx < 1:
        raise ValueError("input must hold context and target frames")
    xp = reshape_patch(x, p)
    hp, wp = xp.shape[-2:]
    hid = [sd[f"cell_list.{i}.conv_h.0.weight"].shape[1] for i in range(L)]
    h_t = [torch.zeros(b, hid[i], hp, wp, device=x.device) for i in range(L)]
    c_t = [torch.zeros(b, hid[i], hp, wp, device=x.device) for i in range(L)]
    memory = torch.zeros(b, hid[L - 1], hp, wp, device=x.device)         # layer 0 reads the top layer's memory
    z_t = torch.zeros(b, hid[0], hp, wp, device=x.device)
    ws = [cell_weights(sd, f"cell_list.{i}.") for i in range(L)]
    x_gen, frames = None, []
    for t in range(total - 1):
        net = xp[:, t] if t < ctx else x_gen
        h_t[0], c_t[0], memory = causal_lstm_step(net, h_t[0], c_t[0], memory, ws[0], q)
        z_t = ghu_step(h_t[0], z_t, sd["gradient_highway.x_concat.0.weight"], sd["gradient_highway.z_concat.0.weight"], q)
        h_t[1], c_t[1], memory = causal_lstm_step(z_t, h_t[1], c_t[1], memory, ws[1], q)
        for i in range(2, L):
            h_t[i], c_t[i], memory = causal_lstm_step(h_t[i - 1], h_t[i], c_t[i], memory, ws[i], q)
        x_gen = _conv(h_t[L - 1], sd["conv_last.weight"], 0, q)
        frames.append(x_gen)
    return reshape_patch_back(torch.stack(frames[-pred_frames:], dim=1), p), {}
