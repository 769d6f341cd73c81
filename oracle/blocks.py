"""
Functional CPU restatement of the reference's recurrent cells and conv blocks (test infrastructure).

All tensors are fp32 NCHW, as in the reference.  Parameters are passed explicitly (taken from a
``state_dict`` by the callers in ``oracle/models.py``) so that the same functions check both the
reference's modules and the drop-in modules of ``vp_suite_b200``.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# Shi et al. ConvLSTM with peepholes          (vp_suite/model_blocks/conv_lstm_hzzone.py:38-70)
# --------------------------------------------------------------------------------------------------
def convlstm_shi_step(x, h, c, w, b, wci, wcf, wco, padding=1):
    """One timestep.  Gate conv over cat(x, h) (conv_lstm_hzzone.py:59-60), chunk order i,f,g,o
    (:62), peephole terms on c_{t-1} for i,f and on c_t for o (:64-67), h = o*tanh(c) (:68)."""
    z = F.conv2d(torch.cat([x, h], dim=1), w, b, stride=1, padding=padding)
    zi, zf, zg, zo = torch.chunk(z, 4, dim=1)
    i = torch.sigmoid(zi + wci * c)
    f = torch.sigmoid(zf + wcf * c)
    c_new = f * c + i * torch.tanh(zg)
    o = torch.sigmoid(zo + wco * c_new)
    h_new = o * torch.tanh(c_new)
    return h_new, c_new


def convlstm_shi_sequence(inputs, states, seq_len, w, b, wci, wcf, wco, in_channels, padding=1):
    """Whole-sequence driver (conv_lstm_hzzone.py:38-70): zero initial state when ``states`` is None
    (:39-45), all-zero input when ``inputs`` is None (:53-56), returns (stack of h, (h, c)) (:70)."""
    C = w.shape[0] // 4
    sh, sw = wci.shape[-2:]
    if states is None:
        bsz = inputs.shape[0]
        h = torch.zeros(bsz, C, sh, sw, dtype=torch.float32)
        c = torch.zeros(bsz, C, sh, sw, dtype=torch.float32)
    else:
        h, c = states
        bsz = h.shape[0]
    outs = []
    for t in range(seq_len):
        x = torch.zeros(bsz, in_channels, sh, sw) if inputs is None else inputs[:, t]
        h, c = convlstm_shi_step(x, h, c, w, b, wci, wcf, wco, padding)
        outs.append(h)
    return torch.stack(outs, dim=1), (h, c)


# --------------------------------------------------------------------------------------------------
# ndrplz ConvLSTM cell                          (vp_suite/model_blocks/conv_lstm_ndrplz.py:28-43)
# --------------------------------------------------------------------------------------------------
def convlstm_cell_step(x, h, c, w, b):
    """Gate conv over cat(x, h) with 'same' padding (:31-33); split order i,f,o,g (:34);
    no peepholes (:35-41)."""
    pad = (w.shape[2] // 2, w.shape[3] // 2)
    z = F.conv2d(torch.cat([x, h], dim=1), w, b, padding=pad)
    C = w.shape[0] // 4
    zi, zf, zo, zg = torch.split(z, C, dim=1)
    c_new = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
    h_new = torch.sigmoid(zo) * torch.tanh(c_new)
    return h_new, c_new


# --------------------------------------------------------------------------------------------------
# ST-LSTM v2 cell, layer_norm=False             (vp_suite/model_blocks/predrnn.py:57-83)
# --------------------------------------------------------------------------------------------------
def stlstm_step(x, h, c, m, w_x, w_h, w_m, w_o, w_last, forget_bias=1.0, ln=None):
    """conv_x/conv_h/conv_m are bias-free 'same' convs (:58-60); split orders (:61-63); gate math
    (:65-76); o uses conv_o over cat(c', m') (:78-79); h' = o * tanh(conv_last(mem)) (:80).
    Returns (h', c', m', delta_c, delta_m) (:82).
    ``ln`` (layer_norm=True, :24-40): dict with 'x', 'h', 'm', 'o' -> (weight, bias) of the nn.LayerNorm([C', H, W])
    that follows conv_x / conv_h / conv_m / conv_o (eps 1e-5, statistics over C', H, W of each sample)."""
    C = w_h.shape[0] // 4
    pad = w_x.shape[-1] // 2

    def norm(t, key):
        if ln is None:
            return t
        w, b = ln[key]
        return F.layer_norm(t, tuple(w.shape), w, b, eps=1e-5)

    X = norm(F.conv2d(x, w_x, None, padding=pad), "x")
    H = norm(F.conv2d(h, w_h, None, padding=pad), "h")
    M = norm(F.conv2d(m, w_m, None, padding=pad), "m")
    i_x, f_x, g_x, ip_x, fp_x, gp_x, o_x = torch.split(X, C, dim=1)
    i_h, f_h, g_h, o_h = torch.split(H, C, dim=1)
    i_m, f_m, g_m = torch.split(M, C, dim=1)
    i_t = torch.sigmoid(i_x + i_h)
    f_t = torch.sigmoid(f_x + f_h + forget_bias)
    g_t = torch.tanh(g_x + g_h)
    delta_c = i_t * g_t
    c_new = f_t * c + delta_c
    ip = torch.sigmoid(ip_x + i_m)
    fp = torch.sigmoid(fp_x + f_m + forget_bias)
    gp = torch.tanh(gp_x + g_m)
    delta_m = ip * gp
    m_new = fp * m + delta_m
    mem = torch.cat([c_new, m_new], dim=1)
    o_t = torch.sigmoid(o_x + o_h + norm(F.conv2d(mem, w_o, None, padding=pad), "o"))
    h_new = o_t * torch.tanh(F.conv2d(mem, w_last, None))
    return h_new, c_new, m_new, delta_c, delta_m


# --------------------------------------------------------------------------------------------------
# TrajGRU                                        (vp_suite/model_blocks/traj_gru.py:134-214)
# --------------------------------------------------------------------------------------------------
def trajgru_warp(inp, flow):
    """TrajGRU._warp (:150-166): bilinear sampling of ``inp`` at (x + flow_x, y + flow_y).  The reference normalises the
    sampling grid with (W - 1) / (H - 1) but calls F.grid_sample with its default align_corners=False -- kept as is."""
    b, c, h, w = inp.shape
    xx = torch.arange(0, w).view(1, -1).repeat(h, 1).view(1, 1, h, w).repeat(b, 1, 1, 1)
    yy = torch.arange(0, h).view(-1, 1).repeat(1, w).view(1, 1, h, w).repeat(b, 1, 1, 1)
    vgrid = torch.cat((xx, yy), 1).float() + flow
    vx = 2.0 * vgrid[:, 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * vgrid[:, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(inp, torch.stack([vx, vy], dim=-1), mode="bilinear", padding_mode="zeros", align_corners=False)


def trajgru_sequence(inputs, states, seq_len, p, act):
    """TrajGRU.forward (:170-214) with zoneout 0.  ``p``: i2h, i2f_conv1, h2f_conv1, flows_conv, ret (.weight / .bias).
    Per step: flows from leaky(i2f(x) + h2f(h)) (:137-147), L warps of h by -flow (:192-195), 1x1 ``ret`` over their concat
    (:196-197), GRU gates (:198-206): r = sig(i2h_0 + h2h_0), u = sig(i2h_1 + h2h_1), m = act(i2h_2 + r * h2h_2),
    h' = u * h + (1 - u) * m.  ``inputs`` None: no i2h / i2f terms (:181-182, 203-205)."""
    C = p["ret.weight"].shape[0] // 3
    if states is None:
        i2h_pad = p["i2h.weight"].shape[-1] // 2
        b, _, _, hh, ww = inputs.shape
        states = torch.zeros(b, C, hh, ww)
    h = states
    outs = []
    for t in range(seq_len):
        f1 = F.conv2d(h, p["h2f_conv1.weight"], p["h2f_conv1.bias"], padding=2)
        i2h = None
        if inputs is not None:
            x = inputs[:, t]
            i2h = F.conv2d(x, p["i2h.weight"], p["i2h.bias"], padding=p["i2h.weight"].shape[-1] // 2)
            f1 = F.conv2d(x, p["i2f_conv1.weight"], p["i2f_conv1.bias"], padding=2) + f1
        flows = F.conv2d(act(f1), p["flows_conv.weight"], p["flows_conv.bias"], padding=2)
        warped = torch.cat([trajgru_warp(h, -fl) for fl in torch.split(flows, 2, dim=1)], dim=1)
        h2h = F.conv2d(warped, p["ret.weight"], p["ret.bias"])
        a, bb, cc = torch.split(h2h, C, dim=1)
        if i2h is not None:
            ia, ib, ic = torch.split(i2h, C, dim=1)
            r, u = torch.sigmoid(ia + a), torch.sigmoid(ib + bb)
            m = act(ic + r * cc)
        else:
            r, u = torch.sigmoid(a), torch.sigmoid(bb)
            m = act(r * cc)
        h = u * h + (1 - u) * m
        outs.append(h)
    return torch.stack(outs, dim=1), h


# --------------------------------------------------------------------------------------------------
# Action-conditional ST-LSTM v2 cell            (vp_suite/model_blocks/predrnn.py:86-169)
# --------------------------------------------------------------------------------------------------
def stlstm_ac_step(x, h, c, m, a, sd, forget_bias=1.0):
    """ActionConditionalSpatioTemporalLSTMCell.forward (:142-169) from the block's own state dict ``sd``.
    Differences to ``stlstm_step``: every conv has a bias (:104-139); a fifth conv ``conv_a`` (4C outputs) runs on the
    action tensor ``a`` [b, C, H, W] and MULTIPLIES conv_h's output before the (i, f, g, o) split (:145,149); with
    layer_norm=True each of conv_x / conv_h / conv_a / conv_m / conv_o is followed by its own LayerNorm([C', H, W]).
    Oracle only so far: libvpk has no kernel path for this cell yet (second half)."""
    C = sd["conv_h.0.weight"].shape[0] // 4
    pad = sd["conv_x.0.weight"].shape[-1] // 2

    def conv(t, key):
        out = F.conv2d(t, sd[f"conv_{key}.0.weight"], sd[f"conv_{key}.0.bias"], padding=pad)
        w = sd.get(f"conv_{key}.1.weight")
        return out if w is None else F.layer_norm(out, tuple(w.shape), w, sd[f"conv_{key}.1.bias"], eps=1e-5)

    X, H, A, M = conv(x, "x"), conv(h, "h"), conv(a, "a"), conv(m, "m")
    i_x, f_x, g_x, ip_x, fp_x, gp_x, o_x = torch.split(X, C, dim=1)
    i_h, f_h, g_h, o_h = torch.split(H * A, C, dim=1)
    i_m, f_m, g_m = torch.split(M, C, dim=1)
    i_t = torch.sigmoid(i_x + i_h)
    f_t = torch.sigmoid(f_x + f_h + forget_bias)
    g_t = torch.tanh(g_x + g_h)
    delta_c = i_t * g_t
    c_new = f_t * c + delta_c
    ip = torch.sigmoid(ip_x + i_m)
    fp = torch.sigmoid(fp_x + f_m + forget_bias)
    gp = torch.tanh(gp_x + g_m)
    delta_m = ip * gp
    m_new = fp * m + delta_m
    mem = torch.cat([c_new, m_new], dim=1)
    o_t = torch.sigmoid(o_x + o_h + conv(mem, "o"))
    h_new = o_t * torch.tanh(F.conv2d(mem, sd["conv_last.weight"], sd["conv_last.bias"]))
    return h_new, c_new, m_new, delta_c, delta_m


# --------------------------------------------------------------------------------------------------
# PhyCell cell, action_conditional=False        (vp_suite/model_blocks/phydnet.py:49-62)
# --------------------------------------------------------------------------------------------------
def find_divisor_for_group_norm(x):
    """vp_suite/model_blocks/phydnet.py:348-362 -- largest co-divisor of the divisor closest below sqrt(x)."""
    sq = math.floor(math.sqrt(x))
    while x % sq != 0:
        sq -= 1
    return x // sq


def phycell_step(x, h, p, action=None):
    """``p`` holds F.conv1.{weight,bias}, F.bn1.{weight,bias}, F.conv2.{weight,bias}, convgate.{weight,bias}.
    K = sigmoid(convgate(cat[x, h])) (:57-59); h~ = h + F(h) with F = conv1 -> GroupNorm -> conv2 (:33-39, :60);
    h' = h~ + K * (x - h~) (:61).  action_conditional=True (``p`` also holds frame_action_conv / hidden_action_conv,
    :44-48): the action vector [b, a] is inflated to the frame size, concatenated to frame and to hidden, and each goes
    through its own 1x1 conv first (:50-55); everything after uses the convolved frame / hidden."""
    if "frame_action_conv.weight" in p:
        infl = action[:, :, None, None].expand(-1, -1, *x.shape[-2:])
        x = F.conv2d(torch.cat([x, infl], dim=1), p["frame_action_conv.weight"], p["frame_action_conv.bias"])
        h = F.conv2d(torch.cat([h, infl], dim=1), p["hidden_action_conv.weight"], p["hidden_action_conv.bias"])
    k1 = p["F.conv1.weight"]
    hid = k1.shape[0]
    groups = find_divisor_for_group_norm(hid)
    f = F.conv2d(h, k1, p["F.conv1.bias"], padding=(k1.shape[2] // 2, k1.shape[3] // 2))
    f = F.group_norm(f, groups, p["F.bn1.weight"], p["F.bn1.bias"], eps=1e-5)
    f = F.conv2d(f, p["F.conv2.weight"], p["F.conv2.bias"])
    gate = torch.sigmoid(F.conv2d(torch.cat([x, h], dim=1), p["convgate.weight"], p["convgate.bias"], padding=1))
    h_tilde = h + f
    return h_tilde + gate * (x - h_tilde)


# --------------------------------------------------------------------------------------------------
# DCGAN conv blocks                              (vp_suite/model_blocks/conv.py:58-95)
# --------------------------------------------------------------------------------------------------
def dcgan_conv(x, w, b, gn_w, gn_b, stride):
    """Conv3x3(stride, pad 1) -> GroupNorm(16) -> LeakyReLU(0.2)  (conv.py:66-70)."""
    y = F.conv2d(x, w, b, stride=stride, padding=1)
    y = F.group_norm(y, 16, gn_w, gn_b, eps=1e-5)
    return F.leaky_relu(y, 0.2)


def dcgan_conv_transpose(x, w, b, gn_w, gn_b, stride):
    """ConvT3x3(stride, pad 1, output_padding = [stride == 2]) -> GroupNorm(16) -> LeakyReLU(0.2)  (conv.py:85-92)."""
    y = F.conv_transpose2d(x, w, b, stride=stride, padding=1, output_padding=int(stride == 2))
    y = F.group_norm(y, 16, gn_w, gn_b, eps=1e-5)
    return F.leaky_relu(y, 0.2)


def _sub(sd, prefix):
    """View of a state dict below ``prefix`` (keys with the prefix stripped)."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def dcgan_block(x, sd, prefix, stride, transpose):
    p = _sub(sd, prefix)
    fn = dcgan_conv_transpose if transpose else dcgan_conv
    return fn(x, p["main.0.weight"], p["main.0.bias"], p["main.1.weight"], p["main.1.bias"], stride)


def dcgan_encoder(x, sd, prefix):
    """DCGANEncoder: c1 stride 2, c2 stride 1, c3 stride 2  (enc.py:107-118)."""
    h = dcgan_block(x, sd, prefix + "c1.", 2, False)
    h = dcgan_block(h, sd, prefix + "c2.", 1, False)
    return dcgan_block(h, sd, prefix + "c3.", 2, False)


def dcgan_decoder(x, sd, prefix):
    """DCGANDecoder: upc1 stride 2, upc2 stride 1, upc3 = bare ConvT stride 2 (enc.py:128-141).
    The trailing Resize is the identity whenever the stride arithmetic already lands on the image size
    (SURVEY.md App. D); other sizes are rejected here rather than silently interpolated."""
    d = dcgan_block(x, sd, prefix + "upc1.", 2, True)
    d = dcgan_block(d, sd, prefix + "upc2.", 1, True)
    return F.conv_transpose2d(d, sd[prefix + "upc3.weight"], sd[prefix + "upc3.bias"],
                              stride=2, padding=1, output_padding=1)


def encoder_split(x, sd, prefix):
    """EncoderSplit: two stride-1 DCGANConv (phydnet.py:184-192)."""
    return dcgan_block(dcgan_block(x, sd, prefix + "c1.", 1, False), sd, prefix + "c2.", 1, False)


def decoder_split(x, sd, prefix):
    """DecoderSplit: two stride-1 DCGANConvTranspose (phydnet.py:201-209)."""
    return dcgan_block(dcgan_block(x, sd, prefix + "upc1.", 1, True), sd, prefix + "upc2.", 1, True)
