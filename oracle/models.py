"""
Functional CPU restatement of the three rollout loops on the hot path (test infrastructure).

Each function takes a ``state_dict`` with the reference's key layout, the input
frames ``x[b, t, c, h, w]`` (fp32) and ``pred_frames`` and returns what the reference's
``VPModel.forward`` returns in eval mode.
"""
import torch
import torch.nn.functional as F

from . import blocks as B

# Default hyper-parameters of the reference models (class attributes there).
EF_DEFAULTS = dict(                       # models/precipitation_nowcasting/ef_conv_lstm.py:31-65
    num_layers=3, enc_c=[16, 64, 64, 96, 96, 96], dec_c=[96, 96, 96, 96, 64, 16],
    enc_conv_k=[3, 3, 3], enc_conv_s=[1, 2, 2], enc_conv_p=[1, 1, 1],
    dec_conv_k=[4, 4, 3], dec_conv_s=[2, 2, 1], dec_conv_p=[1, 1, 1],
    enc_rnn_k=[3, 3, 3], enc_rnn_p=[1, 1, 1], dec_rnn_k=[3, 3, 3], dec_rnn_p=[1, 1, 1],
    enc_conv_names=["conv1_leaky_1", "conv2_leaky_1", "conv3_leaky_1"],
    dec_conv_names=["deconv1_leaky_1", "deconv2_leaky_1", "deconv3_leaky_1"],
    final_conv_2_name="conv3_3",
)
PREDRNN_DEFAULTS = dict(                  # models/predrnn_v2.py:34-43
    patch_size=4, num_layers=3, num_hidden=[128, 128, 128, 128], filter_size=5,
    decoupling_loss_scale=100.0,
)
PHYDNET_DEFAULTS = dict(                  # models/phydnet.py:28-33
    phycell_n_layers=1, phycell_channels=49, phycell_kernel_size=(7, 7),
    convlstm_n_layers=3, convlstm_hidden_dims=[128, 128, 64], convlstm_kernel_size=(3, 3),
)


# --------------------------------------------------------------------------------------------------
# convlstm-shi  (EF_ConvLSTM)
# --------------------------------------------------------------------------------------------------
def _ef_rnn(sd, prefix, inputs, states, seq_len, in_channels, pad):
    """One hzzone ConvLSTM driver with parameters below ``prefix``.  Peepholes missing from the state
    dict (CUDA-constructed reference) are zeros."""
    w, b = sd[prefix + "_conv.weight"], sd[prefix + "_conv.bias"]
    C = w.shape[0] // 4
    if states is not None:
        hw = states[0].shape[-2:]
    else:
        hw = inputs.shape[-2:]
    zeros = torch.zeros(1, C, *hw)
    wci = sd.get(prefix + "Wci", zeros)
    wcf = sd.get(prefix + "Wcf", zeros)
    wco = sd.get(prefix + "Wco", zeros)
    return B.convlstm_shi_sequence(inputs, states, seq_len, w, b, wci, wcf, wco, in_channels, pad)


def ef_convlstm_forward(sd, x, pred_frames, cfg=None):
    """EF_ConvLSTM.forward (ef_blocks.py:184-187): layer-major encoder (:67-82) then layer-major
    forecaster (:100-114) whose top RNN is fed zeros (``inputs=None``)."""
    cfg = {**EF_DEFAULTS, **(cfg or {})}
    L = cfg["num_layers"]
    b, t = x.shape[:2]
    seq = x
    states = []
    for n in range(L):                                                  # Encoder.forward  :76-82
        name = cfg["enc_conv_names"][n]
        flat = seq.reshape(b * t, *seq.shape[2:])                       # forward_by_stage :68-71
        flat = F.conv2d(flat, sd[f"encoder.stage{n + 1}.{name}.weight"], sd[f"encoder.stage{n + 1}.{name}.bias"],
                        stride=cfg["enc_conv_s"][n], padding=cfg["enc_conv_p"][n])
        flat = F.leaky_relu(flat, 0.2)                                  # _make_layers :45
        seq = flat.reshape(b, t, *flat.shape[1:])
        seq, st = _ef_rnn(sd, f"encoder.rnn{n + 1}.", seq, None, t, seq.shape[2], cfg["enc_rnn_p"][n])
        states.append(st)

    seq = None
    for n in range(L):                                                  # Forecaster.forward :108-114
        idx = L - n                                                     # rnn3/stage3 first
        in_c = cfg["enc_c"][-1] if n == 0 else cfg["dec_c"][2 * n - 1]
        seq, _ = _ef_rnn(sd, f"forecaster.rnn{idx}.", seq, states[idx - 1], pred_frames, in_c, cfg["dec_rnn_p"][n])
        name = cfg["dec_conv_names"][n]
        flat = seq.reshape(b * pred_frames, *seq.shape[2:])
        flat = F.conv_transpose2d(flat, sd[f"forecaster.stage{idx}.{name}.weight"],
                                  sd[f"forecaster.stage{idx}.{name}.bias"],
                                  stride=cfg["dec_conv_s"][n], padding=cfg["dec_conv_p"][n])
        flat = F.leaky_relu(flat, 0.2)
        if n == L - 1:                                                  # identity + final 1x1 conv  (ef_conv_lstm.py:99-104)
            fname = cfg["final_conv_2_name"]
            flat = F.conv2d(flat, sd[f"forecaster.stage{idx}.{fname}.weight"], sd[f"forecaster.stage{idx}.{fname}.bias"])
        seq = flat.reshape(b, pred_frames, *flat.shape[1:])
    return seq, None


def ef_trajgru_forward(sd, x, pred_frames, cfg=None):
    """EF_TrajGRU.forward (models/precipitation_nowcasting/ef_traj_gru.py + ef_blocks.py:52-187): the Encoder-Forecaster
    skeleton of ``ef_convlstm_forward`` with TrajGRU drivers (blocks.trajgru_sequence) instead of ConvLSTMs."""
    cfg = {**EF_DEFAULTS, **(cfg or {})}
    L = cfg["num_layers"]
    act = lambda v: F.leaky_relu(v, 0.2)                                # noqa: E731  (ef_traj_gru.py:31)
    b, t = x.shape[:2]
    seq = x
    states = []
    for n in range(L):
        name = cfg["enc_conv_names"][n]
        flat = seq.reshape(b * t, *seq.shape[2:])
        flat = F.leaky_relu(F.conv2d(flat, sd[f"encoder.stage{n + 1}.{name}.weight"], sd[f"encoder.stage{n + 1}.{name}.bias"],
                                     stride=cfg["enc_conv_s"][n], padding=cfg["enc_conv_p"][n]), 0.2)
        seq = flat.reshape(b, t, *flat.shape[1:])
        seq, st = B.trajgru_sequence(seq, None, t, B._sub(sd, f"encoder.rnn{n + 1}."), act)
        states.append(st)
    seq = None
    for n in range(L):
        idx = L - n
        seq, _ = B.trajgru_sequence(seq, states[idx - 1], pred_frames, B._sub(sd, f"forecaster.rnn{idx}."), act)
        name = cfg["dec_conv_names"][n]
        flat = seq.reshape(b * pred_frames, *seq.shape[2:])
        flat = F.leaky_relu(F.conv_transpose2d(flat, sd[f"forecaster.stage{idx}.{name}.weight"],
                                               sd[f"forecaster.stage{idx}.{name}.bias"], stride=cfg["dec_conv_s"][n],
                                               padding=cfg["dec_conv_p"][n]), 0.2)
        if n == L - 1:
            fname = cfg["final_conv_2_name"]
            flat = F.conv2d(flat, sd[f"forecaster.stage{idx}.{fname}.weight"], sd[f"forecaster.stage{idx}.{fname}.bias"])
        seq = flat.reshape(b, pred_frames, *flat.shape[1:])
    return seq, None


# --------------------------------------------------------------------------------------------------
# predrnn-pp  (PredRNN_V2, non action-conditional, layer_norm=False, eval)
# --------------------------------------------------------------------------------------------------
def reshape_patch(x, p):
    """predrnn_v2.py:232-240 -- channel order (p_h, p_w, c)."""
    b, t, c, h, w = x.shape
    x = x.reshape(b, t, c, h // p, p, w // p, p).permute(0, 1, 4, 6, 2, 3, 5)
    return x.reshape(b, t, p * p * c, h // p, w // p)


def reshape_patch_back(xp, p):
    """predrnn_v2.py:242-250."""
    b, t, cpp, hp, wp = xp.shape
    c = cpp // (p * p)
    xp = xp.reshape(b, t, p, p, c, hp, wp).permute(0, 1, 4, 5, 2, 6, 3)
    return xp.reshape(b, t, c, hp * p, wp * p)


def predrnn_v2_ac_forward(sd, x, pred_frames, actions, cfg=None):
    """PredRNN_V2.forward in eval mode with action_conditional=True (predrnn_v2.py:131-230).  The constructor then forces
    conv_actions_on_input and reverse_scheduled_sampling (:65-67): frames and the spatially inflated actions each pass two
    stride-2 k x k convs (no bias, no activation, :74-81, :178-188) down to a (patch_h / 4) x (patch_w / 4) latent, the
    cells are ActionConditionalSpatioTemporalLSTMCell (:94, blocks.stlstm_ac_step), every layer sees the SAME convolved
    action (:191, :200), and x_gen comes back through two transposed convs with the input-side residuals
    (residual_on_action_conv, :213-218).  The reverse-sampling eval mask is 1 for t < context (:306-308), so real frames feed
    the context steps and x_gen the rest, as in the non action-conditional rollout.  ``actions``: [b, >= total - 1, a]."""
    cfg = {**PREDRNN_DEFAULTS, "residual_on_action_conv": True, **(cfg or {})}
    p, L, hid = cfg["patch_size"], cfg["num_layers"], cfg["num_hidden"]
    b, total = x.shape[:2]
    ctx = total - pred_frames
    if ctx < 1:
        raise ValueError("input must hold context and target frames")
    a_size = sd["action_conv_input1.weight"].shape[1]
    if actions is None or actions.shape[-1] != a_size:
        raise ValueError("Given actions are None or of the wrong size!")
    xp = reshape_patch(x, p)
    hp, wp = xp.shape[-2:]
    pad = sd["conv_input1.weight"].shape[-1] // 2
    rh, rw = hp // 4, wp // 4
    h_t = [torch.zeros(b, hid[i], rh, rw) for i in range(L)]
    c_t = [torch.zeros(b, hid[i], rh, rw) for i in range(L)]
    memory = torch.zeros(b, hid[0], rh, rw)
    w_adapter = sd["adapter.weight"]
    x_gen = None
    frames, dec = [], []
    for t in range(total - 1):
        net = xp[:, t] if t < ctx else x_gen
        action = actions[:, t, :, None, None].expand(-1, -1, hp, wp)
        net1 = F.conv2d(net, sd["conv_input1.weight"], stride=2, padding=pad)
        net2 = F.conv2d(net1, sd["conv_input2.weight"], stride=2, padding=pad)
        action = F.conv2d(F.conv2d(action, sd["action_conv_input1.weight"], stride=2, padding=pad),
                          sd["action_conv_input2.weight"], stride=2, padding=pad)
        for i in range(L):
            inp = net2 if i == 0 else h_t[i - 1]
            h_t[i], c_t[i], memory, dc, dm = B.stlstm_ac_step(inp, h_t[i], c_t[i], memory, action, B._sub(sd, f"cell_list.{i}."))
            dcn = F.normalize(F.conv2d(dc, w_adapter).flatten(2), dim=2)
            dmn = F.normalize(F.conv2d(dm, w_adapter).flatten(2), dim=2)
            dec.append(torch.mean(torch.abs(F.cosine_similarity(dcn, dmn, dim=2))))
        top = h_t[L - 1]
        if cfg["residual_on_action_conv"]:
            g = F.conv_transpose2d(top + net2, sd["deconv_output1.weight"], stride=2, padding=pad, output_padding=1)
            x_gen = F.conv_transpose2d(g + net1, sd["deconv_output2.weight"], stride=2, padding=pad, output_padding=1)
        else:
            g = F.conv_transpose2d(top, sd["deconv_output1.weight"], stride=2, padding=pad, output_padding=1)
            x_gen = F.conv_transpose2d(g, sd["deconv_output2.weight"], stride=2, padding=pad, output_padding=1)
        frames.append(x_gen)
    pred = reshape_patch_back(torch.stack(frames[-pred_frames:], dim=1), p)
    loss = cfg["decoupling_loss_scale"] * torch.mean(torch.stack(dec))
    return pred, {"ST-LSTM decouple loss": loss}


def predrnn_v2_forward(sd, x, pred_frames, cfg=None, return_states=False, actions=None):
    """PredRNN_V2.forward in eval mode (predrnn_v2.py:131-230).  ``x`` holds context + target frames (:134-137);
    the eval mask is all zeros (:300-309) so from t >= context_frames the model's own x_gen is the input
    (:172-176); zig-zag ``memory`` (:196-204); decouple loss over all (t, layer) (:197-211, :229)."""
    if "conv_input1.weight" in sd:              # action-conditional checkpoint layout
        return predrnn_v2_ac_forward(sd, x, pred_frames, actions, cfg)
    cfg = {**PREDRNN_DEFAULTS, **(cfg or {})}
    p, L, hid = cfg["patch_size"], cfg["num_layers"], cfg["num_hidden"]
    b, total = x.shape[:2]
    ctx = total - pred_frames
    if ctx < 1:
        raise ValueError("input must hold context and target frames")
    xp = reshape_patch(x, p)
    hp, wp = xp.shape[-2:]
    h_t = [torch.zeros(b, hid[i], hp, wp) for i in range(L)]
    c_t = [torch.zeros(b, hid[i], hp, wp) for i in range(L)]
    memory = torch.zeros(b, hid[0], hp, wp)
    w_adapter = sd["adapter.weight"]
    x_gen = None
    frames, dec = [], []
    for t in range(total - 1):
        net = xp[:, t] if t < ctx else x_gen
        for i in range(L):
            inp = net if i == 0 else h_t[i - 1]
            pre = f"cell_list.{i}."
            ln = None
            if pre + "conv_x.1.weight" in sd:       # layer_norm=True (model_blocks/predrnn.py:24-40)
                ln = {k: (sd[f"{pre}conv_{k}.1.weight"], sd[f"{pre}conv_{k}.1.bias"]) for k in "xhmo"}
            h_t[i], c_t[i], memory, dc, dm = B.stlstm_step(
                inp, h_t[i], c_t[i], memory, sd[pre + "conv_x.0.weight"], sd[pre + "conv_h.0.weight"],
                sd[pre + "conv_m.0.weight"], sd[pre + "conv_o.0.weight"], sd[pre + "conv_last.weight"], ln=ln)
            dcn = F.normalize(F.conv2d(dc, w_adapter).flatten(2), dim=2)
            dmn = F.normalize(F.conv2d(dm, w_adapter).flatten(2), dim=2)
            dec.append(torch.mean(torch.abs(F.cosine_similarity(dcn, dmn, dim=2))))
        x_gen = F.conv2d(h_t[L - 1], sd["conv_last.weight"])
        frames.append(x_gen)
    pred = reshape_patch_back(torch.stack(frames[-pred_frames:], dim=1), p)
    loss = cfg["decoupling_loss_scale"] * torch.mean(torch.stack(dec))
    out = (pred, {"ST-LSTM decouple loss": loss})
    if return_states:
        return out + ((h_t, c_t, memory),)
    return out


# --------------------------------------------------------------------------------------------------
# phy  (PhyDNet, non action-conditional, eval)  and the cfg-2 composition (its residual branch alone)
# --------------------------------------------------------------------------------------------------
def _convcell_stack(sd, prefix, frame, H, C, n_layers, action=None):
    """SingleStepConvLSTM.forward (model_blocks/phydnet.py:147-163) for one frame; H, C are lists (mutated).
    action_conditional: the inflated action is concatenated to the bottom layer's input (:153-155)."""
    inp = frame
    if action is not None:
        inp = torch.cat([inp, action[:, :, None, None].expand(-1, -1, *frame.shape[-2:])], dim=1)
    for j in range(n_layers):
        H[j], C[j] = B.convlstm_cell_step(inp, H[j], C[j], sd[f"{prefix}cell_list.{j}.conv.weight"],
                                          sd[f"{prefix}cell_list.{j}.conv.bias"])
        inp = H[j]
    return H[-1]


def phydnet_forward(sd, x, pred_frames, cfg=None, actions=None):
    """PhyDNet.forward in eval mode (models/phydnet.py:94-137) with encoder_fwd (:73-89): T_in-1 warm-up steps
    on context frames, then autoregression from the last context frame.  Only ``output_image`` (:87-88) is
    computed -- the two visualisation-only decoder passes (:84-85) do not influence the result."""
    cfg = {**PHYDNET_DEFAULTS, **(cfg or {})}
    b, t_in = x.shape[:2]
    hd = cfg["convlstm_hidden_dims"]
    nl = cfg["convlstm_n_layers"]
    n_phy = cfg["phycell_n_layers"]
    state = {}
    ac = "phycell.cell_list.0.frame_action_conv.weight" in sd          # action-conditional checkpoint layout
    if ac:
        a_size = sd["phycell.cell_list.0.frame_action_conv.weight"].shape[1] - sd["phycell.cell_list.0.convgate.weight"].shape[0]
        if actions is None or actions.shape[-1] != a_size:              # models/phydnet.py:103-105
            raise ValueError("Given actions are None or of the wrong size!")

    def step(frame, first, action=None):
        e = B.dcgan_encoder(frame, sd, "encoder_E.")
        ep = B.encoder_split(e, sd, "encoder_Ep.")
        er = B.encoder_split(e, sd, "encoder_Er.")
        if first:                                                        # init_hidden (:107-111, :165-171)
            state["Hp"] = [torch.zeros_like(ep) for _ in range(n_phy)]
            state["H"] = [torch.zeros(b, hd[j], *er.shape[-2:]) for j in range(nl)]
            state["C"] = [torch.zeros(b, hd[j], *er.shape[-2:]) for j in range(nl)]
        inp = ep
        for j in range(n_phy):                                           # PhyCell.forward :95-105
            state["Hp"][j] = B.phycell_step(inp, state["Hp"][j], B._sub(sd, f"phycell.cell_list.{j}."), action)
            inp = state["Hp"][j]
        out_r = _convcell_stack(sd, "convcell.", er, state["H"], state["C"], nl, action)
        dp = B.decoder_split(state["Hp"][-1], sd, "decoder_Dp.")
        dr = B.decoder_split(out_r, sd, "decoder_Dr.")
        return torch.sigmoid(B.dcgan_decoder(dp + dr, sd, "decoder_D."))

    idx = 0                                                             # ac_index (models/phydnet.py:108-122)
    for ei in range(t_in - 1):
        step(x[:, ei], idx == 0, actions[:, idx] if ac else None)
        idx += 1
    frame = x[:, t_in - 1]
    outs = []
    for _ in range(pred_frames):
        frame = step(frame, idx == 0, actions[:, idx] if ac else None)
        outs.append(frame)
        idx += 1
    return torch.stack(outs, dim=1), None


def convlstm_branch_forward(sd, x, pred_frames, cfg=None):
    """BASELINE config 2 ("custom ConvLSTM: encoder + stacked ConvLSTM cells"): OUR composition of reference
    blocks -- PhyDNet's residual branch alone:
    DCGANEncoder -> EncoderSplit -> SingleStepConvLSTM[128,128,64] -> DecoderSplit -> DCGANDecoder -> sigmoid,
    driven with PhyDNet's rollout schedule (T_in-1 warm-up steps, then autoregression).  The cells and blocks
    are pinned by the reference; the composition is not a registered reference model."""
    cfg = {**PHYDNET_DEFAULTS, **(cfg or {})}
    b, t_in = x.shape[:2]
    hd, nl = cfg["convlstm_hidden_dims"], cfg["convlstm_n_layers"]
    state = {}

    def step(frame, first):
        er = B.encoder_split(B.dcgan_encoder(frame, sd, "encoder_E."), sd, "encoder_Er.")
        if first:
            state["H"] = [torch.zeros(b, hd[j], *er.shape[-2:]) for j in range(nl)]
            state["C"] = [torch.zeros(b, hd[j], *er.shape[-2:]) for j in range(nl)]
        out_r = _convcell_stack(sd, "convcell.", er, state["H"], state["C"], nl)
        return torch.sigmoid(B.dcgan_decoder(B.decoder_split(out_r, sd, "decoder_Dr."), sd, "decoder_D."))

    for ei in range(t_in - 1):
        step(x[:, ei], ei == 0)
    frame = x[:, t_in - 1]
    outs = []
    for di in range(pred_frames):
        frame = step(frame, t_in == 1 and di == 0)
        outs.append(frame)
    return torch.stack(outs, dim=1), None


# --------------------------------------------------------------------------------------------------
# st-phy  (STPhy, non action-conditional, eval)          vp_suite/models/st_phy.py:90-181
# --------------------------------------------------------------------------------------------------
STPHY_DEFAULTS = dict(num_layers=3, phycell_channels=49, phycell_kernel_size=(7, 7), st_cell_channels=64)


def autoencoder_encode(x, sd, prefix="autoencoder.encoder."):
    """model_blocks/enc.py:64-69: three valid (unpadded) convs with ReLU -- k5 s2, k3 s2, k3 s1 -- then an L2 normalisation
    along the LAST axis (width), eps 1e-8."""
    x = F.relu(F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], stride=2))
    x = F.relu(F.conv2d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], stride=2))
    x = F.relu(F.conv2d(x, sd[prefix + "mean_layer.weight"], sd[prefix + "mean_layer.bias"]))
    return F.normalize(x, p=2, dim=-1, eps=1e-8)


def autoencoder_decode(z, sd, out_hw, prefix="autoencoder.decoder."):
    """model_blocks/enc.py:93-98: 1x1 conv, two k6 s2 transposed convs (ReLU after each), a k5 s1 transposed conv; the final
    torchvision Resize is the identity when the size already matches (64 x 64 images), which is the case restated here."""
    x = F.relu(F.conv2d(z, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]))
    x = F.relu(F.conv_transpose2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], stride=2))
    x = F.relu(F.conv_transpose2d(x, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], stride=2))
    x = F.conv_transpose2d(x, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"])
    if tuple(x.shape[-2:]) != tuple(out_hw):
        raise ValueError(f"decoder output {tuple(x.shape[-2:])} != image size {tuple(out_hw)}: needs the reference's Resize")
    return x


def stphy_forward(sd, x, pred_frames, cfg=None, actions=None):
    """STPhy.forward in eval mode (models/st_phy.py:90-181).  Per step: the encoded frame (context
    steps) or the previous x_gen feeds EVERY layer (``next_input`` is not updated inside the layer loop, :139-158); layer i
    runs its PhyCell_Cell on (next_input, phy_h[i]) and its LayerNorm ST-LSTM cell on (next_input, h[i], c[i], shared
    st_memory), and merges them with a 1x1 conv over cat[st_h, phy_h] (:158); only the last layer's merge survives as x_gen.
    Frames are decoded from t = context - 1 on (:160-162).  Losses are training-only: returns (frames, None).
    action_conditional=True (``sd`` holds ``action_inflate.weight``, :48-56, 142-150): the step's action vector goes through a
    bias-free Linear to an [inflated_action_dim, enc_h, enc_w] map and the SUM of a (5,1) and a (1,5) conv of it is the
    action tensor of every layer's ActionConditionalSpatioTemporalLSTMCell; the PhyCells get the raw action vector."""
    cfg = {**STPHY_DEFAULTS, **(cfg or {})}
    ac = "action_inflate.weight" in sd
    if ac and (actions is None or actions.shape[-1] != sd["action_inflate.weight"].shape[1]):     # st_phy.py:100-103
        raise ValueError("Given actions are None or of the wrong size!")
    L, Cs = cfg["num_layers"], cfg["st_cell_channels"]
    b, ctx = x.shape[:2]
    enc0 = autoencoder_encode(x[:, 0], sd)
    eh, ew = enc0.shape[-2:]
    phy_h = [torch.zeros(b, Cs, eh, ew) for _ in range(L)]
    st_h = [torch.zeros(b, Cs, eh, ew) for _ in range(L)]
    st_c = [torch.zeros(b, Cs, eh, ew) for _ in range(L)]
    memory = torch.zeros(b, Cs, eh, ew)
    x_gen, outs = None, []
    for t in range(ctx + pred_frames - 1):
        nxt = autoencoder_encode(x[:, t], sd) if t < ctx else x_gen
        if ac:
            a_t = actions[:, t]
            amap = F.linear(a_t, sd["action_inflate.weight"]).view(b, -1, eh, ew)
            infl = F.conv2d(amap, sd["action_conv_h.weight"], padding=(2, 0)) + F.conv2d(amap, sd["action_conv_w.weight"], padding=(0, 2))
        for i in range(L):
            phy_h[i] = B.phycell_step(nxt, phy_h[i], B._sub(sd, f"phycell_list.{i}."), action=a_t if ac else None)
            pre = f"st_cell_list.{i}."
            if ac:
                st_h[i], st_c[i], memory, _, _ = B.stlstm_ac_step(nxt, st_h[i], st_c[i], memory, infl, B._sub(sd, pre))
                x_gen = F.conv2d(torch.cat([st_h[i], phy_h[i]], dim=1), sd[f"hidden_conv_list.{i}.weight"],
                                 sd.get(f"hidden_conv_list.{i}.bias"))
                continue
            ln = {k: (sd[f"{pre}conv_{k}.1.weight"], sd[f"{pre}conv_{k}.1.bias"]) for k in "xhmo"}
            st_h[i], st_c[i], memory, _, _ = B.stlstm_step(
                nxt, st_h[i], st_c[i], memory, sd[pre + "conv_x.0.weight"], sd[pre + "conv_h.0.weight"],
                sd[pre + "conv_m.0.weight"], sd[pre + "conv_o.0.weight"], sd[pre + "conv_last.weight"], ln=ln)
            x_gen = F.conv2d(torch.cat([st_h[i], phy_h[i]], dim=1), sd[f"hidden_conv_list.{i}.weight"],
                             sd.get(f"hidden_conv_list.{i}.bias"))
        if t >= ctx - 1:
            outs.append(autoencoder_decode(x_gen, sd, x.shape[-2:]))
    return torch.stack(outs, dim=1), None


FORWARDS = {
    "convlstm-shi": ef_convlstm_forward,
    "predrnn-pp": predrnn_v2_forward,
    "phy": phydnet_forward,
    "convlstm-branch": convlstm_branch_forward,
    "st-phy": stphy_forward,
    "trajgru": ef_trajgru_forward,
}
