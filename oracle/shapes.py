"""state_dict key -> shape of the reference models at their default hyper-parameters (test infrastructure).

Lets the CPU baseline build weights for any image size without importing the reference or the product package.
Checked against the reference's own listings in tests/golden/manifest.json (tests/test_oracle_golden.py).
"""
from .models import EF_DEFAULTS, PREDRNN_DEFAULTS, PHYDNET_DEFAULTS, STPHY_DEFAULTS


def _conv_out(v, k, s, p):
    return (v + 2 * p - (k - 1) - 1) // s + 1


def ef_shapes(img_shape, cfg=None):
    """SURVEY.md App. B; models/precipitation_nowcasting/ef_conv_lstm.py:70-108."""
    cfg = {**EF_DEFAULTS, **(cfg or {})}
    c, h, w = img_shape
    out = {}
    in_c = c
    sizes = []
    for n in range(3):
        mid, oc = cfg["enc_c"][2 * n], cfg["enc_c"][2 * n + 1]
        k, s, p = cfg["enc_conv_k"][n], cfg["enc_conv_s"][n], cfg["enc_conv_p"][n]
        h, w = _conv_out(h, k, s, p), _conv_out(w, k, s, p)
        sizes.append((h, w))
        name = cfg["enc_conv_names"][n]
        out[f"encoder.stage{n + 1}.{name}.weight"] = (mid, in_c, k, k)
        out[f"encoder.stage{n + 1}.{name}.bias"] = (mid,)
        for pk in ("Wci", "Wcf", "Wco"):
            out[f"encoder.rnn{n + 1}.{pk}"] = (1, oc, h, w)
        rk = cfg["enc_rnn_k"][n]
        out[f"encoder.rnn{n + 1}._conv.weight"] = (4 * oc, mid + oc, rk, rk)
        out[f"encoder.rnn{n + 1}._conv.bias"] = (4 * oc,)
        in_c = oc
    for n in range(3):
        idx = 3 - n
        mid, oc = cfg["dec_c"][2 * n], cfg["dec_c"][2 * n + 1]
        hh, ww = sizes[2 - n]
        for pk in ("Wci", "Wcf", "Wco"):
            out[f"forecaster.rnn{idx}.{pk}"] = (1, mid, hh, ww)
        rk = cfg["dec_rnn_k"][n]
        out[f"forecaster.rnn{idx}._conv.weight"] = (4 * mid, in_c + mid, rk, rk)
        out[f"forecaster.rnn{idx}._conv.bias"] = (4 * mid,)
        k = cfg["dec_conv_k"][n]
        name = cfg["dec_conv_names"][n]
        out[f"forecaster.stage{idx}.{name}.weight"] = (mid, oc, k, k)
        out[f"forecaster.stage{idx}.{name}.bias"] = (oc,)
        in_c = oc
    out[f"forecaster.stage1.{cfg['final_conv_2_name']}.weight"] = (c, in_c, 1, 1)
    out[f"forecaster.stage1.{cfg['final_conv_2_name']}.bias"] = (c,)
    return out


def predrnn_shapes(img_shape, cfg=None):
    """SURVEY.md App. B; models/predrnn_v2.py:92-119, model_blocks/predrnn.py:41-55."""
    cfg = {**PREDRNN_DEFAULTS, **(cfg or {})}
    c = img_shape[0]
    p, L, hid, k = cfg["patch_size"], cfg["num_layers"], cfg["num_hidden"], cfg["filter_size"]
    out = {}
    ac = bool(cfg.get("action_conditional"))            # predrnn_v2.py:65-90 and model_blocks/predrnn.py:97-139
    hp, wp = img_shape[1] // p, img_shape[2] // p
    if ac:
        a, C0, CL = cfg["action_size"], hid[0], hid[L - 1]
        hp, wp = hp // 4, wp // 4
        out["conv_input1.weight"] = (C0 // 2, p * p * c, k, k)
        out["conv_input2.weight"] = (C0, C0 // 2, k, k)
        out["action_conv_input1.weight"] = (C0 // 2, a, k, k)
        out["action_conv_input2.weight"] = (C0, C0 // 2, k, k)
        out["deconv_output1.weight"] = (CL, CL // 2, k, k)
        out["deconv_output2.weight"] = (CL // 2, p * p * c, k, k)
    for i in range(L):
        cin = (hid[0] if ac else p * p * c) if i == 0 else hid[i - 1]
        C = hid[i]
        convs = [("x", 7, cin), ("h", 4, C)] + ([("a", 4, C)] if ac else []) + [("m", 3, C), ("o", 1, 2 * C)]
        for name, mult, ci in convs:
            out[f"cell_list.{i}.conv_{name}.0.weight"] = (mult * C, ci, k, k)
            if ac:
                out[f"cell_list.{i}.conv_{name}.0.bias"] = (mult * C,)
            if cfg.get("layer_norm"):                   # nn.LayerNorm([k * C, H', W']) after every conv
                out[f"cell_list.{i}.conv_{name}.1.weight"] = (mult * C, hp, wp)
                out[f"cell_list.{i}.conv_{name}.1.bias"] = (mult * C, hp, wp)
        out[f"cell_list.{i}.conv_last.weight"] = (C, 2 * C, 1, 1)
        if ac:
            out[f"cell_list.{i}.conv_last.bias"] = (C,)
    if not ac:
        out["conv_last.weight"] = (p * p * c, hid[L - 1], 1, 1)
    ca = hid[L - 1] if ac else hid[0]
    out["adapter.weight"] = (ca, ca, 1, 1)
    return out


def phydnet_shapes(img_shape, cfg=None):
    """SURVEY.md App. B; models/phydnet.py:38-63, model_blocks/{enc,conv,phydnet}.py."""
    cfg = {**PHYDNET_DEFAULTS, **(cfg or {})}
    c = img_shape[0]
    out = {}

    def dcgan(prefix, cin, cout, transpose):
        out[prefix + "main.0.weight"] = (cin, cout, 3, 3) if transpose else (cout, cin, 3, 3)
        out[prefix + "main.0.bias"] = (cout,)
        out[prefix + "main.1.weight"] = (cout,)
        out[prefix + "main.1.bias"] = (cout,)

    dcgan("encoder_E.c1.", c, 32, False)
    dcgan("encoder_E.c2.", 32, 32, False)
    dcgan("encoder_E.c3.", 32, 64, False)
    for e in ("encoder_Ep.", "encoder_Er."):
        dcgan(e + "c1.", 64, 64, False)
        dcgan(e + "c2.", 64, 64, False)
    for d in ("decoder_Dp.", "decoder_Dr."):
        dcgan(d + "upc1.", 64, 64, True)
        dcgan(d + "upc2.", 64, 64, True)
    dcgan("decoder_D.upc1.", 64, 32, True)
    dcgan("decoder_D.upc2.", 32, 32, True)
    out["decoder_D.upc3.weight"] = (32, c, 3, 3)
    out["decoder_D.upc3.bias"] = (c,)
    hidp, kp = cfg["phycell_channels"], cfg["phycell_kernel_size"][0]
    for j in range(cfg["phycell_n_layers"]):
        pre = f"phycell.cell_list.{j}."
        out[pre + "F.conv1.weight"] = (hidp, 64, kp, kp)
        out[pre + "F.conv1.bias"] = (hidp,)
        out[pre + "F.bn1.weight"] = (hidp,)
        out[pre + "F.bn1.bias"] = (hidp,)
        out[pre + "F.conv2.weight"] = (64, hidp, 1, 1)
        out[pre + "F.conv2.bias"] = (64,)
        out[pre + "convgate.weight"] = (64, 128, 3, 3)
        out[pre + "convgate.bias"] = (64,)
        if cfg.get("action_conditional"):               # model_blocks/phydnet.py:44-48
            for nm in ("frame_action_conv", "hidden_action_conv"):
                out[pre + nm + ".weight"] = (64, 64 + cfg["action_size"], 1, 1)
                out[pre + nm + ".bias"] = (64,)
    cin = 64 + (cfg["action_size"] if cfg.get("action_conditional") else 0)
    kc = cfg["convlstm_kernel_size"][0]
    for j, hd in enumerate(cfg["convlstm_hidden_dims"][:cfg["convlstm_n_layers"]]):
        out[f"convcell.cell_list.{j}.conv.weight"] = (4 * hd, cin + hd, kc, kc)
        out[f"convcell.cell_list.{j}.conv.bias"] = (4 * hd,)
        cin = hd
    return out


def stphy_shapes(img_shape, cfg=None):
    """models/st_phy.py:38-83, model_blocks/enc.py:14-98; action_conditional=True adds action_inflate / action_conv_h /
    action_conv_w (:48-56), the ActionConditionalSpatioTemporalLSTMCell layout (conv biases, conv_a;
    model_blocks/predrnn.py:97-139) and PhyCell's two 1x1 action convs (model_blocks/phydnet.py:44-48)."""
    cfg = {**STPHY_DEFAULTS, **(cfg or {})}
    c, h, w = img_shape
    L, C, hid, kp = cfg["num_layers"], cfg["st_cell_channels"], cfg["phycell_channels"], cfg["phycell_kernel_size"][0]
    eh = ((h - 5) // 2 + 1 - 3) // 2 + 1 - 2
    ew = ((w - 5) // 2 + 1 - 3) // 2 + 1 - 2
    ac = bool(cfg.get("action_conditional"))
    a = cfg.get("action_size", 0)
    out = {}
    for name, shp in (("encoder.conv1", (32, c, 5, 5)), ("encoder.conv2", (64, 32, 3, 3)), ("encoder.mean_layer", (C, 64, 3, 3)),
                      ("decoder.fc1", (C, C, 1, 1)), ("decoder.conv1", (C, 64, 6, 6)), ("decoder.conv2", (64, 32, 6, 6)),
                      ("decoder.conv3", (32, c, 5, 5))):
        out[f"autoencoder.{name}.weight"] = shp
        out[f"autoencoder.{name}.bias"] = (shp[1],) if "decoder.conv" in name else (shp[0],)
    if ac:
        ia = cfg.get("inflated_action_dim", 3)
        out["action_inflate.weight"] = (ia * eh * ew, a)
        out["action_conv_h.weight"] = (C, ia, 5, 1)
        out["action_conv_w.weight"] = (C, ia, 1, 5)
    for i in range(L):
        convs = (("x", 7, C), ("h", 4, C), ("a", 4, C), ("m", 3, C), ("o", 1, 2 * C)) if ac else \
                (("x", 7, C), ("h", 4, C), ("m", 3, C), ("o", 1, 2 * C))
        for nm, mult, ci in convs:
            out[f"st_cell_list.{i}.conv_{nm}.0.weight"] = (mult * C, ci, 5, 5)
            if ac:
                out[f"st_cell_list.{i}.conv_{nm}.0.bias"] = (mult * C,)
            out[f"st_cell_list.{i}.conv_{nm}.1.weight"] = (mult * C, eh, ew)
            out[f"st_cell_list.{i}.conv_{nm}.1.bias"] = (mult * C, eh, ew)
        out[f"st_cell_list.{i}.conv_last.weight"] = (C, 2 * C, 1, 1)
        if ac:
            out[f"st_cell_list.{i}.conv_last.bias"] = (C,)
    for i in range(L):
        pre = f"phycell_list.{i}."
        out[pre + "F.conv1.weight"] = (hid, C, kp, kp)
        out[pre + "F.conv1.bias"] = (hid,)
        out[pre + "F.bn1.weight"] = (hid,)
        out[pre + "F.bn1.bias"] = (hid,)
        out[pre + "F.conv2.weight"] = (C, hid, 1, 1)
        out[pre + "F.conv2.bias"] = (C,)
        out[pre + "convgate.weight"] = (C, 2 * C, 3, 3)
        out[pre + "convgate.bias"] = (C,)
        if ac:
            for nm in ("frame_action_conv", "hidden_action_conv"):
                out[pre + nm + ".weight"] = (C, C + a, 1, 1)
                out[pre + nm + ".bias"] = (C,)
    for i in range(L):
        out[f"hidden_conv_list.{i}.weight"] = (C, 2 * C, 1, 1)
        if i < L - 1:
            out[f"hidden_conv_list.{i}.bias"] = (C,)
    out["adapter.weight"] = (C, C, 1, 1)
    return out


def trajgru_shapes(img_shape, cfg=None):
    """models/precipitation_nowcasting/ef_traj_gru.py:30-119 at its default hyper-parameters (L = 13 flows, i2h k3)."""
    base = ef_shapes(img_shape, cfg)
    cfg = {**EF_DEFAULTS, **(cfg or {})}
    out = {}
    for k, v in base.items():
        parts = k.split(".")
        if parts[1].startswith("rnn"):
            continue
        out[k] = v

    def rnn(prefix, in_c, c, L=13, k=3):
        out[prefix + "i2h.weight"] = (3 * c, in_c, k, k)
        out[prefix + "i2h.bias"] = (3 * c,)
        out[prefix + "i2f_conv1.weight"] = (32, in_c, 5, 5)
        out[prefix + "i2f_conv1.bias"] = (32,)
        out[prefix + "h2f_conv1.weight"] = (32, c, 5, 5)
        out[prefix + "h2f_conv1.bias"] = (32,)
        out[prefix + "flows_conv.weight"] = (2 * L, 32, 5, 5)
        out[prefix + "flows_conv.bias"] = (2 * L,)
        out[prefix + "ret.weight"] = (3 * c, c * L, 1, 1)
        out[prefix + "ret.bias"] = (3 * c,)
    in_c = img_shape[0]
    for n in range(3):
        mid, oc = cfg["enc_c"][2 * n], cfg["enc_c"][2 * n + 1]
        rnn(f"encoder.rnn{n + 1}.", mid, oc)
        in_c = oc
    for n in range(3):
        mid, oc = cfg["dec_c"][2 * n], cfg["dec_c"][2 * n + 1]
        rnn(f"forecaster.rnn{3 - n}.", in_c, mid)
        in_c = oc
    return out


SHAPES = {"trajgru": trajgru_shapes, "st-phy": stphy_shapes, "convlstm-shi": ef_shapes, "predrnn-pp": predrnn_shapes, "phy": phydnet_shapes,
          "convlstm-branch": phydnet_shapes}
