"""The oracle (CPU restatement) against golden vectors produced by the real reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import blocks as OB
from oracle import models as OM
from oracle.weights import synth_state_dict, synth_frames, synth_actions

ATOL = 2e-5      # same arithmetic (ATen CPU fp32) in a different call structure


def _case(manifest, name):
    meta = manifest["models"][name]
    sd = synth_state_dict(meta["shapes"], meta["wseed"], meta["gain"])
    t = meta["context"] + (meta["pred"] if meta["key"] == "predrnn-pp" else 0)
    x = synth_frames(meta["batch"], t, *meta["img_shape"], seed=meta["xseed"])
    return meta, sd, x


@pytest.mark.parametrize("name", ["ef_1x64", "ef_3x32", "predrnn_1x64", "predrnn_3x32", "phy_3x64", "phy_1x64",
                                  "branch_1x64", "predrnn_ln_1x64", "predrnn_ln_3x32", "stphy_3x64", "trajgru_1x64", "trajgru_3x32"])
def test_model_rollout_matches_reference(manifest, name):
    meta, sd, x = _case(manifest, name)
    gold = load_golden(name)
    with torch.no_grad():
        pred, aux = OM.FORWARDS[meta["key"]](sd, x, meta["pred"])
    assert pred.shape == gold["pred"].shape
    err = np.abs(pred.numpy() - gold["pred"]).max()
    assert err <= ATOL, f"{name}: max abs err {err}"
    assert gold["pred"].std() > 0.02           # the vector is not near-constant (tolerances are meaningful)
    if "loss" in gold:
        (k, v), = aux.items()
        assert k == "ST-LSTM decouple loss"
        assert abs(float(v) - float(gold["loss"])) <= 1e-4 * max(1.0, abs(float(gold["loss"])))
    else:
        assert aux is None


@pytest.mark.parametrize("name", ["predrnn_ac_1x64", "predrnn_acln_3x32", "phy_ac_3x64", "stphy_ac_3x64"])
def test_action_conditional_rollout_matches_reference(manifest, name):
    """model(x, pred_frames, actions=a) of the reference's action-conditional predrnn-pp (incl. layer_norm) and phy."""
    meta, sd, x = _case(manifest, name)
    actions = synth_actions(meta["batch"], meta["context"] + meta["pred"] - 1, meta["action_size"], seed=meta["aseed"])
    gold = load_golden(name)
    with torch.no_grad():
        pred, aux = OM.FORWARDS[meta["key"]](sd, x, meta["pred"], actions=actions)
    assert np.abs(pred.numpy() - gold["pred"]).max() <= ATOL
    assert gold["pred"].std() > 0.02
    if "loss" in gold:
        assert abs(float(list(aux.values())[0]) - float(gold["loss"])) <= 1e-4 * max(1.0, abs(float(gold["loss"])))
    with pytest.raises(ValueError):                      # predrnn_v2.py:149-151, models/phydnet.py:103-105
        OM.FORWARDS[meta["key"]](sd, x, meta["pred"], actions=actions[..., :-1])


def test_ef_missing_peepholes_are_zero(manifest):
    """CUDA-constructed reference checkpoints lack Wci/Wcf/Wco: treated as zeros."""
    meta, sd, x = _case(manifest, "ef_3x32")
    sd0 = {k: (torch.zeros_like(v) if k.rsplit(".", 1)[-1] in ("Wci", "Wcf", "Wco") else v) for k, v in sd.items()}
    sdm = {k: v for k, v in sd.items() if k.rsplit(".", 1)[-1] not in ("Wci", "Wcf", "Wco")}
    with torch.no_grad():
        a, _ = OM.ef_convlstm_forward(sd0, x, 2)
        b, _ = OM.ef_convlstm_forward(sdm, x, 2)
    assert torch.equal(a, b)


def test_blocks_match_reference(manifest):
    gold = load_golden("blocks")
    mb = manifest["blocks"]
    with torch.no_grad():
        # hzzone ConvLSTM sequence driver
        sd = synth_state_dict(mb["hzzone"]["shapes"], mb["hzzone"]["wseed"])
        xin = torch.rand(tuple(mb["hzzone"]["x_shape"]), generator=torch.Generator().manual_seed(5)) * 2 - 1
        args = (sd["_conv.weight"], sd["_conv.bias"], sd["Wci"], sd["Wcf"], sd["Wco"], 8)
        o1, (h1, c1) = OB.convlstm_shi_sequence(xin, None, 3, *args)
        o2, (h2, c2) = OB.convlstm_shi_sequence(None, (h1, c1), 2, *args)
        for k, v in dict(hz_out1=o1, hz_h1=h1, hz_c1=c1, hz_out2=o2, hz_c2=c2).items():
            assert np.abs(v.numpy() - gold[k]).max() <= ATOL, k

        # ndrplz cell
        sd = synth_state_dict(mb["ndrplz"]["shapes"], mb["ndrplz"]["wseed"])
        g = torch.Generator().manual_seed(6)
        x, h, c = (torch.rand((2, 8, 10, 10), generator=g) * 2 - 1, torch.rand((2, 16, 10, 10), generator=g) * 2 - 1,
                   torch.rand((2, 16, 10, 10), generator=g) * 2 - 1)
        hn, cn = OB.convlstm_cell_step(x, h, c, sd["conv.weight"], sd["conv.bias"])
        assert np.abs(hn.numpy() - gold["nd_h"]).max() <= ATOL
        assert np.abs(cn.numpy() - gold["nd_c"]).max() <= ATOL

        # ST-LSTM cell
        sd = synth_state_dict(mb["stlstm"]["shapes"], mb["stlstm"]["wseed"])
        g = torch.Generator().manual_seed(7)
        x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
        h, c, m = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
        res = OB.stlstm_step(x, h, c, m, sd["conv_x.0.weight"], sd["conv_h.0.weight"], sd["conv_m.0.weight"],
                             sd["conv_o.0.weight"], sd["conv_last.weight"])
        for nm, v in zip(("st_h", "st_c", "st_m", "st_dc", "st_dm"), res):
            assert np.abs(v.numpy() - gold[nm]).max() <= ATOL, nm

        # ST-LSTM cell with layer_norm=True
        sd = synth_state_dict(mb["stlstm_ln"]["shapes"], mb["stlstm_ln"]["wseed"])
        g = torch.Generator().manual_seed(9)
        x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
        h, c, m = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
        ln = {k: (sd[f"conv_{k}.1.weight"], sd[f"conv_{k}.1.bias"]) for k in "xhmo"}
        res = OB.stlstm_step(x, h, c, m, sd["conv_x.0.weight"], sd["conv_h.0.weight"], sd["conv_m.0.weight"],
                             sd["conv_o.0.weight"], sd["conv_last.weight"], ln=ln)
        for nm, v in zip(("stln_h", "stln_c", "stln_m", "stln_dc", "stln_dm"), res):
            assert np.abs(v.numpy() - gold[nm]).max() <= ATOL, nm

        # PhyCell cell
        sd = synth_state_dict(mb["phycell"]["shapes"], mb["phycell"]["wseed"])
        g = torch.Generator().manual_seed(8)
        x, h = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1, torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
        hn = OB.phycell_step(x, h, sd)
        assert np.abs(hn.numpy() - gold["phy_h"]).max() <= ATOL


@pytest.mark.parametrize("tag, key, xseed", [("stac", "stlstm_ac", 10), ("stacln", "stlstm_acln", 11)])
def test_action_conditional_stlstm_cell_matches_reference(manifest, tag, key, xseed):
    """ActionConditionalSpatioTemporalLSTMCell (predrnn.py:86-169), layer_norm off / on: the oracle restatement against
    vectors the reference produced (oracle/make_golden.py run_blocks_ac).  Oracle only: no kernel path yet."""
    gold = load_golden("blocks_ac")
    mb = manifest["blocks"][key]
    sd = synth_state_dict(mb["shapes"], mb["wseed"])
    g = torch.Generator().manual_seed(xseed)
    x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
    h, c, m, a = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(4)]
    with torch.no_grad():
        res = OB.stlstm_ac_step(x, h, c, m, a, sd)
    for nm, v in zip(("h", "c", "m", "dc", "dm"), res):
        assert np.abs(v.numpy() - gold[f"{tag}_{nm}"]).max() <= ATOL, nm


def test_group_norm_divisor():
    # phydnet.py:348-362: 49 -> 7 groups, 64 -> 8
    assert OB.find_divisor_for_group_norm(49) == 7
    assert OB.find_divisor_for_group_norm(64) == 8


def test_patch_roundtrip():
    x = torch.rand(2, 3, 3, 16, 24)
    assert torch.equal(OM.reshape_patch_back(OM.reshape_patch(x, 4), 4), x)


def test_shape_listing_matches_reference(manifest):
    from oracle.shapes import SHAPES
    for name, meta in manifest["models"].items():
        want = {k: tuple(v) for k, v in meta["shapes"].items()}
        kw = meta.get("model_kwargs") or {}
        got = SHAPES[meta["key"]](tuple(meta["img_shape"]), kw) if kw else SHAPES[meta["key"]](tuple(meta["img_shape"]))
        assert got == want, (name, set(got.items()) ^ set(want.items()))
