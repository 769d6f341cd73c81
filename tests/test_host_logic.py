"""Host side of the drop-in boundary (no GPU needed): constructor contract, config, state_dict layouts, error
behaviour -- compared with the reference's own listings (tests/golden/manifest.json) and, when the reference checkout
is mounted (authoring container), with the reference classes themselves."""
import pytest
import torch

import vp_suite_b200 as V
from oracle import ref_shim

KW = dict(action_size=0, tensor_value_range=[0.0, 1.0])


@pytest.mark.parametrize("name", ["ef_1x64", "ef_3x32", "predrnn_1x64", "predrnn_3x32", "phy_3x64", "phy_1x64",
                                  "branch_1x64", "predrnn_ln_1x64", "predrnn_ln_3x32", "predrnn_ac_1x64",
                                  "predrnn_acln_3x32", "phy_ac_3x64", "stphy_3x64", "stphy_ac_3x64", "trajgru_1x64", "trajgru_3x32"])
def test_state_dict_layout_matches_reference(manifest, name):
    meta = manifest["models"][name]
    m = V.MODEL_CLASSES[meta["key"]]("cpu", img_shape=tuple(meta["img_shape"]), **{**KW, **(meta.get("model_kwargs") or {})})
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == meta["shapes"]
    # and the native library expects exactly these tensors
    native = m.native_param_layout()
    assert {m._native_key(k): tuple(v) for k, v in got.items()} == native


def test_registry_keys_follow_the_reference():
    assert {"convlstm-shi", "predrnn-pp", "phy"} <= set(V.MODEL_CLASSES)
    reg = V.register_into({"copy": object})
    assert reg["phy"] is V.MODEL_CLASSES["phy"] and "copy" in reg


def test_constructor_contract():
    cls = V.MODEL_CLASSES["convlstm-shi"]
    with pytest.raises(ValueError):                                     # missing required arg (base_model.py:51-64)
        cls("cpu", img_shape=(1, 64, 64), action_size=0)
    with pytest.raises(ValueError):                                     # type check against the class default
        cls("cpu", img_shape=(1, 64, 64), num_layers="3", **KW)
    with pytest.raises(ValueError):
        cls("cpu", img_shape=(1, 64, 64), action_size=0, tensor_value_range=1.0)
    with pytest.raises(AttributeError):                                 # per-layer list length (ef_blocks.py:134-143)
        cls("cpu", img_shape=(1, 64, 64), enc_conv_k=[3, 3], **KW)
    with pytest.raises(AttributeError):                                 # wrong output size (ef_blocks.py:160-167)
        cls("cpu", img_shape=(1, 64, 64), dec_conv_k=[4, 4, 5], **KW)
    m = cls("cpu", img_shape=(3, 32, 32), **KW)
    cfg = m.config
    assert cfg["img_c"] == 3 and cfg["img_h"] == 32 and cfg["NAME"] == cls.NAME
    assert cfg["enc_c"] == [16, 64, 64, 96, 96, 96] and cfg["tensor_value_range"] == [0.0, 1.0]
    assert not any(k.startswith("_") for k in cfg)                      # native handles stay out of run_cfg.json
    assert "encoder" not in cfg and "forecaster" not in cfg
    assert m.enc_rnn_state_h == [32, 16, 8] and m.dec_rnn_state_h == [8, 16, 32]


def test_class_constants_match_reference():
    P = V.MODEL_CLASSES["predrnn-pp"]
    assert P.NEEDS_COMPLETE_INPUT and not P.CAN_HANDLE_ACTIONS and P.NAME == "PredRNN++"
    assert V.MODEL_CLASSES["phy"].CAN_HANDLE_ACTIONS
    assert V.MODEL_CLASSES["convlstm-shi"].MIN_CONTEXT_FRAMES == 1


def test_forward_refuses_cpu_tensors():
    m = V.MODEL_CLASSES["convlstm-shi"]("cpu", img_shape=(1, 32, 32), **KW)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.rand(1, 2, 1, 32, 32), pred_frames=1)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vp_suite_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "libvpk.so"))
    with pytest.raises(_native.NativeError, match="no fallback"):
        _native.lib()


def test_ef_loads_cuda_built_checkpoint_without_peepholes(manifest):
    meta = manifest["models"]["ef_3x32"]
    m = V.MODEL_CLASSES["convlstm-shi"]("cpu", img_shape=(3, 32, 32), **KW)
    sd = {k: torch.ones(v) for k, v in meta["shapes"].items() if k.rsplit(".", 1)[-1] not in ("Wci", "Wcf", "Wco")}
    assert len(sd) == 26
    m.load_state_dict(sd)
    assert float(m.encoder.rnn1.Wci.detach().abs().sum()) == 0.0


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
@pytest.mark.parametrize("key,img,extra", [
    ("convlstm-shi", (1, 64, 64), {}), ("predrnn-pp", (1, 64, 64), {}), ("phy", (3, 64, 64), {}),
    ("predrnn-pp", (1, 64, 64), {"action_conditional": True, "action_size": 3}),
    ("predrnn-pp", (3, 32, 32), {"action_conditional": True, "action_size": 4, "layer_norm": True}),
    ("phy", (3, 64, 64), {"action_conditional": True, "action_size": 3}),
    ("st-phy", (3, 64, 64), {}), ("st-phy", (3, 64, 64), {"action_conditional": True, "action_size": 3}),
    ("trajgru", (1, 64, 64), {})])
def test_same_seed_gives_the_reference_init_and_config(key, img, extra):
    ref_cls = ref_shim.load_reference()[key]
    kw = {**KW, **extra}
    torch.manual_seed(7)
    ref = ref_cls("cpu", img_shape=img, **kw)
    torch.manual_seed(7)
    ours = V.MODEL_CLASSES[key]("cpu", img_shape=img, **kw)
    rsd, osd = ref.state_dict(), ours.state_dict()
    assert list(rsd) == list(osd)
    for k in rsd:
        assert torch.equal(rsd[k], osd[k]), k
    ours.load_state_dict(rsd)                                           # reference weights load unchanged
    rc, oc = ref.config, ours.config
    for k, v in rc.items():                                             # every reference config entry is present
        if k in ("device", "shape_Ep", "shape_Er", "constraints"):
            continue
        assert k in oc, k
        assert oc[k] == v or k in ("NAME", "recurrent_cell", "activation"), (k, oc[k], v)      # (recurrent_cell: the drop-in block class)
