"""The reference-side half of the drop-in boundary (no GPU needed): base class, `unpack_data`, `eval_iter`, `train_iter`,
registration into the reference's MODEL_CLASSES, whole-module pickling (vpsuite.py:394,135)."""
import io
import os
import subprocess
import sys

import pytest
import torch

import vp_suite_b200 as V
from vp_suite_b200 import base as VB
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(action_size=0, tensor_value_range=[0.0, 1.0])
needs_ref = pytest.mark.skipif(not ref_shim.available() or bool(os.environ.get("VPK_NO_REFERENCE")),
                               reason="reference package not present")


def _vpdata(b, t, c, h, w, a=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"frames": torch.rand((b, t, c, h, w), generator=g), "actions": torch.rand((b, t - 1, a), generator=g),
            "origin": "synthetic"}


@needs_ref
def test_dropins_subclass_the_real_base_and_register():
    import vp_suite.base
    from vp_suite.models import MODEL_CLASSES as REF_CLASSES
    assert VB.REFERENCE_BASE
    for key in ("convlstm-shi", "predrnn-pp", "phy"):
        m = V.MODEL_CLASSES[key]("cpu", img_shape=(3, 64, 64), **KW)
        assert isinstance(m, vp_suite.base.VPModel)
        assert isinstance(m, torch.nn.Module)
    from vp_suite_b200 import model_blocks as MB
    assert issubclass(MB.ConvLSTM, vp_suite.base.VPModelBlock)
    reg = dict(REF_CLASSES)
    V.register_into(reg)
    assert reg["phy"] is V.MODEL_CLASSES["phy"] and reg["copy"] is REF_CLASSES["copy"]
    # the reference's own creation call (vpsuite.py:170): model_class(device, **model_kwargs).to(device)
    m = reg["convlstm-shi"]("cpu", img_shape=(1, 32, 32), **KW).to("cpu")
    assert m.config["img_c"] == 1 and m.NAME == REF_CLASSES["convlstm-shi"].NAME


@needs_ref
@pytest.mark.parametrize("key,complete", [("convlstm-shi", False), ("predrnn-pp", True), ("phy", False)])
def test_unpack_data_is_the_references(key, complete):
    ref_cls = ref_shim.load_reference()[key]
    ours = V.MODEL_CLASSES[key]("cpu", img_shape=(1, 32, 32), **KW)
    ref = ref_cls("cpu", img_shape=(1, 32, 32), **KW)
    cfg = {"device": "cpu", "context_frames": 3, "pred_frames": 2}
    for data in (_vpdata(2, 6, 1, 32, 32), {k: (v[0] if torch.is_tensor(v) else v) for k, v in _vpdata(1, 6, 1, 32, 32).items()}):
        for reverse in (False, True):
            a, b = ours.unpack_data(data, cfg, reverse=reverse), ref.unpack_data(data, cfg, reverse=reverse)
            assert all(torch.equal(x, y) for x, y in zip(a, b))
            assert a[0].shape[1] == (5 if complete else 3) and a[1].shape[1] == 2


def test_train_iter_refuses_clearly_and_model_is_marked_untrainable():
    m = V.MODEL_CLASSES["phy"]("cpu", img_shape=(3, 64, 64), **KW)
    assert m.TRAINABLE is False                        # VPSuite.train skips training for such models (vpsuite.py:312)
    with pytest.raises(NotImplementedError, match="inference-only"):
        m.train_iter({}, [], None, None, 0)
    # the one trainable drop-in: EF_ConvLSTM (differentiable ConvLSTM layers; tests/test_gpu_backward.py trains it)
    assert V.MODEL_CLASSES["convlstm-shi"].TRAINABLE is True


@pytest.mark.parametrize("key", ["convlstm-shi", "predrnn-pp", "phy", "convlstm-branch"])
def test_whole_module_pickle_round_trip(key):
    """torch.save(model) / torch.load (vpsuite.py:394,135): the native handle is dropped and rebuilt lazily."""
    m = V.MODEL_CLASSES[key]("cpu", img_shape=(1, 32, 32), **KW)
    m.native_param_layout()                           # creates a native handle: it must not end up in the pickle
    assert m._handle is not None
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert type(m2) is type(m) and m2._handle is None and m2._workspaces == {}
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert m2.config == m.config
    assert m2.native_param_layout() == m.native_param_layout()


def test_block_pickle_round_trip():
    from vp_suite_b200 import model_blocks as MB
    blk = MB.ConvLSTM("cpu", 8, 16, 12, 10, 3)
    buf = io.BytesIO()
    torch.save(blk, buf)
    buf.seek(0)
    blk2 = torch.load(buf, weights_only=False)
    assert blk2._cell is None and torch.equal(blk2._conv.weight, blk._conv.weight)


MIRROR_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %(root)r)
import vp_suite_b200 as V
from vp_suite_b200 import base as VB
assert not VB.REFERENCE_BASE and "vp_suite" not in sys.modules

class Stub(V.VPModel):                      # a CPU stand-in for forward(): copies the last frame
    NAME = "stub"
    def forward(self, x, pred_frames=1, **kw):
        return x[:, -1:].repeat(1, pred_frames, 1, 1, 1), None

class Losses:                               # what eval_iter needs of PredictionLossProvider (loss_provider.py:30-53)
    def get_losses(self, pred, target):
        v = ((pred - target) ** 2).sum(dim=(4, 3, 2)).mean(dim=1).mean(dim=0)
        return {"mse": v}, v

m = Stub("cpu", img_shape=(1, 8, 8), action_size=0, tensor_value_range=[0.0, 1.0])
g = torch.Generator().manual_seed(0)
loader = [{"frames": torch.rand((2, 5, 1, 8, 8), generator=g), "actions": torch.zeros(2, 4, 0)} for _ in range(3)]
cfg = {"device": "cpu", "context_frames": 3, "pred_frames": 2, "val_rec_criterion": "mse"}
inp, tgt, act = m.unpack_data(loader[0], cfg)
assert inp.shape == (2, 3, 1, 8, 8) and tgt.shape == (2, 2, 1, 8, 8) and act.shape == (2, 4, 0)
assert torch.equal(torch.cat([inp, tgt], 1), loader[0]["frames"])
inp_r, _, _ = m.unpack_data(loader[0], cfg, reverse=True)
assert torch.equal(inp_r, torch.flip(loader[0]["frames"], dims=[1])[:, :3])
inp_c, tgt_c, _ = m.unpack_data(loader[0], cfg, complete=True)
assert inp_c.shape[1] == 5 and torch.equal(tgt_c, inp_c[:, 3:])
all_losses, indicator = m.eval_iter(cfg, loader, Losses())
want = torch.stack([Losses().get_losses(d["frames"][:, 2:3].repeat(1, 2, 1, 1, 1), d["frames"][:, 3:5])[1] for d in loader]).mean()
assert abs(all_losses["mse"] - float(want)) < 1e-6 and abs(float(indicator) - float(want)) < 1e-6
assert m.training                            # eval_iter puts the module back into train mode (base_model.py:214)
try:
    m.train_iter(cfg, loader, None, Losses(), 0)
    raise SystemExit("train_iter did not refuse")
except NotImplementedError:
    pass
print("MIRROR_OK")
"""


def test_mirror_base_without_the_reference_package():
    """Without an importable vp_suite the drop-ins carry their own unpack_data / eval_iter (base_model.py:87-114,181-216)."""
    env = dict(os.environ, VPK_NO_REFERENCE="1")
    out = subprocess.run([sys.executable, "-c", MIRROR_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "MIRROR_OK" in out.stdout, out.stdout + out.stderr


@needs_ref
def test_eval_iter_matches_reference_eval_iter_on_a_stub_forward():
    """The inherited eval_iter is the reference's: same numbers as the reference class with the same stub forward and the
    reference's own PredictionLossProvider."""
    import vp_suite.base
    from vp_suite.measure.loss_provider import PredictionLossProvider

    def fwd(self, x, pred_frames=1, **kw):
        return x[:, -1:].repeat(1, pred_frames, 1, 1, 1), None
    Ours = type("Ours", (V.VPModel,), {"forward": fwd, "NAME": "stub"})
    Ref = type("Ref", (vp_suite.base.VPModel,), {"forward": fwd, "NAME": "stub"})
    cfg = {"device": "cpu", "context_frames": 3, "pred_frames": 2, "val_rec_criterion": "mse",
           "losses_and_scales": {"mse": 1.0, "l1": 0.5}, "img_c": 3}
    loader = [_vpdata(2, 5, 3, 8, 8, seed=s) for s in range(3)]
    res = []
    for cls in (Ours, Ref):
        m = cls("cpu", img_shape=(3, 8, 8), **KW)
        res.append(m.eval_iter(cfg, loader, PredictionLossProvider(dict(cfg))))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
