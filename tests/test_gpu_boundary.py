"""GPU tests of the reference-side boundary: the reference's own call sequences drive the registered drop-ins.

  * vpsuite.py:536-550 (`unpack_data -> model.eval() -> model(input, pred_frames) -> get_metrics`) with the reference's
    PredictionMetricProvider and with the on-device NativeMetricProvider;
  * the reference's model test protocol (tests/test_models.py:19-35: 3x64x64, b = 2, p = 5, t = 3 or p + 3, randn input,
    `pred_1` and `forward`), with values checked against the oracle in both precision modes;
  * whole-module pickling (vpsuite.py:394,135) and `eval_iter` (base_model.py:181-216) on the device;
  * the stateful `PhyCell` block drop-in against the reference's block (or the oracle's cell step).
"""
import io
import os

import numpy as np
import pytest
import torch

from oracle import blocks as OB, models as OM, ref_shim
from oracle.shapes import SHAPES
from oracle.weights import synth_state_dict, synth_frames

pytestmark = pytest.mark.gpu
KW = dict(action_size=0, tensor_value_range=[0.0, 1.0])
HAVE_REF = ref_shim.available() and not os.environ.get("VPK_NO_REFERENCE")
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference package not present (baseline/_ref)")
GAIN = {"convlstm-shi": 2.5}


def _weights(key, img, seed=3, **kw):
    return synth_state_dict(SHAPES[key](img, kw) if kw else SHAPES[key](img), seed=seed, gain=GAIN.get(key, 1.5))


def _errs(a, b):
    d = (a - b).abs()
    return [float(d[:, t].max()) for t in range(d.shape[1])]


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["convlstm-shi", "predrnn-pp", "phy", "convlstm-branch", "st-phy", "trajgru"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reference_model_test_protocol(key, precision):
    """tests/test_models.py:19-35 of the reference, plus values."""
    import vp_suite_b200 as V
    img, b, p = (3, 64, 64), 2, 5
    cls = V.MODEL_CLASSES[key]
    model = cls("cuda:0", img_shape=img, action_size=3, temporal_dim=3, action_conditional=False,
                tensor_value_range=[0.0, 1.0], precision=precision).to("cuda:0")
    sd = _weights(key, img)
    model.load_state_dict(sd)
    t = p + 3 if cls.NEEDS_COMPLETE_INPUT else 3
    x = torch.randn((b, t, *img), generator=torch.Generator().manual_seed(17))
    with torch.no_grad():
        pred_1 = model.pred_1(x.cuda())
        pred_5, _ = model(x.cuda(), pred_frames=p)
    assert pred_1.shape == (b, *img)
    assert pred_5.shape == (b, 5, *img)
    with torch.no_grad():
        ref_5, _ = OM.FORWARDS[key](sd, x, p)
        # pred_1 = self(x, pred_frames=1)[0].squeeze(1): predrnn-pp then treats the last frame of x as the target
        ref_1 = OM.FORWARDS[key](sd, x, 1)[0][:, 0]
    e5, e1 = _errs(pred_5.cpu(), ref_5), float((pred_1.cpu() - ref_1).abs().max())
    print(f"{key} {precision}: pred_1 err {e1:.2e}, forward per-frame err {['%.1e' % e for e in e5]}")
    if precision == "fp32":
        assert e1 <= 1e-4 and max(e5) <= 1e-4, (e1, e5)
    else:
        assert e1 <= 5e-3 and e5[0] <= 5e-3 and max(e5) <= 2e-2, (e1, e5)


@pytest.mark.parametrize("key", ["convlstm-shi", "predrnn-pp", "phy", "st-phy"])
def test_reference_model_test_protocol_with_actions(key):
    """tests/test_models.py:39-60 of the reference: action_conditional = CAN_HANDLE_ACTIONS, actions passed to every model
    (models that cannot handle them ignore the keyword), shapes of pred_1 / forward -- plus values against the oracle."""
    import vp_suite_b200 as V
    from oracle.weights import synth_actions
    img, b, p, a_size = (3, 64, 64), 2, 5, 3
    cls = V.MODEL_CLASSES[key]
    model = cls("cuda:0", img_shape=img, action_size=a_size, temporal_dim=3, action_conditional=cls.CAN_HANDLE_ACTIONS,
                tensor_value_range=[0.0, 1.0], precision="fp32").to("cuda:0")
    sd = synth_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=3, gain=GAIN.get(key, 1.5))
    model.load_state_dict(sd)
    t_x = p + 3 if cls.NEEDS_COMPLETE_INPUT else 3
    x = torch.randn((b, t_x, *img), generator=torch.Generator().manual_seed(18))
    a = synth_actions(b, p + 3 - 1, a_size, seed=19)
    with torch.no_grad():
        pred_1 = model.pred_1(x.cuda(), actions=a.cuda())
        pred_5, _ = model(x.cuda(), pred_frames=p, actions=a.cuda())
        ref_5, _ = OM.FORWARDS[key](sd, x, p, **({"actions": a} if cls.CAN_HANDLE_ACTIONS else {}))
    assert pred_1.shape == (b, *img)
    assert pred_5.shape == (b, 5, *img)
    errs = _errs(pred_5.cpu(), ref_5)
    print(f"{key} with actions (fp32): forward per-frame err {['%.1e' % e for e in errs]}")
    assert max(errs) <= 1e-4, errs


# ---------------------------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("key,img,ctx,pred", [("convlstm-shi", (3, 32, 32), 4, 3), ("predrnn-pp", (1, 64, 64), 3, 3),
                                              ("phy", (3, 64, 64), 2, 3)])
def test_vpsuite_test_loop_sequence_on_registered_dropins(key, img, ctx, pred):
    """vpsuite.py:536-550 with the drop-in registered into the reference's MODEL_CLASSES, next to the reference model
    itself on the CPU; metrics from the reference's own PredictionMetricProvider and from NativeMetricProvider."""
    import vp_suite.base
    import vp_suite_b200 as V
    from vp_suite.models import MODEL_CLASSES as REF
    from vp_suite.measure.metric_provider import PredictionMetricProvider
    from vp_suite_b200.evaluation import NativeMetricProvider, evaluate_loader
    reg = V.register_into(dict(REF))
    ours = reg[key]("cuda:0", img_shape=img, precision="fp32", **KW).to("cuda:0")      # vpsuite.py:170
    ref = REF[key]("cpu", img_shape=img, **KW)
    assert isinstance(ours, vp_suite.base.VPModel)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed=5, gain=GAIN.get(key, 1.5))
    ref.load_state_dict(sd)
    ours.load_state_dict(ref.state_dict())
    metrics = ["mse", "psnr"] + (["fvd"] if img[0] not in (2, 3) else [])
    base_cfg = {"context_frames": ctx, "pred_frames": pred, "metrics": metrics, "img_c": img[0]}
    loader = [{"frames": synth_frames(1, ctx + pred, *img, seed=100 + i), "actions": torch.zeros(1, ctx + pred - 1, 0)}
              for i in range(2)]
    rows = {}
    for name, model, dev in (("ours", ours, "cuda:0"), ("ref", ref, "cpu")):
        cfg = {**base_cfg, "device": dev}
        prov = PredictionMetricProvider(dict(cfg))
        per_dp = []
        with torch.no_grad():
            for data in loader:
                inp, target, actions = model.unpack_data(data, cfg)
                model.eval()
                out, _ = model(inp, pred_frames=pred)
                model.train()
                assert out.device == target.device and out.shape == target.shape
                per_dp.append(prov.get_metrics(out, target, all_frame_cnts=True))
        rows[name] = per_dp
    for dp_o, dp_r in zip(rows["ours"], rows["ref"]):
        for ho, hr in zip(dp_o, dp_r):
            assert sorted(ho) == sorted(hr)
            for k in hr:
                assert abs(ho[k] - hr[k]) <= 1e-3 * abs(hr[k]) + 1e-5, (k, ho[k], hr[k])
    # the on-device evaluation caller: same loop, one pass of the reduction kernels per datapoint
    cfg = {**base_cfg, "device": "cuda:0"}
    means, per_dp = evaluate_loader(ours, loader, cfg)
    for dp_n, dp_r in zip(per_dp, rows["ref"]):
        for hn, hr in zip(dp_n, dp_r):
            for k in hr:
                assert abs(hn[k] - hr[k]) <= 1e-3 * abs(hr[k]) + 1e-5, (k, hn[k], hr[k])
    want = np.mean([dp[-1]["mse (↓)"] for dp in rows["ref"]])                          # vpsuite.py:580-585
    assert abs(means[-1]["mse (↓)"] - want) <= 1e-3 * want


@needs_ref
def test_eval_iter_on_device_with_the_reference_loss_provider():
    import vp_suite_b200 as V
    from vp_suite.measure.loss_provider import PredictionLossProvider
    img, ctx, pred = (3, 32, 32), 3, 2
    m = V.MODEL_CLASSES["convlstm-shi"]("cuda:0", img_shape=img, precision="fp32", **KW)
    sd = _weights("convlstm-shi", img)
    m.load_state_dict(sd)
    cfg = {"device": "cuda:0", "context_frames": ctx, "pred_frames": pred, "val_rec_criterion": "mse",
           "losses_and_scales": {"mse": 1.0, "l1": 1.0}, "img_c": 3}
    loader = [{"frames": synth_frames(2, ctx + pred, *img, seed=50 + i), "actions": torch.zeros(2, ctx + pred - 1, 0)}
              for i in range(2)]
    all_losses, indicator = m.eval_iter(cfg, loader, PredictionLossProvider(dict(cfg)))
    want = []
    for d in loader:
        with torch.no_grad():
            ref, _ = OM.ef_convlstm_forward(sd, d["frames"][:, :ctx], pred)
        want.append(float(((ref - d["frames"][:, ctx:]) ** 2).sum(dim=(4, 3, 2)).mean()))
    assert abs(all_losses["mse"] - np.mean(want)) <= 1e-4 * np.mean(want)
    assert abs(float(indicator) - np.mean(want)) <= 1e-4 * np.mean(want)
    assert m.training


@pytest.mark.parametrize("key", ["convlstm-shi", "predrnn-pp", "phy"])
def test_whole_module_pickle_round_trip_on_device(key):
    """torch.save(model) / torch.load (vpsuite.py:394,135): same frames after the round trip, bit for bit."""
    import vp_suite_b200 as V
    img = (1, 32, 32)
    m = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, **KW).eval()
    m.load_state_dict(_weights(key, img))
    x = synth_frames(2, 5 if key == "predrnn-pp" else 3, *img, seed=9).cuda()
    with torch.no_grad():
        a, aux_a = m(x, pred_frames=2)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False).eval()
    assert m2._handle is None
    with torch.no_grad():
        b, aux_b = m2(x, pred_frames=2)
    assert torch.equal(a, b)
    if aux_a is not None:
        assert float(list(aux_a.values())[0]) == float(list(aux_b.values())[0])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol,tol_end", [("fp32", 1e-4, 1e-4), ("bf16", 5e-3, 2e-2)])
def test_phycell_block_stack_values(precision, tol, tol_end):
    """PhyCell block drop-in (model_blocks/phydnet.py:65-114): module-held state over three timesteps, reset by
    first_timestep, against the reference's block when present, else the oracle's cell step."""
    from vp_suite_b200 import model_blocks as MB
    dev = "cuda:0"
    blk = MB.PhyCell((8, 8), 16, [49], 1, (7, 7), False, 0, dev).to(dev)
    for cell in blk.cell_list:
        cell.precision = precision
    shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    sd = synth_state_dict(shapes, seed=13)
    blk.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand((2, 16, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
    with torch.no_grad():
        got = []
        for t, x in enumerate(xs):
            H, out = blk(x.to(dev), None, first_timestep=(t == 0))
            assert out is H and len(H) == 1
            got.append(out[-1].cpu().clone())
        again, _ = blk(xs[0].to(dev), None, first_timestep=True)              # reset: same as the first call
    assert torch.equal(again[-1].cpu(), got[0])
    if HAVE_REF:
        from vp_suite.model_blocks.phydnet import PhyCell as RefPhyCell
        ref_blk = RefPhyCell((8, 8), 16, [49], 1, (7, 7), False, 0, "cpu")
        ref_blk.load_state_dict(sd)
        with torch.no_grad():
            want = [ref_blk(x, None, first_timestep=(t == 0))[1][-1].clone() for t, x in enumerate(xs)]
    else:
        h = torch.zeros(2, 16, 8, 8)
        want = []
        with torch.no_grad():
            for x in xs:
                h = OB.phycell_step(x, h, OB._sub(sd, "cell_list.0."))
                want.append(h)
    errs = [float((a - b).abs().max()) for a, b in zip(got, want)]
    print(f"PhyCell block {precision}: per-step max abs err {['%.1e' % e for e in errs]}")
    assert errs[0] <= tol and max(errs) <= tol_end, errs      # single step / end of the (three-step) rollout


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,key,xseed,ln", [("stac", "stlstm_ac", 10, False), ("stacln", "stlstm_acln", 11, True)])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 5e-3)])
def test_action_conditional_stlstm_block_matches_reference_golden(manifest, tag, key, xseed, ln, precision, tol):
    """ActionConditionalSpatioTemporalLSTMCell block drop-in (model_blocks/predrnn.py:86-169) through
    vpk_stlstm_ac_cell_step against the vectors the reference block produced (tests/golden/blocks_ac.npz)."""
    from conftest import load_golden
    from vp_suite_b200 import model_blocks as MB
    gold = load_golden("blocks_ac")
    mb = manifest["blocks"][key]
    cell = MB.ActionConditionalSpatioTemporalLSTMCell(16, 32, 8, 8, 5, 1, ln).to("cuda:0")
    cell.precision = precision
    cell.load_state_dict(synth_state_dict(mb["shapes"], mb["wseed"]))
    g = torch.Generator().manual_seed(xseed)
    x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
    h, c, m, a = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(4)]
    with torch.no_grad():
        res = cell(*[t.cuda() for t in (x, h, c, m, a)])
    errs = {nm: float(np.abs(v.cpu().numpy() - gold[f"{tag}_{nm}"]).max()) for nm, v in zip(("h", "c", "m", "dc", "dm"), res)}
    print(f"AC ST-LSTM block layer_norm={ln} {precision}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) <= tol, errs


@pytest.mark.parametrize("precision,tol,tol_end", [("fp32", 1e-4, 1e-4), ("bf16", 5e-3, 2e-2)])
def test_action_conditional_phydnet_blocks(precision, tol, tol_end):
    """PhyCell (frame / hidden action convs, model_blocks/phydnet.py:44-55) and SingleStepConvLSTM (action channels,
    :137, 153-155) block drop-ins with action_conditional=True over three timesteps, against the reference's blocks when
    present, else the oracle's cell steps."""
    from oracle.weights import synth_actions
    from vp_suite_b200 import model_blocks as MB
    dev, a_sz = "cuda:0", 3
    phy = MB.PhyCell((8, 8), 16, [49], 1, (7, 7), True, a_sz, dev).to(dev)
    lstm = MB.SingleStepConvLSTM((8, 8), 16, [32, 16], 2, (3, 3), True, a_sz, dev).to(dev)
    for cell in list(phy.cell_list) + list(lstm.cell_list):
        cell.precision = precision
    sd_p = synth_state_dict({k: tuple(v.shape) for k, v in phy.state_dict().items()}, seed=14)
    sd_l = synth_state_dict({k: tuple(v.shape) for k, v in lstm.state_dict().items()}, seed=15)
    phy.load_state_dict(sd_p)
    lstm.load_state_dict(sd_l)
    g = torch.Generator().manual_seed(4)
    xs = [torch.rand((2, 16, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
    acts = synth_actions(2, 3, a_sz, seed=21)
    got_p, got_l = [], []
    with torch.no_grad():
        for t, x in enumerate(xs):
            _, out = phy(x.to(dev), acts[:, t].to(dev), first_timestep=(t == 0))
            got_p.append(out[-1].cpu().clone())
            _, out = lstm(x.to(dev), acts[:, t].to(dev), first_timestep=(t == 0))
            got_l.append(out[-1].cpu().clone())
    if HAVE_REF:
        from vp_suite.model_blocks.phydnet import PhyCell as RefPhyCell, SingleStepConvLSTM as RefLstm
        rp = RefPhyCell((8, 8), 16, [49], 1, (7, 7), True, a_sz, "cpu")
        rl = RefLstm((8, 8), 16, [32, 16], 2, (3, 3), True, a_sz, "cpu")
        rp.load_state_dict(sd_p)
        rl.load_state_dict(sd_l)
        with torch.no_grad():
            want_p = [rp(x, acts[:, t], first_timestep=(t == 0))[1][-1].clone() for t, x in enumerate(xs)]
            want_l = [rl(x, acts[:, t], first_timestep=(t == 0))[1][-1].clone() for t, x in enumerate(xs)]
    else:
        hp = torch.zeros(2, 16, 8, 8)
        H = [torch.zeros(2, 32, 8, 8), torch.zeros(2, 16, 8, 8)]
        Cs = [torch.zeros(2, 32, 8, 8), torch.zeros(2, 16, 8, 8)]
        want_p, want_l = [], []
        with torch.no_grad():
            for t, x in enumerate(xs):
                hp = OB.phycell_step(x, hp, OB._sub(sd_p, "cell_list.0."), acts[:, t])
                want_p.append(hp)
                want_l.append(OM._convcell_stack(sd_l, "", x, H, Cs, 2, acts[:, t]).clone())
    ep = [float((a - b).abs().max()) for a, b in zip(got_p, want_p)]
    el = [float((a - b).abs().max()) for a, b in zip(got_l, want_l)]
    print(f"AC PhyCell block {precision}: {['%.1e' % e for e in ep]}; AC SingleStepConvLSTM block: {['%.1e' % e for e in el]}")
    assert ep[0] <= tol and max(ep) <= tol_end and el[0] <= tol and max(el) <= tol_end, (ep, el)
