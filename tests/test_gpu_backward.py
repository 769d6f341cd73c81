"""Gradients of the differentiable ConvLSTMCell and Shi-et-al. ConvLSTM drop-ins (vpk_convlstm_cell_backward /
vpk_convlstm_cell_backward_peep behind torch.autograd.Functions) against torch autograd on the reference modules
(baseline/_ref) or, without them, on the oracle's restatement -- fp32 mode to 1e-4 relative, 16-bit mode to a few percent;
and BPTT over two chained steps of a two-layer stack / over a peephole ConvLSTM sequence."""
import os

import pytest
import torch

from oracle import blocks as OB, ref_shim
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu
HAVE_REF = ref_shim.available() and not os.environ.get("VPK_NO_REFERENCE")


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _reference_cell(cin, ch, sd):
    if HAVE_REF:
        from vp_suite.model_blocks.conv_lstm_ndrplz import ConvLSTMCell as RefCell
        ref = RefCell(input_dim=cin, hidden_dim=ch, kernel_size=(3, 3), bias=True)
        ref.load_state_dict(sd)
        return ref, (lambda x, h, c: ref(x, (h, c))), [ref.conv.weight, ref.conv.bias]
    w = sd["conv.weight"].clone().requires_grad_(True)
    b = sd["conv.bias"].clone().requires_grad_(True)
    return None, (lambda x, h, c: OB.convlstm_cell_step(x, h, c, w, b)), [w, b]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("cin,ch,hw,batch", [(8, 16, (10, 10), 2), (64, 128, (16, 16), 3)])
def test_convlstm_cell_gradients_match_autograd(precision, tol, cin, ch, hw, batch):
    from vp_suite_b200 import model_blocks as MB
    dev = "cuda:0"
    cell = MB.ConvLSTMCell(input_dim=cin, hidden_dim=ch, kernel_size=(3, 3), bias=True).to(dev)
    cell.precision = precision
    sd = synth_state_dict({k: tuple(v.shape) for k, v in cell.state_dict().items()}, seed=41)
    cell.load_state_dict(sd)
    g = torch.Generator().manual_seed(42)
    x, h, c = (torch.rand((batch, n, *hw), generator=g) * 2 - 1 for n in (cin, ch, ch))
    r1, r2 = (torch.rand((batch, ch, *hw), generator=g) * 2 - 1 for _ in range(2))
    # reference gradients (CPU autograd)
    _, ref_step, ref_params = _reference_cell(cin, ch, sd)
    xr, hr, cr = (t.clone().requires_grad_(True) for t in (x, h, c))
    hn, cn = ref_step(xr, hr, cr)
    ((hn * r1).sum() + (cn * r2).sum()).backward()
    want = [xr.grad, hr.grad, cr.grad, ref_params[0].grad, ref_params[1].grad]
    # ours
    xo, ho, co = (t.clone().to(dev).requires_grad_(True) for t in (x, h, c))
    hn_o, cn_o = cell(xo, (ho, co))
    assert hn_o.requires_grad and cn_o.requires_grad
    ((hn_o * r1.to(dev)).sum() + (cn_o * r2.to(dev)).sum()).backward()
    got = [xo.grad, ho.grad, co.grad, cell.conv.weight.grad, cell.conv.bias.grad]
    errs = {n: _rel(a.cpu(), b) for n, a, b in zip(("dx", "dh", "dc", "dW", "db"), got, want)}
    print(f"ConvLSTMCell backward {precision} cin={cin} ch={ch}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) <= tol, errs
    # only dh_out upstream (dc_out absent) and no-grad inference still work
    xo2 = x.clone().to(dev).requires_grad_(True)
    hn2, _ = cell(xo2, (h.to(dev), c.to(dev)))
    hn2.sum().backward()
    assert xo2.grad is not None and torch.isfinite(xo2.grad).all()
    with torch.no_grad():
        hn3, _ = cell(x.to(dev), (h.to(dev), c.to(dev)))
    assert not hn3.requires_grad and torch.equal(hn3, hn2.detach())


def test_bptt_through_a_two_layer_stack_over_two_steps():
    """SingleStepConvLSTM (model_blocks/phydnet.py:117-175) built from the differentiable cells: gradients flow through both
    layers and both timesteps exactly as with the reference cells."""
    from vp_suite_b200 import model_blocks as MB
    dev = "cuda:0"
    blk = MB.SingleStepConvLSTM((8, 8), 8, [16, 8], 2, (3, 3), False, 0, dev).to(dev)
    for cell in blk.cell_list:
        cell.precision = "fp32"
    sd = synth_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=43)
    blk.load_state_dict(sd)
    g = torch.Generator().manual_seed(44)
    xs = [torch.rand((2, 8, 8, 8), generator=g) * 2 - 1 for _ in range(2)]
    r = torch.rand((2, 8, 8, 8), generator=g)
    # reference: the oracle's cell step under autograd
    ws = [sd[f"cell_list.{j}.conv.weight"].clone().requires_grad_(True) for j in range(2)]
    bs = [sd[f"cell_list.{j}.conv.bias"].clone().requires_grad_(True) for j in range(2)]
    xr = [x.clone().requires_grad_(True) for x in xs]
    H = [torch.zeros(2, 16, 8, 8), torch.zeros(2, 8, 8, 8)]
    C = [torch.zeros(2, 16, 8, 8), torch.zeros(2, 8, 8, 8)]
    for x in xr:
        inp = x
        for j in range(2):
            H[j], C[j] = OB.convlstm_cell_step(inp, H[j], C[j], ws[j], bs[j])
            inp = H[j]
    (H[1] * r).sum().backward()
    # ours
    xo = [x.clone().to(dev).requires_grad_(True) for x in xs]
    out = None
    for t, x in enumerate(xo):
        _, out = blk(x, None, first_timestep=(t == 0))
    (out[-1] * r.to(dev)).sum().backward()
    errs = {"dx0": _rel(xo[0].grad.cpu(), xr[0].grad), "dx1": _rel(xo[1].grad.cpu(), xr[1].grad)}
    for j in range(2):
        errs[f"dW{j}"] = _rel(blk.cell_list[j].conv.weight.grad.cpu(), ws[j].grad)
        errs[f"db{j}"] = _rel(blk.cell_list[j].conv.bias.grad.cpu(), bs[j].grad)
    print("BPTT 2 layers x 2 steps (fp32): " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) <= 1e-4, errs


def _shi_reference(cin, ch, hw, sd, inputs, h0, c0, seq_len):
    """Shi-et-al. ConvLSTM over a sequence under CPU autograd: the reference module when installed, else a restatement of
    conv_lstm_hzzone.py:52-69.  Returns (outputs, (h, c), leaves)."""
    w, b = sd["_conv.weight"].clone().requires_grad_(True), sd["_conv.bias"].clone().requires_grad_(True)
    peep = [sd[k].clone().requires_grad_(True) for k in ("Wci", "Wcf", "Wco")]
    if HAVE_REF:
        from vp_suite.model_blocks.conv_lstm_hzzone import ConvLSTM as RefShi
        ref = RefShi("cpu", in_channels=cin, enc_channels=ch, state_h=hw[0], state_w=hw[1], kernel_size=3, stride=1, padding=1)
        ref.load_state_dict({k: sd[k] for k in ("_conv.weight", "_conv.bias", "Wci", "Wcf", "Wco")})
        out, (h, c) = ref(inputs, (h0, c0), seq_len)
        return out, (h, c), [ref._conv.weight, ref._conv.bias, ref.Wci, ref.Wcf, ref.Wco]
    h, c, outs = h0, c0, []
    for t in range(seq_len):
        x = torch.zeros(h.shape[0], cin, *hw) if inputs is None else inputs[:, t]
        z = torch.nn.functional.conv2d(torch.cat([x, h], dim=1), w, b, padding=1)
        i, f, g, o = torch.chunk(z, 4, dim=1)
        i, f = torch.sigmoid(i + peep[0] * c), torch.sigmoid(f + peep[1] * c)
        c = f * c + i * torch.tanh(g)
        o = torch.sigmoid(o + peep[2] * c)
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1), (h, c), [w, b] + peep


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("with_inputs", [True, False])
def test_shi_convlstm_sequence_gradients_match_autograd(precision, tol, with_inputs):
    """BPTT through three timesteps of the peephole ConvLSTM block (conv_lstm_hzzone.py:38-70): gradients of the inputs, the
    initial state, the conv weight / bias and the three peepholes; with inputs=None (the forecaster's zero input) too."""
    from vp_suite_b200 import model_blocks as MB
    dev = "cuda:0"
    cin, ch, hw, batch, T = 8, 16, (12, 10), 2, 3
    blk = MB.ConvLSTM(dev, in_channels=cin, enc_channels=ch, state_h=hw[0], state_w=hw[1], kernel_size=3)
    blk.precision = precision
    sd = synth_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=51)
    blk.load_state_dict(sd)
    g = torch.Generator().manual_seed(52)
    inputs = (torch.rand((batch, T, cin, *hw), generator=g) * 2 - 1) if with_inputs else None
    h0, c0 = (torch.rand((batch, ch, *hw), generator=g) * 2 - 1 for _ in range(2))
    r_out, r_h, r_c = torch.rand((batch, T, ch, *hw), generator=g), torch.rand((batch, ch, *hw), generator=g), \
        torch.rand((batch, ch, *hw), generator=g)
    # reference gradients (CPU autograd)
    xi = None if inputs is None else inputs.clone().requires_grad_(True)
    hr, cr = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    out, (hT, cT), leaves = _shi_reference(cin, ch, hw, sd, xi, hr, cr, T)
    ((out * r_out).sum() + (hT * r_h).sum() + (cT * r_c).sum()).backward()
    want = {"dh0": hr.grad, "dc0": cr.grad, "dW": leaves[0].grad, "db": leaves[1].grad, "dWci": leaves[2].grad,
            "dWcf": leaves[3].grad, "dWco": leaves[4].grad}
    if xi is not None:
        want["dx"] = xi.grad
    # ours
    xo = None if inputs is None else inputs.clone().to(dev).requires_grad_(True)
    ho, co = h0.clone().to(dev).requires_grad_(True), c0.clone().to(dev).requires_grad_(True)
    out_o, (hT_o, cT_o) = blk(xo, (ho, co), T)
    assert out_o.requires_grad
    ((out_o * r_out.to(dev)).sum() + (hT_o * r_h.to(dev)).sum() + (cT_o * r_c.to(dev)).sum()).backward()
    got = {"dh0": ho.grad, "dc0": co.grad, "dW": blk._conv.weight.grad, "db": blk._conv.bias.grad, "dWci": blk.Wci.grad,
           "dWcf": blk.Wcf.grad, "dWco": blk.Wco.grad}
    if xo is not None:
        got["dx"] = xo.grad
    errs = {k: _rel(got[k].cpu(), want[k]) for k in want}
    print(f"Shi ConvLSTM BPTT {precision} inputs={with_inputs}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) <= tol, errs
    with torch.no_grad():                                  # the inference path is unchanged and agrees with the forward above
        out_n, _ = blk(None if inputs is None else inputs.to(dev), (h0.to(dev), c0.to(dev)), T)
    assert not out_n.requires_grad and torch.equal(out_n, out_o.detach())


def test_ef_convlstm_trains_like_the_reference():
    """EF_ConvLSTM in train() mode (the differentiable forward: native ConvLSTM steps forward + backward, torch convs for the
    stages): the frames equal the native inference rollout, the parameter gradients of an MSE loss equal those of the
    reference model under CPU autograd (fp32 mode), and base_model.py:148-179's train_iter lowers the loss."""
    import vp_suite_b200 as V
    from oracle import models as OM
    from oracle.weights import synth_frames
    dev, img, b, ctx, pred = "cuda:0", (1, 32, 32), 2, 3, 2
    m = V.MODEL_CLASSES["convlstm-shi"](dev, img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0], precision="fp32")
    assert m.TRAINABLE
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=61, gain=2.5)
    m.load_state_dict(sd)
    x = synth_frames(b, ctx, *img, seed=62)
    target = synth_frames(b, pred, *img, seed=63)
    # reference gradients: the reference model when installed, else the oracle's functional forward over leaf tensors
    if HAVE_REF:
        ref = ref_shim.load_reference()["convlstm-shi"]("cpu", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0])
        ref.load_state_dict(sd)
        ref.train()
        out_r, _ = ref(x, pred_frames=pred)
        torch.nn.functional.mse_loss(out_r, target).backward()
        want = {k: p.grad for k, p in ref.named_parameters()}
    else:
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        out_r, _ = OM.ef_convlstm_forward(leaves, x, pred)
        torch.nn.functional.mse_loss(out_r, target).backward()
        want = {k: v.grad for k, v in leaves.items()}
    m.train()
    torch.backends.cudnn.allow_tf32 = False            # the stage convs of the training forward are torch's: plain fp32 here
    torch.backends.cuda.matmul.allow_tf32 = False
    out, aux = m(x.to(dev), pred_frames=pred)
    assert aux is None and out.requires_grad
    torch.nn.functional.mse_loss(out, target.to(dev)).backward()
    errs = {k: _rel(p.grad.cpu(), want[k]) for k, p in m.named_parameters() if want.get(k) is not None}
    worst = max(errs, key=errs.get)
    print("largest relative gradient errors:", sorted(((round(v, 6), k) for k, v in errs.items()), reverse=True)[:6])
    print(f"EF_ConvLSTM training gradients (fp32 mode): {len(errs)} tensors, worst {worst} {errs[worst]:.1e}; "
          f"forward err {float((out.detach().cpu() - out_r.detach()).abs().max()):.1e}")
    assert len(errs) == 44 and max(errs.values()) <= 1e-4, errs
    m.eval()
    with torch.no_grad():
        infer, _ = m(x.to(dev), pred_frames=pred)
    assert float((infer - out.detach()).abs().max()) <= 1e-5          # the training forward IS the rollout
    # train_iter (base_model.py:148-179) with a minimal loss provider: the loss goes down
    class _Loss:
        def get_losses(self, p, t):
            v = torch.nn.functional.mse_loss(p, t)
            return {"mse": v.detach()}, v
    m.train()
    m.zero_grad()
    loader = [{"frames": torch.cat([x, target], dim=1), "actions": torch.zeros(b, ctx + pred - 1, 0)}] * 6
    cfg = {"context_frames": ctx, "pred_frames": pred, "device": dev, "use_actions": False}
    cfg.update(m.config)
    opt = torch.optim.Adam(m.parameters(), lr=2e-3)
    with torch.no_grad():
        m.eval()
        before = float(torch.nn.functional.mse_loss(m(x.to(dev), pred_frames=pred)[0], target.to(dev)))
        m.train()
    m.train_iter(cfg, loader, opt, _Loss(), epoch=0)
    m.eval()
    with torch.no_grad():
        after = float(torch.nn.functional.mse_loss(m(x.to(dev), pred_frames=pred)[0], target.to(dev)))
    print(f"train_iter: mse {before:.5f} -> {after:.5f}")
    assert after < before


def test_ef_convlstm_train_iter_in_16_bit_mode():
    """The default (bf16-operand) mode trains too: finite gradients on every parameter, the loss falls over a few Adam steps,
    and eval() afterwards is the native rollout again (weights re-mirrored into the library after the optimizer steps)."""
    import vp_suite_b200 as V
    from oracle.weights import synth_frames
    dev, img, b, ctx, pred = "cuda:0", (3, 32, 32), 2, 2, 2
    m = V.MODEL_CLASSES["convlstm-shi"](dev, img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0])
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=71, gain=2.5))
    x, target = synth_frames(b, ctx, *img, seed=72), synth_frames(b, pred, *img, seed=73)

    class _Loss:
        def get_losses(self, p, t):
            v = torch.nn.functional.mse_loss(p, t)
            return {"mse": v.detach()}, v

    def mse():
        m.eval()
        with torch.no_grad():
            v = float(torch.nn.functional.mse_loss(m(x.to(dev), pred_frames=pred)[0], target.to(dev)))
        m.train()
        return v

    before = mse()
    out, _ = m(x.to(dev), pred_frames=pred)
    torch.nn.functional.mse_loss(out, target.to(dev)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    m.zero_grad()
    loader = [{"frames": torch.cat([x, target], dim=1), "actions": torch.zeros(b, ctx + pred - 1, 0)}] * 8
    cfg = {"context_frames": ctx, "pred_frames": pred, "device": dev, "use_actions": False, **m.config}
    m.train_iter(cfg, loader, torch.optim.Adam(m.parameters(), lr=2e-3), _Loss(), epoch=0)
    after = mse()
    print(f"bf16-mode train_iter: mse {before:.5f} -> {after:.5f}")
    assert after < before
