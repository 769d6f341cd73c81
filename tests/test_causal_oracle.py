"""CPU checks of the PredRNN++ restatement (oracle/causal.py; parity unpinned, see its header) and of the drop-in's host side.

With no reference module to compare with, the oracle is checked against an independent second statement of the paper's
equations written gate by gate on explicit channel slices (no torch.split / shared helper), on properties the equations
imply, and the drop-in's state_dict / native parameter layout against the oracle's."""
import pytest
import torch
import torch.nn.functional as F

from oracle import causal
from oracle.weights import synth_state_dict, synth_frames

KW = dict(action_size=0, tensor_value_range=[0.0, 1.0])


def _second_statement(x, h, c, m, w):
    """Eq. (1) of the PredRNN++ paper, one conv call per (gate, source) pair on sliced weights."""
    C = h.shape[1]
    pad = w["conv_x"].shape[-1] // 2

    def part(src, key, blk):
        return F.conv2d(src, w[key][blk * C:(blk + 1) * C], padding=pad)

    i = torch.sigmoid(part(x, "conv_x", 0) + part(h, "conv_h", 0) + part(c, "conv_c", 0))
    f = torch.sigmoid(part(x, "conv_x", 1) + part(h, "conv_h", 1) + part(c, "conv_c", 1) + 1.0)
    g = torch.tanh(part(x, "conv_x", 2) + part(h, "conv_h", 2) + part(c, "conv_c", 2))
    cn = f * c + i * g
    i2 = torch.sigmoid(part(x, "conv_x", 3) + part(m, "conv_m", 0) + part(cn, "conv_c2m", 0))
    f2 = torch.sigmoid(part(x, "conv_x", 4) + part(m, "conv_m", 1) + part(cn, "conv_c2m", 2) + 1.0)
    g2 = torch.tanh(part(x, "conv_x", 5) + part(cn, "conv_c2m", 1))
    mn = f2 * torch.tanh(part(m, "conv_m", 2)) + i2 * g2
    o = torch.tanh(part(x, "conv_x", 6) + part(h, "conv_h", 3) + part(cn, "conv_c2m", 3) + part(mn, "conv_om", 0))
    hn = o * torch.tanh(F.conv2d(cn, w["conv_last"][:, :C]) + F.conv2d(mn, w["conv_last"][:, C:]))
    return hn, cn, mn


def test_cell_restatements_agree():
    shapes = causal.state_dict_shapes(1, 2, 8, 2, 3)
    sd = synth_state_dict(shapes, seed=1, gain=2.5)
    w = causal.cell_weights(sd, "cell_list.1.")
    g = torch.Generator().manual_seed(0)
    x, h, c, m = (torch.randn(2, 8, 6, 5, generator=g) for _ in range(4))
    a = causal.causal_lstm_step(x, h, c, m, w)
    b = _second_statement(x, h, c, m, w)
    for u, v in zip(a, b):
        assert float((u - v).abs().max()) <= 2e-6


def test_ghu_limits():
    g = torch.Generator().manual_seed(3)
    x, z = torch.randn(1, 4, 5, 5, generator=g), torch.randn(1, 4, 5, 5, generator=g)
    w0 = torch.zeros(8, 4, 3, 3)
    # zero weights: switch gate = 1/2, p = 0  ->  z' = z / 2
    assert torch.allclose(causal.ghu_step(x, z, w0, w0), 0.5 * z)
    # a strongly positive switch gate keeps z, a strongly negative one replaces it with tanh(p)
    wu = torch.zeros(8, 4, 3, 3)
    wu[4:, :, 1, 1] = 50.0
    pos = torch.ones_like(x)
    assert torch.allclose(causal.ghu_step(pos, z, wu, w0), z, atol=1e-6)
    assert torch.allclose(causal.ghu_step(-pos, z, wu, w0), torch.zeros_like(z), atol=1e-6)


def test_unequal_widths_rollout_runs_and_uses_the_right_memory():
    shapes = causal.state_dict_shapes(1, 3, [8, 4, 6], 4, 3)
    sd = synth_state_dict(shapes, seed=5, gain=2.5)
    x = synth_frames(1, 5, 1, 16, 16, seed=2)
    p, _ = causal.predrnnpp_forward(sd, x, 2, {"num_layers": 3})
    assert p.shape == (1, 2, 1, 16, 16) and torch.isfinite(p).all() and float(p.abs().max()) > 0


def test_rollout_shape_and_feedback():
    shapes = causal.state_dict_shapes(1, 2, 8, 4, 3)
    sd = synth_state_dict(shapes, seed=4, gain=2.5)
    x = synth_frames(2, 6, 1, 16, 16, seed=1)
    p3, aux = causal.predrnnpp_forward(sd, x, 3, {"num_layers": 2})
    assert p3.shape == (2, 3, 1, 16, 16) and aux == {}
    # eval feeds x_gen back: the target frames behind the context never influence the prediction
    y = x.clone()
    y[:, 3:] = 0.0
    q3, _ = causal.predrnnpp_forward(sd, y, 3, {"num_layers": 2})
    assert torch.equal(p3, q3)
    # sequences are independent
    r3, _ = causal.predrnnpp_forward(sd, x[1:], 3, {"num_layers": 2})
    assert torch.allclose(r3, p3[1:], atol=1e-5)
    with pytest.raises(ValueError):
        causal.predrnnpp_forward(sd, x[:, :3], 3, {"num_layers": 2})


def test_dropin_layout_matches_oracle_and_native_library():
    import vp_suite_b200 as V
    cls = V.MODEL_CLASSES["predrnn-pp-causal"]
    m = cls("cpu", img_shape=(3, 32, 32), num_layers=3, num_hidden=[16, 16, 16], filter_size=3, **KW)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == causal.state_dict_shapes(3, 3, 16, 4, 3)
    assert got == {k: tuple(v) for k, v in m.native_param_layout().items()}
    assert m.NEEDS_COMPLETE_INPUT and not m.TRAINABLE and m.config["num_layers"] == 3
    with pytest.raises(AttributeError):
        cls("cpu", img_shape=(1, 32, 32), num_layers=1, num_hidden=[16], **KW)
    with pytest.raises(NotImplementedError):
        cls("cpu", img_shape=(1, 32, 32), layer_norm=True, **KW)
    # unequal widths (the paper's stacks are 128-64-64-64): conv_m reads the memory of the layer that wrote it
    u = cls("cpu", img_shape=(1, 32, 32), num_layers=3, num_hidden=[32, 16, 24], filter_size=3, **KW)
    got = {k: tuple(v.shape) for k, v in u.state_dict().items()}
    assert got == causal.state_dict_shapes(1, 3, [32, 16, 24], 4, 3)
    assert got == {k: tuple(v) for k, v in u.native_param_layout().items()}
    assert got["cell_list.0.conv_m.0.weight"] == (96, 24, 3, 3) and got["cell_list.1.conv_m.0.weight"] == (48, 32, 3, 3)
    assert got["cell_list.1.conv_x.0.weight"] == (7 * 16, 32, 3, 3) and got["conv_last.weight"] == (16, 24, 1, 1)
    with pytest.raises(AttributeError):
        cls("cpu", img_shape=(1, 32, 32), num_layers=3, num_hidden=[16, 32], **KW)


def test_block_dropins_layout_and_refusals():
    """CausalLSTMCell / GHU as VPModelBlock drop-ins: parameter layout of the paper's cell (conv_m reads a memory of
    num_hidden_in channels), constructor refusals, and no CPU path."""
    from vp_suite_b200.model_blocks import CausalLSTMCell, GHU
    from vp_suite_b200 import _native as N
    cell = CausalLSTMCell(16, 32, 8, 8, 5, 1, False, num_hidden_in=48)
    got = {k: tuple(v.shape) for k, v in cell.state_dict().items()}
    assert got == {"conv_x.0.weight": (224, 16, 5, 5), "conv_h.0.weight": (128, 32, 5, 5), "conv_c.0.weight": (96, 32, 5, 5),
                   "conv_m.0.weight": (96, 48, 5, 5), "conv_c2m.0.weight": (128, 32, 5, 5), "conv_om.0.weight": (32, 32, 5, 5),
                   "conv_last.weight": (32, 64, 1, 1)}
    ghu = GHU(32, 8, 8, 5)
    assert {k: tuple(v.shape) for k, v in ghu.state_dict().items()} == {"x_concat.0.weight": (64, 32, 5, 5),
                                                                          "z_concat.0.weight": (64, 32, 5, 5)}
    with pytest.raises(ValueError):
        CausalLSTMCell(16, 32, 8, 8, 4, 1, False)                      # even filter size
    with pytest.raises(NotImplementedError):
        GHU(32, 8, 8, 5, 1, True)                                       # layer_norm
    with pytest.raises(N.NativeError):                                  # CUDA tensors only: there is no CPU path
        cell(torch.zeros(1, 16, 8, 8), torch.zeros(1, 32, 8, 8), torch.zeros(1, 32, 8, 8), torch.zeros(1, 48, 8, 8))
    # pickling drops the native handle
    import pickle
    c2 = pickle.loads(pickle.dumps(cell))
    assert c2.num_hidden_in == 48 and c2._cell is None
