"""
Drop-in VPModel classes whose ``forward`` runs the whole rollout in libvpk (hand-written sm_100a kernels).

Each class keeps the reference's constructor signature, hyper-parameter attributes, ``forward(x, pred_frames)`` /
``pred_1`` contract and ``state_dict`` layout, so reference checkpoints and random-init weights
load unchanged:  ``ours.load_state_dict(reference_model.state_dict())``.

    EF_ConvLSTM      <- vp_suite/models/precipitation_nowcasting/ef_conv_lstm.py:7-108  (key "convlstm-shi")
    PredRNN_V2       <- vp_suite/models/predrnn_v2.py:11-230                           (key "predrnn-pp")
    PhyDNet          <- vp_suite/models/phydnet.py:12-137                              (key "phy")
    ConvLSTMBranch   <- BASELINE config 2: PhyDNet's residual branch alone (our composition of reference blocks)
    STPhy            <- vp_suite/models/st_phy.py:16-181                               (key "st-phy")
    EF_TrajGRU       <- vp_suite/models/precipitation_nowcasting/ef_traj_gru.py:8-119  (key "trajgru")
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _native as N
from .base import VPModel, NativeRollout


def _conv_out(hw, k, s, p):
    """vp_suite/utils/models.py:131-161."""
    return tuple((v + 2 * p - (k - 1) - 1) // s + 1 for v in hw)


def _convt_out(hw, k, s, p):
    """vp_suite/utils/models.py:164-193 (the reference's own formula)."""
    return tuple((v - 1) * s - 2 * p + (k - 1) + p for v in hw)


class _Params(nn.Module):
    """Parameter holder: gives sub-modules the attribute names the reference's state_dict keys are built from."""


def _stage(layers):
    return nn.Sequential(OrderedDict(layers))


def _act_code(name):
    """ef_blocks.py:32-46 tests 'relu' BEFORE 'leaky': a layer called '*leaky_relu*' is a plain ReLU there."""
    if "relu" in name:
        return 3
    if "leaky" in name:
        return 1
    return 0


class EF_ConvLSTM(NativeRollout, VPModel):
    """ef_conv_lstm.py:7-108 / ef_blocks.py:53-187.  Inference (``torch.no_grad`` / ``eval()``): the whole rollout is one
    native launch program.  Training (``train()`` with gradients enabled; ``train_iter`` of the base class, so
    ``VPSuite.train`` runs): a differentiable forward in the reference's own layer-major order whose six ConvLSTM layers are
    the drop-in ``ConvLSTM`` blocks -- native forward AND backward per timestep (``vpk_convlstm_cell_backward_peep``),
    BPTT composed by autograd -- and whose stage convs / deconvs / head are torch's own conv ops (cuDNN under autograd)."""
    NAME = "EF-ConvLSTM (Shi et al.)"
    TRAINABLE = True
    PAPER_REFERENCE = "https://arxiv.org/abs/1506.04214"
    CODE_REFERENCE = "https://github.com/Hzzone/Precipitation-Nowcasting"
    MATCHES_REFERENCE = "Yes"

    # hyper-parameters: ef_conv_lstm.py:31-65
    num_layers = 3
    enc_c = [16, 64, 64, 96, 96, 96]
    dec_c = [96, 96, 96, 96, 64, 16]
    enc_conv_names = ["conv1_leaky_1", "conv2_leaky_1", "conv3_leaky_1"]
    enc_conv_k = [3, 3, 3]
    enc_conv_s = [1, 2, 2]
    enc_conv_p = [1, 1, 1]
    dec_conv_names = ["deconv1_leaky_1", "deconv2_leaky_1", "deconv3_leaky_1"]
    dec_conv_k = [4, 4, 3]
    dec_conv_s = [2, 2, 1]
    dec_conv_p = [1, 1, 1]
    enc_rnn_k = [3, 3, 3]
    enc_rnn_s = [1, 1, 1]
    enc_rnn_p = [1, 1, 1]
    dec_rnn_k = [3, 3, 3]
    dec_rnn_s = [1, 1, 1]
    dec_rnn_p = [1, 1, 1]
    final_conv_1_name = "identity"
    final_conv_1_c = 16
    final_conv_1_k = 3
    final_conv_1_s = 1
    final_conv_1_p = 1
    final_conv_2_name = "conv3_3"
    final_conv_2_k = 1
    final_conv_2_s = 1
    final_conv_2_p = 0

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        L = self.num_layers
        if L != 3:                                     # Forecaster.forward hard-codes rnn3/stage3 (ef_blocks.py:109-110)
            raise AttributeError("the Encoder-Forecaster structure needs num_layers == 3")
        for name, val in [(k, v) for k, v in vars(self).items() if k.startswith(("enc_", "dec_"))]:
            want = 2 * L if name in ("enc_c", "dec_c") else L            # ef_blocks.py:134-143
            if isinstance(val, (list, tuple)) and len(val) != want:
                raise AttributeError(f"Speficied {L} layers, but len of attribute '{name}' doesn't match that ({val}).")
        if any(s != 1 for s in self.enc_rnn_s + self.dec_rnn_s) or \
                any(p != k // 2 for p, k in zip(self.enc_rnn_p + self.dec_rnn_p, self.enc_rnn_k + self.dec_rnn_k)):
            raise AttributeError("rnn stride must be 1 and padding k//2 (anything else changes the state size)")
        if self.final_conv_1_name != "identity" or self.final_conv_2_k != 1 or self.final_conv_2_s != 1 \
                or self.final_conv_2_p != 0:
            raise AttributeError("only the reference's final block (identity + 1x1 conv) is supported")
        acts = {_act_code(n) for n in self.enc_conv_names + self.dec_conv_names}
        if len(acts) != 1:
            raise AttributeError("all stage convs must share one activation")
        self._ef_act = acts.pop()

        # state sizes (ef_blocks.py:145-167)
        hw = (self.img_h, self.img_w)
        enc_hw = []
        for n in range(L):
            hw = _conv_out(hw, self.enc_conv_k[n], self.enc_conv_s[n], self.enc_conv_p[n])
            enc_hw.append(hw)
        dec_hw = [hw]
        for n in range(L - 1):
            hw = _convt_out(hw, self.dec_conv_k[n], self.dec_conv_s[n], self.dec_conv_p[n])
            dec_hw.append(hw)
        final = _convt_out(hw, self.dec_conv_k[-1], self.dec_conv_s[-1], self.dec_conv_p[-1])
        if final != (self.img_h, self.img_w):
            raise AttributeError(f"Model layer hyperparameters yield wrong output size: {final} "
                                 f"(expected: {(self.img_h, self.img_w)}). All hidden sizes: {enc_hw + dec_hw}")
        self.enc_rnn_state_h = [v[0] for v in enc_hw]
        self.enc_rnn_state_w = [v[1] for v in enc_hw]
        self.dec_rnn_state_h = [v[0] for v in dec_hw]
        self.dec_rnn_state_w = [v[1] for v in dec_hw]

        def rnn(in_c, c, hw_, k):
            m = _Params()
            m._conv = nn.Conv2d(in_c + c, 4 * c, k, 1, k // 2)
            for nm in ("Wci", "Wcf", "Wco"):                             # registered on every device (sec. 0.4)
                setattr(m, nm, nn.Parameter(torch.zeros(1, c, *hw_)))
            return m

        # Same module-construction order as the reference, so that a given torch seed yields the same init:
        # _build_encoder_decoder creates the six ConvLSTMs first (ef_conv_lstm.py:70-108); the stage convs are created
        # later by _make_layers inside Encoder.__init__ / Forecaster.__init__ (ef_blocks.py:63-65, 96-98).
        self.encoder = _Params()
        self.forecaster = _Params()
        in_c = self.img_c
        enc_rnn, enc_io = [], []
        for n in range(L):
            mid, out_c = self.enc_c[2 * n], self.enc_c[2 * n + 1]
            enc_rnn.append(rnn(mid, out_c, enc_hw[n], self.enc_rnn_k[n]))
            enc_io.append((in_c, mid))
            in_c = out_c
        dec_rnn, dec_io = [], []
        for n in range(L):
            mid, out_c = self.dec_c[2 * n], self.dec_c[2 * n + 1]
            dec_rnn.append(rnn(in_c, mid, dec_hw[n], self.dec_rnn_k[n]))
            dec_io.append((mid, out_c))
            in_c = out_c
        for n in range(L):                                               # Encoder.__init__   ef_blocks.py:63-65
            ci, co = enc_io[n]
            st = _stage([(self.enc_conv_names[n], nn.Conv2d(ci, co, self.enc_conv_k[n], self.enc_conv_s[n],
                                                            self.enc_conv_p[n]))])
            setattr(self.encoder, f"stage{n + 1}", st)
            setattr(self.encoder, f"rnn{n + 1}", enc_rnn[n])
        for n in range(L):                                               # Forecaster.__init__ ef_blocks.py:96-98
            ci, co = dec_io[n]
            layers = [(self.dec_conv_names[n], nn.ConvTranspose2d(ci, co, self.dec_conv_k[n], self.dec_conv_s[n],
                                                                  self.dec_conv_p[n]))]
            if n == L - 1:
                layers.append((self.final_conv_1_name, nn.Identity()))
                layers.append((self.final_conv_2_name, nn.Conv2d(self.final_conv_1_c, self.img_c, 1, 1, 0)))
            setattr(self.forecaster, f"rnn{L - n}", dec_rnn[n])
            setattr(self.forecaster, f"stage{L - n}", _stage(layers))
        self.NON_CONFIG_VARS = list(self.NON_CONFIG_VARS) + ["encoder", "forecaster"]
        self.to(device)

    # -- native glue ------------------------------------------------------------------------------------------------
    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = N.VPK_MODEL_CONVLSTM_SHI
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        for f in ("enc_c", "dec_c", "enc_conv_k", "enc_conv_s", "enc_conv_p", "dec_conv_k", "dec_conv_s",
                  "dec_conv_p", "enc_rnn_k", "dec_rnn_k"):
            arr = getattr(d, f)
            for i, v in enumerate(getattr(self, f)):
                arr[i] = int(v)
        d.final_conv_c = int(self.final_conv_1_c)
        d.ef_act = self._ef_act
        return d

    def _native_key(self, key):
        parts = key.split(".")
        if parts[1].startswith("stage"):
            if parts[2] == self.final_conv_2_name:
                parts[2] = "final"
            elif parts[0] == "encoder":
                parts[2] = "conv"
            else:
                parts[2] = "deconv"
        return ".".join(parts)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Checkpoints of a CUDA-constructed reference model lack Wci/Wcf/Wco (conv_lstm_hzzone.py:30-32 registers
        them only on CPU): missing peepholes mean zeros, as in the reference."""
        peep = [k for k in self.state_dict() if k.rsplit(".", 1)[-1] in ("Wci", "Wcf", "Wco")]
        if strict and not any(k in state_dict for k in peep):
            state_dict = dict(state_dict)
            own = self.state_dict()
            for k in peep:
                state_dict[k] = torch.zeros_like(own[k])
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # -- VPModel contract -------------------------------------------------------------------------------------------
    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)        # ef_blocks.py:181-182

    def forward(self, x, pred_frames: int = 1, **kwargs):
        b, t, c, h, w = x.shape
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_differentiable(x, int(pred_frames)), None
        pred, _ = self._native_forward(x, int(pred_frames), t)
        return pred, None                                                # ef_blocks.py:184-187

    # -- training path ----------------------------------------------------------------------------------------------
    def _train_block(self, holder, in_c, enc_c, hw, k):
        """A differentiable drop-in ConvLSTM block that SHARES ``holder``'s parameters (kept out of the module tree: the
        state_dict layout must stay the reference's)."""
        blocks = self.__dict__.setdefault("_train_blocks", {})
        blk = blocks.get(id(holder))
        if blk is None:
            from .model_blocks import ConvLSTM
            blk = ConvLSTM(self._native_device(), in_c, enc_c, hw[0], hw[1], k, 1, k // 2)
            blk._conv, blk.Wci, blk.Wcf, blk.Wco = holder._conv, holder.Wci, holder.Wcf, holder.Wco
            blocks[id(holder)] = blk
        blk.precision, blk.backend = self.precision, self.backend
        return blk

    def _forward_differentiable(self, x, pred_frames):
        import torch.nn.functional as F
        if not x.is_cuda:
            raise N.NativeError("vp_suite_b200 models run on CUDA tensors only (there is no CPU path)")
        act = {0: (lambda v: v), 1: (lambda v: F.leaky_relu(v, 0.2)), 3: F.relu}[self._ef_act]
        L = self.num_layers

        def stage(mods, v):                                              # ef_blocks.py:67-73 / 100-106: time-batched convs
            b_, t_ = v.shape[:2]
            v = v.reshape(-1, *v.shape[2:])
            for name, m in mods.named_children():
                if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                    v = m(v)
                    if _act_code(name) and name != self.final_conv_2_name:
                        v = act(v)
            return v.reshape(b_, t_, *v.shape[1:])

        inp, states = x.to(torch.float32), []
        for n in range(L):                                               # Encoder.forward (ef_blocks.py:76-82)
            inp = stage(getattr(self.encoder, f"stage{n + 1}"), inp)
            holder = getattr(self.encoder, f"rnn{n + 1}")
            hw = (self.enc_rnn_state_h[n], self.enc_rnn_state_w[n])
            blk = self._train_block(holder, self.enc_c[2 * n], self.enc_c[2 * n + 1], hw, self.enc_rnn_k[n])
            inp, st = blk(inp, None, inp.shape[1])
            states.append(st)
        out = None
        for n in range(L):                                               # Forecaster.forward (ef_blocks.py:108-114)
            idx = L - n                                                  # rnn3, rnn2, rnn1
            holder = getattr(self.forecaster, f"rnn{idx}")
            hw = (self.dec_rnn_state_h[n], self.dec_rnn_state_w[n])
            in_c = self.enc_c[-1] if n == 0 else self.dec_c[2 * n - 1]
            blk = self._train_block(holder, in_c, self.dec_c[2 * n], hw, self.dec_rnn_k[n])
            out, _ = blk(out, states[idx - 1], pred_frames)
            out = stage(getattr(self.forecaster, f"stage{idx}"), out)
        return out


class PredRNN_V2(NativeRollout, VPModel):
    NAME = "PredRNN++"
    PAPER_REFERENCE = "https://arxiv.org/abs/2103.09504"
    CODE_REFERENCE = "https://github.com/thuml/predrnn-pytorch"
    MATCHES_REFERENCE: str = "Yes"
    CAN_HANDLE_ACTIONS = False
    NEEDS_COMPLETE_INPUT = True

    # hyper-parameters: predrnn_v2.py:34-54 (training-only ones are kept so that configs round-trip)
    patch_size = 4
    num_layers = 3
    num_hidden = [128, 128, 128, 128]
    filter_size = 5
    stride = 1
    inflated_action_dim = 3
    layer_norm: bool = False
    conv_actions_on_input: bool = True
    residual_on_action_conv: bool = True
    reverse_input: bool = True
    decoupling_loss_scale = 100.0
    scheduled_sampling: bool = True
    sampling_stop_iter: int = 50000
    sampling_changing_rate = 2e-5
    reverse_scheduled_sampling: bool = False
    r_sampling_step_1: int = 25000
    r_sampling_step_2: int = 50000
    r_exp_alpha: int = 5000
    training_iteration: int = None
    sampling_eta: float = None

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        if self.stride != 1:
            raise AttributeError("ST-LSTM stride must be 1")
        self.patch_c = self.patch_size * self.patch_size * self.img_c          # predrnn_v2.py:59-62
        self.patch_a = self.action_size
        self.patch_h = self.rnn_h = self.img_h // self.patch_size
        self.patch_w = self.rnn_w = self.img_w // self.patch_size
        ac = bool(self.action_conditional)
        k, C0 = self.filter_size, self.num_hidden[0]
        if ac:                                                                 # predrnn_v2.py:65-67
            self.conv_actions_on_input = True
            self.reverse_scheduled_sampling = True
        else:                                                                  # predrnn_v2.py:68-70
            self.conv_actions_on_input = False
            self.residual_on_action_conv = False
        if self.conv_actions_on_input:                                         # predrnn_v2.py:73-90
            self.rnn_h //= 4
            self.rnn_w //= 4
            CL = self.num_hidden[self.num_layers - 1]
            self.conv_input1 = nn.Conv2d(self.patch_c, C0 // 2, k, stride=2, padding=k // 2, bias=False)
            self.conv_input2 = nn.Conv2d(C0 // 2, C0, k, stride=2, padding=k // 2, bias=False)
            self.action_conv_input1 = nn.Conv2d(self.patch_a, C0 // 2, k, stride=2, padding=k // 2, bias=False)
            self.action_conv_input2 = nn.Conv2d(C0 // 2, C0, k, stride=2, padding=k // 2, bias=False)
            self.deconv_output1 = nn.ConvTranspose2d(CL, CL // 2, k, stride=2, padding=k // 2, bias=False)
            self.deconv_output2 = nn.ConvTranspose2d(CL // 2, self.patch_c, k, stride=2, padding=k // 2, bias=False)
        cells = []
        for i in range(self.num_layers):                                       # predrnn_v2.py:92-108
            cin = (C0 if ac else self.patch_c) if i == 0 else self.num_hidden[i - 1]
            C = self.num_hidden[i]
            cell = _Params()
            # ActionConditionalSpatioTemporalLSTMCell: convs with bias, a fifth conv_a (model_blocks/predrnn.py:97-139)
            names = (("conv_x", (7 * C, cin)), ("conv_h", (4 * C, C))) + ((("conv_a", (4 * C, C)),) if ac else ()) + \
                    (("conv_m", (3 * C, C)), ("conv_o", (C, 2 * C)))
            for name, (o, ci) in names:
                layers = [nn.Conv2d(ci, o, k, 1, k // 2, bias=ac)]
                if self.layer_norm:                                            # model_blocks/predrnn.py:24-40
                    layers.append(nn.LayerNorm([o, self.rnn_h, self.rnn_w]))
                setattr(cell, name, nn.Sequential(*layers))
            cell.conv_last = nn.Conv2d(2 * C, C, 1, 1, 0, bias=ac)
            cells.append(cell)
        self.cell_list = nn.ModuleList(cells)
        if not ac:                                 # (non-existent when conv_actions_on_input is True, predrnn_v2.py:110-116)
            self.conv_last = nn.Conv2d(self.num_hidden[self.num_layers - 1], self.patch_c, 1, 1, 0, bias=False)
        CA = self.num_hidden[self.num_layers - 1] if ac else C0               # predrnn_v2.py:119-120
        self.adapter = nn.Conv2d(CA, CA, 1, 1, 0, bias=False)
        self.training_iteration = 1                                            # predrnn_v2.py:124-126
        self.sampling_eta = 1.0
        self.to(device)

    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = N.VPK_MODEL_PREDRNN_PP
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        d.patch_size, d.num_layers, d.filter_size = self.patch_size, self.num_layers, self.filter_size
        for i, v in enumerate(self.num_hidden[:8]):
            d.num_hidden[i] = int(v)
        d.decoupling_loss_scale = float(self.decoupling_loss_scale)
        d.layer_norm = int(bool(self.layer_norm))
        d.action_conditional = int(bool(self.action_conditional))
        d.action_size = int(self.action_size)
        d.residual_on_action_conv = int(bool(self.residual_on_action_conv))
        return d

    def _native_key(self, key):
        return key

    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)              # predrnn_v2.py:128-129

    def forward(self, x, pred_frames: int = 1, **kwargs):
        b, total, c, h, w = x.shape
        if total - pred_frames < 1:                                            # predrnn_v2.py:134-137
            raise ValueError(f"Model {self.NAME} needs input sequences that also include the target frames!")
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):                  # predrnn_v2.py:234-236
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        if kwargs.get("train", False):
            raise NotImplementedError("the native rollout is inference-only")
        actions = self._native_actions(kwargs, b, x.device)                    # predrnn_v2.py:147-152
        pred, aux = self._native_forward(x, int(pred_frames), total, want_aux=True, actions=actions)
        return pred, {"ST-LSTM decouple loss": aux[0]}                         # predrnn_v2.py:229-230


class PredRNNpp(NativeRollout, VPModel):
    """PredRNN++ as published (Wang et al., ICML 2018): Causal LSTM stack with a gradient highway unit between the first and
    the second layer -- the cells BASELINE.json's north star names.  The reference checkout has no such model (its
    ``predrnn-pp`` key is PredRNN_V2 above), so this class has no reference twin: PARITY UNPINNED, checked against
    oracle/causal.py only.  The VPModel contract (patches, complete input, eval rollout, ``pred_1``) is PredRNN_V2's
    (models/predrnn_v2.py:128-137, 223-228); parameter names follow its Sequential-per-conv layout."""
    NAME = "PredRNN++ (Causal LSTM + GHU)"
    PAPER_REFERENCE = "https://arxiv.org/abs/1804.06300"
    CODE_REFERENCE = "https://github.com/Yunbo426/predrnn-pp"
    MATCHES_REFERENCE: str = "Not Yet"
    CAN_HANDLE_ACTIONS = False
    NEEDS_COMPLETE_INPUT = True
    TRAINABLE = False

    patch_size = 4
    num_layers = 4
    num_hidden = [128, 128, 128, 128]
    filter_size = 5
    stride = 1
    layer_norm: bool = False

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        if self.stride != 1:
            raise AttributeError("Causal LSTM stride must be 1")
        if self.layer_norm or self.action_conditional:
            raise NotImplementedError("PredRNN++ drop-in: layer_norm / action_conditional are not built")
        if self.num_layers < 2:
            raise AttributeError("PredRNN++ needs at least two layers (the GHU sits between the first two)")
        self.patch_c = self.patch_size * self.patch_size * self.img_c
        self.patch_h = self.rnn_h = self.img_h // self.patch_size
        self.patch_w = self.rnn_w = self.img_w // self.patch_size
        if len(self.num_hidden) < self.num_layers:
            raise AttributeError("num_hidden needs one entry per layer")
        k, L, hid = self.filter_size, self.num_layers, self.num_hidden

        def conv(ci, co):
            return nn.Sequential(nn.Conv2d(ci, co, k, 1, k // 2, bias=False))

        cells = []
        for i in range(L):
            # widths may differ per layer (the paper: 128-64-64-64); the spatial memory a cell reads has the width of the
            # cell that wrote it: the previous layer, or the top layer for layer 0
            C, cin, cm = hid[i], (self.patch_c if i == 0 else hid[i - 1]), hid[(i - 1) % L]
            cell = _Params()
            cell.conv_x = conv(cin, 7 * C)
            cell.conv_h = conv(C, 4 * C)
            cell.conv_c = conv(C, 3 * C)
            cell.conv_m = conv(cm, 3 * C)
            cell.conv_c2m = conv(C, 4 * C)
            cell.conv_om = conv(C, C)
            cell.conv_last = nn.Conv2d(2 * C, C, 1, 1, 0, bias=False)
            cells.append(cell)
        self.cell_list = nn.ModuleList(cells)
        self.gradient_highway = _Params()
        self.gradient_highway.x_concat = conv(hid[0], 2 * hid[0])
        self.gradient_highway.z_concat = conv(hid[0], 2 * hid[0])
        self.conv_last = nn.Conv2d(hid[L - 1], self.patch_c, 1, 1, 0, bias=False)
        self.to(device)

    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = N.VPK_MODEL_PREDRNN_PP_CAUSAL
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        d.patch_size, d.num_layers, d.filter_size = self.patch_size, self.num_layers, self.filter_size
        for i, v in enumerate(self.num_hidden[:8]):
            d.num_hidden[i] = int(v)
        return d

    def _native_key(self, key):
        return key

    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)

    def forward(self, x, pred_frames: int = 1, **kwargs):
        b, total, c, h, w = x.shape
        if total - pred_frames < 1:
            raise ValueError(f"Model {self.NAME} needs input sequences that also include the target frames!")
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        if kwargs.get("train", False):
            raise NotImplementedError("the native rollout is inference-only")
        pred, _ = self._native_forward(x, int(pred_frames), total)
        return pred, {}


def _dcgan(cin, cout, stride, transpose):
    """DCGANConv / DCGANConvTranspose parameter layout (model_blocks/conv.py:58-95): main.0 conv, main.1 GroupNorm."""
    m = _Params()
    if transpose:
        conv = nn.ConvTranspose2d(cin, cout, (3, 3), stride, 1, output_padding=int(stride == 2))
    else:
        conv = nn.Conv2d(cin, cout, (3, 3), stride, 1)
    m.main = nn.Sequential(conv, nn.GroupNorm(16, cout), nn.LeakyReLU(0.2, inplace=True))
    return m


def _gn_divisor(x):
    """model_blocks/phydnet.py:348-362."""
    sq = int(x ** 0.5)
    while (sq + 1) * (sq + 1) <= x:
        sq += 1
    while x % sq != 0:
        sq -= 1
    return x // sq


class PhyDNet(NativeRollout, VPModel):
    NAME = "PhyDNet"
    PAPER_REFERENCE = "https://arxiv.org/abs/2003.01460"
    CODE_REFERENCE = "https://github.com/vincent-leguen/PhyDNet"
    MATCHES_REFERENCE: str = "Not Yet"
    CAN_HANDLE_ACTIONS = True

    # hyper-parameters: models/phydnet.py:28-36
    phycell_n_layers = 1
    phycell_channels = 49
    phycell_kernel_size = (7, 7)
    convlstm_n_layers = 3
    convlstm_hidden_dims = [128, 128, 64]
    convlstm_kernel_size = (3, 3)
    moment_loss_scale = 1.0
    teacher_forcing_decay = 0.003

    _KIND = N.VPK_MODEL_PHY

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        ac = bool(self.action_conditional) and self._KIND == N.VPK_MODEL_PHY
        if bool(self.action_conditional) and not ac:
            raise AttributeError("the ConvLSTM-branch composition is not action-conditional")
        a = int(self.action_size) if ac else 0
        if self.img_h % 4 or self.img_w % 4:
            raise AttributeError("image size must be a multiple of 4 (other sizes need the reference's Resize)")
        c = self.img_c
        # construction order of the reference (models/phydnet.py:41-63) so that a given seed yields the same init
        self.encoder_E = _Params()
        self.encoder_E.c1 = _dcgan(c, 32, 2, False)
        self.encoder_E.c2 = _dcgan(32, 32, 1, False)
        self.encoder_E.c3 = _dcgan(32, 64, 2, False)
        for name in ("encoder_Ep", "encoder_Er"):
            e = _Params()
            e.c1 = _dcgan(64, 64, 1, False)
            e.c2 = _dcgan(64, 64, 1, False)
            setattr(self, name, e)
        self.shape_Ep = self.shape_Er = torch.Size((64, self.img_h // 4, self.img_w // 4))
        for name in ("decoder_Dp", "decoder_Dr"):
            dd = _Params()
            dd.upc1 = _dcgan(64, 64, 1, True)
            dd.upc2 = _dcgan(64, 64, 1, True)
            setattr(self, name, dd)
        self.decoder_D = _Params()
        self.decoder_D.upc1 = _dcgan(64, 32, 2, True)
        self.decoder_D.upc2 = _dcgan(32, 32, 1, True)
        self.decoder_D.upc3 = nn.ConvTranspose2d(32, c, (3, 3), 2, 1, output_padding=1)
        hid, kp = self.phycell_channels, self.phycell_kernel_size
        self.phycell = _Params()
        cells = []
        for _ in range(self.phycell_n_layers):
            cell = _Params()
            cell.F = nn.Sequential()
            cell.F.add_module("conv1", nn.Conv2d(64, hid, kp, (1, 1), (kp[0] // 2, kp[1] // 2)))
            cell.F.add_module("bn1", nn.GroupNorm(_gn_divisor(hid), hid))
            cell.F.add_module("conv2", nn.Conv2d(hid, 64, (1, 1)))
            cell.convgate = nn.Conv2d(128, 64, (3, 3), padding=(1, 1))
            if ac:                                                             # model_blocks/phydnet.py:44-48
                cell.frame_action_conv = nn.Conv2d(64 + a, 64, (1, 1))
                cell.hidden_action_conv = nn.Conv2d(64 + a, 64, (1, 1))
            cells.append(cell)
        self.phycell.cell_list = nn.ModuleList(cells)
        self.convcell = _Params()
        cells, cin = [], 64 + a                                                # model_blocks/phydnet.py:137
        kc = self.convlstm_kernel_size
        for hd in self.convlstm_hidden_dims[:self.convlstm_n_layers]:
            cell = _Params()
            cell.conv = nn.Conv2d(cin + hd, 4 * hd, kc, padding=(kc[0] // 2, kc[1] // 2))
            cells.append(cell)
            cin = hd
        self.convcell.cell_list = nn.ModuleList(cells)
        self.to(device)

    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = self._KIND
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        d.phycell_n_layers = self.phycell_n_layers
        d.phycell_channels = self.phycell_channels
        if self.phycell_kernel_size[0] != self.phycell_kernel_size[1] or \
                self.convlstm_kernel_size[0] != self.convlstm_kernel_size[1]:
            raise AttributeError("square kernels only")
        d.phycell_kernel_size = self.phycell_kernel_size[0]
        d.convlstm_n_layers = self.convlstm_n_layers
        for i, v in enumerate(self.convlstm_hidden_dims[:8]):
            d.convlstm_hidden_dims[i] = int(v)
        d.convlstm_kernel_size = self.convlstm_kernel_size[0]
        d.action_conditional = int(bool(self.action_conditional))
        d.action_size = int(self.action_size)
        return d

    def _native_key(self, key):
        return key

    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)              # models/phydnet.py:91-92

    def forward(self, x, pred_frames=1, **kwargs):
        if kwargs.get("train", False):
            raise NotImplementedError("the native rollout is inference-only")
        b, t, c, h, w = x.shape
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        actions = self._native_actions(kwargs, b, x.device)                    # models/phydnet.py:100-105
        pred, _ = self._native_forward(x, int(pred_frames), t, actions=actions)
        return pred, None                                                      # models/phydnet.py:134-137 (eval)


class ConvLSTMBranch(PhyDNet):
    """BASELINE config 2 ("custom ConvLSTM: encoder + stacked ConvLSTM cells"): PhyDNet's residual branch alone --
    DCGANEncoder -> EncoderSplit -> SingleStepConvLSTM[128,128,64] -> DecoderSplit -> DCGANDecoder -> sigmoid.  It is
    OUR composition of reference blocks (the reference registers no such model); it keeps PhyDNet's state_dict layout
    (the PhyCell / Ep / Dp entries are simply unused) so that a PhyDNet checkpoint loads unchanged."""
    NAME = "ConvLSTM branch of PhyDNet"
    CAN_HANDLE_ACTIONS = False
    _KIND = N.VPK_MODEL_CONVLSTM_BRANCH


class Activation:
    """model_blocks/traj_gru.py:8-27: the configurable activation object EF_TrajGRU keeps as its ``activation`` attribute."""

    def __init__(self, act_type, negative_slope=0.2, inplace=True):
        self._act_type, self.negative_slope, self.inplace = act_type, negative_slope, inplace

    def __call__(self, input):
        if self._act_type == "leaky":
            return torch.nn.functional.leaky_relu(input, negative_slope=self.negative_slope)
        if self._act_type == "relu":
            return torch.relu(input)
        if self._act_type == "sigmoid":
            return torch.sigmoid(input)
        raise NotImplementedError


class EF_TrajGRU(NativeRollout, VPModel):
    """models/precipitation_nowcasting/ef_traj_gru.py:8-119: the Encoder-Forecaster skeleton with TrajGRU blocks
    (model_blocks/traj_gru.py:70-214: learned flows, bilinear warps of the hidden state, 1x1 conv, GRU gates)."""
    NAME = "EF-TrajGRU (Shi et al.)"
    PAPER_REFERENCE = "https://arxiv.org/abs/1706.03458"
    CODE_REFERENCE = "https://github.com/Hzzone/Precipitation-Nowcasting"
    MATCHES_REFERENCE: str = "Yes"

    # hyper-parameters: ef_traj_gru.py:30-74
    activation = Activation("leaky", negative_slope=0.2, inplace=True)
    num_layers = 3
    enc_c = [16, 64, 64, 96, 96, 96]
    dec_c = [96, 96, 96, 96, 64, 16]
    enc_conv_names = ["conv1_leaky_1", "conv2_leaky_1", "conv3_leaky_1"]
    enc_conv_k = [3, 3, 3]
    enc_conv_s = [1, 2, 2]
    enc_conv_p = [1, 1, 1]
    dec_conv_names = ["deconv1_leaky_1", "deconv2_leaky_1", "deconv3_leaky_1"]
    dec_conv_k = [4, 4, 3]
    dec_conv_s = [2, 2, 1]
    dec_conv_p = [1, 1, 1]
    enc_rnn_z = [0.0, 0.0, 0.0]
    enc_rnn_L = [13, 13, 13]
    enc_rnn_i2h_k = [(3, 3), (3, 3), (3, 3)]
    enc_rnn_i2h_s = [(1, 1), (1, 1), (1, 1)]
    enc_rnn_i2h_p = [(1, 1), (1, 1), (1, 1)]
    enc_rnn_h2h_k = [(5, 5), (5, 5), (3, 3)]
    enc_rnn_h2h_d = [(1, 1), (1, 1), (1, 1)]
    dec_rnn_z = [0.0, 0.0, 0.0]
    dec_rnn_L = [13, 13, 13]
    dec_rnn_i2h_k = [(3, 3), (3, 3), (3, 3)]
    dec_rnn_i2h_s = [(1, 1), (1, 1), (1, 1)]
    dec_rnn_i2h_p = [(1, 1), (1, 1), (1, 1)]
    dec_rnn_h2h_k = [(3, 3), (5, 5), (5, 5)]
    dec_rnn_h2h_d = [(1, 1), (1, 1), (1, 1)]
    final_conv_1_name = "identity"
    final_conv_1_c = 16
    final_conv_1_k = 3
    final_conv_1_s = 1
    final_conv_1_p = 1
    final_conv_2_name = "conv3_3"
    final_conv_2_k = 1
    final_conv_2_s = 1
    final_conv_2_p = 0

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        L = self.num_layers
        if L != 3:
            raise AttributeError("the Encoder-Forecaster structure needs num_layers == 3")
        for name, val in [(k, v) for k, v in vars(self).items() if k.startswith(("enc_", "dec_"))]:
            want = 2 * L if name in ("enc_c", "dec_c") else L            # ef_blocks.py:134-143
            if isinstance(val, (list, tuple)) and len(val) != want:
                raise AttributeError(f"Speficied {L} layers, but len of attribute '{name}' doesn't match that ({val}).")
        if any(z != 0.0 for z in self.enc_rnn_z + self.dec_rnn_z):
            raise AttributeError("zoneout is a training-time regulariser: the native rollout takes zoneout 0 only")
        for ks, ss, ps in ((self.enc_rnn_i2h_k, self.enc_rnn_i2h_s, self.enc_rnn_i2h_p), (self.dec_rnn_i2h_k, self.dec_rnn_i2h_s, self.dec_rnn_i2h_p)):
            for k, st, pd in zip(ks, ss, ps):
                if k[0] != k[1] or k[0] % 2 == 0 or tuple(st) != (1, 1) or tuple(pd) != (k[0] // 2, k[0] // 2):
                    raise AttributeError("i2h convs must be square, odd, stride 1, 'same' padding (anything else changes the state size)")
        if self.final_conv_1_name != "identity" or self.final_conv_2_k != 1 or self.final_conv_2_s != 1 or self.final_conv_2_p != 0:
            raise AttributeError("only the reference's final block (identity + 1x1 conv) is supported")
        acts = {_act_code(n) for n in self.enc_conv_names + self.dec_conv_names}
        rnn_act = {"leaky": 1, "relu": 3, "sigmoid": 2}.get(getattr(self.activation, "_act_type", None))
        if len(acts) != 1 or rnn_act is None or acts != {rnn_act}:
            raise AttributeError("stage convs and the TrajGRU activation must share one activation")
        self._ef_act = rnn_act
        hw = (self.img_h, self.img_w)
        enc_hw = []
        for n in range(L):
            hw = _conv_out(hw, self.enc_conv_k[n], self.enc_conv_s[n], self.enc_conv_p[n])
            enc_hw.append(hw)
        dec_hw = [hw]
        for n in range(L - 1):
            hw = _convt_out(hw, self.dec_conv_k[n], self.dec_conv_s[n], self.dec_conv_p[n])
            dec_hw.append(hw)
        final = _convt_out(hw, self.dec_conv_k[-1], self.dec_conv_s[-1], self.dec_conv_p[-1])
        if final != (self.img_h, self.img_w):
            raise AttributeError(f"Model layer hyperparameters yield wrong output size: {final} "
                                 f"(expected: {(self.img_h, self.img_w)}). All hidden sizes: {enc_hw + dec_hw}")
        self.enc_rnn_state_h = [v[0] for v in enc_hw]
        self.enc_rnn_state_w = [v[1] for v in enc_hw]
        self.dec_rnn_state_h = [v[0] for v in dec_hw]
        self.dec_rnn_state_w = [v[1] for v in dec_hw]

        def rnn(in_c, c, flows, k):                                      # traj_gru.py:93-131, same creation order
            m = _Params()
            m.i2h = nn.Conv2d(in_c, 3 * c, k, 1, (k[0] // 2, k[1] // 2))
            m.i2f_conv1 = nn.Conv2d(in_c, 32, (5, 5), 1, (2, 2))
            m.h2f_conv1 = nn.Conv2d(c, 32, (5, 5), 1, (2, 2))
            m.flows_conv = nn.Conv2d(32, flows * 2, (5, 5), 1, (2, 2))
            m.ret = nn.Conv2d(c * flows, 3 * c, (1, 1), 1)
            return m

        self.encoder, self.forecaster = _Params(), _Params()
        in_c = self.img_c
        enc_rnn, enc_io, dec_rnn, dec_io = [], [], [], []
        for n in range(L):                                               # ef_traj_gru.py:80-97
            mid, out_c = self.enc_c[2 * n], self.enc_c[2 * n + 1]
            enc_io.append((in_c, mid))
            enc_rnn.append(rnn(mid, out_c, self.enc_rnn_L[n], self.enc_rnn_i2h_k[n]))
            in_c = out_c
        for n in range(L):                                               # ef_traj_gru.py:99-119
            mid, out_c = self.dec_c[2 * n], self.dec_c[2 * n + 1]
            dec_rnn.append(rnn(in_c, mid, self.dec_rnn_L[n], self.dec_rnn_i2h_k[n]))
            dec_io.append((mid, out_c))
            in_c = out_c
        for n in range(L):                                               # Encoder.__init__   ef_blocks.py:63-65
            ci, co = enc_io[n]
            setattr(self.encoder, f"stage{n + 1}", _stage([(self.enc_conv_names[n], nn.Conv2d(ci, co, self.enc_conv_k[n],
                                                                                                self.enc_conv_s[n], self.enc_conv_p[n]))]))
            setattr(self.encoder, f"rnn{n + 1}", enc_rnn[n])
        for n in range(L):                                               # Forecaster.__init__ ef_blocks.py:96-98
            ci, co = dec_io[n]
            layers = [(self.dec_conv_names[n], nn.ConvTranspose2d(ci, co, self.dec_conv_k[n], self.dec_conv_s[n], self.dec_conv_p[n]))]
            if n == L - 1:
                layers.append((self.final_conv_1_name, nn.Identity()))
                layers.append((self.final_conv_2_name, nn.Conv2d(self.final_conv_1_c, self.img_c, 1, 1, 0)))
            setattr(self.forecaster, f"rnn{L - n}", dec_rnn[n])
            setattr(self.forecaster, f"stage{L - n}", _stage(layers))
        self.NON_CONFIG_VARS = list(self.NON_CONFIG_VARS) + ["encoder", "forecaster"]
        self.to(device)

    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = N.VPK_MODEL_TRAJGRU
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        for f in ("enc_c", "dec_c", "enc_conv_k", "enc_conv_s", "enc_conv_p", "dec_conv_k", "dec_conv_s", "dec_conv_p",
                  "enc_rnn_L", "dec_rnn_L"):
            arr = getattr(d, f)
            for i, v in enumerate(getattr(self, f)):
                arr[i] = int(v)
        for i in range(3):
            d.enc_rnn_k[i] = int(self.enc_rnn_i2h_k[i][0])
            d.dec_rnn_k[i] = int(self.dec_rnn_i2h_k[i][0])
        d.final_conv_c = int(self.final_conv_1_c)
        d.ef_act = self._ef_act
        return d

    _native_key = EF_ConvLSTM._native_key

    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)        # ef_blocks.py:181-182

    def forward(self, x, pred_frames: int = 1, **kwargs):
        b, t, c, h, w = x.shape
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        pred, _ = self._native_forward(x, int(pred_frames), t)
        return pred, None                                                # ef_blocks.py:184-187


class STPhy(NativeRollout, VPModel):
    """models/st_phy.py:16-181: Autoencoder + per layer one PhyCell_Cell and one LayerNorm ST-LSTM cell merged by a 1x1 conv.
    With action_conditional=True (st_phy.py:48-56, 142-150) the action vector is inflated by a Linear + a (5,1) and a (1,5)
    conv to the action tensor of the ActionConditionalSpatioTemporalLSTMCells, and the PhyCells take the raw action.
    The native rollout covers eval mode (losses are training-only there: ``forward`` returns ``(frames, None)``)."""
    NAME = "ST-Phy"
    CAN_HANDLE_ACTIONS = True

    # hyper-parameters: st_phy.py:28-36
    num_layers = 3
    phycell_channels = 49
    phycell_kernel_size = (7, 7)
    st_cell_channels = 64
    inflated_action_dim = 3
    decoupling_loss_scale = 100.0
    moment_loss_scale = 1.0
    teacher_forcing_decay = 0.003

    def __init__(self, device, **model_kwargs):
        super().__init__(device, **model_kwargs)
        self._native_init()
        C, c = self.st_cell_channels, self.img_c
        ac, a = bool(self.action_conditional), int(self.action_size)
        self.dim_st_hidden = [C] * self.num_layers                              # st_phy.py:41-42
        self.dim_phy_hidden = [self.phycell_channels] * self.num_layers
        from .model_blocks import SpatioTemporalLSTMCell, ActionConditionalSpatioTemporalLSTMCell   # (model_blocks imports this module)
        self.recurrent_cell = ActionConditionalSpatioTemporalLSTMCell if ac else SpatioTemporalLSTMCell   # st_phy.py:46-49 (in `config`)
        # Autoencoder (model_blocks/enc.py:14-98): same construction order as the reference for same-seed init
        self.autoencoder = _Params()
        enc = _Params()
        enc.conv1 = nn.Conv2d(c, 32, 5, 2)
        enc.conv2 = nn.Conv2d(32, 64, 3, 2)
        enc.mean_layer = nn.Conv2d(64, C, 3, 1)
        dec = _Params()
        dec.fc1 = nn.Conv2d(C, C, 1, 1)
        dec.conv1 = nn.ConvTranspose2d(C, 64, 6, 2, 0)
        dec.conv2 = nn.ConvTranspose2d(64, 32, 6, 2, 0)
        dec.conv3 = nn.ConvTranspose2d(32, c, 5, 1, 0)
        self.autoencoder.encoder, self.autoencoder.decoder = enc, dec
        h1, w1 = (self.img_h - 5) // 2 + 1, (self.img_w - 5) // 2 + 1
        self.enc_h, self.enc_w = (h1 - 3) // 2 + 1 - 2, (w1 - 3) // 2 + 1 - 2   # autoencoder.encoded_shape (st_phy.py:45)
        if ((self.enc_h - 1) * 2 + 6 - 1) * 2 + 6 + 4 != self.img_h or ((self.enc_w - 1) * 2 + 6 - 1) * 2 + 6 + 4 != self.img_w:
            raise AttributeError("image sizes whose decoder output needs the reference's Resize are not supported")
        if ac:                                                                  # st_phy.py:48-56
            self.action_inflate = nn.Linear(a, self.inflated_action_dim * self.enc_h * self.enc_w, bias=False)
            self.action_conv_h = nn.Conv2d(self.inflated_action_dim, C, (5, 1), padding=(2, 0), bias=False)
            self.action_conv_w = nn.Conv2d(self.inflated_action_dim, C, (1, 5), padding=(0, 2), bias=False)
        st_cells, phycells, hidden_convs = [], [], []
        kp = self.phycell_kernel_size
        for i in range(self.num_layers):                                        # st_phy.py:58-70
            cell = _Params()
            # model_blocks/predrnn.py:24-40 / :97-139 (action-conditional: conv biases, conv_a after conv_h)
            convs = (("conv_x", (7 * C, C)), ("conv_h", (4 * C, C)), ("conv_a", (4 * C, C)), ("conv_m", (3 * C, C)),
                     ("conv_o", (C, 2 * C))) if ac else \
                    (("conv_x", (7 * C, C)), ("conv_h", (4 * C, C)), ("conv_m", (3 * C, C)), ("conv_o", (C, 2 * C)))
            for name, (o, ci) in convs:
                setattr(cell, name, nn.Sequential(nn.Conv2d(ci, o, 5, 1, 2, bias=ac), nn.LayerNorm([o, self.enc_h, self.enc_w])))
            cell.conv_last = nn.Conv2d(2 * C, C, 1, 1, 0, bias=ac)
            st_cells.append(cell)
            pc = _Params()
            pc.F = nn.Sequential()
            pc.F.add_module("conv1", nn.Conv2d(C, self.phycell_channels, kp, (1, 1), (kp[0] // 2, kp[1] // 2)))
            pc.F.add_module("bn1", nn.GroupNorm(_gn_divisor(self.phycell_channels), self.phycell_channels))
            pc.F.add_module("conv2", nn.Conv2d(self.phycell_channels, C, (1, 1)))
            pc.convgate = nn.Conv2d(2 * C, C, (3, 3), padding=(1, 1))
            if ac:                                                              # model_blocks/phydnet.py:44-48
                pc.frame_action_conv = nn.Conv2d(C + a, C, (1, 1))
                pc.hidden_action_conv = nn.Conv2d(C + a, C, (1, 1))
            phycells.append(pc)
            hidden_convs.append(nn.Conv2d(2 * C, C, (1, 1), bias=(i < self.num_layers - 1)))
        self.st_cell_list = nn.ModuleList(st_cells)
        self.phycell_list = nn.ModuleList(phycells)
        self.hidden_conv_list = nn.ModuleList(hidden_convs)
        self.adapter = nn.Conv2d(C, C, 1, stride=1, padding=0, bias=False)
        self.to(device)

    def _native_desc(self):
        d = N.ModelDesc()
        d.kind = N.VPK_MODEL_ST_PHY
        d.img_c, d.img_h, d.img_w = self.img_c, self.img_h, self.img_w
        d.num_layers = self.num_layers
        d.num_hidden[0] = int(self.st_cell_channels)
        d.phycell_channels = self.phycell_channels
        if self.phycell_kernel_size[0] != self.phycell_kernel_size[1]:
            raise AttributeError("square kernels only")
        d.phycell_kernel_size = self.phycell_kernel_size[0]
        d.action_conditional = int(bool(self.action_conditional))
        d.action_size = int(self.action_size) if self.action_conditional else 0
        d.inflated_action_dim = int(self.inflated_action_dim)
        return d

    def _native_key(self, key):
        return key

    def pred_1(self, x, **kwargs):
        return self(x, pred_frames=1, **kwargs)[0].squeeze(dim=1)              # st_phy.py:87-88

    def forward(self, x, pred_frames=1, **kwargs):
        if kwargs.get("train", False):
            raise NotImplementedError("the native rollout is inference-only")
        b, t, c, h, w = x.shape
        if (c, h, w) != (self.img_c, self.img_h, self.img_w):
            raise ValueError(f"shape mismatch: expected {(self.img_c, self.img_h, self.img_w)}, got {(c, h, w)}")
        actions = self._native_actions(kwargs, b, x.device)                    # st_phy.py:98-103
        pred, _ = self._native_forward(x, int(pred_frames), t, actions=actions)
        return pred, None                                                      # st_phy.py:176-181 (eval)


MODEL_CLASSES = {
    "convlstm-shi": EF_ConvLSTM,
    "predrnn-pp": PredRNN_V2,
    "phy": PhyDNet,
    "convlstm-branch": ConvLSTMBranch,
    "st-phy": STPhy,
    "trajgru": EF_TrajGRU,
    "predrnn-pp-causal": PredRNNpp,      # the north star's Causal LSTM + GHU stack; no reference twin (parity unpinned)
}
