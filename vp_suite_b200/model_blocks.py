"""
Drop-in VPModelBlock classes: the recurrent cells of the hot path, one native call per timestep.

Signatures, parameter names/shapes and stateful behaviour follow the reference blocks (file:line in each class);
the arithmetic happens in libvpk's single-step cell entry points (include/vpk.h), never in Python.

    ConvLSTM                <- vp_suite/model_blocks/conv_lstm_hzzone.py:7-70     (Shi et al., peepholes)
    ConvLSTMCell            <- vp_suite/model_blocks/conv_lstm_ndrplz.py:7-48     (ndrplz cell)
    SingleStepConvLSTM      <- vp_suite/model_blocks/phydnet.py:117-175
    SpatioTemporalLSTMCell  <- vp_suite/model_blocks/predrnn.py:7-83
    ActionConditionalSpatioTemporalLSTMCell <- vp_suite/model_blocks/predrnn.py:86-169
    PhyCell_Cell / PhyCell  <- vp_suite/model_blocks/phydnet.py:13-114            (with and without actions)
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _native as N
from .base import VPModelBlock
from .models import _gn_divisor


class _NativeCell:
    """Owns a ``vpk_cell`` handle; rebuilt when the module's weights change."""
    precision: str = "fp32"      #: cells default to the fp32-operand mode (the reference's own numerics)
    backend: str = "auto"

    def _cell_init(self):
        self._cell = None
        self._cell_versions = None

    def __getstate__(self):                       # pickling drops the native handle; it is rebuilt on the next call
        state = dict(self.__dict__)
        state.pop("_cell", None)
        state.pop("_cell_versions", None)
        return state

    def __setstate__(self, state):
        nn.Module.__setstate__(self, state)
        self._cell_init()

    def _cell_handle(self, device):
        """Handle with current weights, living on ``device`` (libvpk allocates on the calling thread's current device)."""
        versions = tuple((k, v._version, v.data_ptr()) for k, v in self.state_dict().items()) + (self.precision,
                                                                                               self.backend, str(device))
        if self._cell is None or versions != self._cell_versions:
            self._cell_release()
            with torch.cuda.device(device):
                self._cell = self._cell_create()
            self._cell_versions = versions
        return self._cell

    @staticmethod
    def _step(device, fn, *args):
        """One native cell step with the tensors' device current."""
        with torch.cuda.device(device):
            N.check(fn(*args))

    def _cell_release(self):
        if getattr(self, "_cell", None) is not None:
            N.lib().vpk_cell_destroy(self._cell)
            self._cell = None

    def __del__(self):
        try:
            self._cell_release()
        except Exception:
            pass

    @staticmethod
    def _host(t):
        return t.detach().to("cpu", torch.float32).contiguous()

    @staticmethod
    def _dev(t):
        if not t.is_cuda:
            raise N.NativeError("vp_suite_b200 blocks run on CUDA tensors only (there is no CPU path)")
        return t.detach().to(torch.float32).contiguous()

    @staticmethod
    def _stream(t):
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _ConvLSTMCellFunction(torch.autograd.Function):
    """One ConvLSTMCell step with a native backward (vpk_convlstm_cell_backward): the first differentiable entry of the
    drop-in.  Saves the step's inputs; the backward recomputes the gates in the library."""

    @staticmethod
    def forward(ctx, module, x, h, c, weight, bias):
        ctx.module = module
        ctx.save_for_backward(x, h, c)
        return module._native_step(x, h, c)

    @staticmethod
    def backward(ctx, dh_out, dc_out):
        module = ctx.module
        x, h, c = ctx.saved_tensors
        cell = module._cell_handle(x.device)
        dx, dh, dc = torch.empty_like(x), torch.empty_like(h), torch.empty_like(c)
        dw = torch.empty_like(module.conv.weight, dtype=torch.float32)
        db = torch.empty_like(module.conv.bias, dtype=torch.float32) if module.conv.bias is not None else None
        gh = None if dh_out is None else module._dev(dh_out)
        gc = None if dc_out is None else module._dev(dc_out)
        module._step(x.device, N.lib().vpk_convlstm_cell_backward, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), N.ptr(gh),
                     N.ptr(gc), N.ptr(dx), N.ptr(dh), N.ptr(dc), N.ptr(dw), N.ptr(db), module._stream(x))
        return None, dx, dh, dc, dw, db


class ConvLSTMCell(_NativeCell, nn.Module):
    """conv_lstm_ndrplz.py:7-48: gate conv over cat(x, h), split order (i, f, o, g), no peepholes.  Differentiable: when
    gradients are enabled and any input or parameter requires them, the step runs through an autograd.Function whose
    backward is libvpk's (input, state, weight and bias gradients), so BPTT over a stack / sequence of these cells composes
    as with the reference cell."""

    def __init__(self, input_dim, hidden_dim, kernel_size, bias):
        super().__init__()
        self._cell_init()
        self.input_dim, self.hidden_dim, self.kernel_size, self.bias = input_dim, hidden_dim, kernel_size, bias
        if kernel_size[0] != kernel_size[1] or kernel_size[0] % 2 == 0:
            raise ValueError("square odd kernels only")
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.conv = nn.Conv2d(input_dim + hidden_dim, 4 * hidden_dim, kernel_size, padding=self.padding, bias=bias)
        self._hw = None

    def _cell_create(self):
        h, w = self._hw
        cell = C.c_void_p()
        wt = self._host(self.conv.weight)
        b = self._host(self.conv.bias) if self.conv.bias is not None else None
        N.check(N.lib().vpk_convlstm_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend],
                                                 self.input_dim, self.hidden_dim, h, w, self.kernel_size[0], 1,
                                                 N.ptr(wt), N.ptr(b), C.byref(cell)))
        return cell

    def _native_step(self, x, h, c):
        if self._hw != tuple(x.shape[-2:]):
            self._hw = tuple(x.shape[-2:])
            self._cell_release()
        cell = self._cell_handle(x.device)
        h_next, c_next = torch.empty_like(h), torch.empty_like(c)
        self._step(x.device, N.lib().vpk_convlstm_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), None, None,
                   None, N.ptr(h_next), N.ptr(c_next), self._stream(x))
        return h_next, c_next

    def forward(self, input_tensor, cur_state):
        h_cur, c_cur = cur_state
        needs_grad = torch.is_grad_enabled() and (
            any(t.requires_grad for t in (input_tensor, h_cur, c_cur)) or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            for t in (input_tensor, h_cur, c_cur):
                if not t.is_cuda:
                    raise N.NativeError("vp_suite_b200 blocks run on CUDA tensors only (there is no CPU path)")
            x, h, c = (t.to(torch.float32).contiguous() for t in (input_tensor, h_cur, c_cur))
            return _ConvLSTMCellFunction.apply(self, x, h, c, self.conv.weight, self.conv.bias)
        return self._native_step(self._dev(input_tensor), self._dev(h_cur), self._dev(c_cur))

    def init_hidden(self, batch_size, image_size):                       # conv_lstm_ndrplz.py:45-48
        height, width = image_size
        dev = self.conv.weight.device
        return (torch.zeros(batch_size, self.hidden_dim, height, width, device=dev),
                torch.zeros(batch_size, self.hidden_dim, height, width, device=dev))


class _ShiConvLSTMStepFunction(torch.autograd.Function):
    """One timestep of the Shi-et-al. ConvLSTM (peepholes; x may be None = the forecaster's zero input) with a native backward
    (vpk_convlstm_cell_backward_peep).  Saves the step's inputs; the backward recomputes the gates in the library and returns the
    gradients of x, h, c, the conv weight / bias and the three peepholes; BPTT over the sequence composes through autograd."""

    @staticmethod
    def forward(ctx, module, x, h, c, weight, bias, wci, wcf, wco):
        ctx.module = module
        ctx.has_x = x is not None
        ctx.save_for_backward(*([x] if x is not None else []), h, c, wci, wcf, wco)
        d = module._dev
        return module._native_step(None if x is None else d(x), d(h), d(c), d(wci), d(wcf), d(wco))

    @staticmethod
    def backward(ctx, dh_out, dc_out):
        module = ctx.module
        saved = list(ctx.saved_tensors)
        x = saved.pop(0) if ctx.has_x else None
        h, c, wci, wcf, wco = saved
        cell = module._cell_handle(h.device)
        d = module._dev
        x, h, c, wci, wcf, wco = (None if t is None else d(t) for t in (x, h, c, wci, wcf, wco))
        dx = torch.empty_like(x) if x is not None else None
        dh, dc = torch.empty_like(h), torch.empty_like(c)
        dw = torch.empty_like(module._conv.weight, dtype=torch.float32)
        db = torch.empty_like(module._conv.bias, dtype=torch.float32)
        dp = [torch.empty_like(p, dtype=torch.float32) for p in (wci, wcf, wco)]
        gh = None if dh_out is None else module._dev(dh_out)
        gc = None if dc_out is None else module._dev(dc_out)
        module._step(h.device, N.lib().vpk_convlstm_cell_backward_peep, cell, h.shape[0], N.ptr(x), N.ptr(h), N.ptr(c),
                     N.ptr(wci), N.ptr(wcf), N.ptr(wco), N.ptr(gh), N.ptr(gc), N.ptr(dx), N.ptr(dh), N.ptr(dc), N.ptr(dw),
                     N.ptr(db), N.ptr(dp[0]), N.ptr(dp[1]), N.ptr(dp[2]), module._stream(h))
        return None, dx, dh, dc, dw, db, dp[0], dp[1], dp[2]


class ConvLSTM(_NativeCell, VPModelBlock):
    """conv_lstm_hzzone.py:7-70: whole-sequence driver of the Shi-et-al. ConvLSTM with peepholes.  Differentiable like
    ConvLSTMCell: with gradients enabled every timestep runs through an autograd.Function whose backward is libvpk's."""
    NAME = "ConvLSTM (Shi et al.)"
    PAPER_REFERENCE = "https://arxiv.org/abs/1506.04214"
    CODE_REFERENCE = "https://github.com/Hzzone/Precipitation-Nowcasting"
    MATCHES_REFERENCE = "Yes"

    def __init__(self, device, in_channels, enc_channels, state_h, state_w, kernel_size, stride=1, padding=1):
        super().__init__()
        self._cell_init()
        if stride != 1 or padding != kernel_size // 2 or kernel_size % 2 == 0:
            raise ValueError("only stride 1 / 'same' padding keeps the state size (as all reference configs do)")
        self.device = device
        self._conv = nn.Conv2d(in_channels + enc_channels, enc_channels * 4, kernel_size, stride, padding)
        self.state_h, self.state_w = state_h, state_w
        # registered on every device (the reference registers them on CPU only, conv_lstm_hzzone.py:30-32)
        self.Wci = nn.Parameter(torch.zeros(1, enc_channels, state_h, state_w))
        self.Wcf = nn.Parameter(torch.zeros(1, enc_channels, state_h, state_w))
        self.Wco = nn.Parameter(torch.zeros(1, enc_channels, state_h, state_w))
        self.in_c, self.enc_c, self._k = in_channels, enc_channels, kernel_size
        self.to(device)

    def _cell_create(self):
        cell = C.c_void_p()
        wt, b = self._host(self._conv.weight), self._host(self._conv.bias)
        N.check(N.lib().vpk_convlstm_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend], self.in_c,
                                                 self.enc_c, self.state_h, self.state_w, self._k, 0, N.ptr(wt),
                                                 N.ptr(b), C.byref(cell)))
        return cell

    def forward(self, inputs, states, seq_len):
        dev = self._conv.weight.device
        if states is None:                                               # conv_lstm_hzzone.py:39-45
            b = inputs.shape[0]
            c = torch.zeros((b, self.enc_c, self.state_h, self.state_w), dtype=torch.float, device=dev)
            h = torch.zeros((b, self.enc_c, self.state_h, self.state_w), dtype=torch.float, device=dev)
        else:
            h, c = states
            b = h.shape[0]
        params = (self._conv.weight, self._conv.bias, self.Wci, self.Wcf, self.Wco)
        grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or h.requires_grad or c.requires_grad or
                                            (inputs is not None and inputs.requires_grad))
        outputs = []
        if grad:
            # differentiable path: the tensors stay attached; each timestep is one autograd node with a native backward
            for t_ in (h, c) + (() if inputs is None else (inputs,)):
                if not t_.is_cuda:
                    raise N.NativeError("vp_suite_b200 blocks run on CUDA tensors only (there is no CPU path)")
            h, c = h.to(torch.float32).contiguous(), c.to(torch.float32).contiguous()
            for t in range(seq_len):
                x = None if inputs is None else inputs[:, t].to(torch.float32).contiguous()
                h, c = _ShiConvLSTMStepFunction.apply(self, x, h, c, *params)
                outputs.append(h)
            return torch.stack(outputs, dim=1), (h, c)
        h, c = self._dev(h), self._dev(c)
        peep = [self._dev(p) for p in (self.Wci, self.Wcf, self.Wco)]
        for t in range(seq_len):                                         # conv_lstm_hzzone.py:52-69
            x = None if inputs is None else self._dev(inputs[:, t])      # None = all-zero input (:54-56)
            h, c = self._native_step(x, h, c, *peep)
            outputs.append(h)
        return torch.stack(outputs, dim=1), (h, c)

    def _native_step(self, x, h, c, wci, wcf, wco):
        cell = self._cell_handle(h.device)
        h_new, c_new = torch.empty_like(h), torch.empty_like(c)
        self._step(h.device, N.lib().vpk_convlstm_cell_step, cell, h.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), N.ptr(wci),
                   N.ptr(wcf), N.ptr(wco), N.ptr(h_new), N.ptr(c_new), self._stream(h))
        return h_new, c_new


class SingleStepConvLSTM(nn.Module):
    """model_blocks/phydnet.py:117-175: time-major stack of ConvLSTMCell with module-held state."""

    def __init__(self, input_size, input_dim, hidden_dims, n_layers, kernel_size, action_conditional, action_size,
                 device):
        super().__init__()
        self.input_size, self.input_dim, self.hidden_dims = input_size, input_dim, hidden_dims
        self.n_layers, self.kernel_size = n_layers, kernel_size
        self.H, self.C = [], []
        self.action_size, self.action_conditional, self.device = action_size, action_conditional, device
        cells, cur = [], input_dim + (action_size if action_conditional else 0)      # phydnet.py:137
        for i in range(n_layers):
            cells.append(ConvLSTMCell(cur, hidden_dims[i], kernel_size, True))
            cur = hidden_dims[i]
        self.cell_list = nn.ModuleList(cells)

    def forward(self, frame, action, first_timestep=False):
        if first_timestep:
            self.init_hidden(frame.shape[0], frame.device)
        inp = frame
        if self.action_conditional:                                      # phydnet.py:153-155 (layout only: the inflated
            infl = action.unsqueeze(-1).unsqueeze(-1).expand(-1, -1, *self.input_size)   # action joins the channels)
            inp = torch.cat([inp, infl.to(inp.dtype)], dim=-3)
        for j, cell in enumerate(self.cell_list):                        # phydnet.py:157-161
            self.H[j], self.C[j] = cell(inp, (self.H[j], self.C[j]))
            inp = self.H[j]
        return (self.H, self.C), self.H

    def init_hidden(self, batch_size, device=None):                      # phydnet.py:165-171 (states on the input's device)
        device = device or self.device
        self.H = [torch.zeros(batch_size, hd, *self.input_size, device=device) for hd in self.hidden_dims[:self.n_layers]]
        self.C = [torch.zeros(batch_size, hd, *self.input_size, device=device) for hd in self.hidden_dims[:self.n_layers]]

    def set_hidden(self, hidden):
        self.H, self.C = hidden


class SpatioTemporalLSTMCell(_NativeCell, VPModelBlock):
    """model_blocks/predrnn.py:7-83, layer_norm False or True."""
    NAME = "Spatio-Temporal LSTM Cell"
    PAPER_REFERENCE = "https://arxiv.org/abs/2103.09504"
    CODE_REFERENCE = "https://github.com/thuml/predrnn-pytorch"
    MATCHES_REFERENCE = "Yes"

    def __init__(self, in_channel, num_hidden, height, width, filter_size, stride, layer_norm):
        super().__init__()
        self._cell_init()
        if stride != 1 or filter_size % 2 == 0:
            raise ValueError("stride 1 and odd filter sizes only")
        self.num_hidden, self.padding, self._forget_bias = num_hidden, filter_size // 2, 1.0
        self._shape = (in_channel, height, width, filter_size)
        self._layer_norm = bool(layer_norm)

        def conv(ci, co):
            layers = [nn.Conv2d(ci, co, filter_size, stride, self.padding, bias=False)]
            if layer_norm:                                               # model_blocks/predrnn.py:24-40
                layers.append(nn.LayerNorm([co, height, width]))
            return nn.Sequential(*layers)
        self.conv_x = conv(in_channel, num_hidden * 7)
        self.conv_h = conv(num_hidden, num_hidden * 4)
        self.conv_m = conv(num_hidden, num_hidden * 3)
        self.conv_o = conv(num_hidden * 2, num_hidden)
        self.conv_last = nn.Conv2d(num_hidden * 2, num_hidden, 1, 1, 0, bias=False)

    def _cell_create(self):
        cin, h, w, k = self._shape
        cell = C.c_void_p()
        ws = [self._host(m) for m in (self.conv_x[0].weight, self.conv_h[0].weight, self.conv_m[0].weight,
                                      self.conv_o[0].weight, self.conv_last.weight)]
        N.check(N.lib().vpk_stlstm_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend], cin,
                                               self.num_hidden, h, w, k, *[N.ptr(t) for t in ws], C.byref(cell)))
        if self._layer_norm:
            ln = [self._host(t) for seq in (self.conv_x, self.conv_h, self.conv_m, self.conv_o)
                  for t in (seq[1].weight, seq[1].bias)]
            N.check(N.lib().vpk_stlstm_cell_set_layer_norm(cell, *[N.ptr(t) for t in ln]))
        return cell

    def forward(self, x_t, h_t, c_t, m_t):
        x, h, c, m = (self._dev(t) for t in (x_t, h_t, c_t, m_t))
        cell = self._cell_handle(x.device)
        outs = [torch.empty_like(h) for _ in range(5)]
        self._step(x.device, N.lib().vpk_stlstm_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), N.ptr(m),
                   *[N.ptr(o) for o in outs], self._stream(x))
        return tuple(outs)                                               # h', c', m', delta_c, delta_m (predrnn.py:82)


class CausalLSTMCell(_NativeCell, VPModelBlock):
    """Causal LSTM cell of PredRNN++ (Wang et al., ICML 2018, eq. 1): the temporal memory c and the spatial memory m in
    cascade, tanh output gate.  The north star names it; the reference checkout has no such block, so
    it has no reference twin -- PARITY UNPINNED, checked against oracle/causal.py.  Constructor and forward signatures
    follow SpatioTemporalLSTMCell above (model_blocks/predrnn.py:7-83); bias-free convs, forget bias 1."""
    NAME = "Causal LSTM Cell"
    PAPER_REFERENCE = "https://arxiv.org/abs/1804.06300"
    CODE_REFERENCE = "https://github.com/Yunbo426/predrnn-pp"
    MATCHES_REFERENCE = "Not Yet"
    _CONVS = ("conv_x", "conv_h", "conv_c", "conv_m", "conv_c2m", "conv_om")

    def __init__(self, in_channel, num_hidden, height, width, filter_size, stride=1, layer_norm=False, num_hidden_in=None):
        super().__init__()
        self._cell_init()
        if stride != 1 or filter_size % 2 == 0:
            raise ValueError("stride 1 and odd filter sizes only")
        if layer_norm:
            raise NotImplementedError("Causal LSTM drop-in: layer_norm is not built")
        # num_hidden_in: channels of the spatial memory the cell READS (the width of the cell that wrote it; the paper's
        # stacks have unequal widths, 128-64-64-64); default: this cell's own width
        self.num_hidden, self.padding, self._forget_bias = num_hidden, filter_size // 2, 1.0
        self.num_hidden_in = num_hidden if num_hidden_in is None else int(num_hidden_in)
        self._shape = (in_channel, height, width, filter_size)

        def conv(ci, co):
            return nn.Sequential(nn.Conv2d(ci, co, filter_size, stride, self.padding, bias=False))
        self.conv_x = conv(in_channel, num_hidden * 7)
        self.conv_h = conv(num_hidden, num_hidden * 4)
        self.conv_c = conv(num_hidden, num_hidden * 3)
        self.conv_m = conv(self.num_hidden_in, num_hidden * 3)
        self.conv_c2m = conv(num_hidden, num_hidden * 4)
        self.conv_om = conv(num_hidden, num_hidden)
        self.conv_last = nn.Conv2d(num_hidden * 2, num_hidden, 1, 1, 0, bias=False)

    def _cell_create(self):
        cin, h, w, k = self._shape
        cell = C.c_void_p()
        ws = [self._host(getattr(self, n)[0].weight) for n in self._CONVS] + [self._host(self.conv_last.weight)]
        wp = (C.c_void_p * 7)(*[t.data_ptr() for t in ws])
        N.check(N.lib().vpk_causal_lstm_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend], cin,
                                                    self.num_hidden_in, self.num_hidden, h, w, k, wp, C.byref(cell)))
        return cell

    def forward(self, x_t, h_t, c_t, m_t):
        x, h, c, m = (self._dev(t) for t in (x_t, h_t, c_t, m_t))
        cell = self._cell_handle(x.device)
        outs = [torch.empty_like(h) for _ in range(3)]
        self._step(x.device, N.lib().vpk_causal_lstm_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), N.ptr(m),
                   *[N.ptr(o) for o in outs], self._stream(x))
        return tuple(outs)                                               # h', c', m'


class GHU(_NativeCell, VPModelBlock):
    """Gradient highway unit of PredRNN++ (eq. 2): z' = s z + (1 - s) tanh(p) with (p, s) from convs over x and z; ``z``
    None means zeros (first timestep).  No reference twin (see CausalLSTMCell): PARITY UNPINNED."""
    NAME = "Gradient Highway Unit"
    PAPER_REFERENCE = "https://arxiv.org/abs/1804.06300"
    CODE_REFERENCE = "https://github.com/Yunbo426/predrnn-pp"
    MATCHES_REFERENCE = "Not Yet"

    def __init__(self, num_hidden, height, width, filter_size, stride=1, layer_norm=False):
        super().__init__()
        self._cell_init()
        if stride != 1 or filter_size % 2 == 0:
            raise ValueError("stride 1 and odd filter sizes only")
        if layer_norm:
            raise NotImplementedError("GHU drop-in: layer_norm is not built")
        self.num_hidden, self.padding = num_hidden, filter_size // 2
        self._shape = (height, width, filter_size)
        self.x_concat = nn.Sequential(nn.Conv2d(num_hidden, num_hidden * 2, filter_size, stride, self.padding, bias=False))
        self.z_concat = nn.Sequential(nn.Conv2d(num_hidden, num_hidden * 2, filter_size, stride, self.padding, bias=False))

    def _cell_create(self):
        h, w, k = self._shape
        cell = C.c_void_p()
        wx, wz = self._host(self.x_concat[0].weight), self._host(self.z_concat[0].weight)
        N.check(N.lib().vpk_ghu_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend], self.num_hidden, h, w, k,
                                            N.ptr(wx), N.ptr(wz), C.byref(cell)))
        return cell

    def forward(self, x, z=None):
        x = self._dev(x)
        z = torch.zeros_like(x) if z is None else self._dev(z)
        cell = self._cell_handle(x.device)
        out = torch.empty_like(x)
        self._step(x.device, N.lib().vpk_ghu_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(z), N.ptr(out), self._stream(x))
        return out


class ActionConditionalSpatioTemporalLSTMCell(_NativeCell, VPModelBlock):
    """model_blocks/predrnn.py:86-169, layer_norm False or True: every conv has a bias, and a fifth conv ``conv_a`` over the
    action tensor multiplies conv_h's output before the gate split (:149)."""
    NAME = "Spatio-Temporal LSTM Cell (Action-Conditional)"
    PAPER_REFERENCE = "https://arxiv.org/abs/2103.09504"
    CODE_REFERENCE = "https://github.com/thuml/predrnn-pytorch"
    MATCHES_REFERENCE = "Yes"
    _CONVS = ("conv_x", "conv_h", "conv_a", "conv_m", "conv_o")

    def __init__(self, in_channel, num_hidden, height, width, filter_size, stride, layer_norm):
        super().__init__()
        self._cell_init()
        if stride != 1 or filter_size % 2 == 0:
            raise ValueError("stride 1 and odd filter sizes only")
        self.num_hidden, self.padding, self._forget_bias = num_hidden, filter_size // 2, 1.0
        self._shape = (in_channel, height, width, filter_size)
        self._layer_norm = bool(layer_norm)

        def conv(ci, co):
            layers = [nn.Conv2d(ci, co, filter_size, stride, self.padding)]
            if layer_norm:
                layers.append(nn.LayerNorm([co, height, width]))
            return nn.Sequential(*layers)
        self.conv_x = conv(in_channel, num_hidden * 7)                   # same construction order as the reference
        self.conv_h = conv(num_hidden, num_hidden * 4)
        self.conv_a = conv(num_hidden, num_hidden * 4)
        self.conv_m = conv(num_hidden, num_hidden * 3)
        self.conv_o = conv(num_hidden * 2, num_hidden)
        self.conv_last = nn.Conv2d(num_hidden * 2, num_hidden, 1, 1, 0)

    def _cell_create(self):
        cin, h, w, k = self._shape
        cell = C.c_void_p()
        mods = [getattr(self, n)[0] for n in self._CONVS] + [self.conv_last]
        ws = [self._host(m.weight) for m in mods]
        bs = [self._host(m.bias) for m in mods]
        wp = (C.c_void_p * 6)(*[t.data_ptr() for t in ws])
        bp = (C.c_void_p * 6)(*[t.data_ptr() for t in bs])
        N.check(N.lib().vpk_stlstm_ac_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend], cin,
                                                  self.num_hidden, h, w, k, wp, bp, C.byref(cell)))
        if self._layer_norm:
            ln = [self._host(t) for n in self._CONVS for t in (getattr(self, n)[1].weight, getattr(self, n)[1].bias)]
            lp = (C.c_void_p * 10)(*[t.data_ptr() for t in ln])
            N.check(N.lib().vpk_stlstm_ac_cell_set_layer_norm(cell, lp))
        return cell

    def forward(self, x_t, h_t, c_t, m_t, a_t):
        x, h, c, m, a = (self._dev(t) for t in (x_t, h_t, c_t, m_t, a_t))
        cell = self._cell_handle(x.device)
        outs = [torch.empty_like(h) for _ in range(5)]
        self._step(x.device, N.lib().vpk_stlstm_ac_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(c), N.ptr(m),
                   N.ptr(a), *[N.ptr(o) for o in outs], self._stream(x))
        return tuple(outs)                                               # h', c', m', delta_c, delta_m (predrnn.py:169)


class PhyCell_Cell(_NativeCell, VPModelBlock):
    """model_blocks/phydnet.py:13-62, with or without the two 1x1 action convs (:44-55)."""
    NAME = "PhyCell - Cell"
    PAPER_REFERENCE = "https://arxiv.org/abs/2003.01460"
    CODE_REFERENCE = "https://github.com/vincent-leguen/PhyDNet"
    MATCHES_REFERENCE = "Not Yet"

    def __init__(self, input_dim, action_conditional, action_size, hidden_dim, kernel_size, bias=True):
        super().__init__()
        self._cell_init()
        if not bias or kernel_size[0] != kernel_size[1] or kernel_size[0] % 2 == 0:
            raise ValueError("bias=True and square odd kernels only")
        self.input_dim, self.action_size, self.action_conditional = input_dim, action_size, action_conditional
        self.F_hidden_dim, self.kernel_size, self.bias = hidden_dim, kernel_size, bias
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.F = nn.Sequential()
        self.F.add_module("conv1", nn.Conv2d(input_dim, hidden_dim, kernel_size, (1, 1), self.padding))
        self.F.add_module("bn1", nn.GroupNorm(_gn_divisor(hidden_dim), hidden_dim))
        self.F.add_module("conv2", nn.Conv2d(hidden_dim, input_dim, (1, 1), (1, 1), (0, 0)))
        self.convgate = nn.Conv2d(2 * input_dim, input_dim, (3, 3), padding=(1, 1), bias=bias)
        if action_conditional:                                           # phydnet.py:44-48
            self.frame_action_conv = nn.Conv2d(input_dim + action_size, input_dim, (1, 1))
            self.hidden_action_conv = nn.Conv2d(input_dim + action_size, input_dim, (1, 1))
        self._hw = None

    def _cell_create(self):
        h, w = self._hw
        cell = C.c_void_p()
        ts = [self._host(t) for t in (self.F.conv1.weight, self.F.conv1.bias, self.F.bn1.weight, self.F.bn1.bias,
                                      self.F.conv2.weight, self.F.conv2.bias, self.convgate.weight, self.convgate.bias)]
        N.check(N.lib().vpk_phycell_cell_create(N.PRECISIONS[self.precision], N.BACKENDS[self.backend],
                                                self.input_dim, self.F_hidden_dim, h, w, self.kernel_size[0],
                                                *[N.ptr(t) for t in ts], C.byref(cell)))
        if self.action_conditional:
            ac = [self._host(t) for t in (self.frame_action_conv.weight, self.frame_action_conv.bias,
                                          self.hidden_action_conv.weight, self.hidden_action_conv.bias)]
            N.check(N.lib().vpk_phycell_cell_set_action_convs(cell, int(self.action_size), *[N.ptr(t) for t in ac]))
        return cell

    def forward(self, frame, action, hidden):
        x, h = self._dev(frame), self._dev(hidden)
        if self._hw != tuple(x.shape[-2:]):
            self._hw = tuple(x.shape[-2:])
            self._cell_release()
        cell = self._cell_handle(x.device)
        out = torch.empty_like(h)
        if self.action_conditional:
            if action is None or action.dim() != 2 or action.shape[-1] != self.action_size:
                raise ValueError("Given actions are None or of the wrong size!")
            a = self._dev(action)
            self._step(x.device, N.lib().vpk_phycell_cell_step_action, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(a),
                       N.ptr(out), self._stream(x))
        else:
            self._step(x.device, N.lib().vpk_phycell_cell_step, cell, x.shape[0], N.ptr(x), N.ptr(h), N.ptr(out), self._stream(x))
        return out


class PhyCell(VPModelBlock):
    """model_blocks/phydnet.py:65-114: stack of PhyCell_Cell with module-held state."""
    NAME = "PhyCell"
    PAPER_REFERENCE = "https://arxiv.org/abs/2003.01460"
    CODE_REFERENCE = "https://github.com/vincent-leguen/PhyDNet"
    MATCHES_REFERENCE = "Not Yet"

    def __init__(self, input_size, input_dim, hidden_dims, n_layers, kernel_size, action_conditional, action_size,
                 device):
        super().__init__()
        self.input_size, self.input_dim, self.hidden_dims = input_size, input_dim, hidden_dims
        self.n_layers, self.kernel_size, self.H, self.device = n_layers, kernel_size, [], device
        self.cell_list = nn.ModuleList([PhyCell_Cell(input_dim, action_conditional, action_size, hidden_dims[i],
                                                     kernel_size) for i in range(n_layers)])

    def forward(self, frame, action, first_timestep=False):
        if first_timestep:
            self.init_hidden(frame.shape[0], frame.device)
        for j, cell in enumerate(self.cell_list):                        # phydnet.py:100-104
            self.H[j] = cell(frame if j == 0 else self.H[j - 1], action, self.H[j])
        return self.H, self.H

    def init_hidden(self, batch_size, device=None):
        device = device or self.device
        self.H = [torch.zeros(batch_size, self.input_dim, self.input_size[0], self.input_size[1], device=device)
                  for _ in range(self.n_layers)]

    def _set_hidden(self, H):
        self.H = H


MODEL_BLOCK_CLASSES = [ConvLSTM, SpatioTemporalLSTMCell, ActionConditionalSpatioTemporalLSTMCell, PhyCell_Cell, PhyCell]
