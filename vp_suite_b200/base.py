"""
Host side of the reference's model contract.

When the reference package is importable (``import vp_suite`` works, e.g. it is installed, or the tests' import shim
registered it), ``VPModel`` / ``VPModelBlock`` here ARE subclasses of ``vp_suite.base.VPModel`` /
``vp_suite.base.VPModelBlock``: constructor, ``config``, ``unpack_data``, ``eval_iter`` are the reference's own code and
``isinstance(model, vp_suite.base.VPModel)`` holds, so ``VPSuite.test()`` (vp_suite/vpsuite.py:536-550) drives a
registered drop-in unchanged.  Otherwise a faithful mirror of vp_suite/base/base_model.py:11-216 is used (constructor
``(device, **model_kwargs)``, ``REQUIRED_ARGS``, kwargs -> type-checked attributes, ``config``, ``unpack_data``,
``eval_iter``, ``train_iter``).  ``train_iter`` refuses for the drop-ins whose rollout is inference-only (``TRAINABLE =
False``, which ``VPSuite.train`` honours, vpsuite.py:312) and is the base class's own loop for the trainable one
(EF_ConvLSTM: differentiable ConvLSTM layers).  ``NativeRollout`` is the glue to libvpk: it mirrors the module's
``state_dict`` into the native handle and runs ``forward`` through the C ABI.  Nothing here computes frames in Python --
without the CUDA library the calls raise.
"""
import ctypes as C
import inspect
import os

import torch
import torch.nn as nn

from . import _native as N


def _reference_bases():
    """(vp_suite.base.VPModel, vp_suite.base.VPModelBlock) when the reference package imports, else (None, None).
    VPK_NO_REFERENCE=1 forces the mirror (tests)."""
    if os.environ.get("VPK_NO_REFERENCE"):
        return None, None
    try:
        from vp_suite.base.base_model import VPModel as ref_model
        from vp_suite.base.base_model_block import VPModelBlock as ref_block
        return ref_model, ref_block
    except Exception:          # not installed, or its own imports (datasets, piqa, ...) fail in this environment
        return None, None


_RefVPModel, _RefVPModelBlock = _reference_bases()
REFERENCE_BASE = _RefVPModel is not None          #: True: the drop-ins subclass the real vp_suite.base.VPModel


def _set_from_kwarg(obj, kwargs, name, required=False):
    """vp_suite/utils/utils.py:113-156 (the parts VPModel uses): class attribute = default, kwargs override it,
    mismatching types raise ValueError."""
    if required and name not in kwargs:
        raise ValueError(f"missing required parameter '{name}' for object '{obj.__class__}'")
    default = getattr(obj, name, None)
    val = kwargs.get(name, default)
    if default is not None and not isinstance(val, type(default)):
        raise ValueError(f"mismatching types for parameter '{name}' for object '{obj.__class__}'")
    setattr(obj, name, val)


class _MirrorVPModelBlock(nn.Module):
    """Marker base of the model blocks (vp_suite/base/base_model_block.py:4-13)."""
    NAME: str = __name__
    PAPER_REFERENCE = None
    CODE_REFERENCE = None
    MATCHES_REFERENCE: str = None


class _MirrorVPModel(nn.Module):
    """vp_suite/base/base_model.py:11-216 restated (used only when the reference package is not importable)."""
    # base_model.py:18-36
    NON_CONFIG_VARS = ["functions", "model_dir", "dump_patches", "training"]
    NAME = None
    PAPER_REFERENCE = None
    CODE_REFERENCE = None
    MATCHES_REFERENCE: str = None
    REQUIRED_ARGS = ["img_shape", "action_size", "tensor_value_range"]
    CAN_HANDLE_ACTIONS = False
    TRAINABLE = True
    NEEDS_COMPLETE_INPUT = False
    MIN_CONTEXT_FRAMES = 1

    model_dir = None
    img_shape = None
    action_size = None
    action_conditional = False
    tensor_value_range = None

    def __init__(self, device, **model_kwargs):
        super().__init__()
        self.device = device
        for arg in self.REQUIRED_ARGS:                                   # base_model.py:51-64
            if arg == "tensor_value_range":
                val = model_kwargs.get(arg, (0, 0))
                if type(val) not in (tuple, list) or len(val) != 2:
                    raise ValueError("value for argument 'tensor_value_range' needs to be tuple or list with 2 elems")
            _set_from_kwarg(self, model_kwargs, arg, required=True)
            if arg == "img_shape":
                self.img_c, self.img_h, self.img_w = self.img_shape
        for arg in model_kwargs:                                         # base_model.py:66-69
            if arg not in self.REQUIRED_ARGS:
                _set_from_kwarg(self, model_kwargs, arg)

    @property
    def config(self):
        """All public, non-callable, non-tensor, non-module attributes (base_model.py:71-85, utils.py:208-234)."""
        out = {}
        for name in set(dir(self)):
            if name.startswith("_") or name[0].isupper() or name == "config":
                continue
            value = getattr(self, name)
            if inspect.isroutine(value) or isinstance(value, (nn.Module, torch.Tensor)):
                continue
            out[name] = value
        for k in self.NON_CONFIG_VARS:
            out.pop(k, None)
        c, h, w = self.img_shape
        out.update({"img_h": h, "img_w": w, "img_c": c, "NAME": self.NAME})
        return out

    def unpack_data(self, data, config, reverse=False, complete=False):
        """base_model.py:87-114: frames / actions of a VPData blob to the run's device, optional time reversal, split
        into context and target frames (or context + target as the input when the model NEEDS_COMPLETE_INPUT)."""
        img_data = data["frames"].to(config["device"])                   # [b, T, c, h, w]
        actions = data["actions"].to(config["device"])                   # [b, T-1, a]
        if img_data.ndim == 4:                                           # no batch dimension
            img_data = img_data.unsqueeze(0)
            actions = actions.unsqueeze(0)
        if reverse:
            img_data = torch.flip(img_data, dims=[1])
            actions = torch.flip(actions, dims=[1])
        t_in, t_pred = config["context_frames"], config["pred_frames"]
        if self.NEEDS_COMPLETE_INPUT or complete:
            input_frames = img_data[:, :t_in + t_pred]
            target_frames = input_frames[:, t_in:].clone()
        else:
            input_frames, target_frames = torch.split(img_data[:, :t_in + t_pred], [t_in, t_pred], dim=1)
        return input_frames, target_frames, actions

    def pred_1(self, x, **kwargs):
        raise NotImplementedError

    def forward(self, x, pred_frames=1, **kwargs):
        raise NotImplementedError

    def train_iter(self, config, loader, optimizer, loss_provider, epoch):
        """base_model.py:148-179: one pass over the training loader (forward, loss incl. the model's own losses, backward,
        optimizer step) -- used by the drop-ins that declare TRAINABLE = True."""
        for data in loader:
            inp, targets, actions = self.unpack_data(data, config)
            predictions, model_losses = self(inp, pred_frames=config["pred_frames"], actions=actions)
            _, total_loss = loss_provider.get_losses(predictions, targets)
            if model_losses is not None:
                for value in model_losses.values():
                    total_loss += value
            optimizer.zero_grad()
            total_loss.backward()
            optimizer.step()

    def eval_iter(self, config, loader, loss_provider):
        """base_model.py:181-216: one pass over the validation loader; returns ({loss name: mean}, indicator loss)."""
        self.eval()
        all_losses, indicator_losses = [], []
        with torch.no_grad():
            for data in loader:
                inp, targets, actions = self.unpack_data(data, config)
                predictions, _ = self(inp, pred_frames=config["pred_frames"], actions=actions)
                loss_values, _ = loss_provider.get_losses(predictions, targets)
                all_losses.append(loss_values)
                indicator_losses.append(loss_values[config["val_rec_criterion"]])
        indicator_loss = torch.stack(indicator_losses).mean()
        all_losses = {k: torch.stack([lv[k] for lv in all_losses]).mean().item() for k in all_losses[0].keys()}
        self.train()
        return all_losses, indicator_loss


class VPModelBlock(_RefVPModelBlock or _MirrorVPModelBlock):
    """Base of the drop-in blocks: the reference's marker class when importable (base_model_block.py:4-13)."""


class VPModel(_RefVPModel or _MirrorVPModel):
    """Base of the drop-in models: the reference's own VPModel when importable, else its mirror above."""
    TRAINABLE = False        # vpsuite.py:312 skips training for such models; the native rollout has no backward pass
    _BASE_TRAIN_ITER = (_RefVPModel or _MirrorVPModel).train_iter        # for the drop-ins that do train (EF_ConvLSTM)

    def train_iter(self, config, loader, optimizer, loss_provider, epoch):
        """base_model.py:148-179 needs gradients through forward(); libvpk is inference-only."""
        if self.TRAINABLE:
            return type(self)._BASE_TRAIN_ITER(self, config, loader, optimizer, loss_provider, epoch)
        raise NotImplementedError(f"{type(self).__name__} (vp_suite_b200) is an inference-only drop-in: forward() runs in "
                                  f"libvpk without autograd, so train_iter() is not available (TRAINABLE = False); train "
                                  f"with the reference class and load its state_dict / checkpoint here")


class NativeRollout:
    """Mixin: owns the libvpk model handle of a drop-in VPModel.

    Sub-classes provide ``_native_desc()`` (a filled ``ModelDesc``) and ``_native_key(key)`` (state_dict key ->
    native parameter name).  Public hyper-parameters live on the module like in the reference; the native state
    is kept in underscore attributes so that ``config`` (and run_cfg.json) stay clean.
    """
    precision: str = "bf16"      #: "bf16" = tcgen05 tensor-core path, "fp32" = fp32-operand CUDA-core path
    backend: str = "auto"        #: "auto" | "simt" (testing aid: same operands on the CUDA-core kernel)
    max_microbatch: int = 0      #: sequences per pass over the layers (0 = library default)
    use_cuda_graph: bool = False  #: replay the per-microbatch launch program as a CUDA graph

    def _native_init(self):
        self._handle = None
        self._handle_device = None
        self._versions = None
        self._workspaces = {}

    # -- pickling (torch.save(model) / torch.load, vpsuite.py:394,135): the native handle is dropped and rebuilt ----------
    _NATIVE_STATE = ("_handle", "_handle_device", "_versions", "_workspaces", "_host_out", "_train_blocks")

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._NATIVE_STATE:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        nn.Module.__setstate__(self, state)
        self._native_init()

    # -- handle lifecycle ---------------------------------------------------------------------------------------
    def _native_device(self):
        """CUDA device the handle lives on: the device of the module's parameters (the constructor moves them to
        ``self.device``; a later ``.to()`` is followed)."""
        p = next(iter(self.parameters()), None)
        if p is None or not p.is_cuda:
            raise N.NativeError("vp_suite_b200 models run on a CUDA device only (there is no CPU path); construct the "
                                "model with device='cuda[:i]'")
        return p.device

    def _native_handle(self, device=None):
        """The libvpk handle with current weights.  libvpk allocates on, and launches on, the calling thread's current
        device: every native call is made with the model's device current."""
        device = torch.device(device) if device is not None else self._native_device()
        N.lib()
        with torch.cuda.device(device):
            if self._handle is not None and self._handle_device != device:
                self._native_release()                 # the module moved: packed weights / programs live on the old device
            self._create_handle()
            self._handle_device = device
            versions = tuple((k, v._version, v.data_ptr()) for k, v in self.state_dict().items())
            if versions != self._versions:
                self._push_weights()
                self._versions = versions
        return self._handle

    def _create_handle(self):
        if self._handle is None:
            desc = self._native_desc()
            desc.precision = N.PRECISIONS[self.precision]
            desc.backend = N.BACKENDS[self.backend]
            desc.max_microbatch = int(self.max_microbatch)
            desc.use_cuda_graph = int(bool(self.use_cuda_graph))
            h = C.c_void_p()
            N.check(N.lib().vpk_model_create(C.byref(desc), C.byref(h)))
            self._handle = h
            self._versions = None

    def _push_weights(self):
        lib = N.lib()
        for key, val in self.state_dict().items():
            host = val.detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * host.dim())(*host.shape)
            N.check(lib.vpk_model_set_param(self._handle, self._native_key(key).encode(), N.ptr(host), shape,
                                            host.dim()))
        stream = torch.cuda.current_stream().cuda_stream
        N.check(lib.vpk_model_finalize(self._handle, C.c_void_p(stream)))

    def native_param_layout(self):
        """{native key: shape} the library expects (for layout tests)."""
        lib = N.lib()
        self._create_handle()
        n = C.c_int32()
        N.check(lib.vpk_model_num_params(self._handle, C.byref(n)))
        out = {}
        for i in range(n.value):
            key = C.c_char_p()
            shape = (C.c_int64 * 4)()
            nd = C.c_int32()
            N.check(lib.vpk_model_param_info(self._handle, i, C.byref(key), shape, C.byref(nd)))
            out[key.value.decode()] = tuple(shape[j] for j in range(nd.value))
        return out

    def _native_release(self):
        if getattr(self, "_handle", None) is not None:
            N.lib().vpk_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._native_release()
        except Exception:
            pass

    # -- forward through the C ABI --------------------------------------------------------------------------------
    def _native_actions(self, kwargs, batch, device):
        """The ``actions`` keyword of an action-conditional model as a contiguous fp32 [b, steps, action_size] tensor on
        ``device`` (host tensor for device=None), validated like the reference (predrnn_v2.py:147-151, models/phydnet.py:100-105:
        missing / all-zero-default or wrongly sized actions raise ValueError); None for other models."""
        if not self.action_conditional:
            return None
        actions = kwargs.get("actions", None)
        if actions is None or actions.dim() != 3 or actions.shape[0] != batch or actions.shape[-1] != self.action_size:
            raise ValueError("Given actions are None or of the wrong size!")
        actions = actions.detach().to(torch.float32)
        return (actions.cpu() if device is None else actions.to(device)).contiguous()

    def _native_forward(self, x, pred_frames, t_in, want_aux=False, actions=None):
        if not x.is_cuda:
            raise N.NativeError("vp_suite_b200 models run on CUDA tensors only (there is no CPU path); "
                                "use forward_host() for host buffers")
        lib = N.lib()
        h = self._native_handle(x.device)
        x = x.detach().to(torch.float32).contiguous()
        b = x.shape[0]
        out = torch.empty((b, pred_frames, self.img_c, self.img_h, self.img_w), dtype=torch.float32, device=x.device)
        aux = torch.zeros(1, dtype=torch.float32, device=x.device) if want_aux else None
        nbytes = C.c_size_t()
        N.check(lib.vpk_model_workspace_bytes(h, b, t_in, pred_frames, C.byref(nbytes)))
        key = (b, t_in, pred_frames, x.device.index)
        ws = self._workspaces.get(key)
        if ws is None or ws.numel() < nbytes.value:
            self._workspaces.clear()
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
            self._workspaces[key] = ws
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if actions is None:
                N.check(lib.vpk_model_forward(h, N.ptr(x), b, t_in, pred_frames, N.ptr(out), N.ptr(aux), N.ptr(ws),
                                              ws.numel(), C.c_void_p(stream)))
            else:
                N.check(lib.vpk_model_forward_actions(h, N.ptr(x), N.ptr(actions), actions.shape[1], b, t_in, pred_frames,
                                                      N.ptr(out), N.ptr(aux), N.ptr(ws), ws.numel(), C.c_void_p(stream)))
        return out, aux

    def forward_host(self, x, pred_frames=1, out=None, **kwargs):
        """Same as ``forward`` for HOST tensors (pinned memory recommended): microbatches are staged through the
        device with copies overlapped with compute; returns host tensors.  ``out`` may be a caller-owned (pinned)
        result buffer; otherwise a pinned buffer owned by the module is reused across calls of the same shape."""
        if x.is_cuda:
            raise ValueError("forward_host takes host tensors")
        lib = N.lib()
        h = self._native_handle()
        x = x.detach().to(torch.float32).contiguous()
        b, t_in = x.shape[:2]
        shape = (b, pred_frames, self.img_c, self.img_h, self.img_w)
        if out is None:
            out = getattr(self, "_host_out", None)
            if out is None or tuple(out.shape) != shape:
                out = torch.empty(shape, dtype=torch.float32, pin_memory=True)
                self._host_out = out
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.is_cuda:
            raise ValueError("out must be a contiguous host fp32 tensor of shape %s" % (shape,))
        aux = torch.zeros(1, dtype=torch.float32)
        actions = self._native_actions(kwargs, b, None)
        with torch.cuda.device(self._handle_device):
            if actions is None:
                N.check(lib.vpk_model_forward_host(h, N.ptr(x), b, self._native_t_in(t_in, pred_frames), pred_frames,
                                                   N.ptr(out), N.ptr(aux)))
            else:
                N.check(lib.vpk_model_forward_host_actions(h, N.ptr(x), N.ptr(actions), actions.shape[1], b,
                                                           self._native_t_in(t_in, pred_frames), pred_frames,
                                                           N.ptr(out), N.ptr(aux)))
        return out, aux

    def _native_t_in(self, t_total, pred_frames):
        return t_total

    def microbatch_size(self, batch):
        """Sequences per pass over the layers for a call with ``batch`` sequences."""
        n = C.c_int32()
        N.check(N.lib().vpk_model_microbatch(self._native_handle(), int(batch), C.byref(n)))
        return n.value

    def last_launch_count(self):
        n = C.c_int64()
        N.check(N.lib().vpk_model_last_launch_count(self._native_handle(), C.byref(n)))
        return n.value

    def set_timing(self, enable):
        """0 off, 1 CUDA events around the gate GEMMs, 2 around every kernel (see ``layer_profile``)."""
        N.check(N.lib().vpk_model_set_timing(self._native_handle(), int(enable)))

    def layer_profile(self):
        """[(layer, launches, ms, gflop)] of the last forward run with ``set_timing(2)``, sorted by time."""
        buf = C.create_string_buffer(1 << 16)
        N.check(N.lib().vpk_model_profile(self._native_handle(), buf, len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            name, n, ms, gf = line.rsplit(" ", 3)
            rows.append((name, int(n), float(ms), float(gf)))
        return rows

    def last_gemm_stats(self):
        ms, n, fl = C.c_float(), C.c_int64(), C.c_double()
        N.check(N.lib().vpk_model_last_gemm_ms(self._native_handle(), C.byref(ms), C.byref(n), C.byref(fl)))
        return {"ms": ms.value, "launches": n.value, "flops": fl.value}
