"""
Host-side mirror of the reference's model contract.

``VPModel`` follows vp_suite/base/base_model.py:11-146 (constructor ``(device, **model_kwargs)``, ``REQUIRED_ARGS``,
kwargs -> type-checked attributes, ``config`` property, ``forward`` / ``pred_1``); ``VPModelBlock`` follows
vp_suite/base/base_model_block.py:4-13.  ``NativeRollout`` is the glue to libvpk: it mirrors the module's
``state_dict`` into the native handle and runs ``forward`` through the C ABI.  Nothing here computes frames in
Python -- without the CUDA library the calls raise.
"""
import ctypes as C
import inspect

import torch
import torch.nn as nn

from . import _native as N


class VPModelBlock(nn.Module):
    """Marker base of the model blocks (vp_suite/base/base_model_block.py:4-13)."""
    NAME: str = __name__
    PAPER_REFERENCE = None
    CODE_REFERENCE = None
    MATCHES_REFERENCE: str = None


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


class VPModel(nn.Module):
    # vp_suite/base/base_model.py:18-36
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

    def pred_1(self, x, **kwargs):
        raise NotImplementedError

    def forward(self, x, pred_frames=1, **kwargs):
        raise NotImplementedError


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
        self._versions = None
        self._workspaces = {}

    # -- handle lifecycle ---------------------------------------------------------------------------------------
    def _native_handle(self):
        lib = N.lib()
        self._create_handle()
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
    def _native_forward(self, x, pred_frames, t_in, want_aux=False):
        if not x.is_cuda:
            raise N.NativeError("vp_suite_b200 models run on CUDA tensors only (there is no CPU path); "
                                "use forward_host() for host buffers")
        lib = N.lib()
        h = self._native_handle()
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
        N.check(lib.vpk_model_forward(h, N.ptr(x), b, t_in, pred_frames, N.ptr(out), N.ptr(aux), N.ptr(ws),
                                      ws.numel(), C.c_void_p(stream)))
        return out, aux

    def forward_host(self, x, pred_frames=1, out=None):
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
        N.check(lib.vpk_model_forward_host(h, N.ptr(x), b, self._native_t_in(t_in, pred_frames), pred_frames,
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
