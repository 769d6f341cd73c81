"""libvpk.so loads without a GPU, exports every symbol include/vpk.h declares, and refuses to compute without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

from vp_suite_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "vpk.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpk_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = N.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/vpk.h but not exported by libvpk.so"
        assert name in N.SYMBOLS, f"{name} has no ctypes binding in vp_suite_b200/_native.py"
    assert set(N.SYMBOLS) == set(names)
    assert lib.vpk_version().startswith(b"libvpk")


def test_desc_struct_matches_header_field_order():
    with open(os.path.join(ROOT, "include", "vpk.h")) as f:
        text = f.read()
    body = text[text.index("typedef struct vpk_model_desc {"):text.index("} vpk_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body):
        for part in decl.split(","):
            fields.append(re.match(r"\s*([A-Za-z_0-9]+)", part).group(1))
    assert fields == [f[0] for f in N.ModelDesc._fields_]


def test_host_side_calls_work_without_gpu_and_compute_fails_loudly():
    lib = N.lib()
    d = N.ModelDesc()
    d.kind = N.VPK_MODEL_PREDRNN_PP
    d.img_c, d.img_h, d.img_w = 1, 64, 64
    d.patch_size, d.num_layers, d.filter_size = 4, 3, 5
    for i in range(4):
        d.num_hidden[i] = 128
    d.decoupling_loss_scale = 100.0
    h = C.c_void_p()
    N.check(lib.vpk_model_create(C.byref(d), C.byref(h)))
    n = C.c_int32()
    N.check(lib.vpk_model_num_params(h, C.byref(n)))
    assert n.value == 17
    nbytes = C.c_size_t()
    N.check(lib.vpk_model_workspace_bytes(h, 4, 20, 10, C.byref(nbytes)))
    assert nbytes.value > 0
    # wrong key / wrong shape are rejected like load_state_dict would
    bad = torch.zeros(3)
    shape = (C.c_int64 * 1)(3)
    with pytest.raises(ValueError):
        N.check(lib.vpk_model_set_param(h, b"no.such.key", N.ptr(bad), shape, 1))
    with pytest.raises(ValueError):
        N.check(lib.vpk_model_set_param(h, b"adapter.weight", N.ptr(bad), shape, 1))
    # too few frames: "needs input sequences that also include the target frames" (predrnn_v2.py:134-137)
    with pytest.raises(ValueError):
        N.check(lib.vpk_model_workspace_bytes(h, 4, 10, 10, C.byref(nbytes)))
    if not torch.cuda.is_available():
        assert lib.vpk_device_ok() == 0
        with pytest.raises(N.NativeError, match="no CUDA device"):
            N.check(lib.vpk_model_finalize(h, None))
    lib.vpk_model_destroy(h)


def test_unknown_model_kind_is_rejected():
    d = N.ModelDesc()
    d.kind = 99
    h = C.c_void_p()
    with pytest.raises(ValueError):
        N.check(N.lib().vpk_model_create(C.byref(d), C.byref(h)))
