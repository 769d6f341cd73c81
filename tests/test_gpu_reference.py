"""The CUDA path against the REAL reference running on the same GPU in fp32 (cuDNN, TF32 off) at the BASELINE
configurations' full shapes and batches, every sequence distinct ("on the GPU box the oracle can also
run on CUDA").  The unmodified reference is imported from baseline/_ref through oracle/ref_shim.

fp32-operand mode must match to <= 1e-4; bf16 mode to <= 5e-3 on the first predicted frame and <= 2e-2 at the end of the
rollout (BASELINE.json north_star) -- over the WHOLE batch, not a three-sequence sample."""
import os

import pytest
import torch

from oracle import ref_shim
from oracle.weights import synth_state_dict, synth_frames

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available() or bool(os.environ.get("VPK_NO_REFERENCE")),
                                 reason="reference package not present (baseline/_ref)")]

# name: (model key, img_shape, context, pred, batch, gain, model kwargs)
CASES = {
    "cfg1": ("convlstm-shi", (1, 64, 64), 10, 10, 8, 2.5, {}),
    "cfg3": ("predrnn-pp", (1, 64, 64), 10, 10, 256, 1.5, {}),
    "cfg3ln": ("predrnn-pp", (1, 64, 64), 10, 10, 256, 1.5, {"layer_norm": True}),
    "cfg4": ("phy", (3, 64, 64), 2, 10, 256, 1.5, {}),
    # cfg 5's shape and rollout length; the reference materialises [b, t, C, H, W] per layer (84 MB per sequence and tensor
    # at 128 x 128), so the batch is what comfortably fits beside our workspace
    "cfg5": ("convlstm-shi", (3, 128, 128), 10, 20, 48, 2.5, {}),
}


def _reference_on_cuda(key, img, sd, kw):
    classes = ref_shim.load_reference()
    ref = classes[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0], **kw).to("cuda:0").eval()
    own = ref.state_dict()
    ref.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    # CUDA-built convlstm-shi does not register its peepholes (conv_lstm_hzzone.py:30-32): set them by hand, as the
    # reference's own test does (tests/test_impl_match/_convlstm_hzzone.py:55-62)
    for k, v in sd.items():
        if k not in own:
            mod = ref
            *path, leaf = k.split(".")
            for p in path:
                mod = getattr(mod, p)
            assert leaf in ("Wci", "Wcf", "Wco"), k
            setattr(mod, leaf, v.to("cuda:0"))
    return ref


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_batch_against_the_real_reference_on_cuda(name):
    import vp_suite_b200 as V
    key, img, ctx, pred, B, gain, kw = CASES[name]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    t_in = ctx + (pred if key == "predrnn-pp" else 0)
    ours32 = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0],
                                  precision="fp32", **kw).eval()
    sd = synth_state_dict({k: tuple(v.shape) for k, v in ours32.state_dict().items()}, seed=11, gain=gain)
    ours32.load_state_dict(sd)
    x = synth_frames(B, t_in, *img, seed=777).cuda()
    ref = _reference_on_cuda(key, img, sd, kw)
    # PredRNN's 5 x 5 convs: cuDNN's fp32 algorithm choice there (FFT / Winograd family) is itself 3e-4 away from the CPU
    # reference after 19 steps (2.6e-2 with LayerNorm), i.e. not an fp32 reference at the 1e-4 level; ATen's native
    # convolution (im2col + fp32 SGEMM, TF32 off) is.  The 3 x 3 models agree with cuDNN to 3e-7 and keep it.
    # (cudnn.flags() resets allow_tf32 to its default True unless told otherwise)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=(key != "predrnn-pp"), allow_tf32=False):
        want, want_aux = ref(x, pred_frames=pred)
    del ref
    torch.cuda.empty_cache()
    for precision, model in (("fp32", ours32), ("bf16", None)):
        if model is None:
            model = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0],
                                         precision=precision, **kw).eval()
            model.load_state_dict(sd)
        with torch.no_grad():
            got, aux = model(x, pred_frames=pred)
        d = (got - want).abs()
        errs = [float(d[:, t].max()) for t in range(pred)]
        print(f"{name} {precision} vs reference on CUDA, batch {B}: first {errs[0]:.2e} last {errs[-1]:.2e} max {max(errs):.2e}")
        if precision == "fp32":
            assert max(errs) <= 1e-4, errs
        else:
            assert errs[0] <= 5e-3 and max(errs) <= 2e-2, errs
        if want_aux is not None:
            (k_, v), = aux.items()
            (_, rv), = want_aux.items()
            tol = 1e-3 if precision == "fp32" else 5e-2
            assert abs(float(v) - float(rv)) <= tol * abs(float(rv)) + 1e-3, (float(v), float(rv))
        del model, got
        torch.cuda.empty_cache()
