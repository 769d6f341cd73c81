"""SSIM partial sums.  Value parity against piqa is UNPINNED (piqa is absent from the reference
checkout and from this image): the checks are the reference's own tests for this measure -- optimal value for equal
tensors and symmetry, on its tensor shape (tests/test_measure.py:17-50: [4, 10, 3, 63, 76], |.| < 1e-4) -- plus agreement
of the three statements of the restated algorithm (numpy oracle, torch definition, CUDA kernels)."""
import numpy as np
import pytest
import torch

from oracle import measure as OM
from vp_suite_b200 import evaluation as E


def _tensors(shape=(4, 10, 3, 63, 76), seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g), torch.rand(*shape, generator=g)


def test_oracle_axioms_on_the_reference_test_shape():
    x, y = _tensors()
    assert abs(OM.ssim_measure(x.numpy(), x.numpy())) < 1e-4            # to_display(0) = 1: the optimal value
    assert abs(OM.ssim_measure(x.numpy(), y.numpy()) - OM.ssim_measure(y.numpy(), x.numpy())) < 1e-4
    v = OM.ssim_images(x.numpy(), y.numpy())
    assert v.shape == (4, 10) and np.all(v < 1.0) and np.all(v > -1.0)


def test_torch_definition_matches_oracle_and_refuses_other_channel_counts():
    x, y = _tensors((3, 4, 3, 40, 33), seed=1)
    x = x * 2.4 - 1.2                                                    # exercises reshape_clamp's clamp
    y = x + 0.3 * (y - 0.5)
    ref = OM.ssim_images(x.numpy(), y.numpy())
    got = E.ssim_partial_sums(x, y)
    assert got.dtype == torch.float64 and got.shape == (4,)
    assert np.abs(got.numpy() - ref.sum(0)).max() < 1e-5
    disp = E.finalize_ssim(got, 3)
    assert abs(disp[-1] - ref.mean()) < 1e-5 and abs(disp[0] - ref[:, 0].mean()) < 1e-5
    with pytest.raises(ValueError):
        E.ssim_partial_sums(x[:, :, :1], y[:, :, :1])
    with pytest.raises(ValueError):
        OM.ssim_measure(x[:, :, :1].numpy(), y[:, :, :1].numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 10, 3, 63, 76), (5, 3, 3, 64, 64), (2, 2, 3, 128, 128), (1, 1, 3, 11, 11),
                                   (300, 1, 3, 27, 12)])
def test_cuda_ssim_against_oracle_and_axioms(shape):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    x, y = _tensors(shape, seed=2)
    x = x * 2.2 - 1.1
    y = (x + 0.4 * (y - 0.5)).contiguous()
    ref = OM.ssim_images(x.numpy(), y.numpy())
    xc, yc = x.cuda(), y.cuda()
    got = E.ssim_partial_sums(xc, yc).cpu().numpy()
    assert np.abs(got - ref.sum(0)).max() <= 2e-5 * shape[0]            # fp32 filters against the fp64 checker
    same = E.ssim_partial_sums(xc, xc).cpu().numpy()
    assert np.abs(same - shape[0]).max() < 1e-4 * shape[0]               # f(x, x) = optimal value
    swapped = E.ssim_partial_sums(yc, xc).cpu().numpy()
    assert np.array_equal(swapped, got)                                  # symmetric, bit for bit
    again = E.ssim_partial_sums(xc, yc).cpu().numpy()
    assert np.array_equal(again, got)                                    # deterministic


@pytest.mark.gpu
def test_cuda_ssim_rejects_images_smaller_than_the_window():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    x = torch.zeros(1, 1, 3, 10, 64, device="cuda")
    with pytest.raises(ValueError):
        E.ssim_partial_sums(x, x)
