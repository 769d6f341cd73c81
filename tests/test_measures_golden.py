"""MSE / PSNR per prediction horizon against vectors produced by the REFERENCE's own measure classes through its
PredictionMetricProvider (oracle/make_golden.py: run_measures -> tests/golden/measures.npz): the oracle's numpy
restatement, the host-side definition / provider, and (gpu) the on-device reduction kernels behind the C ABI."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import measure as OMs
from oracle.weights import measure_inputs
from vp_suite_b200 import evaluation as E

CASES = ["m3", "m1"]
TOL = 2e-5          # relative; the reference accumulates in fp32


def _case(manifest, name):
    meta = manifest["measures"][name]
    pred, target = measure_inputs(tuple(meta["shape"]), meta["seed"])
    return meta, pred, target, load_golden("measures")


@pytest.mark.parametrize("name", CASES)
def test_oracle_restatement_matches_reference_measures(manifest, name):
    meta, pred, target, gold = _case(manifest, name)
    np.testing.assert_allclose(OMs.mse_per_horizon(pred.numpy(), target.numpy()), gold[f"{name}_mse"], rtol=TOL)
    np.testing.assert_allclose(OMs.psnr_per_horizon(pred.numpy(), target.numpy()), gold[f"{name}_psnr"], rtol=TOL)
    # the all-frames forward values (what a loss provider sees): last horizon, PSNR before to_display's negation
    assert abs(gold[f"{name}_mse_fwd"] - gold[f"{name}_mse"][-1]) <= TOL * gold[f"{name}_mse"][-1]
    assert abs(gold[f"{name}_psnr_fwd"] + gold[f"{name}_psnr"][-1]) <= TOL * abs(gold[f"{name}_psnr"][-1])


def _check_provider(rows, gold, name, keys):
    assert [sorted(r) for r in rows] == [keys] * len(rows)                # same dict keys as the reference's provider
    np.testing.assert_allclose([r["mse (↓)"] for r in rows], gold[f"{name}_mse"], rtol=TOL)
    np.testing.assert_allclose([r["psnr (↑)"] for r in rows], gold[f"{name}_psnr"], rtol=TOL)


@pytest.mark.parametrize("name", CASES)
def test_host_definition_and_provider_format(manifest, name):
    meta, pred, target, gold = _case(manifest, name)
    prov = E.NativeMetricProvider({"device": "cpu", "metrics": ["mse", "psnr"], "img_c": meta["shape"][2]})
    _check_provider(prov.get_metrics(pred, target, all_frame_cnts=True), gold, name, meta["keys"])
    last = prov.get_metrics(pred, target)                                 # all_frame_cnts=False: the full horizon only
    assert len(last) == 1 and abs(last[0]["mse (↓)"] - gold[f"{name}_mse"][-1]) <= TOL * gold[f"{name}_mse"][-1]
    two = prov.get_metrics(pred, target, frames=2)
    assert abs(two[0]["psnr (↑)"] - gold[f"{name}_psnr"][1]) <= TOL * gold[f"{name}_psnr"][1]
    with pytest.raises(ValueError):
        prov.get_metrics(pred[0], target[0])
    with pytest.raises(ValueError):
        prov.get_metrics(pred, target[:, :2])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_kernels_match_reference_measures(manifest, name):
    meta, pred, target, gold = _case(manifest, name)
    prov = E.NativeMetricProvider({"device": "cuda:0", "metrics": ["mse", "psnr"], "img_c": meta["shape"][2]})
    _check_provider(prov.get_metrics(pred.cuda(), target.cuda(), all_frame_cnts=True), gold, name, meta["keys"])
