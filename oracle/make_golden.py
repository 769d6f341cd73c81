"""
Golden-vector generator (test infrastructure; runs only in the authoring container).

Imports the REAL vp-suite reference from /root/reference through ``oracle.ref_shim``, loads the deterministic
synthetic weights of ``oracle.weights`` into the reference's own modules, runs them on deterministic synthetic
frames and stores the outputs as small ``.npz`` fixtures under ``tests/golden/`` plus ``manifest.json``
(configs, seeds and the reference's state-dict key->shape listing).  Weights and inputs are NOT stored: tests
regenerate them from the seeds.

    python -m oracle.make_golden            # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle.weights import synth_state_dict, synth_frames, measure_inputs, synth_actions   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = [
    # name,           key,            img_shape,   b, t_ctx, pred, wseed, xseed, gain
    ("ef_1x64",       "convlstm-shi", (1, 64, 64), 2, 4,     3,    0,     1234,  2.5),
    ("ef_3x32",       "convlstm-shi", (3, 32, 32), 1, 3,     4,    1,     77,    2.5),
    ("predrnn_1x64",  "predrnn-pp",   (1, 64, 64), 2, 3,     3,    2,     99,    1.5),
    ("predrnn_3x32",  "predrnn-pp",   (3, 32, 32), 1, 2,     2,    3,     5,     1.5),
    ("phy_3x64",      "phy",          (3, 64, 64), 2, 2,     3,    4,     42,    1.5),
    ("phy_1x64",      "phy",          (1, 64, 64), 1, 3,     2,    5,     43,    1.5),
    ("predrnn_ln_1x64", "predrnn-pp", (1, 64, 64), 2, 3,     3,    7,     101,   1.5),
    ("predrnn_ln_3x32", "predrnn-pp", (3, 32, 32), 1, 2,     2,    8,     6,     1.5),
]
# extra constructor kwargs of a case (the manifest records them as `model_kwargs`)
MODEL_KW = {"predrnn_ln_1x64": {"layer_norm": True}, "predrnn_ln_3x32": {"layer_norm": True}}


def shapes_of(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def run_models(classes, manifest):
    for name, key, img, b, t, p, wseed, xseed, gain in MODEL_CASES:
        torch.manual_seed(0)
        kw = MODEL_KW.get(name, {})
        m = classes[key]("cpu", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0], **kw).eval()
        shp = shapes_of(m)
        m.load_state_dict(synth_state_dict(shp, wseed, gain))
        total_t = t + p if key == "predrnn-pp" else t          # NEEDS_COMPLETE_INPUT (predrnn_v2.py:32)
        x = synth_frames(b, total_t, *img, seed=xseed)
        with torch.no_grad():
            pred, aux = m(x, pred_frames=p)
        arrays = {"pred": pred.numpy()}
        if aux is not None:
            (lk, lv), = aux.items()
            arrays["loss"] = np.asarray(float(lv), dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        manifest["models"][name] = dict(key=key, img_shape=list(img), batch=b, context=t, pred=p,
                                        wseed=wseed, xseed=xseed, gain=gain, shapes=shp, model_kwargs=kw,
                                        pred_std=float(pred.std()), pred_mean=float(pred.mean()))
        print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


def run_branch(classes, manifest):
    """BASELINE config 2 composition (ours), built from the reference PhyDNet's own sub-modules."""
    name, img, b, t, p, wseed, xseed = "branch_1x64", (1, 64, 64), 2, 3, 3, 6, 44
    m = classes["phy"]("cpu", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0]).eval()
    shp = shapes_of(m)
    m.load_state_dict(synth_state_dict(shp, wseed))
    x = synth_frames(b, t, *img, seed=xseed)

    def step(frame, first):
        er = m.encoder_Er(m.encoder_E(frame))
        _, out = m.convcell(er, None, first)
        return torch.sigmoid(m.decoder_D(m.decoder_Dr(out[-1])))

    outs = []
    with torch.no_grad():
        for ei in range(t - 1):
            step(x[:, ei], ei == 0)
        frame = x[:, t - 1]
        for di in range(p):
            frame = step(frame, t == 1 and di == 0)
            outs.append(frame)
    pred = torch.stack(outs, 1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred.numpy())
    manifest["models"][name] = dict(key="convlstm-branch", img_shape=list(img), batch=b, context=t, pred=p,
                                    wseed=wseed, xseed=xseed, gain=1.5, shapes=shp,
                                    pred_std=float(pred.std()), pred_mean=float(pred.mean()))
    print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


def run_blocks(manifest):
    """Single-block vectors from the reference's own block classes."""
    from vp_suite.model_blocks import ConvLSTM, SpatioTemporalLSTMCell, PhyCell_Cell
    from vp_suite.model_blocks.conv_lstm_ndrplz import ConvLSTMCell
    arrays = {}

    # hzzone ConvLSTM: sequence with given inputs / zero state, then inputs=None continuing from that state
    blk = ConvLSTM("cpu", in_channels=8, enc_channels=16, state_h=12, state_w=10, kernel_size=3).eval()
    shp = shapes_of(blk)
    blk.load_state_dict(synth_state_dict(shp, 21))
    xin = torch.rand((2, 3, 8, 12, 10), generator=torch.Generator().manual_seed(5)) * 2 - 1
    with torch.no_grad():
        o1, (h1, c1) = blk(xin, None, seq_len=3)
        o2, (h2, c2) = blk(None, (h1, c1), seq_len=2)
    arrays.update(hz_out1=o1.numpy(), hz_h1=h1.numpy(), hz_c1=c1.numpy(), hz_out2=o2.numpy(), hz_c2=c2.numpy())
    manifest["blocks"]["hzzone"] = dict(shapes=shp, wseed=21, xseed=5, x_shape=[2, 3, 8, 12, 10])

    cell = ConvLSTMCell(input_dim=8, hidden_dim=16, kernel_size=(3, 3), bias=True).eval()
    shp = shapes_of(cell)
    cell.load_state_dict(synth_state_dict(shp, 22))
    g = torch.Generator().manual_seed(6)
    x, h, c = (torch.rand((2, 8, 10, 10), generator=g) * 2 - 1, torch.rand((2, 16, 10, 10), generator=g) * 2 - 1,
               torch.rand((2, 16, 10, 10), generator=g) * 2 - 1)
    with torch.no_grad():
        hn, cn = cell(x, (h, c))
    arrays.update(nd_h=hn.numpy(), nd_c=cn.numpy())
    manifest["blocks"]["ndrplz"] = dict(shapes=shp, wseed=22, xseed=6)

    st = SpatioTemporalLSTMCell(16, 32, 8, 8, 5, 1, False).eval()
    shp = shapes_of(st)
    st.load_state_dict(synth_state_dict(shp, 23))
    g = torch.Generator().manual_seed(7)
    x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
    h, c, mm = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
    with torch.no_grad():
        res = st(x, h, c, mm)
    for nm, v in zip(("st_h", "st_c", "st_m", "st_dc", "st_dm"), res):
        arrays[nm] = v.numpy()
    manifest["blocks"]["stlstm"] = dict(shapes=shp, wseed=23, xseed=7)

    stl = SpatioTemporalLSTMCell(16, 32, 8, 8, 5, 1, True).eval()        # layer_norm=True (predrnn.py:24-40)
    shp = shapes_of(stl)
    stl.load_state_dict(synth_state_dict(shp, 25))
    g = torch.Generator().manual_seed(9)
    x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
    h, c, mm = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(3)]
    with torch.no_grad():
        res = stl(x, h, c, mm)
    for nm, v in zip(("stln_h", "stln_c", "stln_m", "stln_dc", "stln_dm"), res):
        arrays[nm] = v.numpy()
    manifest["blocks"]["stlstm_ln"] = dict(shapes=shp, wseed=25, xseed=9)

    pc = PhyCell_Cell(input_dim=16, action_conditional=False, action_size=0, hidden_dim=49, kernel_size=(7, 7)).eval()
    shp = shapes_of(pc)
    pc.load_state_dict(synth_state_dict(shp, 24))
    g = torch.Generator().manual_seed(8)
    x, h = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1, torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
    with torch.no_grad():
        hn = pc(x, None, h)
    arrays.update(phy_h=hn.numpy())
    manifest["blocks"]["phycell"] = dict(shapes=shp, wseed=24, xseed=8)

    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **arrays)
    print("blocks:", sorted(arrays))


def run_blocks_ac(manifest):
    """Action-conditional ST-LSTM cell (predrnn.py:86-169), both layer_norm settings -> blocks_ac.npz.  Kept in its own
    file so that adding it does not rewrite the other fixtures: ``python -m oracle.make_golden blocks_ac``."""
    from vp_suite.model_blocks.predrnn import ActionConditionalSpatioTemporalLSTMCell as ACCell
    arrays = {}
    for tag, ln, wseed, xseed in (("stac", False, 26, 10), ("stacln", True, 27, 11)):
        cell = ACCell(16, 32, 8, 8, 5, 1, ln).eval()
        shp = shapes_of(cell)
        cell.load_state_dict(synth_state_dict(shp, wseed))
        g = torch.Generator().manual_seed(xseed)
        x = torch.rand((2, 16, 8, 8), generator=g) * 2 - 1
        h, c, mm, a = [torch.rand((2, 32, 8, 8), generator=g) * 2 - 1 for _ in range(4)]
        with torch.no_grad():
            res = cell(x, h, c, mm, a)
        for nm, v in zip(("h", "c", "m", "dc", "dm"), res):
            arrays[f"{tag}_{nm}"] = v.numpy()
        manifest["blocks"][f"stlstm_{tag[2:]}"] = dict(shapes=shp, wseed=wseed, xseed=xseed)
    np.savez_compressed(os.path.join(OUT, "blocks_ac.npz"), **arrays)
    print("blocks_ac:", sorted(arrays))


def run_measures(manifest):
    """MSE / PSNR per prediction horizon from the reference's OWN measure classes, through its PredictionMetricProvider
    (measure/metric_provider.py:34-73, measure/image_wise.py:19-31,53-75) -> measures.npz.  SSIM is not included: the
    reference delegates it to piqa, which is absent (value parity unpinned)."""
    from vp_suite.measure.metric_provider import PredictionMetricProvider
    from vp_suite.measure.image_wise import MSE, PSNR
    arrays = {}
    manifest["measures"] = {}
    for name, shape, seed in (("m3", (3, 5, 3, 24, 20), 31), ("m1", (2, 4, 1, 16, 16), 32)):
        pred, target = measure_inputs(shape, seed)
        # (a 1-channel config must list "fvd": metric_provider.py:29-31 pops it unconditionally when img_c is not 2 or 3)
        metrics = ["mse", "psnr"] + (["fvd"] if shape[2] not in (2, 3) else [])
        prov = PredictionMetricProvider({"device": "cpu", "metrics": metrics, "img_c": shape[2]})
        rows = prov.get_metrics(pred, target, all_frame_cnts=True)
        keys = sorted(rows[0])
        assert keys == ["mse (↓)", "psnr (↑)"], keys
        arrays[f"{name}_mse"] = np.asarray([r["mse (↓)"] for r in rows], dtype=np.float64)
        arrays[f"{name}_psnr"] = np.asarray([r["psnr (↑)"] for r in rows], dtype=np.float64)
        # the raw (lower-is-better) forward values over all frames, as loss providers see them
        arrays[f"{name}_mse_fwd"] = np.asarray(float(MSE("cpu")(pred, target)), dtype=np.float64)
        arrays[f"{name}_psnr_fwd"] = np.asarray(float(PSNR("cpu")(pred, target)), dtype=np.float64)
        manifest["measures"][name] = dict(shape=list(shape), seed=seed, keys=keys)
    np.savez_compressed(os.path.join(OUT, "measures.npz"), **arrays)
    print("measures:", {k: np.round(v, 4).tolist() for k, v in arrays.items()})


AC_CASES = [
    # name,              key,          img_shape,   b, ctx, pred, wseed, xseed, gain, action_size, extra kwargs
    ("predrnn_ac_1x64",  "predrnn-pp", (1, 64, 64), 2, 3,   3,    31,    201,   1.5,  3,           {}),
    ("predrnn_acln_3x32", "predrnn-pp", (3, 32, 32), 2, 2,   2,    32,    202,   1.5,  4,           {"layer_norm": True}),
    ("phy_ac_3x64",      "phy",        (3, 64, 64), 2, 2,   3,    33,    203,   1.5,  3,           {}),
]


def run_models_ac(manifest):
    """Action-conditional rollouts of the reference (predrnn_v2.py:73-90,181-221; models/phydnet.py:94-122 with
    model_blocks/phydnet.py:44-55,153-155): ``model(x, pred_frames, actions=a)`` -> <name>.npz."""
    classes = ref_shim.load_reference()
    for name, key, img, b, t, p, wseed, xseed, gain, a_size, kw in AC_CASES:
        torch.manual_seed(0)
        m = classes[key]("cpu", img_shape=img, action_size=a_size, tensor_value_range=[0.0, 1.0], action_conditional=True,
                         **kw).eval()
        shp = shapes_of(m)
        m.load_state_dict(synth_state_dict(shp, wseed, gain))
        total_t = t + p if key == "predrnn-pp" else t
        x = synth_frames(b, total_t, *img, seed=xseed)
        actions = synth_actions(b, t + p - 1, a_size, seed=xseed + 1)
        with torch.no_grad():
            pred, aux = m(x, pred_frames=p, actions=actions)
        arrays = {"pred": pred.numpy()}
        if aux is not None:
            (lk, lv), = aux.items()
            arrays["loss"] = np.asarray(float(lv), dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        manifest["models"][name] = dict(key=key, img_shape=list(img), batch=b, context=t, pred=p, wseed=wseed, xseed=xseed,
                                        gain=gain, shapes=shp, action_size=a_size, aseed=xseed + 1,
                                        model_kwargs={**kw, "action_conditional": True, "action_size": a_size},
                                        pred_std=float(pred.std()), pred_mean=float(pred.mean()))
        print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


def run_stphy(manifest):
    """ST-Phy (models/st_phy.py), non action-conditional, eval -> stphy_3x64.npz."""
    classes = ref_shim.load_reference()
    name, img, b, t, p, wseed, xseed, gain = "stphy_3x64", (3, 64, 64), 2, 3, 3, 35, 205, 1.5
    torch.manual_seed(0)
    m = classes["st-phy"]("cpu", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0]).eval()
    shp = shapes_of(m)
    m.load_state_dict(synth_state_dict(shp, wseed, gain))
    x = synth_frames(b, t, *img, seed=xseed)
    with torch.no_grad():
        pred, aux = m(x, pred_frames=p)
    assert aux is None
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred.numpy())
    manifest["models"][name] = dict(key="st-phy", img_shape=list(img), batch=b, context=t, pred=p, wseed=wseed, xseed=xseed,
                                    gain=gain, shapes=shp, pred_std=float(pred.std()), pred_mean=float(pred.mean()))
    print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


def run_stphy_ac(manifest):
    """ST-Phy with action_conditional=True (st_phy.py:48-56, 142-150; ActionConditionalSpatioTemporalLSTMCell with
    layer_norm=True and PhyCell's action convs per layer), eval -> stphy_ac_3x64.npz."""
    classes = ref_shim.load_reference()
    name, img, b, t, p, wseed, xseed, gain, a_size = "stphy_ac_3x64", (3, 64, 64), 2, 3, 3, 38, 208, 1.5, 3
    torch.manual_seed(0)
    m = classes["st-phy"]("cpu", img_shape=img, action_size=a_size, action_conditional=True, tensor_value_range=[0.0, 1.0]).eval()
    shp = shapes_of(m)
    m.load_state_dict(synth_state_dict(shp, wseed, gain))
    x = synth_frames(b, t, *img, seed=xseed)
    actions = synth_actions(b, t + p - 1, a_size, seed=xseed + 1)
    with torch.no_grad():
        pred, aux = m(x, pred_frames=p, actions=actions)
    assert aux is None
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred.numpy())
    manifest["models"][name] = dict(key="st-phy", img_shape=list(img), batch=b, context=t, pred=p, wseed=wseed, xseed=xseed,
                                    gain=gain, shapes=shp, action_size=a_size, aseed=xseed + 1,
                                    model_kwargs={"action_conditional": True, "action_size": a_size},
                                    pred_std=float(pred.std()), pred_mean=float(pred.mean()))
    print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


def run_trajgru(manifest):
    """EF-TrajGRU (models/precipitation_nowcasting/ef_traj_gru.py), eval -> trajgru_1x64.npz, trajgru_3x32.npz."""
    classes = ref_shim.load_reference()
    for name, img, b, t, p, wseed, xseed, gain in (("trajgru_1x64", (1, 64, 64), 2, 3, 3, 36, 206, 2.0),
                                                   ("trajgru_3x32", (3, 32, 32), 1, 2, 3, 37, 207, 2.0)):
        torch.manual_seed(0)
        m = classes["trajgru"]("cpu", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0]).eval()
        shp = shapes_of(m)
        m.load_state_dict(synth_state_dict(shp, wseed, gain))
        x = synth_frames(b, t, *img, seed=xseed)
        with torch.no_grad():
            pred, aux = m(x, pred_frames=p)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), pred=pred.numpy())
        manifest["models"][name] = dict(key="trajgru", img_shape=list(img), batch=b, context=t, pred=p, wseed=wseed,
                                        xseed=xseed, gain=gain, shapes=shp, pred_std=float(pred.std()), pred_mean=float(pred.mean()))
        print(f"{name}: pred {tuple(pred.shape)} mean {pred.mean():.4f} std {pred.std():.4f}")


INCREMENTAL = {"trajgru": run_trajgru, "blocks_ac": run_blocks_ac, "measures": run_measures, "models_ac": run_models_ac, "stphy": run_stphy,
               "stphy_ac": run_stphy_ac}


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] and all(a in INCREMENTAL for a in sys.argv[1:]):
        # add fixtures to an existing manifest without rewriting the others: python -m oracle.make_golden measures ...
        ref_shim.load_reference()
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        with open(os.path.join(OUT, "manifest.json")) as f:
            manifest = json.load(f)
        for a in sys.argv[1:]:
            INCREMENTAL[a](manifest)
        with open(os.path.join(OUT, "manifest.json"), "w") as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        return
    classes = ref_shim.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    manifest = {"generator": "oracle/make_golden.py", "reference": "AIS-Bonn/vp-suite v0.0.9 (/root/reference)",
                "torch": torch.__version__, "models": {}, "blocks": {}}
    run_models(classes, manifest)
    run_branch(classes, manifest)
    run_blocks(manifest)
    run_blocks_ac(manifest)
    run_measures(manifest)
    run_models_ac(manifest)
    run_stphy(manifest)
    run_stphy_ac(manifest)
    run_trajgru(manifest)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
