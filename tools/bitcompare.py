"""Developer aid: run one workload with the library selected by VPK_LIB_PATH and save / compare the frames bit for bit.
    VPK_LIB_PATH=tools/trace/libvpk_old.so python tools/bitcompare.py save cfg1 8 /tmp/a.pt
    python tools/bitcompare.py check cfg1 8 /tmp/a.pt"""
import sys

import torch

sys.path.insert(0, ".")
from bench import WORKLOADS, WORKLOAD_KW          # noqa: E402
import vp_suite_b200 as V                        # noqa: E402
from oracle.weights import synth_state_dict, synth_frames   # noqa: E402


def main():
    mode, wl, batch, path = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    graph = len(sys.argv) > 5 and sys.argv[5] == "graph"
    key, img, ctx, pred = WORKLOADS[wl][:4]
    m = V.MODEL_CLASSES[key]("cuda:0", img_shape=img, action_size=0, tensor_value_range=[0.0, 1.0], precision="bf16",
                             use_cuda_graph=graph, **WORKLOAD_KW.get(wl, {})).eval()
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=5, gain=2.5))
    t = ctx + (pred if key == "predrnn-pp" else 0)
    x = synth_frames(batch, t, *img, seed=6).cuda()
    with torch.no_grad():
        out = m(x, pred_frames=pred)[0].cpu()
    if mode == "save":
        torch.save(out, path)
        print("saved", tuple(out.shape), float(out.abs().mean()))
    else:
        ref = torch.load(path)
        print("bit-identical:", torch.equal(out, ref), "max abs diff", float((out - ref).abs().max()))


if __name__ == "__main__":
    main()
