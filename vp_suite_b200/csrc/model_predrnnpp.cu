// predrnn-pp-causal: the PredRNN++ rollout the north star names -- a stack of Causal LSTM cells with a gradient highway
// unit between the first and the second layer (Wang et al., ICML 2018, sec. 3; figure 3), eval mode, behind the same
// VPModel contract as the reference's `predrnn-pp` (PredRNN_V2, models/predrnn_v2.py:131-230: patchified frames, t_in =
// context + predicted frames, the model's own x_gen fed back after the context, 1x1 head, un-patchify).
//
// PARITY UNPINNED: /root/reference holds no CausalLSTMCell / GHU, so the checker is oracle/causal.py, a
// restatement of the published equations (causal.h).  Three fused tcgen05 launches per cell step + one for the GHU.
//
// Per step t (total_frames - 1 steps): layer 0 reads the patch frame x_t (t < context) or x_gen; the spatial memory m
// zig-zags through the layers (top layer of step t -> layer 0 of step t + 1); z_t = GHU(h_t^1, z_{t-1}) is layer 1's input.
#include <algorithm>
#include <cstdlib>

#include "builders.h"
#include "causal.h"
#include "elementwise.h"
#include "model.h"

namespace vpk {

namespace {

class PredRnnPP : public Model {
 public:
  explicit PredRnnPP(const vpk_model_desc& d) : Model(d) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h > 0 && d.img_w > 0, "bad img_shape");
    p = d.patch_size;
    L = d.num_layers;
    k = d.filter_size;
    VPK_REQUIRE(p > 0 && d.img_h % p == 0 && d.img_w % p == 0, "image size must be a multiple of patch_size");
    VPK_REQUIRE(L >= 2 && L <= 8 && k % 2 == 1, "predrnn-pp-causal needs 2..8 layers (the GHU sits between layers 0 and 1) and an odd filter_size");
    VPK_REQUIRE(d.layer_norm == 0 && d.action_conditional == 0, "predrnn-pp-causal: layer_norm / action_conditional are not built");
    // widths may differ per layer (the paper's 128-64-64-64): the spatial memory a cell reads has the width of the cell that
    // wrote it -- the previous layer, or the top layer of the previous step for layer 0
    for (int i = 0; i < L; ++i) {
      VPK_REQUIRE(d.num_hidden[i] > 0, "bad num_hidden");
      Cs[i] = d.num_hidden[i];
      Cmax = std::max(Cmax, Cs[i]);
    }
    cp = p * p * d.img_c;
    hp_ = d.img_h / p;
    wp_ = d.img_w / p;
    for (int i = 0; i < L; ++i) {
      const std::string pre = "cell_list." + std::to_string(i) + ".";
      const int C = Cs[i], cin = (i == 0) ? cp : Cs[i - 1], cm = Cs[(i + L - 1) % L];
      declare(pre + "conv_x.0.weight", {7 * C, cin, k, k});
      declare(pre + "conv_h.0.weight", {4 * C, C, k, k});
      declare(pre + "conv_c.0.weight", {3 * C, C, k, k});
      declare(pre + "conv_m.0.weight", {3 * C, cm, k, k});
      declare(pre + "conv_c2m.0.weight", {4 * C, C, k, k});
      declare(pre + "conv_om.0.weight", {C, C, k, k});
      declare(pre + "conv_last.weight", {C, 2 * C, 1, 1});
    }
    declare("gradient_highway.x_concat.0.weight", {2 * Cs[0], Cs[0], k, k});
    declare("gradient_highway.z_concat.0.weight", {2 * Cs[0], Cs[0], k, k});
    declare("conv_last.weight", {cp, Cs[L - 1], 1, 1});
  }

 protected:
  void validate(int t_in, int pred) const override {
    if (t_in - pred < 1) VPK_THROW(1, "predrnn-pp-causal needs input sequences that also include the target frames");
  }
  int default_microbatch() const override { return 256; }
  int used_in_frames(int t_in, int pred) const override { return t_in - pred; }   // eval: context frames only
  bool streams_input() const override { return !desc.use_cuda_graph && getenv("VPK_NO_INPUT_STREAM") == nullptr; }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const ActInfo act{dtype, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int ctx = t_in - pred;
    const size_t px = static_cast<size_t>(B) * hp_ * wp_;

    char* xp = static_cast<char*>(arena.alloc(px * cp * esz * ctx));
    float* out_stage = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * pred * c * h * w * sizeof(float)));
    std::vector<void*> hb(2 * L), memb(2 * L);
    std::vector<float*> cb(L);
    for (int i = 0; i < L; ++i) {
      const int C = Cs[i];
      hb[2 * i] = arena.alloc(px * C * esz);
      hb[2 * i + 1] = arena.alloc(px * C * esz);
      memb[2 * i] = arena.alloc(px * 2 * C * esz);          // (c', m') of layer i, even / odd steps
      memb[2 * i + 1] = arena.alloc(px * 2 * C * esz);
      cb[i] = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    }
    const int C = Cmax, C0 = Cs[0];      // scratch shared by the layers is sized for the widest one
    float* mstate = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    float* opart = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    // launch O as two launches once the layer is tensor-bound (stlstm.h: o_raw; bit-identical); VPK_SPLIT_O=0/1 overrides
    bool split_o = dtype != DT_F32 && backend == 0 && px / 128 >= 2 * static_cast<size_t>(num_sms) &&
                   getenv("VPK_NO_REGIONS") != nullptr;   // superseded by accumulator regions in the fused launch (lowering.cu)
    if (const char* env = getenv("VPK_SPLIT_O")) split_o = atoi(env) != 0;
    float* oraw = split_o ? static_cast<float*>(arena.alloc(px * C * sizeof(float))) : nullptr;
    void* zb[2] = {arena.alloc(px * C0 * esz), arena.alloc(px * C0 * esz)};
    float* zstate = static_cast<float*>(arena.alloc(px * C0 * sizeof(float)));
    float* xgen32 = static_cast<float*>(arena.alloc(px * cp * sizeof(float)));
    void* xgen_act = (dtype == DT_F32) ? static_cast<void*>(xgen32) : arena.alloc(px * cp * esz);

    if (!measure) {
      const int ns = num_sms, dt = dtype, pp = p;
      if (!streams_input()) {     // CUDA-graph replay: the conversion reads the call's input pointer, so it stays outside
        Op pre;
        pre.name = "patchify";
        pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
          launch_patchify_strided(rc.x, static_cast<long long>(t_in) * c * h * w, xp, dt, B, ctx, c, h, w, pp, ns, s);
        };
        prog.pre.push_back(std::move(pre));
      }
      for (int i = 0; i < L; ++i) {
        add_memset(prog, hb[2 * i], px * Cs[i] * esz, "zero_h");
        add_memset(prog, cb[i], px * Cs[i] * sizeof(float), "zero_c");
        add_memset(prog, memb[2 * i + 1], px * 2 * Cs[i] * esz, "zero_mem");     // c_{-1} of every layer, m seen by layer 0
      }
      add_memset(prog, mstate, px * C * sizeof(float), "zero_m32");      // write-only state, read (and ignored) by the prefetch
      add_memset(prog, zb[0], px * C0 * esz, "zero_z");
      add_memset(prog, zstate, px * C0 * sizeof(float), "zero_z32");
    }

    std::vector<int> par(L, 0);
    for (int t = 0; t < t_in - 1; ++t) {
      const void* net = (t < ctx) ? static_cast<const void*>(xp + static_cast<size_t>(t) * px * cp * esz) : xgen_act;
      if (!measure && t < ctx && streams_input()) {
        const int ns = num_sms, dt = dtype, pp = p;
        char* dst = xp + static_cast<size_t>(t) * px * cp * esz;
        const long long bstride = static_cast<long long>(t_in) * c * h * w, foff = static_cast<long long>(t) * c * h * w;
        Op cv;
        cv.name = "patchify";
        cv.needs_input = t;
        cv.fn = [=](cudaStream_t s, const RunCtx& rc) {
          launch_patchify_strided(rc.x + foff, bstride, dst, dt, B, 1, c, h, w, pp, ns, s);
        };
        prog.body.push_back(std::move(cv));
      }
      for (int i = 0; i < L; ++i) {
        const std::string pre = "cell_list." + std::to_string(i) + ".";
        const void* inp = (i == 0) ? net : (i == 1) ? zb[(t + 1) & 1] : hb[2 * (i - 1) + par[i - 1]];
        const int C = Cs[i], cin = (i == 0) ? cp : Cs[i - 1], cm = Cs[(i + L - 1) % L];
        const void* mem_prev = (i == 0) ? memb[2 * (L - 1) + ((t + 1) & 1)] : memb[2 * (i - 1) + (t & 1)];
        CausalArgs a{pre, B, hp_, wp_, cin, C, k, inp, hb[2 * i + par[i]],
                     make_channel_view(memb[2 * i + ((t + 1) & 1)], hp_, wp_, 2 * C, 0, C, esz),
                     make_channel_view(mem_prev, hp_, wp_, 2 * cm, cm, cm, esz), hb[2 * i + (par[i] ^ 1)], cb[i], mstate, opart,
                     memb[2 * i + (t & 1)],
                     hp(pre + "conv_x.0.weight"), hp(pre + "conv_h.0.weight"), hp(pre + "conv_c.0.weight"),
                     hp(pre + "conv_m.0.weight"), hp(pre + "conv_c2m.0.weight"), hp(pre + "conv_om.0.weight"),
                     hp(pre + "conv_last.weight")};
        a.c4 = true;
        a.Cm = cm;
        a.o_raw = oraw;
        for (const ConvSpec& sp : causal_lstm_specs(a, act)) add_conv(prog, sp, measure, stream, dtype);
        par[i] ^= 1;
        if (i == 0) {   // z_t = GHU(h_t^1, z_{t-1}): read zb[t & 1], write zb[(t + 1) & 1]
          GhuArgs g{"gradient_highway.", B, hp_, wp_, C0, k, hb[par[0]], zb[t & 1], zb[(t + 1) & 1], zstate,
                    hp("gradient_highway.x_concat.0.weight"), hp("gradient_highway.z_concat.0.weight")};
          g.c4 = true;
          add_conv(prog, ghu_spec(g, act), measure, stream, dtype);
        }
      }
      // head: x_gen = conv_last(h_top)  (1x1, no bias), kept in fp32 so that output frames carry no extra rounding
      int oh, ow;
      ConvArgs hd{"conv_last.", B, hp_, wp_, Cs[L - 1], cp, 1, 1, 0, hb[2 * (L - 1) + par[L - 1]], hp("conv_last.weight"),
                  nullptr, ACT_NONE, xgen32};
      hd.f32_strided = true;
      hd.oB = static_cast<long long>(hp_) * wp_ * cp;
      hd.oY = static_cast<long long>(wp_) * cp;
      hd.oX = cp;
      hd.oC = 1;
      add_conv(prog, conv_spec(hd, act, &oh, &ow), measure, stream, dtype);
      if (!measure) {
        const int ns = num_sms;
        if (dtype != DT_F32 && t + 1 >= ctx && t + 1 < t_in - 1) {
          const long long n = static_cast<long long>(px) * cp;
          Op op;
          op.name = "cast_xgen";
          op.fn = [=](cudaStream_t s, const RunCtx&) { launch_cast_f32_to_bf16(xgen32, xgen_act, n, ns, s); };
          prog.body.push_back(std::move(op));
        }
        const int first_out = t_in - 1 - pred;
        if (t >= first_out) {
          const int fo = t - first_out, pp = p;
          Op op;
          op.name = "unpatchify";
          op.fn = [=](cudaStream_t s, const RunCtx&) {
            launch_unpatchify(xgen32, out_stage, DT_F32, B, pred, fo, c, h, w, pp, ns, s);
          };
          op.frame = fo;
          op.frame_src = out_stage + static_cast<size_t>(fo) * c * h * w;
          op.frame_pitch = static_cast<long long>(pred) * c * h * w;
          op.frame_elems = static_cast<long long>(c) * h * w;
          prog.body.push_back(std::move(op));
        }
      }
    }
    if (!measure) {
      const size_t bytes = static_cast<size_t>(B) * pred * c * h * w * sizeof(float);
      Op post;
      post.name = "copy_out";
      post.is_kernel = false;
      post.fn = [=](cudaStream_t s, const RunCtx& rc) {
        if (rc.on_frame != nullptr) return;
        VPK_CUDA(cudaMemcpyAsync(rc.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s));
      };
      prog.post.push_back(std::move(post));
    }
  }

 private:
  int p = 4, L = 4, k = 5, cp = 16, hp_ = 16, wp_ = 16;
  int Cs[8] = {0, 0, 0, 0, 0, 0, 0, 0}, Cmax = 0;
};

}  // namespace

Model* make_predrnnpp_causal(const vpk_model_desc& d) { return new PredRnnPP(d); }

}  // namespace vpk
