// predrnn-pp: PredRNN-V2 rollout, non action-conditional, layer_norm=False, eval mode
// (reference: models/predrnn_v2.py:131-230; cell: model_blocks/predrnn.py:57-83).
//
// Per step t (total_frames - 1 steps): layer 0 reads the patchified frame x_t (t < context) or the model's own
// x_gen (eval mask is all zero, predrnn_v2.py:172-176, 300-309); the spatio-temporal memory m zig-zags through the
// layers (:196-204); every (t, layer) contributes a decoupling-loss term over adapter(delta_c), adapter(delta_m)
// (:197-211); the 1x1 head gives x_gen (:223); the last `pred` x_gen are un-patchified (:227-228).
#include "builders.h"
#include "elementwise.h"
#include "model.h"
#include "stlstm.h"

namespace vpk {

namespace {

class PredRnnV2 : public Model {
 public:
  explicit PredRnnV2(const vpk_model_desc& d) : Model(d) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h > 0 && d.img_w > 0, "bad img_shape");
    p = d.patch_size;
    L = d.num_layers;
    k = d.filter_size;
    VPK_REQUIRE(p > 0 && d.img_h % p == 0 && d.img_w % p == 0, "image size must be a multiple of patch_size");
    VPK_REQUIRE(L >= 1 && L <= 8 && k % 2 == 1, "bad num_layers / filter_size");
    C = d.num_hidden[0];
    for (int i = 0; i < L; ++i)   // the shared memory tensor and the single adapter imply equal widths
      VPK_REQUIRE(d.num_hidden[i] == C, "all ST-LSTM layers must have the same num_hidden");
    cp = p * p * d.img_c;
    hp_ = d.img_h / p;
    wp_ = d.img_w / p;
    for (int i = 0; i < L; ++i) {
      const std::string pre = "cell_list." + std::to_string(i) + ".";
      const int cin = (i == 0) ? cp : C;
      declare(pre + "conv_x.0.weight", {7 * C, cin, k, k});
      declare(pre + "conv_h.0.weight", {4 * C, C, k, k});
      declare(pre + "conv_m.0.weight", {3 * C, C, k, k});
      declare(pre + "conv_o.0.weight", {C, 2 * C, k, k});
      declare(pre + "conv_last.weight", {C, 2 * C, 1, 1});
    }
    declare("conv_last.weight", {cp, C, 1, 1});
    declare("adapter.weight", {C, C, 1, 1});
  }
  ~PredRnnV2() override {
    if (d_loss) cudaFree(d_loss);
  }

 protected:
  void validate(int t_in, int pred) const override {
    // "needs input sequences that also include the target frames" (predrnn_v2.py:134-137)
    if (t_in - pred < 1) VPK_THROW(1, "predrnn-pp needs input sequences that also include the target frames");
  }
  int default_microbatch() const override { return 256; }

  void begin_call(int, float*, cudaStream_t stream) override {
    if (!d_loss) VPK_CUDA(cudaMalloc(&d_loss, sizeof(double)));
    VPK_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(double), stream));
  }
  void end_call(int batch, float* aux, cudaStream_t stream) override {
    if (aux == nullptr) return;
    // 100 * mean over (t, layer) of mean over (b, ch)   (predrnn_v2.py:209-211, 229-230)
    const double scale = static_cast<double>(desc.decoupling_loss_scale) /
                         (static_cast<double>(n_terms) * batch * C);
    launch_decouple_finalize(d_loss, aux, scale, stream);
  }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const ActInfo act{dtype, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int ctx = t_in - pred;
    const size_t px = static_cast<size_t>(B) * hp_ * wp_;
    n_terms = (t_in - 1) * L;
    if (!measure && !d_loss) VPK_CUDA(cudaMalloc(&d_loss, sizeof(double)));

    // only the context frames are ever read (eval mask = 0): patchify those
    char* xp = static_cast<char*>(arena.alloc(px * cp * esz * ctx));
    float* out_stage = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * pred * c * h * w * sizeof(float)));
    std::vector<void*> hb(2 * L), memb(2 * L);
    std::vector<float*> cb(L);
    for (int i = 0; i < L; ++i) {
      hb[2 * i] = arena.alloc(px * C * esz);
      hb[2 * i + 1] = arena.alloc(px * C * esz);
      memb[2 * i] = arena.alloc(px * 2 * C * esz);
      memb[2 * i + 1] = arena.alloc(px * 2 * C * esz);
      cb[i] = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    }
    float* mstate = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    float* opart = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
    char* dcdm = static_cast<char*>(arena.alloc(2 * px * C * esz));          // [delta_c ; delta_m] stacked on batch
    float* adapt = static_cast<float*>(arena.alloc(2 * px * C * sizeof(float)));
    float* xgen32 = static_cast<float*>(arena.alloc(px * cp * sizeof(float)));
    void* xgen_act = (dtype == DT_F32) ? static_cast<void*>(xgen32) : arena.alloc(px * cp * esz);

    if (!measure) {
      const int ns = num_sms, dt = dtype, pp = p;
      Op pre;
      pre.name = "patchify";
      // x holds t_in frames per sequence; frames >= ctx are ignored
      pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
        launch_patchify_strided(rc.x, static_cast<long long>(t_in) * c * h * w, xp, dt, B, ctx, c, h, w, pp, ns, s);
      };
      prog.pre.push_back(std::move(pre));
      for (int i = 0; i < L; ++i) {
        add_memset(prog, hb[2 * i], px * C * esz, "zero_h");
        add_memset(prog, cb[i], px * C * sizeof(float), "zero_c");
      }
      add_memset(prog, mstate, px * C * sizeof(float), "zero_m");
      add_memset(prog, memb[2 * (L - 1) + 1], px * 2 * C * esz, "zero_mem");   // m seen by layer 0 at t = 0
    }

    std::vector<int> par(L, 0);
    for (int t = 0; t < t_in - 1; ++t) {
      const void* net = (t < ctx) ? static_cast<const void*>(xp + static_cast<size_t>(t) * px * cp * esz) : xgen_act;
      for (int i = 0; i < L; ++i) {
        const std::string pre = "cell_list." + std::to_string(i) + ".";
        const void* inp = (i == 0) ? net : hb[2 * (i - 1) + par[i - 1]];
        const int cin = (i == 0) ? cp : C;
        // memory comes from the previous layer of this step, or from the top layer of the previous step
        const void* mem_prev = (i == 0) ? memb[2 * (L - 1) + ((t + 1) & 1)] : memb[2 * (i - 1) + (t & 1)];
        StLstmArgs a{pre, B, hp_, wp_, cin, C, k, inp, hb[2 * i + par[i]],
                     make_channel_view(mem_prev, hp_, wp_, 2 * C, C, C, esz), hb[2 * i + (par[i] ^ 1)], cb[i], mstate,
                     opart, memb[2 * i + (t & 1)], dcdm, dcdm + px * C * esz,
                     hp(pre + "conv_x.0.weight"), hp(pre + "conv_h.0.weight"), hp(pre + "conv_m.0.weight"),
                     hp(pre + "conv_o.0.weight"), hp(pre + "conv_last.weight")};
        a.c4 = true;
        for (const ConvSpec& sp : stlstm_specs(a, act)) add_conv(prog, sp, measure, stream);
        par[i] ^= 1;
        // decoupling-loss term: adapter (1x1, no bias) over [delta_c ; delta_m], then the per-(b, ch) cosine
        int oh, ow;
        ConvArgs ad{"adapter.", 2 * B, hp_, wp_, C, C, 1, 1, 0, dcdm, hp("adapter.weight"), nullptr, ACT_NONE, adapt};
        ad.f32_strided = true;
        ad.oB = static_cast<long long>(hp_) * wp_ * C;
        ad.oY = static_cast<long long>(wp_) * C;
        ad.oX = C;
        ad.oC = 1;
        add_conv(prog, conv_spec(ad, act, &oh, &ow), measure, stream);
        if (!measure) {
          const int HW = hp_ * wp_, CC = C;
          double* acc = d_loss;
          Op op;
          op.name = "decouple_reduce";
          op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_reduce(adapt, B, HW, CC, acc, s); };
          prog.body.push_back(std::move(op));
        }
      }
      // head: x_gen = conv_last(h_top)  (1x1, no bias), kept in fp32 so that output frames carry no extra rounding
      int oh, ow;
      ConvArgs hd{"conv_last.", B, hp_, wp_, C, cp, 1, 1, 0, hb[2 * (L - 1) + par[L - 1]], hp("conv_last.weight"),
                  nullptr, ACT_NONE, xgen32};
      hd.f32_strided = true;
      hd.oB = static_cast<long long>(hp_) * wp_ * cp;
      hd.oY = static_cast<long long>(wp_) * cp;
      hd.oX = cp;
      hd.oC = 1;
      add_conv(prog, conv_spec(hd, act, &oh, &ow), measure, stream);
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
        VPK_CUDA(cudaMemcpyAsync(rc.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s));
      };
      prog.post.push_back(std::move(post));
    }
  }

 private:
  int p = 4, L = 3, k = 5, C = 128, cp = 16, hp_ = 16, wp_ = 16;
  int n_terms = 1;
  double* d_loss = nullptr;
};

}  // namespace

Model* make_predrnn(const vpk_model_desc& d) { return new PredRnnV2(d); }

}  // namespace vpk
