// phy: PhyDNet rollout, non action-conditional, eval mode (reference: models/phydnet.py:73-137), and the BASELINE
// config-2 composition "convlstm-branch" = its residual branch alone (DCGANEncoder -> EncoderSplit ->
// SingleStepConvLSTM -> DecoderSplit -> DCGANDecoder -> sigmoid; our composition of reference blocks).
//
// Per encoder_fwd call (models/phydnet.py:73-89):
//   E (3 DCGANConv) -> Ep / Er (2 DCGANConv each) -> PhyCell (model_blocks/phydnet.py:49-62, 95-105) and the stacked
//   ConvLSTMCell (model_blocks/phydnet.py:147-163; cell: conv_lstm_ndrplz.py:28-43) -> Dp / Dr (2 DCGANConvTranspose
//   each) -> D (2 DCGANConvTranspose + ConvTranspose2d) -> sigmoid.
// Warm-up steps on context frames only advance the recurrent state (their images are discarded in eval, :108-113), so
// the decoders are skipped there; of the three decoder_D passes per step only `output_image` (:87-88) is computed.
//
// Precision in bf16 mode: the recurrent cells (85 % of the FLOPs) run bf16 on the tensor cores; the DCGAN encoder /
// decoder convs feed GroupNorm, which amplifies operand rounding past the 5e-3 single-step bound,
// so they keep fp32 operands.
#include <cmath>

#include "builders.h"
#include "elementwise.h"
#include "model.h"
#include "phycell.h"

namespace vpk {

namespace {

int group_norm_divisor(int x) {   // model_blocks/phydnet.py:348-362
  int sq = static_cast<int>(std::floor(std::sqrt(static_cast<double>(x))));
  while (x % sq != 0) --sq;
  return x / sq;
}

class PhyDNetModel : public Model {
 public:
  PhyDNetModel(const vpk_model_desc& d, bool branch) : Model(d), branch_only(branch) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h % 4 == 0 && d.img_w % 4 == 0 && d.img_h > 0 && d.img_w > 0,
                "img size must be a multiple of 4 (other sizes need the reference's Resize)");
    n_phy = d.phycell_n_layers;
    hid = d.phycell_channels;
    kp = d.phycell_kernel_size;
    n_lstm = d.convlstm_n_layers;
    kl = d.convlstm_kernel_size;
    VPK_REQUIRE(n_phy >= 1 && n_phy <= 4 && hid > 0 && kp % 2 == 1, "bad PhyCell hyper-parameters");
    VPK_REQUIRE(n_lstm >= 1 && n_lstm <= 8 && kl % 2 == 1, "bad ConvLSTM hyper-parameters");
    const int c = d.img_c;
    auto dcgan = [&](const std::string& p, int cin, int cout, bool transpose) {
      declare(p + "main.0.weight", transpose ? std::vector<int64_t>{cin, cout, 3, 3} : std::vector<int64_t>{cout, cin, 3, 3});
      declare(p + "main.0.bias", {cout});
      declare(p + "main.1.weight", {cout});
      declare(p + "main.1.bias", {cout});
    };
    dcgan("encoder_E.c1.", c, 32, false);
    dcgan("encoder_E.c2.", 32, 32, false);
    dcgan("encoder_E.c3.", 32, 64, false);
    for (const char* e : {"encoder_Ep.", "encoder_Er."}) {
      dcgan(std::string(e) + "c1.", 64, 64, false);
      dcgan(std::string(e) + "c2.", 64, 64, false);
    }
    for (const char* e : {"decoder_Dp.", "decoder_Dr."}) {
      dcgan(std::string(e) + "upc1.", 64, 64, true);
      dcgan(std::string(e) + "upc2.", 64, 64, true);
    }
    dcgan("decoder_D.upc1.", 64, 32, true);
    dcgan("decoder_D.upc2.", 32, 32, true);
    declare("decoder_D.upc3.weight", {32, c, 3, 3});
    declare("decoder_D.upc3.bias", {c});
    for (int j = 0; j < n_phy; ++j) {
      const std::string p = "phycell.cell_list." + std::to_string(j) + ".";
      declare(p + "F.conv1.weight", {hid, 64, kp, kp});
      declare(p + "F.conv1.bias", {hid});
      declare(p + "F.bn1.weight", {hid});
      declare(p + "F.bn1.bias", {hid});
      declare(p + "F.conv2.weight", {64, hid, 1, 1});
      declare(p + "F.conv2.bias", {64});
      declare(p + "convgate.weight", {64, 128, 3, 3});
      declare(p + "convgate.bias", {64});
    }
    int cin = 64;
    for (int j = 0; j < n_lstm; ++j) {
      const int hd = d.convlstm_hidden_dims[j];
      VPK_REQUIRE(hd > 0, "bad convlstm_hidden_dims");
      const std::string p = "convcell.cell_list." + std::to_string(j) + ".conv.";
      declare(p + "weight", {4 * hd, cin + hd, kl, kl});
      declare(p + "bias", {4 * hd});
      cin = hd;
    }
    // DecoderSplit consumes the top ConvLSTM layer (64 channels in the reference)
    VPK_REQUIRE(cin == 64, "the top ConvLSTM layer must have 64 channels (decoder_Dr input)");
  }

 protected:
  int default_microbatch() const override { return 256; }

  std::vector<float> vec(const std::string& key) const {
    const HostParam& p = params.at(key);
    return p.data;
  }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const int edt = DT_F32;               // encoder / decoder operand type (see header comment)
    const int cdt = dtype;                // recurrent-cell operand type
    const ActInfo ea{edt, 4}, ca{cdt, esize()};
    const int esz_c = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int h2 = h / 2, w2 = w / 2, h4 = h / 4, w4 = w / 4;
    const size_t px1 = static_cast<size_t>(B) * h * w, px2 = static_cast<size_t>(B) * h2 * w2,
                 px4 = static_cast<size_t>(B) * h4 * w4;
    const int Cp = phycell_padded_channels(hid);
    const int ns = num_sms;

    float* frames_in = static_cast<float*>(arena.alloc(px1 * c * 4 * t_in));
    float* out_stage = static_cast<float*>(arena.alloc(px1 * c * 4 * pred));
    float* frame_fb = static_cast<float*>(arena.alloc(px1 * c * 4));          // fed-back frame, NHWC
    float* raw = static_cast<float*>(arena.alloc(std::max(px2 * 32, px4 * 64) * 4));   // pre-GroupNorm conv output
    float* e1 = static_cast<float*>(arena.alloc(px2 * 32 * 4));
    float* e2 = static_cast<float*>(arena.alloc(px2 * 32 * 4));
    float* e3 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    float* mid = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    void* ep = arena.alloc(px4 * 64 * esz_c);
    void* er = arena.alloc(px4 * 64 * esz_c);
    // PhyCell state
    std::vector<float*> hp_master(n_phy), htilde(n_phy);
    std::vector<void*> hp_act(2 * n_phy);
    float* f1raw = static_cast<float*>(arena.alloc(px4 * Cp * 4));
    void* f1n = arena.alloc(px4 * Cp * esz_c);
    if (!branch_only)
      for (int j = 0; j < n_phy; ++j) {
        hp_master[j] = static_cast<float*>(arena.alloc(px4 * 64 * 4));
        htilde[j] = static_cast<float*>(arena.alloc(px4 * 64 * 4));
        hp_act[2 * j] = arena.alloc(px4 * 64 * esz_c);
        hp_act[2 * j + 1] = arena.alloc(px4 * 64 * esz_c);
      }
    // ConvLSTM state
    std::vector<void*> hb(2 * n_lstm);
    std::vector<float*> cb(n_lstm);
    for (int j = 0; j < n_lstm; ++j) {
      const int hd = d.convlstm_hidden_dims[j];
      hb[2 * j] = arena.alloc(px4 * hd * esz_c);
      hb[2 * j + 1] = arena.alloc(px4 * hd * esz_c);
      cb[j] = static_cast<float*>(arena.alloc(px4 * hd * 4));
    }
    float* h_top32 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    float* dp = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    float* dsum = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    float* d1 = static_cast<float*>(arena.alloc(px2 * 32 * 4));
    float* d2 = static_cast<float*>(arena.alloc(px2 * 32 * 4));

    auto gn_op = [&](const std::string& key, const float* in, void* out, int out_dt, const void* add, int HW, int C,
                     int Cs_in, int Cs_out, int groups, int actk) {
      if (measure) return;
      const float* g = dev_f32(key + "weight", vec(key + "weight"), stream);
      const float* bta = dev_f32(key + "bias", vec(key + "bias"), stream);
      Op op;
      op.name = "groupnorm " + key;
      op.fn = [=](cudaStream_t s, const RunCtx&) {
        launch_groupnorm_act(in, DT_F32, out, out_dt, add, B, HW, C, Cs_in, Cs_out, groups, g, bta, 1e-5f, actk, s);
      };
      prog.body.push_back(std::move(op));
    };
    // DCGANConv / DCGANConvTranspose: conv -> GroupNorm(16) -> LeakyReLU(0.2)   (model_blocks/conv.py:58-95)
    auto dcgan = [&](const std::string& p, bool transpose, const void* in, int H, int W, int Cin, int Cout, int stride,
                     void* out, int out_dt, const void* add) {
      int oh, ow;
      if (!transpose) {
        ConvArgs a{p + "main.0.", B, H, W, Cin, Cout, 3, stride, 1, in, hp(p + "main.0.weight"), hp(p + "main.0.bias"),
                   ACT_NONE, raw};
        a.out_f32_dense = true;
        add_conv(prog, conv_spec(a, ea, &oh, &ow), measure, stream, edt);
      } else {
        DeconvArgs a{p + "main.0.", B, H, W, Cin, Cout, 3, stride, 1, stride == 2 ? 1 : 0, in, hp(p + "main.0.weight"),
                     hp(p + "main.0.bias"), ACT_NONE, raw};
        a.out_f32 = true;
        add_conv(prog, deconv_spec(a, ea, &oh, &ow), measure, stream, edt);
      }
      gn_op(p + "main.1.", raw, out, out_dt, add, oh * ow, Cout, Cout, Cout, 16, ACT_LEAKY);
    };

    if (!measure) {
      Op pre;
      pre.name = "frames_to_nhwc";
      pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
        launch_frames_to_nhwc(rc.x, frames_in, DT_F32, B, t_in, c, h, w, ns, s);
      };
      prog.pre.push_back(std::move(pre));
      if (!branch_only) {
        for (int j = 0; j < n_phy; ++j) {
          add_memset(prog, hp_master[j], px4 * 64 * 4, "zero_hp");
          add_memset(prog, hp_act[2 * j], px4 * 64 * esz_c, "zero_hp_act");
        }
        add_memset(prog, f1n, px4 * Cp * esz_c, "zero_f1n_pad");
      }
      for (int j = 0; j < n_lstm; ++j) {
        add_memset(prog, hb[2 * j], px4 * d.convlstm_hidden_dims[j] * esz_c, "zero_h");
        add_memset(prog, cb[j], px4 * d.convlstm_hidden_dims[j] * 4, "zero_c");
      }
    }

    std::vector<int> ppar(n_phy, 0), lpar(n_lstm, 0);
    const int n_steps = (t_in - 1) + pred;
    for (int st = 0; st < n_steps; ++st) {
      const bool decode = st >= t_in - 1;                        // produces a predicted frame
      const int di = st - (t_in - 1);
      const float* frame = (st < t_in) ? frames_in + static_cast<size_t>(st) * px1 * c : frame_fb;
      // ---- encoders ----
      dcgan("encoder_E.c1.", false, frame, h, w, c, 32, 2, e1, DT_F32, nullptr);
      dcgan("encoder_E.c2.", false, e1, h2, w2, 32, 32, 1, e2, DT_F32, nullptr);
      dcgan("encoder_E.c3.", false, e2, h2, w2, 32, 64, 2, e3, DT_F32, nullptr);
      if (!branch_only) {
        dcgan("encoder_Ep.c1.", false, e3, h4, w4, 64, 64, 1, mid, DT_F32, nullptr);
        dcgan("encoder_Ep.c2.", false, mid, h4, w4, 64, 64, 1, ep, cdt, nullptr);
      }
      dcgan("encoder_Er.c1.", false, e3, h4, w4, 64, 64, 1, mid, DT_F32, nullptr);
      dcgan("encoder_Er.c2.", false, mid, h4, w4, 64, 64, 1, er, cdt, nullptr);

      // ---- PhyCell stack (model_blocks/phydnet.py:95-105) ----
      if (!branch_only) {
        const void* xin = ep;
        for (int j = 0; j < n_phy; ++j) {
          const std::string p = "phycell.cell_list." + std::to_string(j) + ".";
          const void* h_act = hp_act[2 * j + ppar[j]];
          void* h_act_new = hp_act[2 * j + (ppar[j] ^ 1)];
          PhyCellArgs pa{p, B, h4, w4, 64, hid, kp, xin, h_act, h_act_new, hp_master[j], htilde[j], f1raw, f1n,
                         hp(p + "F.conv1.weight"), hp(p + "F.conv1.bias"), hp(p + "F.conv2.weight"),
                         hp(p + "F.conv2.bias"), hp(p + "convgate.weight"), hp(p + "convgate.bias")};
          std::vector<ConvSpec> specs = phycell_specs(pa, ca);
          add_conv(prog, specs[0], measure, stream, cdt);
          // F.bn1 = GroupNorm(find_divisor(hid), hid), no activation
          gn_op(p + "F.bn1.", f1raw, f1n, cdt, nullptr, h4 * w4, hid, Cp, Cp, group_norm_divisor(hid), ACT_NONE);
          add_conv(prog, specs[1], measure, stream, cdt);
          add_conv(prog, specs[2], measure, stream, cdt);
          ppar[j] ^= 1;
          xin = h_act_new;
        }
      }
      // ---- ConvLSTM stack (model_blocks/phydnet.py:147-163) ----
      {
        const void* xin = er;
        int cin = 64;
        for (int j = 0; j < n_lstm; ++j) {
          const int hd = d.convlstm_hidden_dims[j];
          const std::string p = "convcell.cell_list." + std::to_string(j) + ".conv.";
          LstmArgs la{p, B, h4, w4, cin, hd, kl, xin, hb[2 * j + lpar[j]], hb[2 * j + (lpar[j] ^ 1)], cb[j],
                      hp(p + "weight"), hp(p + "bias"), true, nullptr, nullptr, nullptr};
          la.c4 = true;
          ConvSpec s = lstm_spec(la, ca);
          if (j == n_lstm - 1) s.phases[0].epi.h32 = h_top32;
          add_conv(prog, s, measure, stream, cdt);
          lpar[j] ^= 1;
          xin = hb[2 * j + lpar[j]];
          cin = hd;
        }
      }
      if (!decode) continue;
      // ---- decoders ----
      const float* dec_in = h_top32;
      if (!branch_only) {
        dcgan("decoder_Dp.upc1.", true, hp_master[n_phy - 1], h4, w4, 64, 64, 1, mid, DT_F32, nullptr);
        dcgan("decoder_Dp.upc2.", true, mid, h4, w4, 64, 64, 1, dp, DT_F32, nullptr);
      }
      dcgan("decoder_Dr.upc1.", true, dec_in, h4, w4, 64, 64, 1, mid, DT_F32, nullptr);
      // concat = decoded_phys + decoded_conv (models/phydnet.py:87) folded into the last GroupNorm pass
      dcgan("decoder_Dr.upc2.", true, mid, h4, w4, 64, 64, 1, dsum, DT_F32, branch_only ? nullptr : dp);
      dcgan("decoder_D.upc1.", true, dsum, h4, w4, 64, 32, 2, d1, DT_F32, nullptr);
      dcgan("decoder_D.upc2.", true, d1, h2, w2, 32, 32, 1, d2, DT_F32, nullptr);
      {
        int oh, ow;
        DeconvArgs a{"decoder_D.upc3.", B, h2, w2, 32, c, 3, 2, 1, 1, d2, hp("decoder_D.upc3.weight"),
                     hp("decoder_D.upc3.bias"), ACT_SIGMOID, out_stage + static_cast<size_t>(di) * c * h * w};
        a.nchw = true;
        a.oB_nchw = static_cast<long long>(pred) * c * h * w;
        add_conv(prog, deconv_spec(a, ea, &oh, &ow), measure, stream, edt);
        VPK_REQUIRE(oh == h && ow == w, "decoder output size mismatch");
      }
      if (!measure && di + 1 < pred) {   // next decoder input = output_image (models/phydnet.py:121)
        const float* src = out_stage + static_cast<size_t>(di) * c * h * w;
        const long long bs = static_cast<long long>(pred) * c * h * w;
        Op op;
        op.name = "feedback_frame";
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          launch_frames_to_nhwc_strided(src, bs, frame_fb, DT_F32, B, 1, c, h, w, ns, s);
        };
        prog.body.push_back(std::move(op));
      }
    }
    if (!measure) {
      const size_t bytes = px1 * c * sizeof(float) * pred;
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
  bool branch_only;
  int n_phy = 1, hid = 49, kp = 7, n_lstm = 3, kl = 3;
};

}  // namespace

Model* make_phydnet(const vpk_model_desc& d, bool branch_only) { return new PhyDNetModel(d, branch_only); }

}  // namespace vpk
