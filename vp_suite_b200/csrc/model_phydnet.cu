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
// Precision in bf16 mode: the recurrent cells (85 % of the FLOPs) run bf16 on the tensor cores.  The DCGAN encoder /
// decoder convs feed GroupNorm, which amplifies operand rounding past the 5e-3 single-step bound with bf16 operands
// (tests/tools/precision_probe.py: 4.3e-3 .. 8.1e-3 per frame), so they need more mantissa bits.
// Default: FP16 operands (11 bits; the feature maps are O(1) after GroupNorm + LeakyReLU, the frames lie in [0, 1] and
// the accumulators stay fp32 in TMEM) -- one tcgen05 product per conv, 2-byte feature maps, first-frame error 1.1e-3
// (fp32 convs next to bf16 cells: 1.0e-3, the cells dominate).  VPK_FEAT_SPLIT=1 selects the earlier SPLIT-bf16 form
// instead (hi + lo, 16 mantissa bits, three bf16 products A_hi W_hi + A_hi W_lo + A_lo W_hi per conv), kept for A/B runs.
#include <cmath>
#include <cstdlib>

#include "builders.h"
#include "conv_stem.h"
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
    if (const char* env = getenv("VPK_FEAT_SPLIT")) feat_split = atoi(env) != 0;
    VPK_REQUIRE(d.img_c > 0 && d.img_h % 4 == 0 && d.img_w % 4 == 0 && d.img_h > 0 && d.img_w > 0,
                "img size must be a multiple of 4 (other sizes need the reference's Resize)");
    n_phy = d.phycell_n_layers;
    hid = d.phycell_channels;
    kp = d.phycell_kernel_size;
    n_lstm = d.convlstm_n_layers;
    kl = d.convlstm_kernel_size;
    ac = d.action_conditional != 0 && !branch;
    a_sz = ac ? d.action_size : 0;
    VPK_REQUIRE(!ac || a_sz > 0, "action-conditional phy needs action_size > 0");
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
      if (ac) {                 // model_blocks/phydnet.py:44-48
        declare(p + "frame_action_conv.weight", {64, 64 + a_sz, 1, 1});
        declare(p + "frame_action_conv.bias", {64});
        declare(p + "hidden_action_conv.weight", {64, 64 + a_sz, 1, 1});
        declare(p + "hidden_action_conv.bias", {64});
      }
    }
    int cin = 64 + (ac ? a_sz : 0);    // SingleStepConvLSTM: the inflated action joins the bottom layer's input (:137, 153-155)
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
  bool streams_input() const override { return !desc.use_cuda_graph && getenv("VPK_NO_INPUT_STREAM") == nullptr; }
  // one action per encoder_fwd call: ac_index runs over the context and the predicted steps (models/phydnet.py:108-122)
  int action_steps_needed(int t_in, int pred) const override { return ac ? t_in - 1 + pred : 0; }

  std::vector<float> vec(const std::string& key) const {
    const HostParam& p = params.at(key);
    return p.data;
  }

  // An encoder / decoder feature map.  fp32 mode: `a` is a float tensor.  bf16 mode: split-bf16, `a` holds the high
  // parts and `lo` the low parts (hi = bf16(v), lo = bf16(v - hi)), both NHWC bf16.
  struct Feat {
    void* a = nullptr;
    void* lo = nullptr;
  };
  enum OutKind { OUT_FEAT, OUT_CELL, OUT_F32 };

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const int cdt = dtype;                // recurrent-cell operand type
    const bool split = (cdt == DT_BF16) && feat_split;   // encoder / decoder convs: three bf16 products
    const bool f16 = (cdt == DT_BF16) && !feat_split;    // encoder / decoder convs: one fp16 product
    const bool tcfeat = split || f16;                     // feature maps are 16-bit tensor-core operands
    const ActInfo f32a{DT_F32, 4}, sa{f16 ? DT_F16 : DT_BF16, 2}, ca{cdt, esize()};
    const int esz_c = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int h2 = h / 2, w2 = w / 2, h4 = h / 4, w4 = w / 4;
    const size_t px1 = static_cast<size_t>(B) * h * w, px2 = static_cast<size_t>(B) * h2 * w2,
                 px4 = static_cast<size_t>(B) * h4 * w4;
    const int Cp = phycell_padded_channels(hid);
    const int ns = num_sms;

    auto feat = [&](size_t elems) {
      Feat f;
      if (split) {
        f.a = arena.alloc(elems * 2);
        f.lo = arena.alloc(elems * 2);
      } else if (f16) {
        f.a = arena.alloc(elems * 2);
      } else {
        f.a = arena.alloc(elems * 4);
      }
      return f;
    };
    // split mode: frames as split-bf16 with 8 channels per pixel (zero padded): TMA-addressable, so encoder_E.c1 runs on
    // the tensor cores like every other conv; fp32 mode: plain fp32 NHWC frames for the CUDA-core kernel
    const bool pad8 = tcfeat && backend == 0 && c <= 8;
    const int cs = pad8 ? 8 : c;
    const bool direct_ok = f16 && pad8 && getenv("VPK_NO_STEM") == nullptr;
    const bool stem_direct = direct_ok && conv_stem_supported(3, 2, 1, c, 32, h, w);
    const bool tail_direct = direct_ok && deconv_tail_supported(3, 2, 1, 1, 32, c, h2, w2);
    Feat frames_in, frame_fb;            // [t_in] frames / the fed-back frame
    if (pad8) {
      frames_in.a = arena.alloc(px1 * 8 * 2 * t_in);
      frame_fb.a = arena.alloc(px1 * 8 * 2);
      if (split) {
        frames_in.lo = arena.alloc(px1 * 8 * 2 * t_in);
        frame_fb.lo = arena.alloc(px1 * 8 * 2);
      }
    } else {
      frames_in.a = arena.alloc(px1 * c * 4 * t_in);
      frame_fb.a = arena.alloc(px1 * c * 4);
    }
    float* out_stage = static_cast<float*>(arena.alloc(px1 * c * 4 * pred));
    // The encoders of the CONTEXT frames do not depend on the recurrence: they run once, time-batched over all t_in
    // frames (batch t_in * B -- GroupNorm is per sample, so the frames are just more samples), instead of t_in times on
    // B sequences; the step loop then only consumes their outputs.  Fed-back frames are encoded per step as before.
    // (tcgen05 / fp16-feature path; ~100 short launches fewer per cfg-2 rollout.)
    // (action-conditional: PhyCell's frame input is needed in fp32 per step -- kept per step for simplicity)
    const bool batch_ctx = f16 && pad8 && t_in >= 2 && !ac && getenv("VPK_NO_CTX_BATCH") == nullptr;
    const size_t tb = batch_ctx ? static_cast<size_t>(t_in) : 1;     // buffers shared by both uses are sized for t_in * B
    float* raw = static_cast<float*>(arena.alloc(std::max(px2 * 32, px4 * 64) * 4 * tb));   // pre-GroupNorm conv output
    Feat e1 = feat(px2 * 32 * tb), e2 = feat(px2 * 32 * tb), e3 = feat(px4 * 64 * tb), mid = feat(px4 * 64 * tb);
    void* ep = arena.alloc(px4 * 64 * esz_c * tb);     // batch_ctx: [t_in][B, h4, w4, 64], slice st = step st's cell input
    void* er = arena.alloc(px4 * 64 * esz_c * tb);
    int Bcur = B;                                       // batch the dcgan / gn_op / stem builders below emit launches for
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
    // action-conditional: inflated actions of every step (fp32 for PhyCell's 1x1 action convs, cell type for the ConvLSTM
    // gate conv), PhyCell's fp32 frame input and the convolved frame / hidden (fp32 + cell-type copies)
    const int n_steps_all = (t_in - 1) + pred;
    const int a_pad = 8;
    float* act32 = nullptr;
    char* act16 = nullptr;
    float *ep32 = nullptr, *fa32 = nullptr, *ha32 = nullptr;
    void *fa_act = nullptr, *ha_act = nullptr;
    if (ac) {
      VPK_REQUIRE(a_sz <= a_pad, "action_size above 8 is not supported by the action-conditional phy rollout");
      act32 = static_cast<float*>(arena.alloc(px4 * a_pad * 4 * n_steps_all));
      act16 = (cdt == DT_F32) ? reinterpret_cast<char*>(act32) : static_cast<char*>(arena.alloc(px4 * a_pad * esz_c * n_steps_all));
      ep32 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
      fa32 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
      ha32 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
      fa_act = (cdt == DT_F32) ? static_cast<void*>(fa32) : arena.alloc(px4 * 64 * esz_c);
      ha_act = (cdt == DT_F32) ? static_cast<void*>(ha32) : arena.alloc(px4 * 64 * esz_c);
    }
    float* h_top32 = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    float* dp = static_cast<float*>(arena.alloc(px4 * 64 * 4));
    Feat dsum = feat(px4 * 64), d1 = feat(px2 * 32), d2 = feat(px2 * 32);
    Feat dec_p, dec_r;                   // decoder inputs: PhyCell h and top ConvLSTM h as feature maps
    if (tcfeat) {
      dec_p = feat(px4 * 64);
      dec_r = feat(px4 * 64);
    }

    // GroupNorm partial statistics written by the producing conv's epilogue (tcgen05 path, fp16 feature maps): one
    // region [B][slots][16][2] fp32 reused by every DCGAN block (stream order separates producer / consumer pairs);
    // slots = (phases) x (8x16 tiles per image) x 4 warps, at most 4 x 8 x 4 for the 32 x 32 maps of a 64 x 64 image
    const char* halo_env = getenv("VPK_TC_HALO");          // tests force the per-tap kernels with VPK_TC_HALO=0
    const bool fuse_gn = f16 && backend == 0 && getenv("VPK_NO_FUSED_GN") == nullptr && (halo_env == nullptr || atoi(halo_env) != 0);
    auto gn_tiles = [](int H, int W) { return ((H + 15) / 16) * ((W + 7) / 8); };
    const int gn_max_slots = 4 * 4 * std::max(gn_tiles(h2, w2), gn_tiles(h4, w4));
    float* gn_region = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * tb * gn_max_slots * 16 * 2 * sizeof(float)));
    // GroupNorm (+ LeakyReLU) of the fp32 conv output `in` into a feature map / cell operand / fp32 tensor
    auto gn_op = [&](const std::string& key, const float* in, OutKind kind, Feat out, const float* add, int HW, int C,
                     int groups, int actk, const float* sums = nullptr, int nslots = 0, bool raw16 = false) {
      if (measure) return;
      const float* g = dev_f32(key + "weight", vec(key + "weight"), stream);
      const float* bta = dev_f32(key + "bias", vec(key + "bias"), stream);
      const int Bx = Bcur;
      const int out_dt = (kind == OUT_F32) ? DT_F32 : (kind == OUT_CELL) ? cdt : (split ? DT_BF16 : f16 ? DT_F16 : DT_F32);
      const bool two = (kind == OUT_FEAT) && split;
      Op op;
      op.name = "groupnorm " + key;
      if (sums != nullptr) {
        const int ok = (out_dt == DT_BF16 ? 1 : out_dt == DT_F16 ? 3 : 0);
        const int in16 = raw16 ? 1 : 0;     // the fused-statistics convs leave their raw output as fp16 (see dcgan below)
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          launch_groupnorm_apply(in, in16, out.a, ok, add, sums, nslots, Bx, HW, C, groups, g, bta, 1e-5f, actk, ns, s);
        };
      } else if (groupnorm_smem_supported(HW, C, groups)) {
        const int ok = two ? 2 : (out_dt == DT_BF16 ? 1 : out_dt == DT_F16 ? 3 : 0);
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          launch_groupnorm_smem(in, out.a, out.lo, ok, add, Bx, HW, C, groups, g, bta, 1e-5f, actk, s);
        };
      } else {
        VPK_REQUIRE(!two && (add == nullptr || out_dt == DT_F32), "groupnorm: shape needs the shared-memory kernel");
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          launch_groupnorm_act(in, DT_F32, out.a, out_dt, add, Bx, HW, C, C, C, groups, g, bta, 1e-5f, actk, s);
        };
      }
      prog.body.push_back(std::move(op));
    };
    // DCGANConv / DCGANConvTranspose: conv -> GroupNorm(16) -> LeakyReLU(0.2)   (model_blocks/conv.py:58-95)
    // `in_f32`: the input is a plain fp32 tensor (image frames: CUDA-core kernel), otherwise a feature map
    auto dcgan = [&](const std::string& p, bool transpose, Feat in, bool in_f32, int H, int W, int Cin, int Cout,
                     int stride, OutKind kind, Feat out, const float* add, int cin_w = -1) {
      int oh, ow;
      const bool sp = split && !in_f32;
      const ActInfo& ai = (tcfeat && !in_f32) ? sa : f32a;
      const int gsz = Cout / 16;
      float* sums = nullptr;
      if (fuse_gn && !in_f32 && Cout % 16 == 0 && (gsz == 2 || gsz == 4 || gsz == 8) && Cout <= 64) sums = gn_region;
      int nslots = 0;
      auto with_stats = [&](ConvSpec sp_) {
        if (sums != nullptr) {
          int slot0 = 0;
          for (PhaseSpec& ph : sp_.phases) {
            ph.epi.gn_sums = sums;
            ph.epi.gn_group_size = gsz;
            ph.epi.gn_slot0 = slot0;
            slot0 += 4 * gn_tiles(ph.H, ph.W);
          }
          nslots = slot0;
          VPK_REQUIRE(nslots <= gn_max_slots, "GroupNorm statistics region too small");
          for (PhaseSpec& ph : sp_.phases) ph.epi.gn_nslots = nslots;
        }
        return sp_;
      };
      // With the statistics fused into the epilogue (taken from the fp32 accumulators), the raw conv output only feeds
      // the GroupNorm apply pass: stored as fp16 it halves that round trip, and the conv runs the lean epilogue with staged
      // bulk stores (16-bit outputs only).  The separate-statistics paths keep fp32 raw values.
      const char* fe_env = getenv("VPK_TC_FAST_EPI");
      const bool raw16 = sums != nullptr && getenv("VPK_GN_RAW32") == nullptr && (fe_env == nullptr || atoi(fe_env) != 0);
      if (!transpose) {
        ConvArgs a{p + "main.0.", Bcur, H, W, Cin, Cout, 3, stride, 1, in.a, hp(p + "main.0.weight"), hp(p + "main.0.bias"),
                   ACT_NONE, raw};
        a.out_f32_dense = !raw16;
        a.out_f16 = raw16;
        a.split = sp;
        a.x_lo = in.lo;
        a.cin_w = cin_w;
        add_conv(prog, with_stats(conv_spec(a, ai, &oh, &ow)), measure, stream, ai.dtype);
      } else {
        DeconvArgs a{p + "main.0.", Bcur, H, W, Cin, Cout, 3, stride, 1, stride == 2 ? 1 : 0, in.a, hp(p + "main.0.weight"),
                     hp(p + "main.0.bias"), ACT_NONE, raw};
        a.out_f32 = !raw16;
        a.out_f16 = raw16;
        a.split = sp;
        a.x_lo = in.lo;
        add_conv(prog, with_stats(deconv_spec(a, ai, &oh, &ow)), measure, stream, ai.dtype);
      }
      if (sums != nullptr && !groupnorm_apply_supported(oh * ow, Cout, 16)) VPK_THROW(1, "groupnorm_apply: unsupported shape");
      gn_op(p + "main.1.", raw, kind, out, add, oh * ow, Cout, 16, ACT_LEAKY, sums, nslots, raw16);
    };
    auto split_op = [&](const float* src, Feat dst, size_t n, const char* name) {
      if (measure) return;
      Op op;
      op.name = name;
      op.fn = [=](cudaStream_t s, const RunCtx&) {
        if (f16) launch_cast_f32_to_f16(src, dst.a, static_cast<long long>(n), ns, s);
        else launch_split_bf16(src, dst.a, dst.lo, static_cast<long long>(n), ns, s);
      };
      prog.body.push_back(std::move(op));
    };

    // Input frames [f0, f0 + n) -> channels-last.  Device entry: one pre op over all frames.  Host entry
    // (streams_input()): one body op per group of frames, marked with the last frame it reads, so that the op waits for
    // that frame's host-to-device copy only.
    const bool stream_in = host_build && streams_input();
    auto convert_frames = [&](std::vector<Op>& dst, int f0, int n, bool mark) {
      if (measure) return;
      const long long chw = static_cast<long long>(c) * h * w;
      const size_t off = static_cast<size_t>(f0) * px1 * cs * (pad8 ? 2 : 4);
      void* oa = static_cast<char*>(frames_in.a) + off;
      void* olo = frames_in.lo != nullptr ? static_cast<char*>(frames_in.lo) + off : nullptr;
      const int sdt = sa.dtype;
      Op cv;
      cv.name = "frames_to_nhwc";
      if (mark) cv.needs_input = f0 + n - 1;
      cv.fn = [=](cudaStream_t s, const RunCtx& rc) {
        if (pad8)
          launch_frames_to_nhwc8(rc.x + f0 * chw, t_in * chw, oa, olo, sdt, B, n, c, h, w, ns, s);
        else
          launch_frames_to_nhwc_strided(rc.x + f0 * chw, t_in * chw, oa, DT_F32, B, n, c, h, w, ns, s);
      };
      dst.push_back(std::move(cv));
    };
    if (!measure && ac) {
      const int HW4 = h4 * w4, steps = n_steps_all, asz = a_sz;
      Op inf;
      inf.name = "inflate_actions";
      inf.fn = [=](cudaStream_t s, const RunCtx& rc) {
        VPK_REQUIRE(rc.actions != nullptr && rc.action_steps >= steps, "Given actions are None or of the wrong size!");
        const long long bs = static_cast<long long>(rc.action_steps) * asz;
        launch_inflate_actions(rc.actions, bs, asz, act32, DT_F32, B, steps, HW4, a_pad, ns, s);
        if (cdt != DT_F32) launch_inflate_actions(rc.actions, bs, asz, act16, cdt, B, steps, HW4, a_pad, ns, s);
      };
      prog.pre.push_back(std::move(inf));
    }
    if (!measure) {
      if (!stream_in) convert_frames(prog.pre, 0, t_in, false);
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

    // encoder_E -> encoder_Ep / encoder_Er of `Bcur` frames `fr`; cell inputs go to ep_out / er_out
    auto encoders = [&](Feat fr, void* ep_out, void* er_out) {
      if (stem_direct) {   // image-channel stem: direct CUDA-core kernel (HBM-bound), then the usual GroupNorm pass
        if (!measure) {
          const std::string p = "encoder_E.c1.";
          StemArgs sa_{fr.a, 1, Bcur, h, w, c, 2, dev_f32(p + "stem.w", conv_stem_pack(hp(p + "main.0.weight"), 32, c, DT_F16), stream),
                       dev_f32(p + "stem.b", vec(p + "main.0.bias"), stream), 32, ACT_NONE, raw, 1};
          Op op;
          op.name = p + "main.0.stem";
          op.flops = 2.0 * static_cast<double>(px2) * (Bcur / B) * 32 * 9 * c;
          op.fn = [=](cudaStream_t s, const RunCtx&) { launch_conv_stem(sa_, ns, s); };
          prog.body.push_back(std::move(op));
        }
        gn_op("encoder_E.c1.main.1.", raw, OUT_FEAT, e1, nullptr, h2 * w2, 32, 16, ACT_LEAKY);
      } else {
        dcgan("encoder_E.c1.", false, fr, !pad8, h, w, cs, 32, 2, OUT_FEAT, e1, nullptr, c);
      }
      dcgan("encoder_E.c2.", false, e1, false, h2, w2, 32, 32, 1, OUT_FEAT, e2, nullptr);
      dcgan("encoder_E.c3.", false, e2, false, h2, w2, 32, 64, 2, OUT_FEAT, e3, nullptr);
      if (!branch_only) {
        dcgan("encoder_Ep.c1.", false, e3, false, h4, w4, 64, 64, 1, OUT_FEAT, mid, nullptr);
        if (ac) dcgan("encoder_Ep.c2.", false, mid, false, h4, w4, 64, 64, 1, OUT_F32, Feat{ep32, nullptr}, nullptr);
        else dcgan("encoder_Ep.c2.", false, mid, false, h4, w4, 64, 64, 1, OUT_CELL, Feat{ep_out, nullptr}, nullptr);
      }
      dcgan("encoder_Er.c1.", false, e3, false, h4, w4, 64, 64, 1, OUT_FEAT, mid, nullptr);
      dcgan("encoder_Er.c2.", false, mid, false, h4, w4, 64, 64, 1, OUT_CELL, Feat{er_out, nullptr}, nullptr);

    };
    // group_at[st] = number of context frames converted (and, with batch_ctx, encoded) right before step st.  Host
    // entry: groups of 1, 2, 4, ... frames -- the copy engine delivers frames several times faster than the steps consume
    // them, so only the first frame's copy is exposed and later groups keep most of the time-batching.
    std::vector<int> group_at(t_in, 0);
    if (stream_in) {
      for (int s0 = 0, g = 1; s0 < t_in; s0 += g, g *= 2) group_at[s0] = batch_ctx ? std::min(g, t_in - s0) : 1;
      if (!batch_ctx) std::fill(group_at.begin(), group_at.end(), 1);
    } else if (batch_ctx) {
      group_at[0] = t_in;
    }

    auto mark_frame = [&](int di) {   // the op just added completes predicted frame di (host entry: starts its D2H)
      if (measure || prog.body.empty()) return;
      Op& o = prog.body.back();
      o.frame = di;
      o.frame_src = out_stage + static_cast<size_t>(di) * c * h * w;
      o.frame_pitch = static_cast<long long>(pred) * c * h * w;
      o.frame_elems = static_cast<long long>(c) * h * w;
    };
    std::vector<int> ppar(n_phy, 0), lpar(n_lstm, 0);
    const int n_steps = (t_in - 1) + pred;
    for (int st = 0; st < n_steps; ++st) {
      const bool decode = st >= t_in - 1;                        // produces a predicted frame
      const int di = st - (t_in - 1);
      Feat frame = frame_fb;
      if (st < t_in) {
        const size_t off = static_cast<size_t>(st) * px1 * cs * (pad8 ? 2 : 4);
        frame.a = static_cast<char*>(frames_in.a) + off;
        frame.lo = pad8 ? static_cast<char*>(frames_in.lo) + off : nullptr;
      }
      // ---- encoders ----
      const void* ep_in = ep;
      const void* er_in = er;
      if (stream_in && st < t_in && group_at[st] > 0) convert_frames(prog.body, st, group_at[st], true);
      if (batch_ctx && st < t_in) {       // context frame: encoded by a time-batched pass over its group
        const size_t slot = static_cast<size_t>(st) * px4 * 64 * esz_c;
        if (group_at[st] > 0) {
          Bcur = B * group_at[st];
          encoders(frame, static_cast<char*>(ep) + slot, static_cast<char*>(er) + slot);
          Bcur = B;
        }
        ep_in = static_cast<const char*>(ep) + slot;
        er_in = static_cast<const char*>(er) + slot;
      } else {
        encoders(frame, ep, er);
      }

      // ---- PhyCell stack (model_blocks/phydnet.py:95-105) ----
      if (!branch_only) {
        const void* xin = ep_in;
        for (int j = 0; j < n_phy; ++j) {
          const std::string p = "phycell.cell_list." + std::to_string(j) + ".";
          const void* h_act = hp_act[2 * j + ppar[j]];
          void* h_act_new = hp_act[2 * j + (ppar[j] ^ 1)];
          const float* h_res = nullptr;
          if (ac) {
            // frame = frame_action_conv(cat[frame, action]), hidden = hidden_action_conv(cat[hidden, action]) (1x1 convs,
            // model_blocks/phydnet.py:50-55): fp32 operands on the CUDA cores straight from the fp32 frame / state (K = 64 +
            // a: negligible work; rounding the carried state to 16 bits here would enter h' directly)
            const float* frame32 = (j == 0) ? ep32 : hp_master[j - 1];
            const float* a32 = act32 + static_cast<size_t>(st) * px4 * a_pad;
            auto action_conv = [&](const std::string& key, const float* src, float* out32, void* out_act) {
              ConvSpec s1;
              s1.name = p + key + ".";
              s1.B = B;
              s1.G = 1;
              s1.C = 64;
              WeightRef wr;
              wr.w = hp(p + key + ".weight");
              wr.O = 64;
              wr.I = 64 + a_sz;
              wr.KH = wr.KW = 1;
              s1.wrefs.push_back(wr);
              BiasRef br;
              br.b = hp(p + key + ".bias");
              s1.biases.push_back(br);
              ConvInput i0{make_view(src, h4, w4, 64), 0, 0};
              ConvInput i1{make_view(a32, h4, w4, a_pad), 0, 64};
              i1.wc_count = a_sz;
              int oh_, ow_;
              lower_conv(s1, 1, 1, 0, {i0, i1}, h4, w4, 4, &oh_, &ow_);
              EpiParams& e = s1.phases[0].epi;
              e.kind = EPI_BIAS_ACT;
              e.act = ACT_NONE;
              e.out_f32 = 1;
              dense_out(e, out32, h4, w4, 64);
              add_conv(prog, s1, measure, stream, DT_F32);
              if (!measure && cdt != DT_F32) {
                const long long n = static_cast<long long>(px4) * 64;
                Op op;
                op.name = p + key + ".cast";
                op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(out32, DT_F32, nullptr, out_act, cdt, n, ns, s); };
                prog.body.push_back(std::move(op));
              }
            };
            action_conv("frame_action_conv", frame32, fa32, fa_act);
            action_conv("hidden_action_conv", hp_master[j], ha32, ha_act);
            xin = fa_act;
            h_act = ha_act;
            h_res = ha32;
          }
          PhyCellArgs pa{p, B, h4, w4, 64, hid, kp, xin, h_act, h_act_new, hp_master[j], htilde[j], f1raw, f1n,
                         hp(p + "F.conv1.weight"), hp(p + "F.conv1.bias"), hp(p + "F.conv2.weight"),
                         hp(p + "F.conv2.bias"), hp(p + "convgate.weight"), hp(p + "convgate.bias")};
          pa.h_res = h_res;
          std::vector<ConvSpec> specs = phycell_specs(pa, ca);
          add_conv(prog, specs[0], measure, stream, cdt);
          // tcgen05 path: GroupNorm + 1x1 conv2 + residual in one kernel (phy_f_tail_kernel), then the gate conv
          const int f_groups = group_norm_divisor(hid);
          if (backend == 0 && getenv("VPK_NO_PHY_TAIL") == nullptr && phy_f_tail_supported(h4 * w4, hid, Cp, f_groups, 64)) {
            if (!measure) {
              const float* g = dev_f32(p + "F.bn1.weight", vec(p + "F.bn1.weight"), stream);
              const float* bta = dev_f32(p + "F.bn1.bias", vec(p + "F.bn1.bias"), stream);
              const float* w2 = dev_f32(p + "F.conv2.weight", vec(p + "F.conv2.weight"), stream);
              const float* b2 = dev_f32(p + "F.conv2.bias", vec(p + "F.conv2.bias"), stream);
              const int HW = h4 * w4, hid_ = hid;
              const float* hm = h_res ? h_res : hp_master[j];
              float* ht = htilde[j];
              Op op;
              op.name = p + "F.tail (GroupNorm + conv2 + h)";
              op.flops = 2.0 * static_cast<double>(px4) * hid * 64;
              op.fn = [=](cudaStream_t s, const RunCtx&) {
                launch_phy_f_tail(f1raw, hm, ht, g, bta, w2, b2, B, HW, hid_, Cp, f_groups, 64, 1e-5f, s);
              };
              prog.body.push_back(std::move(op));
            }
            add_conv(prog, specs[2], measure, stream, cdt);
            ppar[j] ^= 1;
            xin = h_act_new;
            continue;
          }
          // F.bn1 = GroupNorm(find_divisor(hid), hid), no activation
          if (!measure) {
            const std::string key = p + "F.bn1.";
            const float* g = dev_f32(key + "weight", vec(key + "weight"), stream);
            const float* bta = dev_f32(key + "bias", vec(key + "bias"), stream);
            const int groups = group_norm_divisor(hid), HW = h4 * w4, hid_ = hid;
            Op op;
            op.name = "groupnorm " + key;
            op.fn = [=](cudaStream_t s, const RunCtx&) {
              launch_groupnorm_act(f1raw, DT_F32, f1n, cdt, nullptr, B, HW, hid_, Cp, Cp, groups, g, bta, 1e-5f,
                                   ACT_NONE, s);
            };
            prog.body.push_back(std::move(op));
          }
          add_conv(prog, specs[1], measure, stream, cdt);
          add_conv(prog, specs[2], measure, stream, cdt);
          ppar[j] ^= 1;
          xin = h_act_new;
        }
      }
      // ---- ConvLSTM stack (model_blocks/phydnet.py:147-163) ----
      {
        const void* xin = er_in;
        int cin = 64;
        for (int j = 0; j < n_lstm; ++j) {
          const int hd = d.convlstm_hidden_dims[j];
          const std::string p = "convcell.cell_list." + std::to_string(j) + ".conv.";
          LstmArgs la{p, B, h4, w4, cin, hd, kl, xin, hb[2 * j + lpar[j]], hb[2 * j + (lpar[j] ^ 1)], cb[j],
                      hp(p + "weight"), hp(p + "bias"), true, nullptr, nullptr, nullptr};
          la.c4 = true;
          if (ac && j == 0) {
            la.x2 = act16 + static_cast<size_t>(st) * px4 * a_pad * esz_c;
            la.C2 = a_sz;
            la.C2p = a_pad;
          }
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
      Feat in_p, in_r;
      if (tcfeat) {
        if (!branch_only) split_op(hp_master[n_phy - 1], dec_p, px4 * 64, "split_phy_h");
        split_op(h_top32, dec_r, px4 * 64, "split_lstm_h");
        in_p = dec_p;
        in_r = dec_r;
      } else {
        if (!branch_only) in_p.a = hp_master[n_phy - 1];
        in_r.a = h_top32;
      }
      if (!branch_only) {
        dcgan("decoder_Dp.upc1.", true, in_p, false, h4, w4, 64, 64, 1, OUT_FEAT, mid, nullptr);
        dcgan("decoder_Dp.upc2.", true, mid, false, h4, w4, 64, 64, 1, OUT_F32, Feat{dp, nullptr}, nullptr);
      }
      dcgan("decoder_Dr.upc1.", true, in_r, false, h4, w4, 64, 64, 1, OUT_FEAT, mid, nullptr);
      // concat = decoded_phys + decoded_conv (models/phydnet.py:87) folded into the last GroupNorm pass
      dcgan("decoder_Dr.upc2.", true, mid, false, h4, w4, 64, 64, 1, OUT_FEAT, dsum, branch_only ? nullptr : dp);
      dcgan("decoder_D.upc1.", true, dsum, false, h4, w4, 64, 32, 2, OUT_FEAT, d1, nullptr);
      dcgan("decoder_D.upc2.", true, d1, false, h2, w2, 32, 32, 1, OUT_FEAT, d2, nullptr);
      if (tail_direct) {   // all four output parities + sigmoid + the fed-back 8-channel frame in one direct kernel
        if (!measure) {
          const std::string p = "decoder_D.upc3.";
          TailArgs ta{d2.a, B, h2, w2, 32, c, dev_f32(p + "tail.w", deconv_tail_pack(hp(p + "weight"), 32, c, DT_F16), stream),
                      dev_f32(p + "tail.b", vec(p + "bias"), stream), ACT_SIGMOID,
                      out_stage + static_cast<size_t>(di) * c * h * w, static_cast<long long>(pred) * c * h * w,
                      (di + 1 < pred) ? frame_fb.a : nullptr};
          Op op;
          op.name = p + "tail";
          op.flops = 2.0 * static_cast<double>(px2) * 32 * 9 * c;
          op.fn = [=](cudaStream_t s, const RunCtx&) { launch_deconv_tail(ta, ns, s); };
          prog.body.push_back(std::move(op));
          mark_frame(di);
        }
        continue;
      }
      {
        int oh, ow;
        DeconvArgs a{"decoder_D.upc3.", B, h2, w2, 32, c, 3, 2, 1, 1, d2.a, hp("decoder_D.upc3.weight"),
                     hp("decoder_D.upc3.bias"), ACT_SIGMOID, out_stage + static_cast<size_t>(di) * c * h * w};
        a.nchw = true;
        a.oB_nchw = static_cast<long long>(pred) * c * h * w;
        a.split = split;
        a.x_lo = d2.lo;
        const ActInfo& ai = tcfeat ? sa : f32a;
        add_conv(prog, deconv_spec(a, ai, &oh, &ow), measure, stream, ai.dtype);
        VPK_REQUIRE(oh == h && ow == w, "decoder output size mismatch");
        mark_frame(di);
      }
      if (!measure && di + 1 < pred) {   // next decoder input = output_image (models/phydnet.py:121)
        const float* src = out_stage + static_cast<size_t>(di) * c * h * w;
        const long long bs = static_cast<long long>(pred) * c * h * w;
        Op op;
        op.name = "feedback_frame";
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          if (pad8) launch_frames_to_nhwc8(src, bs, frame_fb.a, frame_fb.lo, sa.dtype, B, 1, c, h, w, ns, s);
          else launch_frames_to_nhwc_strided(src, bs, frame_fb.a, DT_F32, B, 1, c, h, w, ns, s);
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
        if (rc.on_frame != nullptr) return;      // host entry with frame streaming: every frame has been copied already
        VPK_CUDA(cudaMemcpyAsync(rc.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s));
      };
      prog.post.push_back(std::move(post));
    }
  }

 private:
  bool branch_only;
  bool feat_split = false;     // VPK_FEAT_SPLIT=1: split-bf16 feature maps instead of fp16
  bool ac = false;             // action_conditional (models/phydnet.py; model_blocks/phydnet.py:44-55, 153-155)
  int a_sz = 0;
  int n_phy = 1, hid = 49, kp = 7, n_lstm = 3, kl = 3;
};

}  // namespace

Model* make_phydnet(const vpk_model_desc& d, bool branch_only) { return new PhyDNetModel(d, branch_only); }

}  // namespace vpk
