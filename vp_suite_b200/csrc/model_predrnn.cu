// predrnn-pp: PredRNN-V2 rollout, non action-conditional, eval mode (reference: models/predrnn_v2.py:131-230; cell:
// model_blocks/predrnn.py:57-83).  layer_norm=False: three fused tcgen05 launches per cell step (stlstm.h);
// layer_norm=True (model_blocks/predrnn.py:24-40): raw convs + per-sample statistics + fused gate kernels (stlstm_ln.h).
//
// Per step t (total_frames - 1 steps): layer 0 reads the patchified frame x_t (t < context) or the model's own
// x_gen (eval mask is all zero, predrnn_v2.py:172-176, 300-309); the spatio-temporal memory m zig-zags through the
// layers (:196-204); every (t, layer) contributes a decoupling-loss term over adapter(delta_c), adapter(delta_m)
// (:197-211); the 1x1 head gives x_gen (:223); the last `pred` x_gen are un-patchified (:227-228).
#include <cstdlib>

#include "builders.h"
#include "elementwise.h"
#include "model.h"
#include "stlstm.h"
#include "stlstm_ln.h"
#include "stlstm_model.h"

namespace vpk {

namespace {

class PredRnnV2 : public StLstmModelBase {
 public:
  explicit PredRnnV2(const vpk_model_desc& d) : StLstmModelBase(d) {
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
    ac = d.action_conditional != 0;
    if (ac) {
      // action_conditional forces conv_actions_on_input (predrnn_v2.py:65-67): the cells run on a latent a quarter of the
      // patch grid's size (:73-75), reached by two stride-2 convs and left by two stride-2 transposed convs (:76-90)
      VPK_REQUIRE(d.action_size > 0, "action-conditional predrnn-pp needs action_size > 0");
      VPK_REQUIRE(hp_ % 4 == 0 && wp_ % 4 == 0, "action-conditional predrnn-pp needs a patch grid that is a multiple of 4");
      VPK_REQUIRE(C % 2 == 0, "num_hidden must be even");
      rh = hp_ / 4;
      rw = wp_ / 4;
      declare("conv_input1.weight", {C / 2, cp, k, k});
      declare("conv_input2.weight", {C, C / 2, k, k});
      declare("action_conv_input1.weight", {C / 2, d.action_size, k, k});
      declare("action_conv_input2.weight", {C, C / 2, k, k});
      declare("deconv_output1.weight", {C, C / 2, k, k});            // ConvTranspose2d: [Cin, Cout, k, k]
      declare("deconv_output2.weight", {C / 2, cp, k, k});
    } else {
      rh = hp_;
      rw = wp_;
    }
    for (int i = 0; i < L; ++i) {
      const std::string pre = "cell_list." + std::to_string(i) + ".";
      const int cin = (i == 0 && !ac) ? cp : C;
      // ActionConditionalSpatioTemporalLSTMCell's convs have biases and a fifth conv, conv_a (model_blocks/predrnn.py:104-139)
      std::vector<std::pair<std::string, std::pair<int, int>>> convs = {{"conv_x", {7, cin}}, {"conv_h", {4, C}}};
      if (ac) convs.push_back({"conv_a", {4, C}});
      convs.push_back({"conv_m", {3, C}});
      convs.push_back({"conv_o", {1, 2 * C}});
      for (const auto& cv : convs) {
        declare(pre + cv.first + ".0.weight", {cv.second.first * C, cv.second.second, k, k});
        if (ac) declare(pre + cv.first + ".0.bias", {cv.second.first * C});
        if (d.layer_norm) {
          declare(pre + cv.first + ".1.weight", {cv.second.first * C, rh, rw});
          declare(pre + cv.first + ".1.bias", {cv.second.first * C, rh, rw});
        }
      }
      declare(pre + "conv_last.weight", {C, 2 * C, 1, 1});
      if (ac) declare(pre + "conv_last.bias", {C});
    }
    if (!ac) declare("conv_last.weight", {cp, C, 1, 1});             // (non-existent with conv_actions_on_input, :110-116)
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
  // eval mode reads the context frames only (the mask is all zero, predrnn_v2.py:300-309): the target frames that
  // NEEDS_COMPLETE_INPUT puts behind them never have to reach the device
  int used_in_frames(int t_in, int pred) const override { return t_in - pred; }
  int action_steps_needed(int t_in, int) const override { return ac ? t_in - 1 : 0; }
  bool streams_input() const override { return !desc.use_cuda_graph && getenv("VPK_NO_INPUT_STREAM") == nullptr; }

  void begin_call(int, int t_in, int, float*, cudaStream_t stream) override {
    call_terms = (t_in - 1) * L;      // decoupling-loss terms of THIS call (a cached program may serve other lengths)
    if (!d_loss) VPK_CUDA(cudaMalloc(&d_loss, sizeof(double)));
    VPK_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(double), stream));
  }
  void end_call(int batch, float* aux, cudaStream_t stream) override {
    if (aux == nullptr) return;
    // 100 * mean over (t, layer) of mean over (b, ch)   (predrnn_v2.py:209-211, 229-230)
    const double scale = static_cast<double>(desc.decoupling_loss_scale) /
                         (static_cast<double>(call_terms) * batch * C);
    launch_decouple_finalize(d_loss, aux, scale, stream);
  }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    // layer_norm=True in bf16 mode: LayerNorm rescales every conv output to unit variance, which amplifies operand
    // rounding -- bf16 operands give 1.3e-2 on the first frame of the golden case (bound 5e-3).  The LN variant therefore
    // runs its convs on FP16 operands (11 mantissa bits; h in (-1, 1), frames in [0, 1], c / m O(1)), like PhyDNet's
    // GroupNorm-fed convs: all of them write fp32 (raw conv outputs, head, adapter), which is what DT_F16 launches do.
    const int adt = (d.layer_norm != 0 && dtype == DT_BF16) ? DT_F16 : dtype;
    const ActInfo act{adt, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int ctx = t_in - pred;
    const size_t px = static_cast<size_t>(B) * hp_ * wp_;
    if (!measure && !d_loss) VPK_CUDA(cudaMalloc(&d_loss, sizeof(double)));
    if (ac) {
      build_ac(prog, arena, B, t_in, pred, measure, stream);
      return;
    }

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
    // launch O as conv_o + conv_last (stlstm.h: o_raw) once the layer is tensor-bound: two position tiles per SM and more.
    // Bit-identical to the fused launch (tests: full batch == repeated small batch), so the size rule is invisible.
    // VPK_SPLIT_O=0/1 overrides (A/B runs)
    bool split_o = d.layer_norm == 0 && dtype != DT_F32 && backend == 0 && px / 128 >= 2 * static_cast<size_t>(num_sms) &&
                   getenv("VPK_NO_REGIONS") != nullptr;   // superseded by accumulator regions in the fused launch (lowering.cu)
    if (const char* env = getenv("VPK_SPLIT_O")) split_o = d.layer_norm == 0 && atoi(env) != 0;
    float* oraw_split = split_o ? static_cast<float*>(arena.alloc(px * C * sizeof(float))) : nullptr;
    char* dcdm = static_cast<char*>(arena.alloc(2 * px * C * esz));          // [delta_c ; delta_m] stacked on batch
    float* adapt = static_cast<float*>(arena.alloc(2 * px * C * sizeof(float)));
    // layer_norm=True: raw conv outputs (fp32), a dense copy of m for conv_m, statistics slots
    const bool ln = d.layer_norm != 0;
    float *xraw = nullptr, *hraw = nullptr, *mraw = nullptr, *oraw = nullptr, *lraw = nullptr, *lnpart = nullptr;
    void* m_act = nullptr;
    if (ln) {
      xraw = static_cast<float*>(arena.alloc(px * 7 * C * sizeof(float)));
      hraw = static_cast<float*>(arena.alloc(px * 4 * C * sizeof(float)));
      mraw = static_cast<float*>(arena.alloc(px * 3 * C * sizeof(float)));
      oraw = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
      lraw = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
      // statistics partials: (tiles per image) x (N tiles) x 8 slots per tensor and sample when the conv epilogues write
      // them (three regions: conv_x, conv_h, conv_m; conv_o reuses conv_x's), kLnSlices per tensor otherwise
      const size_t slots = std::max<size_t>(3 * kLnSlices, static_cast<size_t>(ln_slots(7 * C)) + ln_slots(4 * C) + ln_slots(3 * C));
      lnpart_floats = static_cast<size_t>(B) * slots * 2;
      lnpart = static_cast<float*>(arena.alloc(lnpart_floats * sizeof(float)));
      m_act = arena.alloc(px * C * esz);
    }
    // layer_norm=True, 16-bit mode: number of fp16 products per conv_x / conv_h / conv_m tap (VPK_LN_PRODUCTS, default 3):
    //   1  A * W                                   plain fp16 operands: 2.4e-2 on the tenth frame of cfg 3's shape
    //   2  A * W_hi + A * W_lo                     split weights: 1.5e-2 on three sequences, 2.1e-2 worst of 256
    //   3  A_hi * W_hi + A_hi * W_lo + A_lo * W_hi  split weights and activations (~22 bits each)
    // LayerNorm renormalises every conv output, so the rollout amplifies operand rounding ~100x over 19 steps (even the
    // fp32 mode is 6e-5 from the reference at the end): only split weights everywhere AND split activations for conv_x
    // keep every sequence of a 256-sequence batch inside north_star's 2e-2 (conv_h / conv_m run form 2, see
    // stlstm_model.h).  The low parts of x / h are written by their producers.
    int ln_products = 1;
    if (ln && adt == DT_F16) {
      ln_products = 3;
      if (const char* env = getenv("VPK_LN_PRODUCTS")) ln_products = std::max(1, std::min(3, atoi(env)));
      if (((C + 63) / 64) * ln_products * k * k > kMaxSteps) ln_products = 1;      // step-table capacity
    }
    const bool lo3 = ln_products == 3;
    char* xp_lo = lo3 ? static_cast<char*>(arena.alloc(px * cp * esz * ctx)) : nullptr;
    float* xp32 = lo3 ? static_cast<float*>(arena.alloc(px * cp * sizeof(float))) : nullptr;
    std::vector<void*> hb_lo(2 * L, nullptr);
    void* xgen_lo = nullptr;
    if (lo3) {
      for (int i = 0; i < 2 * L; ++i) hb_lo[i] = arena.alloc(px * C * esz);
      xgen_lo = arena.alloc(px * cp * esz);
    }
    // fused decoupling loss (tcgen05 path): per-warp partial slots + one term per (step, layer, sample)
    const char* halo_env = getenv("VPK_TC_HALO");
    const bool fuse_dec = adt != DT_F32 && backend == 0 && C % 8 == 0 && C <= 128 && getenv("VPK_NO_FUSED_DECOUPLE") == nullptr &&
                          (halo_env == nullptr || atoi(halo_env) != 0);
    const int dec_nslots = 4 * ((hp_ + 15) / 16) * ((wp_ + 7) / 8);
    float* dec_slots = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * dec_nslots * C * 3 * sizeof(float)));
    float* dec_terms = static_cast<float*>(arena.alloc(static_cast<size_t>(t_in - 1) * L * B * sizeof(float)));
    float* xgen32 = static_cast<float*>(arena.alloc(px * cp * sizeof(float)));
    void* xgen_act = (dtype == DT_F32) ? static_cast<void*>(xgen32) : arena.alloc(px * cp * esz);

    if (!measure) {
      const int ns = num_sms, dt = adt, pp = p;
      if (!streams_input()) {     // CUDA-graph replay: the conversion reads the call's input pointer, so it stays outside
        Op pre;
        pre.name = "patchify";
        // x holds t_in frames per sequence; frames >= ctx are ignored
        const size_t fbytes = px * cp * esz;
        const long long fn_ = static_cast<long long>(px) * cp;
        pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
          if (!lo3) {
            launch_patchify_strided(rc.x, static_cast<long long>(t_in) * c * h * w, xp, dt, B, ctx, c, h, w, pp, ns, s);
            return;
          }
          for (int f = 0; f < ctx; ++f) {      // fp32 patches of one frame, then their (hi, lo) fp16 split
            launch_patchify_strided(rc.x + static_cast<long long>(f) * c * h * w, static_cast<long long>(t_in) * c * h * w, xp32,
                                    DT_F32, B, 1, c, h, w, pp, ns, s);
            launch_split_f16(xp32, xp + f * fbytes, xp_lo + f * fbytes, fn_, ns, s);
          }
        };
        prog.pre.push_back(std::move(pre));
      }
      for (int i = 0; i < L; ++i) {
        add_memset(prog, hb[2 * i], px * C * esz, "zero_h");
        add_memset(prog, cb[i], px * C * sizeof(float), "zero_c");
      }
      add_memset(prog, mstate, px * C * sizeof(float), "zero_m");
      add_memset(prog, memb[2 * (L - 1) + 1], px * 2 * C * esz, "zero_mem");   // m seen by layer 0 at t = 0
      if (ln) add_memset(prog, m_act, px * C * esz, "zero_m_act");
      if (lo3) {
        for (int i = 0; i < L; ++i) add_memset(prog, hb_lo[2 * i], px * C * esz, "zero_h_lo");
      }
    }

    std::vector<int> par(L, 0);
    for (int t = 0; t < t_in - 1; ++t) {
      const void* net = (t < ctx) ? static_cast<const void*>(xp + static_cast<size_t>(t) * px * cp * esz) : xgen_act;
      if (!measure && t < ctx && streams_input()) {
        // context frame t is patchified right before the step that reads it; under the host entry this op waits for
        // the frame's own host-to-device copy only
        const int ns = num_sms, dt = adt, pp = p;
        char* dst = xp + static_cast<size_t>(t) * px * cp * esz;
        const long long bstride = static_cast<long long>(t_in) * c * h * w, foff = static_cast<long long>(t) * c * h * w;
        Op cv;
        cv.name = "patchify";
        cv.needs_input = t;
        char* dst_lo = lo3 ? xp_lo + static_cast<size_t>(t) * px * cp * esz : nullptr;
        const long long fn_ = static_cast<long long>(px) * cp;
        cv.fn = [=](cudaStream_t s, const RunCtx& rc) {
          if (!lo3) {
            launch_patchify_strided(rc.x + foff, bstride, dst, dt, B, 1, c, h, w, pp, ns, s);
          } else {
            launch_patchify_strided(rc.x + foff, bstride, xp32, DT_F32, B, 1, c, h, w, pp, ns, s);
            launch_split_f16(xp32, dst, dst_lo, fn_, ns, s);
          }
        };
        prog.body.push_back(std::move(cv));
      }
      for (int i = 0; i < L; ++i) {
        const std::string pre = "cell_list." + std::to_string(i) + ".";
        const void* inp = (i == 0) ? net : hb[2 * (i - 1) + par[i - 1]];
        LnLo lo{};
        if (lo3) {
          lo.x = (i == 0) ? ((t < ctx) ? static_cast<const void*>(xp_lo + static_cast<size_t>(t) * px * cp * esz) : xgen_lo)
                          : hb_lo[2 * (i - 1) + par[i - 1]];
          lo.h_in = hb_lo[2 * i + par[i]];
          lo.h_out = hb_lo[2 * i + (par[i] ^ 1)];
          lo.m_act = nullptr;      // conv_m runs two products (split weights): no low part of m needed
        }
        const int cin = (i == 0) ? cp : C;
        // memory comes from the previous layer of this step, or from the top layer of the previous step
        const void* mem_prev = (i == 0) ? memb[2 * (L - 1) + ((t + 1) & 1)] : memb[2 * (i - 1) + (t & 1)];
        StLstmArgs a{pre, B, hp_, wp_, cin, C, k, inp, hb[2 * i + par[i]],
                     make_channel_view(mem_prev, hp_, wp_, 2 * C, C, C, esz), hb[2 * i + (par[i] ^ 1)], cb[i], mstate,
                     opart, memb[2 * i + (t & 1)], dcdm, dcdm + px * C * esz,
                     hp(pre + "conv_x.0.weight"), hp(pre + "conv_h.0.weight"), hp(pre + "conv_m.0.weight"),
                     hp(pre + "conv_o.0.weight"), hp(pre + "conv_last.weight")};
        a.c4 = true;
        a.o_raw = oraw_split;
        if (!ln) {
          for (const ConvSpec& sp : stlstm_specs(a, act)) add_conv(prog, sp, measure, stream, adt);
        } else {
          add_ln_cell(prog, pre, B, cin, inp, hb[2 * i + par[i]], hb[2 * i + (par[i] ^ 1)], cb[i], mstate, opart,
                      memb[2 * i + (t & 1)], m_act, dcdm, dcdm + px * C * esz, xraw, hraw, mraw, oraw, lraw, lnpart, act,
                      measure, stream, ln_products, lo);
        }
        par[i] ^= 1;
        // decoupling-loss term: adapter (1x1, no bias) over delta_c and delta_m, then the per-(b, ch) cosine.
        // tcgen05 path: ONE conv with two gate columns per channel (adapter(delta_c), adapter(delta_m)) whose epilogue
        // reduces dot / norms over the positions -- the adapter outputs (2 x px x C fp32) never reach memory.
        if (fuse_dec) {
          ConvSpec sp;
          sp.name = "adapter.decouple";
          sp.B = B;
          sp.G = 2;
          sp.C = C;
          for (int g = 0; g < 2; ++g) {
            WeightRef wr;
            wr.w = hp("adapter.weight");
            wr.O = C;
            wr.I = C;
            wr.KH = wr.KW = 1;
            for (int q = 0; q < 4; ++q) wr.gate_block[q] = -1;
            wr.gate_block[g] = 0;
            sp.wrefs.push_back(wr);
          }
          int oh2, ow2;
          std::vector<ConvInput> ins;
          ins.push_back(ConvInput{dense_view(dcdm, hp_, wp_, C), 0, 0});
          ins.push_back(ConvInput{dense_view(dcdm + px * C * esz, hp_, wp_, C), 1, 0});
          lower_conv(sp, 1, 1, 0, ins, hp_, wp_, esz, &oh2, &ow2);
          EpiParams& e = sp.phases[0].epi;
          e.kind = EPI_DECOUPLE;
          e.s1 = dec_slots;
          e.gn_slot0 = 0;
          e.gn_nslots = dec_nslots;
          add_conv(prog, sp, measure, stream, adt);
          if (!measure) {
            const int CC = C, nsl = dec_nslots;
            float* term = dec_terms + static_cast<size_t>(t * L + i) * B;
            Op op;
            op.name = "decouple_cos";
            op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_cos(dec_slots, nsl, B, CC, term, s); };
            prog.body.push_back(std::move(op));
          }
          continue;
        }
        int oh, ow;
        ConvArgs ad{"adapter.", 2 * B, hp_, wp_, C, C, 1, 1, 0, dcdm, hp("adapter.weight"), nullptr, ACT_NONE, adapt};
        ad.f32_strided = true;
        ad.oB = static_cast<long long>(hp_) * wp_ * C;
        ad.oY = static_cast<long long>(wp_) * C;
        ad.oX = C;
        ad.oC = 1;
        add_conv(prog, conv_spec(ad, act, &oh, &ow), measure, stream, adt);
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
      add_conv(prog, conv_spec(hd, act, &oh, &ow), measure, stream, adt);
      if (!measure) {
        const int ns = num_sms;
        if (dtype != DT_F32 && t + 1 >= ctx && t + 1 < t_in - 1) {
          const long long n = static_cast<long long>(px) * cp;
          Op op;
          op.name = "cast_xgen";
          op.fn = [=](cudaStream_t s, const RunCtx&) {
            if (lo3) launch_split_f16(xgen32, xgen_act, xgen_lo, n, ns, s);
            else if (adt == DT_F16) launch_cast_f32_to_f16(xgen32, xgen_act, n, ns, s);
            else launch_cast_f32_to_bf16(xgen32, xgen_act, n, ns, s);
          };
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
          op.frame = fo;       // completes predicted frame fo (host entry: its D2H starts here)
          op.frame_src = out_stage + static_cast<size_t>(fo) * c * h * w;
          op.frame_pitch = static_cast<long long>(pred) * c * h * w;
          op.frame_elems = static_cast<long long>(c) * h * w;
          prog.body.push_back(std::move(op));
        }
      }
    }
    if (!measure && fuse_dec) {
      const long long n = static_cast<long long>(t_in - 1) * L * B;
      double* acc = d_loss;
      Op op;
      op.name = "decouple_sum";
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_sum(dec_terms, n, acc, s); };
      prog.body.push_back(std::move(op));
    }
    if (!measure) {
      const size_t bytes = static_cast<size_t>(B) * pred * c * h * w * sizeof(float);
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

  // ==================================================================================================================
  // action_conditional=True (predrnn_v2.py:65-90, 147-152, 178-221; cell: model_blocks/predrnn.py:86-169).
  //
  // Per step t: the patch frame (real for t < context, the model's x_gen afterwards -- the reverse-sampling eval mask is 1
  // exactly on the context steps, :306-308) goes through conv_input1 / conv_input2 (k x k, stride 2, no bias, no
  // activation) down to the (patch_h / 4) x (patch_w / 4) latent; the action vector, inflated to the patch grid (:151),
  // through action_conv_input1 / 2 likewise -- all steps' actions are known up front, so those two convs run ONCE, time-
  // batched over (t_in - 1) * B samples, in the program's pre ops.  Cells: raw convs (bias in the epilogue; fp32 outputs,
  // per-sample LayerNorm statistics from the epilogue when layer_norm) + the action-conditional gate kernel + conv_o /
  // conv_last + the output kernel.  x_gen = deconv_output2(deconv_output1(h_top + net2) + net1) (:213-215, residuals
  // optional).  16-bit mode uses FP16 operands throughout (raw fp32 conv outputs, like the LayerNorm rollout).
  // ==================================================================================================================
  void build_ac(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) {
    const vpk_model_desc& d = desc;
    const bool ln = d.layer_norm != 0;
    const bool resid = d.residual_on_action_conv != 0;
    const int adt = (dtype == DT_BF16) ? DT_F16 : dtype;
    const ActInfo act{adt, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w, a_sz = d.action_size;
    const int ctx = t_in - pred, steps = t_in - 1;
    const int h2 = hp_ / 2, w2 = wp_ / 2;
    const int a_pad = (a_sz + 7) / 8 * 8;
    const size_t px = static_cast<size_t>(B) * hp_ * wp_, px2 = static_cast<size_t>(B) * h2 * w2,
                 px4 = static_cast<size_t>(B) * rh * rw;
    const int ns = num_sms;
    const bool f32 = adt == DT_F32;
    // two fp16 products (split weights) for the LayerNorm-fed convs in 16-bit mode, as VPK_LN_PRODUCTS=2 of the plain rollout
    int products = (ln && !f32) ? 2 : 1;
    if (const char* env = getenv("VPK_LN_PRODUCTS")) products = std::min(products, std::max(1, atoi(env)));
    if (((C + 63) / 64) * products * k * k > kMaxSteps) products = 1;

    char* xp = static_cast<char*>(arena.alloc(px * cp * esz * ctx));
    float* out_stage = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * pred * c * h * w * sizeof(float)));
    // inflated actions and their two convs, all steps at once: [steps][B][...]
    void* a_inf = arena.alloc(px * a_pad * esz * steps);
    float* a1_32 = static_cast<float*>(arena.alloc(px2 * (C / 2) * sizeof(float) * steps));
    void* a1 = f32 ? static_cast<void*>(a1_32) : arena.alloc(px2 * (C / 2) * esz * steps);
    float* a2_32 = static_cast<float*>(arena.alloc(px4 * C * sizeof(float) * steps));
    char* a2 = f32 ? reinterpret_cast<char*>(a2_32) : static_cast<char*>(arena.alloc(px4 * C * esz * steps));
    // per-step input path
    float* n1_32 = static_cast<float*>(arena.alloc(px2 * (C / 2) * sizeof(float)));
    void* n1 = f32 ? static_cast<void*>(n1_32) : arena.alloc(px2 * (C / 2) * esz);
    float* n2_32 = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    void* n2 = f32 ? static_cast<void*>(n2_32) : arena.alloc(px4 * C * esz);
    // cells
    std::vector<void*> hb(2 * L), memb(2 * L);
    std::vector<float*> cb(L);
    for (int i = 0; i < L; ++i) {
      hb[2 * i] = arena.alloc(px4 * C * esz);
      hb[2 * i + 1] = arena.alloc(px4 * C * esz);
      memb[2 * i] = arena.alloc(px4 * 2 * C * esz);
      memb[2 * i + 1] = arena.alloc(px4 * 2 * C * esz);
      cb[i] = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    }
    float* mstate = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    float* opart = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    char* dcdm = static_cast<char*>(arena.alloc(2 * px4 * C * esz));
    float* adapt = static_cast<float*>(arena.alloc(2 * px4 * C * sizeof(float)));
    float* xraw = static_cast<float*>(arena.alloc(px4 * 7 * C * sizeof(float)));
    float* hraw = static_cast<float*>(arena.alloc(px4 * 4 * C * sizeof(float)));
    float* araw = static_cast<float*>(arena.alloc(px4 * 4 * C * sizeof(float)));
    float* mraw = static_cast<float*>(arena.alloc(px4 * 3 * C * sizeof(float)));
    float* oraw = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    float* lraw = static_cast<float*>(arena.alloc(px4 * C * sizeof(float)));
    const size_t slots = std::max<size_t>(4 * kLnSlices, static_cast<size_t>(ln_slots(7 * C)) + 2 * ln_slots(4 * C) + ln_slots(3 * C));
    lnpart_floats = static_cast<size_t>(B) * slots * 2;
    float* lnpart = static_cast<float*>(arena.alloc(lnpart_floats * sizeof(float)));
    void* m_act = arena.alloc(px4 * C * esz);
    // output path
    void* s1 = arena.alloc(px4 * C * esz);                          // h_top (+ net2)
    float* g32 = static_cast<float*>(arena.alloc(px2 * (C / 2) * sizeof(float)));
    void* s2 = f32 ? static_cast<void*>(g32) : arena.alloc(px2 * (C / 2) * esz);   // deconv_output1 (+ net1)
    float* xgen32 = static_cast<float*>(arena.alloc(px * cp * sizeof(float)));
    void* xgen_act = f32 ? static_cast<void*>(xgen32) : arena.alloc(px * cp * esz);
    // decoupling loss
    const char* halo_env = getenv("VPK_TC_HALO");
    const bool fuse_dec = !f32 && backend == 0 && C % 8 == 0 && C <= 128 && getenv("VPK_NO_FUSED_DECOUPLE") == nullptr &&
                          (halo_env == nullptr || atoi(halo_env) != 0);
    const int dec_nslots = 4 * ((rh + 15) / 16) * ((rw + 7) / 8);
    float* dec_slots = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * dec_nslots * C * 3 * sizeof(float)));
    float* dec_terms = static_cast<float*>(arena.alloc(static_cast<size_t>(steps) * L * B * sizeof(float)));

    auto cast_op = [&](std::vector<Op>& dst, const float* src, void* out, size_t n, const char* name) {
      if (measure || f32) return;
      Op op;
      op.name = name;
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(src, DT_F32, nullptr, out, adt, static_cast<long long>(n), ns, s); };
      dst.push_back(std::move(op));
    };
    // a k x k stride-2 conv without bias / activation: fp32 dense output (+ its activation-type copy)
    auto down_conv = [&](std::vector<Op>* dst, const std::string& key, int Bx, int H, int W, int ci, int co, const void* in,
                         float* out32, void* out_act, int cin_w) {
      int oh, ow;
      ConvArgs a{key + ".", Bx, H, W, ci, co, k, 2, k / 2, in, hp(key + ".weight"), nullptr, ACT_NONE, out32};
      a.out_f32_dense = true;
      a.cin_w = cin_w;
      add_conv(prog, conv_spec(a, act, &oh, &ow), measure, stream, adt, dst);
      VPK_REQUIRE(oh == H / 2 && ow == W / 2, "stride-2 input conv size mismatch");
      cast_op(dst != nullptr ? *dst : prog.body, out32, out_act, static_cast<size_t>(Bx) * oh * ow * co, "cast_down_conv");
    };

    // ---- pre ops: patchify (device entry), actions -> latent for every step ----
    if (!measure) {
      if (!streams_input()) {
        Op pre;
        pre.name = "patchify";
        pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
          launch_patchify_strided(rc.x, static_cast<long long>(t_in) * c * h * w, xp, adt, B, ctx, c, h, w, p, ns, s);
        };
        prog.pre.push_back(std::move(pre));
      }
      Op inf;
      inf.name = "inflate_actions";
      const int HWp = hp_ * wp_;
      inf.fn = [=](cudaStream_t s, const RunCtx& rc) {
        VPK_REQUIRE(rc.actions != nullptr && rc.action_steps >= steps, "Given actions are None or of the wrong size!");
        launch_inflate_actions(rc.actions, static_cast<long long>(rc.action_steps) * a_sz, a_sz, a_inf, adt, B, steps, HWp, a_pad, ns, s);
      };
      prog.pre.push_back(std::move(inf));
    }
    down_conv(&prog.pre, "action_conv_input1", B * steps, hp_, wp_, a_pad, C / 2, a_inf, a1_32, a1, a_sz);
    down_conv(&prog.pre, "action_conv_input2", B * steps, h2, w2, C / 2, C, a1, a2_32, a2, -1);
    if (!measure) {
      for (int i = 0; i < L; ++i) {
        add_memset(prog, hb[2 * i], px4 * C * esz, "zero_h");
        add_memset(prog, cb[i], px4 * C * sizeof(float), "zero_c");
      }
      add_memset(prog, mstate, px4 * C * sizeof(float), "zero_m");
      add_memset(prog, m_act, px4 * C * esz, "zero_m_act");
    }

    std::vector<int> par(L, 0);
    for (int t = 0; t < steps; ++t) {
      const void* net = (t < ctx) ? static_cast<const void*>(xp + static_cast<size_t>(t) * px * cp * esz) : xgen_act;
      if (!measure && t < ctx && streams_input()) {
        char* dst = xp + static_cast<size_t>(t) * px * cp * esz;
        const long long bstride = static_cast<long long>(t_in) * c * h * w, foff = static_cast<long long>(t) * c * h * w;
        const int pp = p;
        Op cv;
        cv.name = "patchify";
        cv.needs_input = t;
        cv.fn = [=](cudaStream_t s, const RunCtx& rc) { launch_patchify_strided(rc.x + foff, bstride, dst, adt, B, 1, c, h, w, pp, ns, s); };
        prog.body.push_back(std::move(cv));
      }
      down_conv(nullptr, "conv_input1", B, hp_, wp_, cp, C / 2, net, n1_32, n1, -1);
      down_conv(nullptr, "conv_input2", B, h2, w2, C / 2, C, n1, n2_32, n2, -1);
      const void* action = a2 + static_cast<size_t>(t) * px4 * C * (f32 ? 4 : esz);
      for (int i = 0; i < L; ++i) {
        const std::string pre = "cell_list." + std::to_string(i) + ".";
        const void* inp = (i == 0) ? static_cast<const void*>(n2) : hb[2 * (i - 1) + par[i - 1]];
        add_ac_cell(prog, pre, B, inp, hb[2 * i + par[i]], hb[2 * i + (par[i] ^ 1)], cb[i], mstate, opart, memb[2 * i + (t & 1)],
                    m_act, dcdm, dcdm + px4 * C * esz, action, xraw, hraw, araw, mraw, oraw, lraw, lnpart, act, measure,
                    stream, ln, products);
        par[i] ^= 1;
        add_decouple(prog, B, t, i, dcdm, px4, adapt, dec_slots, dec_nslots, dec_terms, fuse_dec, act, measure, stream);
      }
      // ---- x_gen (predrnn_v2.py:213-218) ----
      const void* top = hb[2 * (L - 1) + par[L - 1]];
      const void* din = top;
      if (resid && !measure) {
        const long long n = static_cast<long long>(px4) * C;
        Op op;
        op.name = "residual_net2";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(top, adt, n2_32, s1, adt, n, ns, s); };
        prog.body.push_back(std::move(op));
      }
      if (resid) din = s1;
      int oh, ow;
      DeconvArgs d1{"deconv_output1.", B, rh, rw, C, C / 2, k, 2, k / 2, 1, din, hp("deconv_output1.weight"), nullptr, ACT_NONE, g32};
      d1.out_f32 = true;
      add_conv(prog, deconv_spec(d1, act, &oh, &ow), measure, stream, adt);
      VPK_REQUIRE(oh == h2 && ow == w2, "deconv_output1 size mismatch");
      if (!measure && (resid || !f32)) {
        const long long n = static_cast<long long>(px2) * (C / 2);
        const float* add = resid ? n1_32 : nullptr;
        Op op;
        op.name = "residual_net1";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(g32, DT_F32, add, s2, adt, n, ns, s); };
        prog.body.push_back(std::move(op));
      }
      DeconvArgs d2{"deconv_output2.", B, h2, w2, C / 2, cp, k, 2, k / 2, 1, s2, hp("deconv_output2.weight"), nullptr, ACT_NONE, xgen32};
      d2.out_f32 = true;
      add_conv(prog, deconv_spec(d2, act, &oh, &ow), measure, stream, adt);
      VPK_REQUIRE(oh == hp_ && ow == wp_, "deconv_output2 size mismatch");
      if (t + 1 >= ctx && t + 1 < steps) cast_op(prog.body, xgen32, xgen_act, px * cp, "cast_xgen");
      if (!measure) {
        const int first_out = steps - pred;
        if (t >= first_out) {
          const int fo = t - first_out, pp = p;
          Op op;
          op.name = "unpatchify";
          op.fn = [=](cudaStream_t s, const RunCtx&) { launch_unpatchify(xgen32, out_stage, DT_F32, B, pred, fo, c, h, w, pp, ns, s); };
          op.frame = fo;
          op.frame_src = out_stage + static_cast<size_t>(fo) * c * h * w;
          op.frame_pitch = static_cast<long long>(pred) * c * h * w;
          op.frame_elems = static_cast<long long>(c) * h * w;
          prog.body.push_back(std::move(op));
        }
      }
    }
    if (!measure && fuse_dec) {
      const long long n = static_cast<long long>(steps) * L * B;
      double* acc = d_loss;
      Op op;
      op.name = "decouple_sum";
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_sum(dec_terms, n, acc, s); };
      prog.body.push_back(std::move(op));
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

  // decoupling-loss term of (step t, layer i) over delta_c / delta_m (predrnn_v2.py:197-211) on the cells' latent grid
  void add_decouple(Program& prog, int B, int t, int i, char* dcdm, size_t pxl, float* adapt, float* dec_slots, int dec_nslots,
                    float* dec_terms, bool fuse_dec, const ActInfo& act, bool measure, cudaStream_t stream) {
    const int esz = act.esize;
    if (fuse_dec) {
      ConvSpec sp;
      sp.name = "adapter.decouple";
      sp.B = B;
      sp.G = 2;
      sp.C = C;
      for (int g = 0; g < 2; ++g) {
        WeightRef wr;
        wr.w = hp("adapter.weight");
        wr.O = C;
        wr.I = C;
        wr.KH = wr.KW = 1;
        for (int q = 0; q < 4; ++q) wr.gate_block[q] = -1;
        wr.gate_block[g] = 0;
        sp.wrefs.push_back(wr);
      }
      int oh2, ow2;
      std::vector<ConvInput> ins;
      ins.push_back(ConvInput{dense_view(dcdm, rh, rw, C), 0, 0});
      ins.push_back(ConvInput{dense_view(dcdm + pxl * C * esz, rh, rw, C), 1, 0});
      lower_conv(sp, 1, 1, 0, ins, rh, rw, esz, &oh2, &ow2);
      EpiParams& e = sp.phases[0].epi;
      e.kind = EPI_DECOUPLE;
      e.s1 = dec_slots;
      e.gn_slot0 = 0;
      e.gn_nslots = dec_nslots;
      add_conv(prog, sp, measure, stream, act.dtype);
      if (!measure) {
        const int CC = C, nsl = dec_nslots;
        float* term = dec_terms + static_cast<size_t>(t * L + i) * B;
        Op op;
        op.name = "decouple_cos";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_cos(dec_slots, nsl, B, CC, term, s); };
        prog.body.push_back(std::move(op));
      }
      return;
    }
    int oh, ow;
    ConvArgs ad{"adapter.", 2 * B, rh, rw, C, C, 1, 1, 0, dcdm, hp("adapter.weight"), nullptr, ACT_NONE, adapt};
    ad.f32_strided = true;
    ad.oB = static_cast<long long>(rh) * rw * C;
    ad.oY = static_cast<long long>(rw) * C;
    ad.oX = C;
    ad.oC = 1;
    add_conv(prog, conv_spec(ad, act, &oh, &ow), measure, stream, act.dtype);
    if (!measure) {
      const int HW = rh * rw, CC = C;
      double* acc = d_loss;
      Op op;
      op.name = "decouple_reduce";
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_decouple_reduce(adapt, B, HW, CC, acc, s); };
      prog.body.push_back(std::move(op));
    }
  }

  // One ActionConditionalSpatioTemporalLSTMCell step (model_blocks/predrnn.py:142-169): raw convs with bias (x, h, a, m),
  // optional per-sample LayerNorm statistics, the action-conditional gate kernel, conv_o / conv_last, the output kernel.

 private:
  int p = 4, L = 3, cp = 16, hp_ = 16, wp_ = 16;
  bool ac = false;
  int call_terms = 1;
  double* d_loss = nullptr;
};

}  // namespace

Model* make_predrnn(const vpk_model_desc& d) { return new PredRnnV2(d); }

}  // namespace vpk
