// st-phy: ST-Phy rollout, eval mode, with and without actions (reference: models/st_phy.py:90-181).
// A hybrid of the two families already on the kernel path: per layer one PhyCell_Cell (model_blocks/phydnet.py:49-62,
// phycell.h) and one SpatioTemporalLSTMCell with layer_norm=True (model_blocks/predrnn.py:24-40, 57-83; the raw-conv +
// statistics + fused-gate pipeline of stlstm_model.h), merged by a 1x1 conv over cat[st_h, phy_h]; an Autoencoder of
// unpadded convs around them (model_blocks/enc.py:14-98).
//
// Per step t (context + pred - 1 steps): next_input = encode(x_t) for the context steps, else the previous x_gen.  EVERY
// layer reads that same next_input (the reference does not update it inside the layer loop, st_phy.py:139-158); layer i
// advances phy_h[i] and (h[i], c[i], shared st_memory) and overwrites x_gen with its merge, so the last layer's merge is the
// step's x_gen.  Frames are decoded from t = context - 1 on.  Losses are training-only: forward returns (frames, None).
//
// action_conditional (st_phy.py:48-56, 142-150): the step's action vector goes through a bias-free Linear to an
// [inflated_action_dim, enc_h, enc_w] map; the sum of a (5,1) and a (1,5) conv of it is the action tensor of every layer's
// ActionConditionalSpatioTemporalLSTMCell (layer_norm=True; stlstm_model.h: add_ac_cell) -- computed for all steps in one
// pre-op launch (it depends on the call's actions only) -- and every PhyCell first sends cat[frame, action] and
// cat[hidden, action] through its own 1x1 convs (model_blocks/phydnet.py:50-55; fp32 CUDA-core convs as in model_phydnet.cu).
//
// Operand types in 16-bit mode: the LayerNorm ST-LSTM convs run on fp16 (three products, stlstm_model.h), the PhyCell on
// bf16, the autoencoder convs on FP16 with fp32 outputs (the encoder ends in an L2 normalisation along W of ReLU outputs:
// rows with a tiny norm blow bf16 operand rounding up to 3e-2 in the frames -- CPU emulation: bf16 encoder 3.2e-2, fp16
// 9e-4), the 1x1 merge on fp32 CUDA cores straight from the fp32 h' / PhyCell state; next_input is kept as one fp32 tensor
// fanned out into the operand copies each consumer needs.
#include <cmath>
#include <cstdlib>

#include "builders.h"
#include "elementwise.h"
#include "phycell.h"
#include "stlstm_model.h"

namespace vpk {

namespace {

int stphy_group_norm_divisor(int x) {   // model_blocks/phydnet.py:348-362
  int sq = static_cast<int>(std::floor(std::sqrt(static_cast<double>(x))));
  while (x % sq != 0) --sq;
  return x / sq;
}

class StPhyModel : public StLstmModelBase {
 public:
  explicit StPhyModel(const vpk_model_desc& d) : StLstmModelBase(d) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h > 0 && d.img_w > 0, "bad img_shape");
    ac = d.action_conditional != 0;
    a_sz = ac ? d.action_size : 0;
    ia = d.inflated_action_dim > 0 ? d.inflated_action_dim : 3;
    VPK_REQUIRE(!ac || (a_sz > 0 && a_sz <= 8), "action-conditional st-phy needs 1 <= action_size <= 8");
    L = d.num_layers;
    C = d.num_hidden[0];                 // st_cell_channels
    k = 5;                               // the ST cells are built with filter_size=5 (st_phy.py:61)
    hid = d.phycell_channels;
    kp = d.phycell_kernel_size;
    VPK_REQUIRE(L >= 1 && L <= 8 && C > 0 && C % 8 == 0 && hid > 0 && kp % 2 == 1, "bad st-phy hyper-parameters");
    // encoder: k5 s2, k3 s2, k3 s1, all unpadded (enc.py:60-62)
    VPK_REQUIRE(d.img_h % 2 == 0 && d.img_w % 2 == 0, "st-phy: image size must be even");
    h1 = (d.img_h - 5) / 2 + 1;
    w1 = (d.img_w - 5) / 2 + 1;
    VPK_REQUIRE(h1 % 2 == 0 && w1 % 2 == 0, "st-phy: the first encoder map must have even size (stride-2 conv on parity views)");
    h2 = (h1 - 3) / 2 + 1;
    w2 = (w1 - 3) / 2 + 1;
    rh = h2 - 2;
    rw = w2 - 2;
    VPK_REQUIRE(rh > 0 && rw > 0, "st-phy: image too small");
    // decoder: k6 s2, k6 s2, k5 s1 transposed, unpadded (enc.py:87-89); Resize must be the identity
    d1h = (rh - 1) * 2 + 6;
    d1w = (rw - 1) * 2 + 6;
    d2h = (d1h - 1) * 2 + 6;
    d2w = (d1w - 1) * 2 + 6;
    VPK_REQUIRE(d2h + 4 == d.img_h && d2w + 4 == d.img_w,
                "st-phy: the decoder does not land on the image size (other sizes need the reference's Resize)");
    const int c = d.img_c;
    declare("autoencoder.encoder.conv1.weight", {32, c, 5, 5});
    declare("autoencoder.encoder.conv1.bias", {32});
    declare("autoencoder.encoder.conv2.weight", {64, 32, 3, 3});
    declare("autoencoder.encoder.conv2.bias", {64});
    declare("autoencoder.encoder.mean_layer.weight", {C, 64, 3, 3});
    declare("autoencoder.encoder.mean_layer.bias", {C});
    declare("autoencoder.decoder.fc1.weight", {C, C, 1, 1});
    declare("autoencoder.decoder.fc1.bias", {C});
    declare("autoencoder.decoder.conv1.weight", {C, 64, 6, 6});
    declare("autoencoder.decoder.conv1.bias", {64});
    declare("autoencoder.decoder.conv2.weight", {64, 32, 6, 6});
    declare("autoencoder.decoder.conv2.bias", {32});
    declare("autoencoder.decoder.conv3.weight", {32, c, 5, 5});
    declare("autoencoder.decoder.conv3.bias", {c});
    if (ac) {
      declare("action_inflate.weight", {ia * rh * rw, a_sz});
      declare("action_conv_h.weight", {C, ia, 5, 1});
      declare("action_conv_w.weight", {C, ia, 1, 5});
    }
    for (int i = 0; i < L; ++i) {
      const std::string s = "st_cell_list." + std::to_string(i) + ".";
      const std::pair<const char*, std::pair<int, int>> convs[5] = {{"conv_x", {7, C}}, {"conv_h", {4, C}}, {"conv_m", {3, C}},
                                                                    {"conv_o", {1, 2 * C}}, {"conv_a", {4, C}}};
      for (int q = 0; q < (ac ? 5 : 4); ++q) {
        const auto& cv = convs[q];
        declare(s + cv.first + ".0.weight", {cv.second.first * C, cv.second.second, k, k});
        if (ac) declare(s + cv.first + ".0.bias", {cv.second.first * C});
        declare(s + cv.first + ".1.weight", {cv.second.first * C, rh, rw});
        declare(s + cv.first + ".1.bias", {cv.second.first * C, rh, rw});
      }
      declare(s + "conv_last.weight", {C, 2 * C, 1, 1});
      if (ac) declare(s + "conv_last.bias", {C});
      const std::string p = "phycell_list." + std::to_string(i) + ".";
      declare(p + "F.conv1.weight", {hid, C, kp, kp});
      declare(p + "F.conv1.bias", {hid});
      declare(p + "F.bn1.weight", {hid});
      declare(p + "F.bn1.bias", {hid});
      declare(p + "F.conv2.weight", {C, hid, 1, 1});
      declare(p + "F.conv2.bias", {C});
      declare(p + "convgate.weight", {C, 2 * C, 3, 3});
      declare(p + "convgate.bias", {C});
      if (ac) {
        declare(p + "frame_action_conv.weight", {C, C + a_sz, 1, 1});
        declare(p + "frame_action_conv.bias", {C});
        declare(p + "hidden_action_conv.weight", {C, C + a_sz, 1, 1});
        declare(p + "hidden_action_conv.bias", {C});
      }
      declare("hidden_conv_list." + std::to_string(i) + ".weight", {C, 2 * C, 1, 1});
      if (i < L - 1) declare("hidden_conv_list." + std::to_string(i) + ".bias", {C});     // st_phy.py:68-70
    }
    declare("adapter.weight", {C, C, 1, 1});       // training-only (decoupling loss); kept for the state_dict layout
  }

 protected:
  int default_microbatch() const override { return 128; }
  // one action per step (st_phy.py:96-103, 142)
  int action_steps_needed(int t_in, int pred) const override { return ac ? t_in + pred - 1 : 0; }

  std::vector<float> vec(const std::string& key) const { return params.at(key).data; }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const int cdt = dtype;                                          // PhyCell / autoencoder operand type
    const int adt = (dtype == DT_BF16) ? DT_F16 : dtype;            // LayerNorm ST-LSTM operand type
    const bool f32 = dtype == DT_F32;
    const ActInfo ca{cdt, esize()}, aa{adt, esize()};
    const int fdt = adt;                                            // autoencoder feature maps: fp16 operands, fp32 conv outputs
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int ns = num_sms;
    const int steps = t_in + pred - 1;
    const size_t pxl = static_cast<size_t>(B) * rh * rw;
    const int Cp = phycell_padded_channels(hid);
    int products = 1;
    if (!f32) {
      products = 3;
      if (const char* env = getenv("VPK_LN_PRODUCTS")) products = std::max(1, std::min(3, atoi(env)));
      if (((C + 63) / 64) * products * k * k > kMaxSteps) products = 1;
    }
    if (ac) products = std::min(products, 2);       // the action-conditional cell runs split weights (as in predrnn-pp)
    const bool lo3 = products == 3;
    const bool pad8 = !f32 && backend == 0 && c <= 8;
    const int cs = pad8 ? 8 : c;

    // ---- buffers ----
    char* frames = static_cast<char*>(arena.alloc(static_cast<size_t>(B) * h * w * cs * esz * t_in));
    float* out_stage = static_cast<float*>(arena.alloc(static_cast<size_t>(B) * pred * c * h * w * sizeof(float)));
    // feature maps: fp32 conv output + its fp16 operand copy (one buffer in fp32 mode)
    auto feat32 = [&](size_t elems) { return static_cast<float*>(arena.alloc(elems * sizeof(float))); };
    auto feat16 = [&](float* f, size_t elems) { return f32 ? static_cast<void*>(f) : arena.alloc(elems * esz); };
    const size_t n_e1 = static_cast<size_t>(B) * h1 * w1 * 32, n_e2 = static_cast<size_t>(B) * h2 * w2 * 64;
    float* e1_32 = feat32(n_e1);
    void* e1 = feat16(e1_32, n_e1);
    float* e2_32 = feat32(n_e2);
    void* e2 = feat16(e2_32, n_e2);
    float* e3 = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));            // mean_layer output before the normalisation
    float* nxt32 = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));         // next_input (fp32): x_gen lands here
    void* nxt_hi = f32 ? static_cast<void*>(nxt32) : arena.alloc(pxl * C * esz);
    void* nxt_lo = lo3 ? arena.alloc(pxl * C * esz) : nullptr;
    void* nxt_cell = f32 ? static_cast<void*>(nxt32) : arena.alloc(pxl * C * esz);
    // ST-LSTM
    std::vector<void*> hb(2 * L), hb_lo(2 * L, nullptr);
    std::vector<float*> cb(L), sth32(L);
    for (int i = 0; i < L; ++i) {
      hb[2 * i] = arena.alloc(pxl * C * esz);
      hb[2 * i + 1] = arena.alloc(pxl * C * esz);
      if (lo3) {
        hb_lo[2 * i] = arena.alloc(pxl * C * esz);
        hb_lo[2 * i + 1] = arena.alloc(pxl * C * esz);
      }
      cb[i] = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
      sth32[i] = f32 ? nullptr : static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
    }
    float* mstate = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
    float* opart = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
    void* mem = arena.alloc(pxl * 2 * C * esz);
    void* m_act = arena.alloc(pxl * C * esz);
    void* dcb = arena.alloc(pxl * C * esz);
    void* dmb = arena.alloc(pxl * C * esz);
    float* xraw = static_cast<float*>(arena.alloc(pxl * 7 * C * sizeof(float)));
    float* hraw = static_cast<float*>(arena.alloc(pxl * 4 * C * sizeof(float)));
    float* mraw = static_cast<float*>(arena.alloc(pxl * 3 * C * sizeof(float)));
    float* oraw = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
    float* lraw = static_cast<float*>(arena.alloc(pxl * C * sizeof(float)));
    float* araw = ac ? static_cast<float*>(arena.alloc(pxl * 4 * C * sizeof(float))) : nullptr;
    const size_t slots = std::max<size_t>(4 * kLnSlices, static_cast<size_t>(ln_slots(7 * C)) + 2 * ln_slots(4 * C) + ln_slots(3 * C));
    lnpart_floats = static_cast<size_t>(B) * slots * 2;
    float* lnpart = static_cast<float*>(arena.alloc(lnpart_floats * sizeof(float)));
    // PhyCell
    std::vector<float*> hp_master(L), htilde(L);
    std::vector<void*> hp_act(2 * L);
    for (int i = 0; i < L; ++i) {
      hp_master[i] = static_cast<float*>(arena.alloc(pxl * C * 4));
      htilde[i] = static_cast<float*>(arena.alloc(pxl * C * 4));
      hp_act[2 * i] = arena.alloc(pxl * C * esz);
      hp_act[2 * i + 1] = arena.alloc(pxl * C * esz);
    }
    // action-conditional: the cells' action tensor of every step, the spatially inflated raw actions (PhyCell), and the
    // outputs of PhyCell's two 1x1 action convs
    const int a_pad = 8;
    void* atens = ac ? arena.alloc(static_cast<size_t>(steps) * pxl * C * esz) : nullptr;
    float* act32 = ac ? static_cast<float*>(arena.alloc(static_cast<size_t>(steps) * pxl * a_pad * sizeof(float))) : nullptr;
    float* fa32 = ac ? static_cast<float*>(arena.alloc(pxl * C * 4)) : nullptr;
    float* ha32 = ac ? static_cast<float*>(arena.alloc(pxl * C * 4)) : nullptr;
    void* fa_act = ac ? (f32 ? static_cast<void*>(fa32) : arena.alloc(pxl * C * esz)) : nullptr;
    void* ha_act = ac ? (f32 ? static_cast<void*>(ha32) : arena.alloc(pxl * C * esz)) : nullptr;
    float* f1raw = static_cast<float*>(arena.alloc(pxl * Cp * 4));
    void* f1n = arena.alloc(pxl * Cp * esz);
    // decoder
    const size_t n_g1 = static_cast<size_t>(B) * d1h * d1w * 64, n_g2 = static_cast<size_t>(B) * d2h * d2w * 32;
    float* g0_32 = feat32(pxl * C);
    void* g0 = feat16(g0_32, pxl * C);
    float* g1_32 = feat32(n_g1);
    void* g1 = feat16(g1_32, n_g1);
    float* g2_32 = feat32(n_g2);
    void* g2 = feat16(g2_32, n_g2);
    auto cast_feat = [&](const float* src, void* dst, size_t n, const char* name) {
      if (measure || f32) return;
      Op op;
      op.name = name;
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(src, DT_F32, nullptr, dst, fdt, static_cast<long long>(n), ns, s); };
      prog.body.push_back(std::move(op));
    };

    if (!measure) {
      Op pre;
      pre.name = "frames_to_nhwc";
      pre.fn = [=](cudaStream_t s, const RunCtx& rc) {
        if (pad8) launch_frames_to_nhwc8(rc.x, static_cast<long long>(t_in) * c * h * w, frames, nullptr, fdt, B, t_in, c, h, w, ns, s);
        else launch_frames_to_nhwc(rc.x, frames, fdt, B, t_in, c, h, w, ns, s);
      };
      prog.pre.push_back(std::move(pre));
      if (ac) {
        const float* wl = dev_f32("action_inflate.weight", vec("action_inflate.weight"), stream);
        const float* wh = dev_f32("action_conv_h.weight", vec("action_conv_h.weight"), stream);
        const float* ww = dev_f32("action_conv_w.weight", vec("action_conv_w.weight"), stream);
        const int asz = a_sz, ia_ = ia, rh_ = rh, rw_ = rw, CC = C, HW = rh * rw;
        Op inf;
        inf.name = "action_tensor";
        inf.fn = [=](cudaStream_t s, const RunCtx& rc) {
          VPK_REQUIRE(rc.actions != nullptr && rc.action_steps >= steps, "Given actions are None or of the wrong size!");
          const long long bs = static_cast<long long>(rc.action_steps) * asz;
          launch_stphy_action_tensor(rc.actions, bs, asz, wl, wh, ww, atens, adt, B, steps, rh_, rw_, CC, ia_, s);
          launch_inflate_actions(rc.actions, bs, asz, act32, DT_F32, B, steps, HW, a_pad, ns, s);
        };
        prog.pre.push_back(std::move(inf));
      }
      for (int i = 0; i < L; ++i) {
        add_memset(prog, hb[2 * i], pxl * C * esz, "zero_h");
        if (lo3) add_memset(prog, hb_lo[2 * i], pxl * C * esz, "zero_h_lo");
        add_memset(prog, cb[i], pxl * C * sizeof(float), "zero_c");
        add_memset(prog, hp_master[i], pxl * C * 4, "zero_hp");
        add_memset(prog, hp_act[2 * i], pxl * C * esz, "zero_hp_act");
      }
      add_memset(prog, mstate, pxl * C * sizeof(float), "zero_m");
      add_memset(prog, m_act, pxl * C * esz, "zero_m_act");
      add_memset(prog, f1n, pxl * Cp * esz, "zero_f1n_pad");
    }
    auto fanout = [&](const float* src, bool norm, const char* name) {      // fp32 -> the operand copies of next_input
      if (measure) return;
      const int rh_ = rh, rw_ = rw, CC = C;
      Op op;
      op.name = name;
      op.fn = [=](cudaStream_t s, const RunCtx&) {
        if (f32) launch_fanout(src, nxt32, nullptr, DT_F32, nullptr, DT_F32, nullptr, B, rh_, rw_, CC, norm, 1e-8f, ns, s);
        else launch_fanout(src, nxt_hi, nxt_lo, adt, nxt_cell, cdt, (src == nxt32) ? nullptr : nxt32, B, rh_, rw_, CC, norm, 1e-8f, ns, s);
      };
      prog.body.push_back(std::move(op));
    };

    std::vector<int> spar(L, 0), ppar(L, 0);
    for (int t = 0; t < steps; ++t) {
      int oh, ow;
      if (t < t_in) {
        // ---- autoencoder.encode(x_t) (enc.py:64-69) ----
        const void* fr = frames + static_cast<size_t>(t) * B * h * w * cs * esz;
        ConvArgs c1{"autoencoder.encoder.conv1.", B, h, w, cs, 32, 5, 2, 0, fr, hp("autoencoder.encoder.conv1.weight"),
                    hp("autoencoder.encoder.conv1.bias"), ACT_RELU, e1_32};
        c1.cin_w = c;
        c1.out_f32_dense = true;
        add_conv(prog, conv_spec(c1, aa, &oh, &ow), measure, stream, fdt);
        VPK_REQUIRE(oh == h1 && ow == w1, "st-phy encoder conv1 size mismatch");
        cast_feat(e1_32, e1, n_e1, "encoder.conv1.cast");
        ConvArgs c2{"autoencoder.encoder.conv2.", B, h1, w1, 32, 64, 3, 2, 0, e1, hp("autoencoder.encoder.conv2.weight"),
                    hp("autoencoder.encoder.conv2.bias"), ACT_RELU, e2_32};
        c2.out_f32_dense = true;
        add_conv(prog, conv_spec(c2, aa, &oh, &ow), measure, stream, fdt);
        VPK_REQUIRE(oh == h2 && ow == w2, "st-phy encoder conv2 size mismatch");
        cast_feat(e2_32, e2, n_e2, "encoder.conv2.cast");
        ConvArgs c3{"autoencoder.encoder.mean_layer.", B, h2, w2, 64, C, 3, 1, 0, e2, hp("autoencoder.encoder.mean_layer.weight"),
                    hp("autoencoder.encoder.mean_layer.bias"), ACT_RELU, e3};
        c3.out_f32_dense = true;
        add_conv(prog, conv_spec(c3, aa, &oh, &ow), measure, stream, fdt);
        VPK_REQUIRE(oh == rh && ow == rw, "st-phy encoder mean_layer size mismatch");
        fanout(e3, true, "encode.normalize");
      } else {
        fanout(nxt32, false, "x_gen.fanout");
      }
      for (int i = 0; i < L; ++i) {
        // ---- PhyCell_Cell(next_input, phy_h[i]) ----
        const std::string p = "phycell_list." + std::to_string(i) + ".";
        const void* h_act = hp_act[2 * i + ppar[i]];
        void* h_act_new = hp_act[2 * i + (ppar[i] ^ 1)];
        const void* xin = nxt_cell;
        const float* h_res = nullptr;
        if (ac) {
          // frame = frame_action_conv(cat[frame, action]), hidden = hidden_action_conv(cat[hidden, action]): fp32 1x1 convs
          // on the CUDA cores straight from the fp32 next_input / state (model_blocks/phydnet.py:50-55, as in model_phydnet.cu)
          const float* a32 = act32 + static_cast<size_t>(t) * pxl * a_pad;
          auto action_conv = [&](const std::string& key, const float* src, float* out32, void* out_act) {
            ConvSpec s1;
            s1.name = p + key + ".";
            s1.B = B;
            s1.G = 1;
            s1.C = C;
            WeightRef wr;
            wr.w = hp(p + key + ".weight");
            wr.O = C;
            wr.I = C + a_sz;
            wr.KH = wr.KW = 1;
            s1.wrefs.push_back(wr);
            BiasRef br;
            br.b = hp(p + key + ".bias");
            s1.biases.push_back(br);
            ConvInput i0{make_view(src, rh, rw, C), 0, 0};
            ConvInput i1{make_view(a32, rh, rw, a_pad), 0, C};
            i1.wc_count = a_sz;
            int oh_, ow_;
            lower_conv(s1, 1, 1, 0, {i0, i1}, rh, rw, 4, &oh_, &ow_);
            EpiParams& e = s1.phases[0].epi;
            e.kind = EPI_BIAS_ACT;
            e.act = ACT_NONE;
            e.out_f32 = 1;
            dense_out(e, out32, rh, rw, C);
            add_conv(prog, s1, measure, stream, DT_F32);
            if (!measure && !f32) {
              const long long n = static_cast<long long>(pxl) * C;
              Op op;
              op.name = p + key + ".cast";
              op.fn = [=](cudaStream_t s, const RunCtx&) { launch_add_to_act(out32, DT_F32, nullptr, out_act, cdt, n, ns, s); };
              prog.body.push_back(std::move(op));
            }
          };
          action_conv("frame_action_conv", nxt32, fa32, fa_act);
          action_conv("hidden_action_conv", hp_master[i], ha32, ha_act);
          xin = fa_act;
          h_act = ha_act;
          h_res = ha32;
        }
        PhyCellArgs pa{p, B, rh, rw, C, hid, kp, xin, h_act, h_act_new, hp_master[i], htilde[i], f1raw, f1n,
                       hp(p + "F.conv1.weight"), hp(p + "F.conv1.bias"), hp(p + "F.conv2.weight"), hp(p + "F.conv2.bias"),
                       hp(p + "convgate.weight"), hp(p + "convgate.bias")};
        pa.h_res = h_res;
        std::vector<ConvSpec> specs = phycell_specs(pa, ca);
        add_conv(prog, specs[0], measure, stream, cdt);
        const int f_groups = stphy_group_norm_divisor(hid);
        if (backend == 0 && C % 16 == 0 && phy_f_tail_supported(rh * rw, hid, Cp, f_groups, C)) {
          if (!measure) {
            const float* g = dev_f32(p + "F.bn1.weight", vec(p + "F.bn1.weight"), stream);
            const float* bta = dev_f32(p + "F.bn1.bias", vec(p + "F.bn1.bias"), stream);
            const float* w2_ = dev_f32(p + "F.conv2.weight", vec(p + "F.conv2.weight"), stream);
            const float* b2_ = dev_f32(p + "F.conv2.bias", vec(p + "F.conv2.bias"), stream);
            const int HW = rh * rw, hid_ = hid, CC = C;
            const float* hm = h_res ? h_res : hp_master[i];
            float* ht = htilde[i];
            Op op;
            op.name = p + "F.tail (GroupNorm + conv2 + h)";
            op.fn = [=](cudaStream_t s, const RunCtx&) {
              launch_phy_f_tail(f1raw, hm, ht, g, bta, w2_, b2_, B, HW, hid_, Cp, f_groups, CC, 1e-5f, s);
            };
            prog.body.push_back(std::move(op));
          }
        } else {
          if (!measure) {
            const float* g = dev_f32(p + "F.bn1.weight", vec(p + "F.bn1.weight"), stream);
            const float* bta = dev_f32(p + "F.bn1.bias", vec(p + "F.bn1.bias"), stream);
            const int HW = rh * rw, hid_ = hid;
            Op op;
            op.name = "groupnorm " + p + "F.bn1.";
            op.fn = [=](cudaStream_t s, const RunCtx&) {
              launch_groupnorm_act(f1raw, DT_F32, f1n, cdt, nullptr, B, HW, hid_, Cp, Cp, f_groups, g, bta, 1e-5f, ACT_NONE, s);
            };
            prog.body.push_back(std::move(op));
          }
          add_conv(prog, specs[1], measure, stream, cdt);
        }
        add_conv(prog, specs[2], measure, stream, cdt);
        ppar[i] ^= 1;
        // ---- SpatioTemporalLSTMCell(layer_norm=True)(next_input, h[i], c[i], st_memory) ----
        const std::string sp = "st_cell_list." + std::to_string(i) + ".";
        LnLo lo{};
        if (lo3) {
          lo.x = nxt_lo;
          lo.h_in = hb_lo[2 * i + spar[i]];
          lo.h_out = hb_lo[2 * i + (spar[i] ^ 1)];
          lo.m_act = nullptr;      // conv_m runs two products (split weights): no low part of m needed
        }
        float* h32 = f32 ? static_cast<float*>(hb[2 * i + (spar[i] ^ 1)]) : sth32[i];
        lo.h_out32 = f32 ? nullptr : sth32[i];
        if (ac)
          add_ac_cell(prog, sp, B, nxt_hi, hb[2 * i + spar[i]], hb[2 * i + (spar[i] ^ 1)], cb[i], mstate, opart, mem, m_act, dcb, dmb,
                      static_cast<const char*>(atens) + static_cast<size_t>(t) * pxl * C * esz, xraw, hraw, araw, mraw, oraw, lraw,
                      lnpart, aa, measure, stream, true, products, f32 ? nullptr : sth32[i]);
        else
        add_ln_cell(prog, sp, B, C, nxt_hi, hb[2 * i + spar[i]], hb[2 * i + (spar[i] ^ 1)], cb[i], mstate, opart, mem, m_act, dcb,
                    dmb, xraw, hraw, mraw, oraw, lraw, lnpart, aa, measure, stream, products, lo);
        spar[i] ^= 1;
        // ---- merge: x_gen = hidden_conv[i](cat[st_h, phy_h]) (1x1, fp32 operands; st_phy.py:158) ----
        {
          const std::string hk = "hidden_conv_list." + std::to_string(i) + ".";
          ConvSpec s1;
          s1.name = hk;
          s1.B = B;
          s1.G = 1;
          s1.C = C;
          WeightRef wr;
          wr.w = hp(hk + "weight");
          wr.O = C;
          wr.I = 2 * C;
          wr.KH = wr.KW = 1;
          s1.wrefs.push_back(wr);
          if (i < L - 1) {
            BiasRef br;
            br.b = hp(hk + "bias");
            s1.biases.push_back(br);
          }
          lower_conv(s1, 1, 1, 0, {ConvInput{make_view(h32, rh, rw, C), 0, 0}, ConvInput{make_view(hp_master[i], rh, rw, C), 0, C}},
                     rh, rw, 4, &oh, &ow);
          EpiParams& e = s1.phases[0].epi;
          e.kind = EPI_BIAS_ACT;
          e.act = ACT_NONE;
          e.out_f32 = 1;
          dense_out(e, nxt32, rh, rw, C);
          // (every layer reads the SAME next_input, and only the last layer's merge is used: the earlier merges are dead
          // values in the reference too -- they are skipped here, their cells are not)
          if (i == L - 1) add_conv(prog, s1, measure, stream, DT_F32);
        }
      }
      if (t < t_in - 1) continue;
      // ---- autoencoder.decode(x_gen) (enc.py:93-98) -> predicted frame t - (t_in - 1) ----
      const int fo = t - (t_in - 1);
      // the decoder's fp16 operand copy of x_gen (the next step's fan-out rewrites nxt_hi with the same values)
      cast_feat(nxt32, nxt_hi, pxl * C, "x_gen.cast");
      ConvArgs f1{"autoencoder.decoder.fc1.", B, rh, rw, C, C, 1, 1, 0, nxt_hi, hp("autoencoder.decoder.fc1.weight"),
                  hp("autoencoder.decoder.fc1.bias"), ACT_RELU, g0_32};
      f1.out_f32_dense = true;
      add_conv(prog, conv_spec(f1, aa, &oh, &ow), measure, stream, fdt);
      cast_feat(g0_32, g0, pxl * C, "decoder.fc1.cast");
      DeconvArgs u1{"autoencoder.decoder.conv1.", B, rh, rw, C, 64, 6, 2, 0, 0, g0, hp("autoencoder.decoder.conv1.weight"),
                    hp("autoencoder.decoder.conv1.bias"), ACT_RELU, g1_32};
      u1.out_f32 = true;
      add_conv(prog, deconv_spec(u1, aa, &oh, &ow), measure, stream, fdt);
      VPK_REQUIRE(oh == d1h && ow == d1w, "st-phy decoder conv1 size mismatch");
      cast_feat(g1_32, g1, n_g1, "decoder.conv1.cast");
      DeconvArgs u2{"autoencoder.decoder.conv2.", B, d1h, d1w, 64, 32, 6, 2, 0, 0, g1, hp("autoencoder.decoder.conv2.weight"),
                    hp("autoencoder.decoder.conv2.bias"), ACT_RELU, g2_32};
      u2.out_f32 = true;
      add_conv(prog, deconv_spec(u2, aa, &oh, &ow), measure, stream, fdt);
      VPK_REQUIRE(oh == d2h && ow == d2w, "st-phy decoder conv2 size mismatch");
      cast_feat(g2_32, g2, n_g2, "decoder.conv2.cast");
      DeconvArgs u3{"autoencoder.decoder.conv3.", B, d2h, d2w, 32, c, 5, 1, 0, 0, g2, hp("autoencoder.decoder.conv3.weight"),
                    hp("autoencoder.decoder.conv3.bias"), ACT_NONE, out_stage + static_cast<size_t>(fo) * c * h * w};
      u3.nchw = true;
      u3.oB_nchw = static_cast<long long>(pred) * c * h * w;
      add_conv(prog, deconv_spec(u3, aa, &oh, &ow), measure, stream, fdt);
      VPK_REQUIRE(oh == h && ow == w, "st-phy decoder conv3 size mismatch");
      if (!measure && !prog.body.empty()) {
        Op& o = prog.body.back();
        o.frame = fo;
        o.frame_src = out_stage + static_cast<size_t>(fo) * c * h * w;
        o.frame_pitch = static_cast<long long>(pred) * c * h * w;
        o.frame_elems = static_cast<long long>(c) * h * w;
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
  int L = 3, hid = 49, kp = 7;
  bool ac = false;             // action_conditional (st_phy.py:48-56)
  int a_sz = 0, ia = 3;        // action_size, inflated_action_dim
  int h1 = 0, w1 = 0, h2 = 0, w2 = 0, d1h = 0, d1w = 0, d2h = 0, d2w = 0;
};

}  // namespace

Model* make_stphy(const vpk_model_desc& d) { return new StPhyModel(d); }

}  // namespace vpk
