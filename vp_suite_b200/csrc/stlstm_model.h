// Shared base of the rollouts built on the LayerNorm ST-LSTM cell pipeline (stlstm_ln.h): PredRNN-V2 (model_predrnn.cu)
// and ST-Phy (model_stphy.cu).  Holds the cell geometry and emits one cell step: raw tcgen05 convs + per-sample statistics
// + the fused gate / output kernels (model_blocks/predrnn.py:24-40, 57-83).
#pragma once
#include <cctype>
#include <cstdlib>

#include "builders.h"
#include "elementwise.h"
#include "model.h"
#include "stlstm_ln.h"

namespace vpk {

class StLstmModelBase : public Model {
 public:
  explicit StLstmModelBase(const vpk_model_desc& d) : Model(d) {}

 protected:
  int k = 5, C = 128;              // filter size / hidden channels of the ST-LSTM cells
  int rh = 16, rw = 16;            // latent size the cells run on
  size_t lnpart_floats = 0;        // capacity of the statistics buffer handed to add_ln_cell

  // statistics slots per sample the tcgen05 epilogue of a G = 1 conv with `co` output channels writes (epilogue slot rule
  // in common.h: ((tile in image) * n_tiles + N tile) * 8 + quadrant * 2 + half)
  int ln_slots(int co) const { return ((rh + 15) / 16) * ((rw + 7) / 8) * conv_n_tiles(co, 1) * 8; }

  // LayerNorm affine of `key` ([kC, H, W] in the reference) repacked to the NHWC order of the raw conv outputs
  const float* ln_param(const std::string& key, int kc, cudaStream_t stream) {
    const float* src = hp(key);
    const int HW = rh * rw;
    std::vector<float> v(static_cast<size_t>(kc) * HW);
    for (int ch = 0; ch < kc; ++ch)
      for (int q = 0; q < HW; ++q) v[static_cast<size_t>(q) * kc + ch] = src[static_cast<size_t>(ch) * HW + q];
    return dev_f32(key + ".nhwc", v, stream);
  }

  // low parts of the split activations of one LayerNorm cell step (three-product mode), all nullptr otherwise
  struct LnLo {
    const void* x = nullptr;
    const void* h_in = nullptr;
    void* h_out = nullptr;
    void* m_act = nullptr;
    float* h_out32 = nullptr;      // optional fp32 copy of h' (consumers that run in fp32; not a low part)
  };

  // One ST-LSTM step with layer_norm=True (stlstm_ln.h): 5 raw convs, 2 statistics launches, 2 fused gate kernels.
  void add_ln_cell(Program& prog, const std::string& pre, int B, int cin, const void* x, const void* h_in, void* h_out,
                   float* c, float* m, float* opart, void* mem, void* m_act, void* dc, void* dm, float* xraw, float* hraw,
                   float* mraw, float* oraw, float* lraw, float* part, const ActInfo& act, bool measure,
                   cudaStream_t stream, int products, const LnLo& lo) {
    int oh, ow;
    // tcgen05 path: the conv epilogues leave the per-sample (sum, sum of squares) partials themselves (one slot per
    // warp, tile and N tile); otherwise a separate statistics launch reads the raw tensors once more
    const char* halo_env = getenv("VPK_TC_HALO");
    const bool fuse_stats = act.dtype != DT_F32 && backend == 0 && getenv("VPK_NO_FUSED_LN_STATS") == nullptr &&
                            (halo_env == nullptr || atoi(halo_env) != 0);
    auto slots_of = [&](int co) { return ln_slots(co); };
    // 16-bit mode: conv_x / conv_h / conv_m run `products` fp16 products per tap (see build(): 2 = split weights over the
    // same activation tile, 3 = split weights and activations), counted once in the algorithmic FLOPs.
    auto raw_conv = [&](const std::string& name, const void* in, int ci, int co, int kk, const std::string& wkey, float* out,
                        float* stat, int nslots, const void* in_lo = nullptr, bool precise = false) {
      ConvArgs a{pre + name, B, rh, rw, ci, co, kk, 1, kk / 2, in, hp(pre + wkey), nullptr, ACT_NONE, out};
      a.out_f32_dense = true;
      // conv_x (its input enters all seven gate pre-activations) needs split activations as well; for conv_h / conv_m
      // split weights are enough: worst of 256 sequences against the reference 1.62e-2 with three products everywhere,
      // 1.65e-2 with two for conv_h / conv_m (2.19e-2 with two for conv_x and conv_h).  VPK_LN_PRODUCTS_X / _H / _M
      // override per conv (developer builds only, capped by `products`).
      int prod = name[5] == 'x' ? products : std::min(products, 2);
      if (precise) {
        const std::string key = std::string("VPK_LN_PRODUCTS_") + static_cast<char>(std::toupper(name[5]));
        if (const char* env = dev_env(key.c_str())) prod = std::max(1, std::min(products, atoi(env)));   // -DVPK_DEV builds only
      }
      if (precise && prod == 2) a.w_split = true;
      if (precise && prod == 3) {
        a.split = true;
        a.x_lo = in_lo;
        a.split_uncounted = true;
      }
      ConvSpec sp = conv_spec(a, act, &oh, &ow);
      sp.is_gate_gemm = true;
      if (stat != nullptr) {
        EpiParams& e = sp.phases[0].epi;
        e.gn_sums = stat;
        e.gn_group_size = -1;
        e.gn_slot0 = 0;
        e.gn_nslots = nslots;
      }
      add_conv(prog, sp, measure, stream, act.dtype);
    };
    // statistics regions inside `part`: X, H, M (and O reuses X's)
    const int nsx = fuse_stats ? slots_of(7 * C) : kLnSlices, nsh = fuse_stats ? slots_of(4 * C) : kLnSlices,
              nsm = fuse_stats ? slots_of(3 * C) : kLnSlices, nso = fuse_stats ? slots_of(C) : kLnSlices;
    float* px_ = part;
    float* ph_ = px_ + static_cast<size_t>(B) * nsx * 2;
    float* pm_ = ph_ + static_cast<size_t>(B) * nsh * 2;
    VPK_REQUIRE(static_cast<size_t>(B) * (static_cast<size_t>(nsx) + nsh + nsm) * 2 <= lnpart_floats && nso <= nsx,
                "LayerNorm statistics regions exceed their buffer");
    raw_conv("conv_x.ln.", x, cin, 7 * C, k, "conv_x.0.weight", xraw, fuse_stats ? px_ : nullptr, nsx, lo.x, true);
    raw_conv("conv_h.ln.", h_in, C, 4 * C, k, "conv_h.0.weight", hraw, fuse_stats ? ph_ : nullptr, nsh, lo.h_in, true);
    raw_conv("conv_m.ln.", m_act, C, 3 * C, k, "conv_m.0.weight", mraw, fuse_stats ? pm_ : nullptr, nsm, lo.m_act, true);
    const int HW = rh * rw, CC = C, ns = num_sms, dt = act.dtype;
    if (!measure) {
      if (!fuse_stats) {
        LnStatsArgs sa{{xraw, hraw, mraw}, {7ll * C * HW, 4ll * C * HW, 3ll * C * HW}, 3, B, part};
        Op op;
        op.name = pre + "ln_stats_xhm";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_ln_stats(sa, s); };
        prog.body.push_back(std::move(op));
      }
      StLnGatesArgs ga{xraw, hraw, mraw, {px_, ph_, pm_}, {nsx, nsh, nsm},
                       ln_param(pre + "conv_x.1.weight", 7 * C, stream), ln_param(pre + "conv_x.1.bias", 7 * C, stream),
                       ln_param(pre + "conv_h.1.weight", 4 * C, stream), ln_param(pre + "conv_h.1.bias", 4 * C, stream),
                       ln_param(pre + "conv_m.1.weight", 3 * C, stream), ln_param(pre + "conv_m.1.bias", 3 * C, stream),
                       c, m, mem, m_act, dc, dm, opart, B, HW, CC, dt, 1.0f};
      ga.m_act_lo = lo.m_act;
      Op og;
      og.name = pre + "ln_gates";
      og.fn = [=](cudaStream_t s, const RunCtx&) { launch_stlstm_ln_gates(ga, ns, s); };
      prog.body.push_back(std::move(og));
    }
    raw_conv("conv_o.ln.", mem, 2 * C, C, k, "conv_o.0.weight", oraw, fuse_stats ? px_ : nullptr, nso);
    raw_conv("conv_last.ln.", mem, 2 * C, C, 1, "conv_last.weight", lraw, nullptr, 0);
    if (!measure) {
      if (!fuse_stats) {
        LnStatsArgs so{{oraw, nullptr, nullptr}, {1ll * C * HW, 0, 0}, 1, B, part};
        Op op;
        op.name = pre + "ln_stats_o";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_ln_stats(so, s); };
        prog.body.push_back(std::move(op));
      }
      StLnOutArgs oa{oraw, lraw, px_, nso, ln_param(pre + "conv_o.1.weight", C, stream),
                     ln_param(pre + "conv_o.1.bias", C, stream), opart, h_out, B, HW, CC, dt};
      oa.h_lo = lo.h_out;
      oa.h32 = lo.h_out32;
      Op oo;
      oo.name = pre + "ln_out";
      oo.fn = [=](cudaStream_t s, const RunCtx&) { launch_stlstm_ln_out(oa, ns, s); };
      prog.body.push_back(std::move(oo));
    }
  }



  // One ActionConditionalSpatioTemporalLSTMCell step (model_blocks/predrnn.py:86-169; stlstm_ln.h): the LayerNorm cell's
  // pipeline with conv biases and a fifth raw conv (conv_a on the action tensor) whose (normalised) output multiplies
  // conv_h's in the i / f / g / o sums; LayerNorm optional (`ln`).  Shared by predrnn-pp and st-phy.
  void add_ac_cell(Program& prog, const std::string& pre, int B, const void* x, const void* h_in, void* h_out, float* c,
                   float* m, float* opart, void* mem, void* m_act, void* dc, void* dm, const void* action, float* xraw,
                   float* hraw, float* araw, float* mraw, float* oraw, float* lraw, float* part, const ActInfo& act,
                   bool measure, cudaStream_t stream, bool ln, int products, float* h32 = nullptr) {
    int oh, ow;
    const char* halo_env = getenv("VPK_TC_HALO");
    const bool fuse_stats = ln && act.dtype != DT_F32 && backend == 0 && getenv("VPK_NO_FUSED_LN_STATS") == nullptr &&
                            (halo_env == nullptr || atoi(halo_env) != 0);
    auto raw_conv = [&](const std::string& name, const void* in, int ci, int co, int kk, const std::string& key, float* out,
                        float* stat, int nslots, bool precise) {
      ConvArgs a{pre + name, B, rh, rw, ci, co, kk, 1, kk / 2, in, hp(pre + key + "weight"), hp(pre + key + "bias"), ACT_NONE, out};
      a.out_f32_dense = true;
      if (precise && products == 2) a.w_split = true;
      ConvSpec sp = conv_spec(a, act, &oh, &ow);
      sp.is_gate_gemm = true;
      if (stat != nullptr) {
        EpiParams& e = sp.phases[0].epi;
        e.gn_sums = stat;
        e.gn_group_size = -1;
        e.gn_slot0 = 0;
        e.gn_nslots = nslots;
      }
      add_conv(prog, sp, measure, stream, act.dtype);
    };
    const int nsx = fuse_stats ? ln_slots(7 * C) : kLnSlices, nsh = fuse_stats ? ln_slots(4 * C) : kLnSlices,
              nsm = fuse_stats ? ln_slots(3 * C) : kLnSlices, nso = fuse_stats ? ln_slots(C) : kLnSlices;
    // regions X, H, M, A in `part` (the non-fused statistics launch writes X, H, M contiguously with kLnSlices each)
    float* px_ = part;
    float* ph_ = px_ + static_cast<size_t>(B) * nsx * 2;
    float* pm_ = ph_ + static_cast<size_t>(B) * nsh * 2;
    float* pa_ = pm_ + static_cast<size_t>(B) * nsm * 2;
    VPK_REQUIRE(static_cast<size_t>(B) * (static_cast<size_t>(nsx) + 2 * nsh + nsm) * 2 <= lnpart_floats && nso <= nsx,
                "LayerNorm statistics regions exceed their buffer");
    raw_conv("conv_x.ac.", x, C, 7 * C, k, "conv_x.0.", xraw, fuse_stats ? px_ : nullptr, nsx, true);
    raw_conv("conv_h.ac.", h_in, C, 4 * C, k, "conv_h.0.", hraw, fuse_stats ? ph_ : nullptr, nsh, true);
    raw_conv("conv_m.ac.", m_act, C, 3 * C, k, "conv_m.0.", mraw, fuse_stats ? pm_ : nullptr, nsm, true);
    raw_conv("conv_a.ac.", action, C, 4 * C, k, "conv_a.0.", araw, fuse_stats ? pa_ : nullptr, nsh, true);
    const int HW = rh * rw, CC = C, ns = num_sms, dt = act.dtype;
    if (!measure) {
      if (ln && !fuse_stats) {
        LnStatsArgs sa{{xraw, hraw, mraw}, {7ll * C * HW, 4ll * C * HW, 3ll * C * HW}, 3, B, part};
        LnStatsArgs sb{{araw, nullptr, nullptr}, {4ll * C * HW, 0, 0}, 1, B, pa_};
        Op op;
        op.name = pre + "ln_stats_xhma";
        op.fn = [=](cudaStream_t s, const RunCtx&) {
          launch_ln_stats(sa, s);
          launch_ln_stats(sb, s);
        };
        prog.body.push_back(std::move(op));
      }
      auto lp = [&](const std::string& key, int kc) -> const float* { return ln ? ln_param(pre + key, kc, stream) : nullptr; };
      StLnGatesArgs ga{xraw, hraw, mraw, {px_, ph_, pm_}, {nsx, nsh, nsm},
                       lp("conv_x.1.weight", 7 * C), lp("conv_x.1.bias", 7 * C), lp("conv_h.1.weight", 4 * C),
                       lp("conv_h.1.bias", 4 * C), lp("conv_m.1.weight", 3 * C), lp("conv_m.1.bias", 3 * C),
                       c, m, mem, m_act, dc, dm, opart, B, HW, CC, dt, 1.0f};
      ga.A = araw;
      ga.part_a = pa_;
      ga.nslots_a = nsh;
      ga.ga = lp("conv_a.1.weight", 4 * C);
      ga.ba = lp("conv_a.1.bias", 4 * C);
      ga.use_ln = ln ? 1 : 0;
      Op og;
      og.name = pre + "ac_gates";
      og.fn = [=](cudaStream_t s, const RunCtx&) { launch_stlstm_ln_gates(ga, ns, s); };
      prog.body.push_back(std::move(og));
    }
    raw_conv("conv_o.ac.", mem, 2 * C, C, k, "conv_o.0.", oraw, fuse_stats ? px_ : nullptr, nso, false);
    {
      ConvArgs a{pre + "conv_last.ac.", B, rh, rw, 2 * C, C, 1, 1, 0, mem, hp(pre + "conv_last.weight"), hp(pre + "conv_last.bias"),
                 ACT_NONE, lraw};
      a.out_f32_dense = true;
      ConvSpec sp = conv_spec(a, act, &oh, &ow);
      sp.is_gate_gemm = true;
      add_conv(prog, sp, measure, stream, act.dtype);
    }
    if (!measure) {
      if (ln && !fuse_stats) {
        LnStatsArgs so{{oraw, nullptr, nullptr}, {1ll * C * HW, 0, 0}, 1, B, part};
        Op op;
        op.name = pre + "ln_stats_o";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_ln_stats(so, s); };
        prog.body.push_back(std::move(op));
      }
      StLnOutArgs oa{oraw, lraw, px_, nso, ln ? ln_param(pre + "conv_o.1.weight", C, stream) : nullptr,
                     ln ? ln_param(pre + "conv_o.1.bias", C, stream) : nullptr, opart, h_out, B, HW, CC, dt};
      oa.use_ln = ln ? 1 : 0;
      oa.h32 = h32;          // optional fp32 copy of h' (ST-Phy's merge conv reads the fp32 state)
      Op oo;
      oo.name = pre + "ac_out";
      oo.fn = [=](cudaStream_t s, const RunCtx&) { launch_stlstm_ln_out(oa, ns, s); };
      prog.body.push_back(std::move(oo));
    }
  }

};

}  // namespace vpk
