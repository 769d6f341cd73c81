// PhyCell cell step as three generalised-conv launches around one GroupNorm pass
// (reference: PhyCell_Cell.forward, action_conditional=False, model_blocks/phydnet.py:49-62).
//
//   conv1 : F.conv1 (k x k, bias) over h                 -> f1 (fp32, channel-padded so that TMA strides are 16-byte)
//   [GroupNorm(groups, hid) of f1 -> f1n is launched by the caller between conv1 and conv2]
//   conv2 : F.conv2 (1 x 1, bias) over f1n, + h (fp32)   -> h~ = h + F(h)  (fp32)
//   gate  : convgate (3 x 3, bias) over (x, h) with the fused blend  h' = h~ + sigmoid(.) * (x - h~)
//           -> h' as fp32 master and as conv-operand copy
#pragma once
#include "builders.h"

namespace vpk {

struct PhyCellArgs {
  std::string name;
  int B, H, W, C, hid, k;
  const void* x;            // [B,H,W,C] activation type
  const void* h_act;        // [B,H,W,C] activation-type copy of h
  void* h_act_out;          // [B,H,W,C] (must differ from h_act)
  float* h_master;          // fp32 [B,H,W,C], updated in place
  float* htilde;            // fp32 [B,H,W,C] scratch
  float* f1raw;             // fp32 [B,H,W,Cp]
  const void* f1n;          // activation type [B,H,W,Cp], padded channels zero
  const float *conv1_w, *conv1_b, *conv2_w, *conv2_b, *gate_w, *gate_b;   // host, reference layouts
  // optional: a separate copy of h for F.conv1 in the operand type of the F path (see phycell_specs' f_act); nullptr: h_act
  const void* h_f = nullptr;
  // optional: the fp32 `hidden` the prediction h~ = hidden + F(hidden) starts from when it is not the carried state
  // (action-conditional: hidden = hidden_action_conv(cat[state, action]), model_blocks/phydnet.py:53-55); nullptr: h_master
  const float* h_res = nullptr;
};

inline int phycell_padded_channels(int hid) { return (hid + 7) / 8 * 8; }

// act: operand type of the gate conv (and of x / h_act / h_act_out); f_act: operand type of F.conv1 / F.conv2 (h_f, f1n).
// The caller builds specs [0], [1] with f_act.dtype and [2] with act.dtype.
inline std::vector<ConvSpec> phycell_specs(const PhyCellArgs& a, const ActInfo& act, const ActInfo& f_act) {
  std::vector<ConvSpec> out;
  const int Cp = phycell_padded_channels(a.hid);
  int oh, ow;
  {
    ConvArgs c1{a.name + "F.conv1.", a.B, a.H, a.W, a.C, a.hid, a.k, 1, a.k / 2, a.h_f ? a.h_f : a.h_act, a.conv1_w,
                a.conv1_b, ACT_NONE, a.f1raw};
    c1.out_f32_dense = true;
    c1.out_pix = Cp;
    out.push_back(conv_spec(c1, f_act, &oh, &ow));
  }
  {
    ConvSpec s;
    s.name = a.name + "F.conv2.";
    s.B = a.B;
    s.G = 1;
    s.C = a.C;
    WeightRef wr;
    wr.w = a.conv2_w;
    wr.O = a.C;
    wr.I = a.hid;
    wr.KH = wr.KW = 1;
    s.wrefs.push_back(wr);
    BiasRef br;
    br.b = a.conv2_b;
    s.biases.push_back(br);
    ConvInput in{make_view(a.f1n, a.H, a.W, Cp), 0, 0};
    in.wc_count = a.hid;
    lower_conv(s, 1, 1, 0, {in}, a.H, a.W, f_act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_BIAS_ACT;
    e.act = ACT_NONE;
    e.out_f32 = 1;
    dense_out(e, a.htilde, a.H, a.W, a.C);
    e.res = a.h_res ? a.h_res : a.h_master;
    out.push_back(std::move(s));
  }
  {
    ConvSpec s;
    s.name = a.name + "convgate.";
    s.B = a.B;
    s.G = 1;
    s.C = a.C;
    s.is_gate_gemm = true;
    WeightRef wr;
    wr.w = a.gate_w;
    wr.O = a.C;
    wr.I = 2 * a.C;
    wr.KH = wr.KW = 3;
    s.wrefs.push_back(wr);
    BiasRef br;
    br.b = a.gate_b;
    s.biases.push_back(br);
    lower_conv(s, 3, 1, 1,
               {ConvInput{make_view(a.x, a.H, a.W, a.C), 0, 0}, ConvInput{make_view(a.h_act, a.H, a.W, a.C), 0, a.C}},
               a.H, a.W, act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_PHY_GATE;
    e.q0 = a.x;
    e.res = a.htilde;
    e.s0 = a.h_master;
    dense_out(e, a.h_act_out, a.H, a.W, a.C);
    out.push_back(std::move(s));
  }
  return out;
}
inline std::vector<ConvSpec> phycell_specs(const PhyCellArgs& a, const ActInfo& act) { return phycell_specs(a, act, act); }

}  // namespace vpk
