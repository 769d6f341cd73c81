// Causal LSTM cell step and gradient highway unit (PredRNN++: Wang et al., "PredRNN++: Towards A Resolution of the
// Deep-in-Time Dilemma in Spatiotemporal Predictive Learning", ICML 2018, eqs. (1)-(3) / sec. 3.1-3.2) as generalised-conv
// launches.  BASELINE.json's north star names these cells; the vp-suite checkout has neither (its `predrnn-pp` key is
// PredRNN-V2's ST-LSTM), so there is no reference module to compare with: PARITY UNPINNED -- the checker is
// oracle/causal.py, a restatement of the published equations in the bias-free form of the public PyTorch
// re-implementations (conv_x 7C, conv_h 4C, conv_c 3C, conv_m 3C over the previous layer's memory, conv_c2m 4C split (i, g, f, o), conv_om C,
// conv_last 1x1).
//
// The cell is a cascade -- c' feeds the spatial memory's gates, c' and m' feed the output gate -- so it takes three
// dependent launches, each a concat-free multi-source contraction with the gate math fused into the epilogue:
//
//   launch C (G=4): (i, f, g, o_x+o_h) <- conv_x rows {0,1,2,6} over x + conv_h rows {0,1,2,3} over h + conv_c rows
//                   {0,1,2} over c           epilogue (EPI_ST_C): c' = sig(f+1) c + sig(i) tanh(g); stores c' (fp32 state +
//                   activation copy into mem[..., 0:C]) and the partial output gate o_x+o_h (fp32)
//   launch M (G=4): (i', f', g', m_m)  <- conv_x rows {3,4,5} over x + conv_m rows {0,1,-,2} over m + conv_c2m rows
//                   {0,2,1} over c'          epilogue (EPI_ST_C variant 1): m' = sig(f'+1) tanh(m_m) + sig(i') tanh(g');
//                   stores m' into mem[..., C:2C]
//   launch O (G=2): (o_c + o_m, last)  <- conv_c2m row 3 over c' + conv_om over m' (k x k) and conv_last (1 x 1) over
//                   mem = cat(c', m')        epilogue (EPI_ST_O variant 1): h' = tanh(o_part + acc0) * tanh(acc1)
//
//   GHU      (G=2): (p, u)             <- x_concat over x + z_concat over z
//                                      epilogue (EPI_ST_O variant 2): z' = sig(u) z + (1 - sig(u)) tanh(p)
#pragma once
#include "builders.h"

namespace vpk {

struct CausalArgs {
  std::string name;
  int B, H, W, Cin, C, k;
  const void* x;          // [B,H,W,Cin]
  const void* h_in;       // [B,H,W,C]
  SrcView c_in;           // activation-type view of c_t (mem[..., 0:C] this layer wrote at the previous step)
  SrcView m_in;           // activation-type view of m_t (mem[..., C:2C] of the previous cell)
  void* h_out;            // [B,H,W,C]
  float* c;               // fp32 [B,H,W,C] in place
  float* m;               // fp32 [B,H,W,C] (written only: m enters the next cell through its convs alone)
  float* o_part;          // fp32 [B,H,W,C] scratch
  void* mem;              // [B,H,W,2C] written here: (c', m')
  const float *w_x, *w_h, *w_c, *w_m, *w_c2m, *w_om, *w_last;   // host, layouts of the header comment
  bool c4 = false;        // c, m, o_part use the channel-quad layout
  float* o_raw = nullptr; // optional fp32 dense [B,H,W,C] scratch: launch O as two launches (EPI_ST_O1, see stlstm.h)
  int Cm = 0;             // channels of m_t (0: C).  The spatial memory comes from the PREVIOUS layer (the top layer for layer 0),
                          // so in a stack of unequal widths (the paper's 128-64-64-64) conv_m is [3C, Cm, k, k]
};

inline WeightRef causal_wref(const float* w, int O, int I, int kk, std::initializer_list<int> blocks) {
  WeightRef r;
  r.w = w;
  r.O = O;
  r.I = I;
  r.KH = r.KW = kk;
  int g = 0;
  for (int b : blocks) r.gate_block[g++] = b;
  for (; g < 4; ++g) r.gate_block[g] = -1;
  return r;
}

inline std::vector<ConvSpec> causal_lstm_specs(const CausalArgs& a, const ActInfo& act) {
  std::vector<ConvSpec> out;
  const int C = a.C, k = a.k, pad = a.k / 2;
  const int c4 = (a.c4 && C % 4 == 0) ? 1 : 0;
  int oh, ow;
  const SrcView cnew = make_channel_view(a.mem, a.H, a.W, 2 * C, 0, C, act.esize);
  const SrcView mnew = make_channel_view(a.mem, a.H, a.W, 2 * C, C, C, act.esize);
  {  // ---- C: temporal memory ----
    ConvSpec s;
    s.name = a.name + "C";
    s.B = a.B;
    s.G = 4;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(causal_wref(a.w_x, 7 * C, a.Cin, k, {0, 1, 2, 6}));
    s.wrefs.push_back(causal_wref(a.w_h, 4 * C, C, k, {0, 1, 2, 3}));
    s.wrefs.push_back(causal_wref(a.w_c, 3 * C, C, k, {0, 1, 2, -1}));
    lower_conv(s, k, 1, pad,
               {ConvInput{make_view(a.x, a.H, a.W, a.Cin), 0, 0}, ConvInput{make_view(a.h_in, a.H, a.W, C), 1, 0},
                ConvInput{a.c_in, 2, 0}},
               a.H, a.W, act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_C;
    e.state_c4 = c4;
    e.forget_bias = 1.0f;
    e.s0 = a.c;
    e.s1 = a.o_part;
    e.t0 = a.mem;
    e.t0_pix = 2 * C;
    e.t1 = nullptr;
    s.region_g0 = 3;        // c does not feed o_part: regions (i, f, g) | (o_x + o_h)
    out.push_back(std::move(s));
  }
  {  // ---- M: spatial memory, cascaded behind c' ----
    ConvSpec s;
    s.name = a.name + "M";
    s.B = a.B;
    s.G = 4;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(causal_wref(a.w_x, 7 * C, a.Cin, k, {3, 4, 5, -1}));
    s.wrefs.push_back(causal_wref(a.w_m, 3 * C, a.Cm > 0 ? a.Cm : C, k, {0, 1, -1, 2}));
    s.wrefs.push_back(causal_wref(a.w_c2m, 4 * C, C, k, {0, 2, 1, -1}));      // conv_c2m splits as (i, g, f, o)
    lower_conv(s, k, 1, pad,
               {ConvInput{make_view(a.x, a.H, a.W, a.Cin), 0, 0}, ConvInput{a.m_in, 1, 0}, ConvInput{cnew, 2, 0}}, a.H, a.W,
               act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_C;
    e.variant = 1;
    e.state_c4 = c4;
    e.forget_bias = 1.0f;
    e.s0 = a.m;
    e.t0 = static_cast<char*>(a.mem) + static_cast<size_t>(C) * act.esize;
    e.t0_pix = 2 * C;
    s.region_g0 = 3;        // x and c' do not feed m_m: regions (i', f', g') | (m_m)
    out.push_back(std::move(s));
  }
  if (a.o_raw != nullptr) {  // ---- O as conv_last (1 x 1, raw) + (conv_c2m[o](c') + conv_om(m')) with the output gate ----
    {
      ConvSpec s;
      s.name = a.name + "O.conv_last";
      s.B = a.B;
      s.G = 1;
      s.C = C;
      s.is_gate_gemm = true;
      s.wrefs.push_back(causal_wref(a.w_last, C, 2 * C, 1, {0}));
      lower_conv(s, 1, 1, 0, {ConvInput{make_view(a.mem, a.H, a.W, 2 * C), 0, 0}}, a.H, a.W, act.esize, &oh, &ow);
      EpiParams& e = s.phases[0].epi;
      e.kind = EPI_BIAS_ACT;
      e.act = ACT_NONE;
      e.out_f32 = 1;
      dense_out(e, a.o_raw, a.H, a.W, C);
      out.push_back(std::move(s));
    }
    {
      ConvSpec s;
      s.name = a.name + "O.conv_o";
      s.B = a.B;
      s.G = 1;
      s.C = C;
      s.is_gate_gemm = true;
      s.wrefs.push_back(causal_wref(a.w_c2m, 4 * C, C, k, {3}));
      s.wrefs.push_back(causal_wref(a.w_om, C, C, k, {0}));
      lower_conv(s, k, 1, pad, {ConvInput{cnew, 0, 0}, ConvInput{mnew, 1, 0}}, a.H, a.W, act.esize, &oh, &ow);
      EpiParams& e = s.phases[0].epi;
      e.kind = EPI_ST_O1;
      e.variant = 3;
      e.state_c4 = c4;
      e.s0 = a.o_part;
      e.res = a.o_raw;
      dense_out(e, a.h_out, a.H, a.W, C);
      out.push_back(std::move(s));
    }
  } else {  // ---- O: output gate over (x, h) [kept from launch C] + c' + m', and the 1 x 1 conv over cat(c', m') ----
    ConvSpec s;
    s.name = a.name + "O";
    s.B = a.B;
    s.G = 2;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(causal_wref(a.w_c2m, 4 * C, C, k, {3, -1}));
    s.wrefs.push_back(causal_wref(a.w_om, C, C, k, {0, -1}));
    s.wrefs.push_back(causal_wref(a.w_last, C, 2 * C, 1, {-1, 0}));
    lower_conv(s, k, 1, pad, {ConvInput{cnew, 0, 0}, ConvInput{mnew, 1, 0}}, a.H, a.W, act.esize, &oh, &ow);
    lower_conv(s, 1, 1, 0, {ConvInput{make_view(a.mem, a.H, a.W, 2 * C), 2, 0}}, a.H, a.W, act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_O;
    e.variant = 1;
    e.state_c4 = c4;
    e.s0 = a.o_part;
    dense_out(e, a.h_out, a.H, a.W, C);
    s.region_g0 = 1;        // k x k taps feed the output gate only, the 1 x 1 taps conv_last only
    out.push_back(std::move(s));
  }
  return out;
}

struct GhuArgs {
  std::string name;
  int B, H, W, C, k;
  const void* x;          // [B,H,W,C]  (h of the first Causal LSTM layer)
  const void* z_in;       // [B,H,W,C]  activation copy of z_{t-1}
  void* z_out;            // [B,H,W,C]  activation copy of z_t (must differ from z_in)
  float* z;               // fp32 [B,H,W,C] in place
  const float *w_x, *w_z; // host [2C, C, k, k] each, rows (p, u)
  bool c4 = false;
};

inline ConvSpec ghu_spec(const GhuArgs& a, const ActInfo& act) {
  ConvSpec s;
  s.name = a.name;
  s.B = a.B;
  s.G = 2;
  s.C = a.C;
  s.is_gate_gemm = true;
  s.wrefs.push_back(causal_wref(a.w_x, 2 * a.C, a.C, a.k, {0, 1}));
  s.wrefs.push_back(causal_wref(a.w_z, 2 * a.C, a.C, a.k, {0, 1}));
  int oh, ow;
  lower_conv(s, a.k, 1, a.k / 2,
             {ConvInput{make_view(a.x, a.H, a.W, a.C), 0, 0}, ConvInput{make_view(a.z_in, a.H, a.W, a.C), 1, 0}}, a.H, a.W,
             act.esize, &oh, &ow);
  EpiParams& e = s.phases[0].epi;
  e.kind = EPI_ST_O;
  e.variant = 2;
  e.state_c4 = (a.c4 && a.C % 4 == 0) ? 1 : 0;
  e.s0 = a.z;
  dense_out(e, a.z_out, a.H, a.W, a.C);
  return s;
}

}  // namespace vpk
