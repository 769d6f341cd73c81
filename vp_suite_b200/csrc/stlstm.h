// ST-LSTM (PredRNN-V2) cell step as three generalised-conv launches
// (reference: SpatioTemporalLSTMCell.forward, layer_norm=False, model_blocks/predrnn.py:57-83).
//
//   launch C (G=4): gates (i, f, g, o_x+o_h)  <- conv_x rows {0,1,2,6} over x  +  conv_h rows {0,1,2,3} over h
//                   epilogue: c' = sig(f+1) c + sig(i) tanh(g);  stores c' (fp32 state + activation copy into
//                   mem[..., 0:C]), delta_c, and the partial output gate o_x+o_h (fp32)
//   launch M (G=3): gates (i', f', g')        <- conv_x rows {3,4,5} over x    +  conv_m rows {0,1,2} over m
//                   epilogue: m' = sig(f'+1) m + sig(i') tanh(g'); stores m' (fp32 + mem[..., C:2C]), delta_m
//   launch O (G=2): (conv_o(mem), conv_last(mem)) over mem = cat(c', m')  -- needs the k x k halo of c', m',
//                   hence a separate launch;  epilogue: h' = sig(o_part + conv_o) * tanh(conv_last)
//
// Splitting x's 7 gate groups between C and M keeps the contraction free of structural zeros (h never feeds the
// primed gates, m never feeds the plain ones); torch.cat((c', m')) is the mem buffer the two epilogues write into.
#pragma once
#include "builders.h"

namespace vpk {

struct StLstmArgs {
  std::string name;
  int B, H, W, Cin, C, k;
  const void* x;          // [B,H,W,Cin]
  const void* h_in;       // [B,H,W,C]
  SrcView m_in;           // activation-type view of m_t (C channels; typically mem[..., C:2C] of the previous cell)
  void* h_out;            // [B,H,W,C]
  float* c;               // fp32 [B,H,W,C] in place
  float* m;               // fp32 [B,H,W,C] in place (the zig-zag memory)
  float* o_part;          // fp32 [B,H,W,C] scratch
  void* mem;              // [B,H,W,2C] written here
  void* dc;               // [B,H,W,C] delta_c
  void* dm;               // [B,H,W,C] delta_m
  const float *w_x, *w_h, *w_m, *w_o, *w_last;   // host, reference layouts
  bool c4 = false;        // c, m, o_part use the channel-quad layout (rollouts); the NCHW-boundary cell keeps NHWC
  // optional fp32 dense [B,H,W,C] scratch: launch O then runs as TWO launches (EPI_ST_O1, common.h) -- conv_o alone with
  // N = C (half the MMA work of the fused form, whose conv_last gate column is zero for all k x k taps) and the 1 x 1
  // conv_last with the output gate in its epilogue.  Bit-identical to the fused launch; pays when the layer is tensor-bound
  float* o_raw = nullptr;
};

inline std::vector<ConvSpec> stlstm_specs(const StLstmArgs& a, const ActInfo& act) {
  std::vector<ConvSpec> out;
  const int C = a.C, k = a.k, pad = a.k / 2;
  int oh, ow;
  auto wref = [&](const float* w, int O, int I, int kk, std::initializer_list<int> blocks) {
    WeightRef r;
    r.w = w;
    r.O = O;
    r.I = I;
    r.KH = r.KW = kk;
    int g = 0;
    for (int b : blocks) r.gate_block[g++] = b;
    for (; g < 4; ++g) r.gate_block[g] = -1;
    return r;
  };
  {  // ---- C ----
    ConvSpec s;
    s.name = a.name + "C";
    s.B = a.B;
    s.G = 4;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(wref(a.w_x, 7 * C, a.Cin, k, {0, 1, 2, 6}));
    s.wrefs.push_back(wref(a.w_h, 4 * C, C, k, {0, 1, 2, 3}));
    lower_conv(s, k, 1, pad,
               {ConvInput{make_view(a.x, a.H, a.W, a.Cin), 0, 0}, ConvInput{make_view(a.h_in, a.H, a.W, C), 1, 0}}, a.H,
               a.W, act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_C;
    e.state_c4 = (a.c4 && C % 4 == 0) ? 1 : 0;
    e.forget_bias = 1.0f;   // predrnn.py:23
    e.s0 = a.c;
    e.s1 = a.o_part;
    e.t0 = a.mem;
    e.t0_pix = 2 * C;
    e.t1 = a.dc;
    out.push_back(std::move(s));
  }
  {  // ---- M ----
    ConvSpec s;
    s.name = a.name + "M";
    s.B = a.B;
    s.G = 3;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(wref(a.w_x, 7 * C, a.Cin, k, {3, 4, 5}));
    s.wrefs.push_back(wref(a.w_m, 3 * C, C, k, {0, 1, 2}));
    lower_conv(s, k, 1, pad, {ConvInput{make_view(a.x, a.H, a.W, a.Cin), 0, 0}, ConvInput{a.m_in, 1, 0}}, a.H, a.W,
               act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_M;
    e.state_c4 = (a.c4 && C % 4 == 0) ? 1 : 0;
    e.forget_bias = 1.0f;
    e.s0 = a.m;
    e.t0 = static_cast<char*>(a.mem) + static_cast<size_t>(C) * act.esize;
    e.t0_pix = 2 * C;
    e.t1 = a.dm;
    out.push_back(std::move(s));
  }
  if (a.o_raw != nullptr) {  // ---- O as conv_last (1 x 1, raw) + conv_o (k x k, N = C) with the output gate ----
    const SrcView mv = make_view(a.mem, a.H, a.W, 2 * C);
    {
      ConvSpec s;
      s.name = a.name + "O.conv_last";
      s.B = a.B;
      s.G = 1;
      s.C = C;
      s.is_gate_gemm = true;
      s.wrefs.push_back(wref(a.w_last, C, 2 * C, 1, {0}));
      lower_conv(s, 1, 1, 0, {ConvInput{mv, 0, 0}}, a.H, a.W, act.esize, &oh, &ow);
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
      s.wrefs.push_back(wref(a.w_o, C, 2 * C, k, {0}));
      lower_conv(s, k, 1, pad, {ConvInput{mv, 0, 0}}, a.H, a.W, act.esize, &oh, &ow);
      EpiParams& e = s.phases[0].epi;
      e.kind = EPI_ST_O1;
      e.variant = 2;
      e.state_c4 = (a.c4 && C % 4 == 0) ? 1 : 0;
      e.s0 = a.o_part;
      e.res = a.o_raw;
      dense_out(e, a.h_out, a.H, a.W, C);
      out.push_back(std::move(s));
    }
  } else {  // ---- O ----
    ConvSpec s;
    s.name = a.name + "O";
    s.B = a.B;
    s.G = 2;
    s.C = C;
    s.is_gate_gemm = true;
    s.wrefs.push_back(wref(a.w_o, C, 2 * C, k, {0, -1}));
    s.wrefs.push_back(wref(a.w_last, C, 2 * C, 1, {-1, 0}));
    const SrcView mv = make_view(a.mem, a.H, a.W, 2 * C);
    lower_conv(s, k, 1, pad, {ConvInput{mv, 0, 0}}, a.H, a.W, act.esize, &oh, &ow);
    lower_conv(s, 1, 1, 0, {ConvInput{mv, 1, 0}}, a.H, a.W, act.esize, &oh, &ow);
    EpiParams& e = s.phases[0].epi;
    e.kind = EPI_ST_O;
    e.state_c4 = (a.c4 && C % 4 == 0) ? 1 : 0;
    e.s0 = a.o_part;
    dense_out(e, a.h_out, a.H, a.W, C);
    s.region_g0 = 1;        // k x k taps feed conv_o only, the 1 x 1 taps conv_last only (ConvLaunch::region_g0)
    out.push_back(std::move(s));
  }
  return out;
}

}  // namespace vpk
