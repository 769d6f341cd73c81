// Spec builders shared by the rollouts and the single-cell entry points: each returns the generalised-convolution
// spec of one reference operation.
#pragma once
#include "lowering.h"

namespace vpk {

struct ActInfo {          // activation storage of the current precision mode
  int dtype;
  int esize;
};

inline SrcView make_view(const void* p, int H, int W, int C) {
  return SrcView{p, H, W, C, static_cast<long long>(H) * W * C, static_cast<long long>(W) * C, C};
}
// channels [c_off, c_off + C) of a dense NHWC tensor with C_total channels
inline SrcView make_channel_view(const void* p, int H, int W, int C_total, int c_off, int C, int esize) {
  return SrcView{static_cast<const char*>(p) + static_cast<size_t>(c_off) * esize, H, W, C,
                 static_cast<long long>(H) * W * C_total, static_cast<long long>(W) * C_total, C_total};
}

inline void dense_out(EpiParams& e, void* out, int H, int W, int C) {
  e.out = out;
  e.oB = static_cast<long long>(H) * W * C;
  e.oY = static_cast<long long>(W) * C;
  e.oX = C;
  e.oC = 1;
}

// ConvLSTM gate conv + fused update.
//   reference: ConvLSTM.forward loop body (model_blocks/conv_lstm_hzzone.py:52-69), rows (i,f,g,o), peepholes
//              ConvLSTMCell.forward (model_blocks/conv_lstm_ndrplz.py:28-43), rows (i,f,o,g), no peepholes
//   x may be null (zero input, conv_lstm_hzzone.py:54-56): its K-steps are dropped, not multiplied.
struct LstmArgs {
  std::string name;
  int B, H, W, Cin, C, k;
  const void* x;            // [B,H,W,Cin] activation type or nullptr
  const void* h_in;         // [B,H,W,C]
  void* h_out;              // [B,H,W,C] (must differ from h_in)
  float* c;                 // [B,H,W,C] fp32, updated in place
  const float* weight;      // host [4C, Cin+C, k, k]
  const float* bias;        // host [4C] or nullptr
  bool order_ifog;          // true: reference rows are (i,f,o,g)
  const float *wci, *wcf, *wco;   // device fp32 [H,W,C] (or [C/4,H,W,4] when c4) or nullptr
  bool c4 = false;          // c and the peepholes use the channel-quad layout [.., C/4, H, W, 4] (needs C % 4 == 0)
  const void* pp16 = nullptr;   // device: the three peepholes as packed bf16 [C/8][H][W][3][8] (tcgen05 epilogue), optional
  // optional second input tensor concatenated BEHIND x (action-conditional SingleStepConvLSTM: the inflated action,
  // model_blocks/phydnet.py:153-155): [B,H,W,C2p] with C2 real channels (C2p - C2 zero padding channels)
  const void* x2 = nullptr;
  int C2 = 0, C2p = 0;
};
inline ConvSpec lstm_spec(const LstmArgs& a, const ActInfo& act) {
  ConvSpec s;
  s.name = a.name;
  s.B = a.B;
  s.G = 4;
  s.C = a.C;
  s.is_gate_gemm = true;
  WeightRef w;
  w.w = a.weight;
  w.O = 4 * a.C;
  w.I = a.Cin + a.C2 + a.C;
  w.KH = w.KW = a.k;
  const int hz[4] = {0, 1, 2, 3}, nd[4] = {0, 1, 3, 2};   // packed gates are always (i, f, g, o)
  for (int g = 0; g < 4; ++g) w.gate_block[g] = a.order_ifog ? nd[g] : hz[g];
  s.wrefs.push_back(w);
  if (a.bias) {
    BiasRef b;
    b.b = a.bias;
    for (int g = 0; g < 4; ++g) b.gate_block[g] = w.gate_block[g];
    s.biases.push_back(b);
  }
  std::vector<ConvInput> in;
  if (a.x) in.push_back(ConvInput{make_view(a.x, a.H, a.W, a.Cin), 0, 0});
  if (a.x2) {
    ConvInput i2{make_view(a.x2, a.H, a.W, a.C2p), 0, a.Cin};
    i2.wc_count = a.C2;
    in.push_back(i2);
  }
  in.push_back(ConvInput{make_view(a.h_in, a.H, a.W, a.C), 0, a.Cin + a.C2});
  int oh, ow;
  lower_conv(s, a.k, 1, a.k / 2, in, a.H, a.W, act.esize, &oh, &ow);
  EpiParams& e = s.phases[0].epi;
  e.kind = EPI_LSTM;
  e.s0 = a.c;
  e.p0 = a.wci;
  e.p1 = a.wcf;
  e.p2 = a.wco;
  e.pp16 = a.pp16;
  e.state_c4 = (a.c4 && a.C % 4 == 0) ? 1 : 0;
  dense_out(e, a.h_out, a.H, a.W, a.C);
  return s;
}

// Conv2d(k, stride, pad) + bias + activation, NHWC activation-type output (dense) or fp32 output with explicit strides.
struct ConvArgs {
  std::string name;
  int B, H, W, Cin, Cout, k, stride, pad;
  const void* x;
  const float* weight;      // host [Cout, Cin, k, k]
  const float* bias;        // host [Cout] or nullptr
  int act;
  void* out;                // dense [B, OH, OW, Cout] activation type unless f32_strided
  bool f32_strided = false; // out is float*, strides below
  long long oB = 0, oY = 0, oX = 0, oC = 0;
  bool out_f32_dense = false;   // out is float*, dense NHWC [B, OH, OW, out_pix] (out_pix >= Cout channels per pixel)
  bool out_f16 = false;         // fp16 operands on the halo kernel's lean epilogue: `out` is __half* (GroupNorm-fed raw convs)
  int out_pix = 0;              // pixel stride of a dense output (0: Cout)
  const float* res = nullptr;   // fp32 dense [B, OH, OW, Cout] residual added after the activation
  bool split = false;           // split-bf16 input: x holds the high parts, x_lo the low parts (same layout)
  const void* x_lo = nullptr;
  int cin_w = -1;               // input channels the weight tensor really has (-1: Cin); x may carry zero padding channels
  bool w_split = false;         // weights as high + low 16-bit parts over the same activations (lowering.h: ConvInput::w_split)
  bool split_uncounted = false; // `split` is for precision: the two extra products are not algorithmic FLOPs
};
inline ConvSpec conv_spec(const ConvArgs& a, const ActInfo& act, int* oh, int* ow) {
  ConvSpec s;
  s.name = a.name;
  s.B = a.B;
  s.G = 1;
  s.C = a.Cout;
  WeightRef w;
  w.w = a.weight;
  w.O = a.Cout;
  w.I = a.cin_w > 0 ? a.cin_w : a.Cin;
  w.KH = w.KW = a.k;
  s.wrefs.push_back(w);
  if (a.bias) {
    BiasRef b;
    b.b = a.bias;
    s.biases.push_back(b);
  }
  ConvInput in{make_view(a.x, a.H, a.W, a.Cin), 0, 0};
  in.wc_count = a.cin_w;
  in.w_split = a.w_split && a.stride == 1;
  if (a.split) in.lo_view = make_view(a.x_lo, a.H, a.W, a.Cin);
  in.extra_uncounted = a.split && a.split_uncounted;
  lower_conv(s, a.k, a.stride, a.pad, {in}, a.H, a.W, act.esize, oh, ow);
  EpiParams& e = s.phases[0].epi;
  e.kind = EPI_BIAS_ACT;
  e.act = a.act;
  if (a.f32_strided) {
    e.out = a.out;
    e.out_f32 = 1;
    e.oB = a.oB; e.oY = a.oY; e.oX = a.oX; e.oC = a.oC;
  } else {
    dense_out(e, a.out, *oh, *ow, a.out_pix > 0 ? a.out_pix : a.Cout);
    e.out_f32 = a.out_f32_dense ? 1 : 0;
    e.out_f16 = (a.out_f16 && !a.out_f32_dense) ? 1 : 0;
  }
  e.res = a.res;
  return s;
}

// ConvTranspose2d(k, stride, pad, output_padding) + bias + activation (weight layout [Cin, Cout, k, k]); one launch per
// output parity.  Output: dense NHWC [B, OH, OW, Cout] of `out_dtype`, or (nchw) fp32 [.., Cout, OH, OW] with batch
// stride oB_nchw (a frame inside a [B, P, c, h, w] tensor).
struct DeconvArgs {
  std::string name;
  int B, H, W, Cin, Cout, k, stride, pad, out_pad;
  const void* x;
  const float* weight;
  const float* bias;
  int act;
  void* out;
  bool out_f32 = false;       // element type of `out` is float regardless of the activation type
  bool out_f16 = false;       // fp16 operands on the halo kernel's lean epilogue: `out` is __half*
  bool nchw = false;          // fp32 NCHW output (implies out_f32)
  long long oB_nchw = 0;
  bool split = false;         // split-bf16 input: x holds the high parts, x_lo the low parts (same layout)
  const void* x_lo = nullptr;
  // fused 1x1 projection of the activated output to proj_n <= 4 channels (tcgen05 path, stride 1 only): `out` is then
  // the fp32 NCHW tensor of the projection (nchw / oB_nchw describe it) and Cout never reaches memory
  const float* proj_w = nullptr;   // device fp32 [proj_n][Cout]
  const float* proj_b = nullptr;   // device fp32 [proj_n]
  int proj_n = 0;
};
// The same stride-2 transposed conv in sub-pixel form (lowering.h): one launch, dense bf16 NHWC output only.
inline ConvSpec deconv_subpix_spec(const DeconvArgs& a, const ActInfo& act, int* oh, int* ow) {
  ConvSpec s;
  s.name = a.name + "subpix.";
  s.B = a.B;
  s.G = 4;
  s.C = a.Cout;
  WeightRef w;
  w.w = a.weight;
  w.O = a.Cout;
  w.I = a.Cin;
  w.KH = w.KW = a.k;
  w.transposed = true;
  for (int g = 0; g < 4; ++g) w.gate_block[g] = 0;
  s.wrefs.push_back(w);
  if (a.bias) {
    BiasRef b;
    b.b = a.bias;
    for (int g = 0; g < 4; ++g) b.gate_block[g] = 0;
    s.biases.push_back(b);
  }
  ConvInput cin{make_view(a.x, a.H, a.W, a.Cin), 0, 0};
  lower_conv_transpose_subpixel(s, a.k, a.pad, a.out_pad, cin, a.H, a.W, oh, ow);
  EpiParams& e = s.phases[0].epi;
  e.kind = EPI_SUBPIX;
  e.act = a.act;
  e.out = a.out;
  e.out_f32 = 0;
  const long long C = a.Cout, OW = *ow, OH = *oh;
  e.oB = OH * OW * C;
  e.oY = 2 * OW * C;
  e.oX = 2 * C;
  e.oC = 1;
  e.ps_row = OW * C;
  return s;
}
inline bool deconv_subpix_ok(const DeconvArgs& a) {
  if (a.stride != 2 || a.split || a.nchw || a.out_f32 || a.proj_n > 0 || a.Cout % 8 != 0 || a.Cin % 8 != 0) return false;
  return (a.H - 1) * 2 - 2 * a.pad + a.k + a.out_pad == 2 * a.H;
}
inline ConvSpec deconv_spec(const DeconvArgs& a, const ActInfo& act, int* oh, int* ow) {
  ConvSpec s;
  s.name = a.name;
  s.B = a.B;
  s.G = 1;
  s.C = a.Cout;
  WeightRef w;
  w.w = a.weight;
  w.O = a.Cout;
  w.I = a.Cin;
  w.KH = w.KW = a.k;
  w.transposed = true;
  s.wrefs.push_back(w);
  if (a.bias) {
    BiasRef b;
    b.b = a.bias;
    s.biases.push_back(b);
  }
  const int C = a.Cout, actk = a.act;
  const bool f32 = a.out_f32 || a.nchw, nchw = a.nchw;
  const bool f16o = a.out_f16 && !f32;
  const size_t esz = f32 ? sizeof(float) : static_cast<size_t>(act.esize);
  const long long oBn = a.oB_nchw;
  void* out = a.out;
  const float* pw = a.proj_w;
  const float* pb = a.proj_b;
  const int pn = a.proj_n;
  ConvInput cin{make_view(a.x, a.H, a.W, a.Cin), 0, 0};
  if (a.split) cin.lo_view = make_view(a.x_lo, a.H, a.W, a.Cin);
  lower_conv_transpose(s, a.k, a.stride, a.pad, a.out_pad, cin, a.H, a.W,
                       oh, ow, [=](int ry, int rx, int stride, int OH, int OW) {
                         EpiParams e{};
                         e.kind = EPI_BIAS_ACT;
                         e.act = actk;
                         e.out_f32 = f32 ? 1 : 0;
                         e.out_f16 = f16o ? 1 : 0;
                         e.proj_w = pw;
                         e.proj_b = pb;
                         e.proj_n = pn;
                         if (nchw) {   // element (b, ch, Y, X) with Y = stride*q_y + ry, X = stride*q_x + rx
                           e.out = static_cast<char*>(out) + (static_cast<size_t>(ry) * OW + rx) * esz;
                           e.oB = oBn;
                           e.oC = static_cast<long long>(OH) * OW;
                           e.oY = static_cast<long long>(stride) * OW;
                           e.oX = stride;
                         } else {
                           e.out = static_cast<char*>(out) + (static_cast<size_t>(ry) * OW + rx) * C * esz;
                           e.oB = static_cast<long long>(OH) * OW * C;
                           e.oY = static_cast<long long>(stride) * OW * C;
                           e.oX = static_cast<long long>(stride) * C;
                           e.oC = 1;
                         }
                         return e;
                       });
  return s;
}

}  // namespace vpk
