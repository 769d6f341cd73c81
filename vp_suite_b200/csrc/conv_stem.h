// Direct CUDA-core kernels for the image-channel ends of the rollouts (conv_stem.cu).
#pragma once
#include <vector>

#include "common.h"

namespace vpk {

struct StemArgs {
  const void* x;      // [B][H][W][8] 16-bit frames (channels >= cin are zero)
  int x_f16;          // 1: fp16, 0: bf16
  int B, H, W, cin, stride;
  const float* w;     // device, conv_stem_pack() layout
  const float* bias;  // device [N] or nullptr
  int N;              // 16 or 32 output channels
  int act;            // ACT_NONE / ACT_LEAKY / ACT_RELU
  void* out;          // dense [B][OH][OW][N], fp32 (out_f32) or bf16
  int out_f32;
};
bool conv_stem_supported(int k, int stride, int pad, int cin, int N, int H, int W);
// round_to: DT_BF16 / DT_F16 rounds the weights like the 16-bit GEMM kernels this replaces do; DT_F32 keeps them
std::vector<float> conv_stem_pack(const float* w /* [N][cin][3][3] */, int N, int cin, int round_to);
void launch_conv_stem(const StemArgs& a, int num_sms, cudaStream_t stream);

struct TailArgs {
  const void* x;      // [B][H][W][CI] fp16
  int B, H, W, CI, CO;
  const float* w;     // device, deconv_tail_pack() layout
  const float* bias;  // device [CO] or nullptr
  int act;
  float* out;         // fp32 NCHW frame: element (b, co, Y, X) at out + b*oB + (co*2H + Y)*2W + X
  long long oB;
  void* fb;           // optional fp16 [B][2H][2W][8] copy of the activated output (channels >= CO zero)
};
bool deconv_tail_supported(int k, int stride, int pad, int out_pad, int cin, int cout, int H, int W);
std::vector<float> deconv_tail_pack(const float* w /* [cin][cout][3][3] */, int cin, int cout, int round_to);
void launch_deconv_tail(const TailArgs& a, int num_sms, cudaStream_t stream);

}  // namespace vpk
