// ST-LSTM with layer_norm=True (model_blocks/predrnn.py:24-40, 57-83): every conv is followed by
// nn.LayerNorm([k*C, H, W]) (statistics over all k*C*H*W elements of a sample, elementwise affine), so the gate
// pre-activations have to exist as whole tensors before they can be normalised.  The cell step becomes
//   conv_x / conv_h / conv_m (tcgen05, fp32 raw outputs) -> per-sample statistics -> ONE fused gate kernel
//   (normalise the 14 gate slices, c' / m' update, mem = cat(c', m'), delta_c / delta_m, o_x + o_h) ->
//   conv_o / conv_last over mem (raw) -> statistics of conv_o -> h' = sigmoid(o_part + LN(conv_o)) * tanh(conv_last).
// The elementwise kernels are HBM-bound: 14C fp32 values per position are written by the convs and read twice.
#pragma once
#include "common.h"

namespace vpk {

constexpr int kLnSlices = 32;      // partial-statistics slots per (tensor, sample), added in order by the consumers

struct LnStatsArgs {
  const float* in[3];   // dense fp32 [B][n[z]]
  long long n[3];
  int ntens, B;
  float* part;          // [ntens][B][kLnSlices][2] (sum, sum of squares)
};
void launch_ln_stats(const LnStatsArgs& a, cudaStream_t stream);

struct StLnGatesArgs {
  const float *X, *H, *M;                      // raw conv outputs, NHWC [B*HW][7C] / [4C] / [3C]
  // partial statistics of X, H, M: part[z][b][nslots[z]][2] (sum, sum of squares), written either by launch_ln_stats
  // (kLnSlices slots) or by the conv epilogues themselves (EpiParams::gn_group_size = -1)
  const float* part[3];
  int nslots[3];
  const float *gx, *bx, *gh, *bh, *gm, *bm;    // LayerNorm affine, repacked [HW][k*C]
  float *c, *m;                                // fp32 state [B*HW][C], updated in place
  void* mem;                                   // activation type [B*HW][2C]: c' | m'
  void* m_act;                                 // activation type [B*HW][C]: m' (dense copy for the next conv_m)
  void *dc, *dm;                               // activation type [B*HW][C]
  float* opart;                                // fp32 [B*HW][C]: LN(conv_x)_o + LN(conv_h)_o
  int B, HW, C, dtype;
  float forget_bias;
  void* m_act_lo = nullptr;                    // optional: low part of m' (m' - m_act), activation type [B*HW][C]
  // action-conditional cell (model_blocks/predrnn.py:86-169): A = raw conv_a output [B*HW][4C]; (LN(H) * LN(A)) replaces
  // LN(H) in the i, f, g, o sums (:149).  nullptr: the plain cell
  const float* A = nullptr;
  const float* part_a = nullptr;               // partial statistics of A, [b][nslots_a][2]
  int nslots_a = 0;
  const float *ga = nullptr, *ba = nullptr;    // LayerNorm affine of conv_a, [HW][4C]
  int use_ln = 1;                              // 0: no LayerNorm (raw conv outputs already carry their bias): identity
};
void launch_stlstm_ln_gates(const StLnGatesArgs& a, int num_sms, cudaStream_t stream);

struct StLnOutArgs {
  const float *O, *Lraw;                       // raw conv_o / conv_last outputs [B*HW][C]
  const float* part;                           // partial statistics of O: [b][nslots][2]
  int nslots;
  const float *go, *bo;                        // LayerNorm affine of conv_o, [HW][C]
  const float* opart;
  void* h;                                     // activation type [B*HW][C]
  int B, HW, C, dtype;
  void* h_lo = nullptr;                        // optional: low part of h' (h' - h), activation type [B*HW][C]
  int use_ln = 1;                              // 0: no LayerNorm on conv_o's output
  float* h32 = nullptr;                        // optional fp32 copy of h' [B*HW][C]
};
void launch_stlstm_ln_out(const StLnOutArgs& a, int num_sms, cudaStream_t stream);

}  // namespace vpk
