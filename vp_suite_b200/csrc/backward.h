// Backward pass of the two ConvLSTM cells (the differentiable entries behind the C ABI).
#pragma once
#include "common.h"

namespace vpk {

// ConvLSTMCell.forward (model_blocks/conv_lstm_ndrplz.py:28-43) differentiated: z = (i, f, o, g) pre-activations in the
// reference's row-block order, dense NHWC fp32 [P][4C]; c [P][C]; upstream gradients dh_out / dc_out [P][C] (nullptr = zero).
// Writes dz [P][4C] (fp32, and an activation-type copy for the dgrad conv when dz_act != nullptr) and dc_in [P][C].
void launch_lstm_gate_backward(const float* z, const float* c, const float* dh_out, const float* dc_out, float* dz,
                               void* dz_act, int act_dtype, float* dc_in, long long P, int C, int num_sms,
                               cudaStream_t stream);

// The Shi et al. ConvLSTM step with peepholes (conv_lstm_hzzone.py:57-69), split order (i, f, g, o): z [B][HW][4C], c [B][HW][C],
// peepholes wci / wcf / wco NHWC [HW][C] (nullptr = zero); also writes the peephole gradients dwci / dwcf / dwco [HW][C]
// (sums over the batch, fixed order; nullptr = not wanted).
void launch_lstm_peep_gate_backward(const float* z, const float* c, const float* wci, const float* wcf, const float* wco,
                                    const float* dh_out, const float* dc_out, float* dz, void* dz_act, int act_dtype,
                                    float* dc_in, float* dwci, float* dwcf, float* dwco, int B, long long HW, int C, int num_sms,
                                    cudaStream_t stream);

// Weight gradient of a k x k stride-1 'same' conv: dw[o][i][ky][kx] = sum over (b, y, x) of dz[b, y, x, o] *
// in[b, y + ky - k/2, x + kx - k/2, i] (zero outside the image); `in` NHWC fp32 [B][H][W][Ci], dz NHWC fp32 [B][H][W][Co];
// dw fp32 in the reference layout [Co][Ci_total][k][k], written at input-channel offset i0 (cat(x, h): two calls).
// Deterministic (no atomics: one CTA owns an output tile and walks all positions in order).
void launch_conv_wgrad(const float* in, const float* dz, float* dw, int B, int H, int W, int Ci, int Co, int k, int Ci_total,
                       int i0, cudaStream_t stream);

// db[o] = sum over positions of dz[p][o]
void launch_bias_grad(const float* dz, float* db, long long P, int Co, cudaStream_t stream);

}  // namespace vpk
