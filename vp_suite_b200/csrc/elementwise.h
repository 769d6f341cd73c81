// Launchers of the memory-bound helper kernels (elementwise.cu).
#pragma once
#include <algorithm>

#include "common.h"

namespace vpk {

// x fp32 [B, T, C, H, W] -> out (activation type) [T][B][H][W][C]
void launch_frames_to_nhwc(const float* x, void* out, int dtype, int B, int T, int C, int H, int W, int num_sms,
                           cudaStream_t stream);
// x fp32 [B, *, C, H, W] (sequence stride `bstride` elements), first T frames -> bf16 [T][B][H][W][8], channels C..7
// zero (TMA-addressable frames); lo != nullptr: split-bf16 (hi, lo) pair; out_dtype == DT_F16: `hi` receives fp16 values
void launch_frames_to_nhwc8(const float* x, long long bstride, void* hi, void* lo, int out_dtype, int B, int T, int C,
                            int H, int W, int num_sms, cudaStream_t stream);
// PredRNN patchify: the first T frames of x fp32 [B, *, c, H, W] (sequence stride `bstride` elements)
// -> out [T][B][H/p][W/p][p*p*c]
void launch_patchify_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H, int W,
                             int p, int num_sms, cudaStream_t stream);
// one patch frame [B][H/p][W/p][p*p*c] -> frame t of fp32 [B, P, c, H, W]
void launch_unpatchify(const void* in, float* out, int dtype, int B, int P, int t, int C, int H, int W, int p,
                       int num_sms, cudaStream_t stream);

// in (activation type or fp32) [B][H][W][C] -> out fp32 [B][C][H][W]
void launch_nhwc_to_nchw(const void* in, int dtype, float* out, int B, int C, int H, int W, int num_sms,
                         cudaStream_t stream);

// same with an explicit batch stride (elements) of x: frames of one sequence stay contiguous
void launch_frames_to_nhwc_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H,
                                   int W, int num_sms, cudaStream_t stream);
void launch_cast_f32_to_bf16(const float* in, void* out, long long n, int num_sms, cudaStream_t stream);
// PredRNN-V2 decouple loss: ad fp32 [2B][HW][C] (adapter(delta_c) then adapter(delta_m)); *acc += sum_{b,ch} |cos|
void launch_decouple_reduce(const float* ad, int B, int HW, int C, double* acc, cudaStream_t stream);
// PhyCell: h~ = h + conv2(GroupNorm(f1)), one block per sample (f1 fp32 [B][HW][Cs], hid real channels; w2 [C][hid], b2 [C])
bool phy_f_tail_supported(int HW, int hid, int Cs, int groups, int C);
void launch_phy_f_tail(const float* f1, const float* h, float* htilde, const float* gamma, const float* beta,
                       const float* w2, const float* b2, int B, int HW, int hid, int Cs, int groups, int C, float eps,
                       cudaStream_t stream);
// fused form: the EPI_DECOUPLE conv epilogue leaves slots[b][slot][C][3]; term[b] = sum_ch |cos|; then *acc += sum(terms)
void launch_decouple_cos(const float* slots, int nslots, int B, int C, float* term, cudaStream_t stream);
void launch_decouple_sum(const float* terms, long long n, double* acc, cudaStream_t stream);
void launch_decouple_finalize(const double* acc, float* aux, double scale, cudaStream_t stream);

// GroupNorm(groups, C) + optional LeakyReLU(0.2) + optional residual add, one sample per CTA; NHWC with pixel strides
// Cs_in / Cs_out (>= C).  nn.GroupNorm semantics (eps inside the sqrt, affine).
void launch_groupnorm_act(const void* in, int in_dtype, void* out, int out_dtype, const void* add, int B, int HW, int C,
                          int Cs_in, int Cs_out, int groups, const float* gamma, const float* beta, float eps,
                          int act, cudaStream_t stream);

// Same operation with the sample staged in shared memory (one HBM read + one write).  in: fp32 dense [B][HW][C];
// out_kind 0: fp32 out; 1: bf16 out; 2: split-bf16 (out = high parts, out_lo = low parts); 3: fp16 out;
// add: optional fp32 dense
// residual added after the activation.
bool groupnorm_smem_supported(int HW, int C, int groups);
void launch_groupnorm_smem(const float* in, void* out, void* out_lo, int out_kind, const float* add, int B, int HW,
                           int C, int groups, const float* gamma, const float* beta, float eps, int act,
                           cudaStream_t stream);
// fp32 -> split-bf16: hi = bf16(v), lo = bf16(v - hi)
// GroupNorm whose partial statistics the producing conv's epilogue wrote: sums[b][slot][group][2] (sum, sum of squares)
bool groupnorm_apply_supported(int HW, int C, int groups);
void launch_groupnorm_apply(const void* in, int in_f16, void* out, int out_kind, const float* add, const float* sums, int nslots,
                            int B, int HW, int C, int groups, const float* gamma, const float* beta, float eps, int act, int num_sms,
                            cudaStream_t stream);
// per-(sample, frame) sum of squared errors (fp64), then per-frame sums over the batch of the SSE and of
// 10 * log10(SSE / chw); out = [frames SSE sums | frames log sums | batch]
void launch_metric_partial_sums(const float* pred, const float* target, int B, int P, long long chw, double* scratch,
                                double* out, cudaStream_t stream);
// SSIM (piqa defaults, restated): out[t] = sum over the batch of SSIM(pred[b, t], target[b, t]); scratch holds
// metric_ssim_scratch_elems() fp64 strip sums (-1: unsupported image size)
long long metric_ssim_scratch_elems(int B, int P, int C, int H, int W);
void launch_metric_ssim_sums(const float* pred, const float* target, int B, int P, int C, int H, int W, double* scratch,
                             double* out, cudaStream_t stream);
void launch_cast_f32_to_f16(const float* in, void* out, long long n, int num_sms, cudaStream_t stream);   // n % 4 == 0
void launch_split_bf16(const float* in, void* hi, void* lo, long long n, int num_sms, cudaStream_t stream);
// action-conditional models: actions fp32 [B, *, a] (sequence stride `bstride` elements), steps 0..T-1 ->
// out (activation type) [T][B][HW][a_pad], channels a..a_pad-1 zero (the action vector inflated to the frame size)
void launch_inflate_actions(const float* actions, long long bstride, int a, void* out, int dtype, int B, int T, int HW,
                            int a_pad, int num_sms, cudaStream_t stream);
// ST-Phy, action-conditional (st_phy.py:48-56, 146-148): actions fp32 [B, *, a] -> out (activation type) [T][B][H][W][C] =
// conv(5,1)(amap) + conv(1,5)(amap) with amap = Linear(action_t) as [IA][H][W]; wl [IA*H*W][a], wh / ww [C][IA][5] (device fp32)
void launch_stphy_action_tensor(const float* actions, long long bstride, int a, const float* wl, const float* wh, const float* ww,
                                void* out, int dtype, int B, int T, int H, int W, int C, int IA, cudaStream_t stream);
// TrajGRU (model_blocks/traj_gru.py): bilinear warps of the fp32 state h [B][H][W][C] by the L flow pairs in
// flows [B][H][W][fpix] (pair l at channels 2l, 2l + 1) -> out (activation type) [B][H][W][L * C]; and the GRU gate update
// from the raw i2h (nullable) / h2h pre-activations [P][3C] (fp32): h_out fp32 [P][C] + an activation-type copy
void launch_trajgru_warp(const float* h, const float* flows, int fpix, void* out, int dtype, int B, int H, int W, int C, int L,
                         int num_sms, cudaStream_t stream);
void launch_trajgru_gates(const float* i2h, const float* h2h, const float* h, float* h_out, void* h_act, int dtype, long long P,
                          int C, int act, int num_sms, cudaStream_t stream);
// in fp32 NHWC [B][H][W][C] -> typed copies: hi / lo (split pair of hi_dtype; either may be nullptr), cell (cell_dtype,
// nullable), f32 (nullable; may alias `in` only when norm_w is false); norm_w: L2-normalise along W first (eps clamp)
void launch_fanout(const float* in, void* hi, void* lo, int hi_dtype, void* cell, int cell_dtype, float* f32, int B, int H,
                   int W, int C, bool norm_w, float eps, int num_sms, cudaStream_t stream);
// out = x + y (y fp32 or nullptr: a plain conversion), x of x_dtype, out of out_dtype
void launch_add_to_act(const void* x, int x_dtype, const float* y, void* out, int out_dtype, long long n, int num_sms,
                       cudaStream_t stream);
// fp32 -> split-fp16: hi = fp16(v), lo = fp16(v - hi)   (n % 4 == 0)
void launch_split_f16(const float* in, void* hi, void* lo, long long n, int num_sms, cudaStream_t stream);

}  // namespace vpk
