// Shared host/device definitions of libvpk (the generalised convolution "launch" every hot-path op lowers to).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

namespace vpk {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define VPK_THROW(code, msg) throw ::vpk::Error((code), std::string(msg))
#define VPK_REQUIRE(cond, msg)                                              \
  do {                                                                      \
    if (!(cond)) VPK_THROW(1, std::string(msg) + " [" #cond "]");           \
  } while (0)
#define VPK_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      VPK_THROW(2, std::string(#expr) + " failed: " + cudaGetErrorString(e_) + " (" __FILE__ ":" +       \
                       std::to_string(__LINE__) + ")");                                                  \
  } while (0)

// DT_F16: fp16 conv OPERANDS (11 mantissa bits; PhyDNet's GroupNorm-fed encoder/decoder convs, whose feature maps are
// O(1) after GroupNorm + LeakyReLU).  Such launches read fp16 activations / packed weights and write fp32 only.
enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
__host__ __device__ inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------------------------------
// Generalised convolution launch.
//
// Output positions form a grid (B, H, W).  The contraction runs over a list of K-steps; step s multiplies the
// channels [c0, c0+kc) of source view `src`, read at spatial offset (dy, dx) from the output position (zero outside
// the view), with columns [wk, wk+kc) of the packed weight matrix Wp[N_pad][K_pad].  Ordinary convs, the concat-free
// gate conv of the recurrent cells (x, h, m are separate sources), stride-2 convs (4 parity views of the input) and
// transposed convs (one launch per output parity, strided output) are all instances of this.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxSrc = 8;
constexpr int kMaxSteps = 256;

struct SrcView {          // NHWC view, channels contiguous; strides in elements
  const void* base;
  int H, W, C;
  long long sB, sY, sX;
};

struct ConvStep {         // 12 bytes
  short src;
  signed char dy, dx;
  short c0;               // first channel of the source covered by this step
  short kc;               // valid channels (<= 64)
  int wk;                 // column offset in the packed weights (multiple of 16)
};

enum EpiKind : int {
  EPI_BIAS_ACT = 0,       // G=1  y = act(acc + bias)
  EPI_LSTM = 1,           // G=4  (i,f,g,o) ConvLSTM update with optional peepholes
  EPI_ST_C = 2,           // G=4  (i,f,g,o_part) ST-LSTM temporal-memory update
  EPI_ST_M = 3,           // G=3  (i',f',g') ST-LSTM spatio-temporal-memory update
  EPI_ST_O = 4,           // G=2  (conv_o, conv_last) ST-LSTM output gate
  EPI_PHY_GATE = 5,       // G=1  PhyCell Kalman-style blend
  EPI_SUBPIX = 7,         // G=4  sub-pixel transposed conv: gate (ry, rx) of a position is output pixel (2y + ry, 2x + rx);
                          //      y = act(acc + bias) stored as activation type at out + b*oB + y*oY + x*oX + ry*ps_row + rx*C + ch
  EPI_ST_O1 = 8,          // G=1  the output gate of ST-LSTM / Causal LSTM split over two launches: a plain G = 1 conv leaves one of
                          //      (conv_o(mem), conv_last(mem)) as raw fp32 dense NHWC in `res`, this launch computes the other and
                          //      h' = gate(o_part + conv_o) * tanh(conv_last), s0 = o_part (state layout).  variant bit 0: tanh gate
                          //      (Causal LSTM) instead of sigmoid; bit 1: acc = conv_o and res = conv_last (the form the rollouts
                          //      use: the memory-bound reads then hide behind the k x k MMAs), else the other way round.
                          //      The fused EPI_ST_O launch spends half of its k x k MMAs on the zero rows of conv_last's gate
                          //      column; worth two launches once the layer is tensor-bound.  Same sums, same order: bit-identical
  EPI_DECOUPLE = 6,       // G=2  (adapter(delta_c), adapter(delta_m)): per-(sample, channel) dot product and squared norms
                          //      over the positions (PredRNN-V2 decoupling loss); nothing else is stored.  tcgen05 halo
                          //      kernel only: s1[b][slot][C][3] gets one warp's 32 positions per slot (slot as gn_slot0 / gn_nslots)
};
enum ActKind : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_SIGMOID = 2, ACT_RELU = 3 };

struct EpiParams {
  int kind;
  int C;                  // real output channels per gate
  int act;
  int out_f32;            // primary output element type: 0 = activation type T, 1 = float
  int out_f16;            // tcgen05 halo kernel, lean BIAS_ACT epilogue only: the 16-bit output is fp16, not bf16
  const float* bias;      // packed order [N_pad] or nullptr
  // primary output (BIAS_ACT: y; LSTM / ST_O / PHY_GATE: h'), address = out + b*oB + y*oY + x*oX + ch*oC
  void* out;
  long long oB, oY, oX, oC;
  // fp32 state tensors, dense NHWC [B, H, W, C]
  float* s0;              // LSTM: c (in/out)   ST_C: c (in/out)   ST_M: m (in/out)   ST_O: o_part (in)
  float* s1;              // ST_C: o_part (out)
  // peepholes, fp32 [H, W, C] (LSTM) or nullptr
  const float *p0, *p1, *p2;
  // LSTM, tcgen05 path only: the same three peepholes as bf16, packed [C/8][H][W][3][8] (48 contiguous bytes per
  // position and 8-channel chunk, positions contiguous: coalesced); nullptr: use p0..p2
  const void* pp16;
  // secondary activation-type outputs
  void* t0;               // ST_C / ST_M: mem buffer [B,H,W,2C] (channel offset applied by caller)
  long long t0_pix;       // elements per pixel of t0 (2C)
  void* t1;               // ST_C: delta_c, ST_M: delta_m   dense [B,H,W,C]
  // fp32 dense [B,H,W,C]: BIAS_ACT: residual added after the activation; PHY_GATE: h~ = h + F(h)
  const float* res;
  // PHY_GATE: x (frame), activation type, dense [B,H,W,C]
  const void* q0;
  // LSTM: optional fp32 copy of h' (dense), for consumers that run in fp32
  float* h32;
  float forget_bias;
  // 1: the fp32 tensors only the epilogue touches (s0, s1, p0..p2) use the channel-quad layout [B][C/4][H][W][4]
  // instead of NHWC, so that the 32 positions of a warp read/write contiguous 16-byte pieces (coalesced); needs C % 4 == 0
  int state_c4;
  int debug;              // perf experiments only (VPK_TC_DEBUG): 8 = skip activation-type stores, 16 = skip bias loads
  // BIAS_ACT on the tcgen05 path only: a fused 1x1 projection of the activated channels to proj_n <= 4 outputs,
  // z[p] = proj_b[p] + sum_ch proj_w[p][ch] * act(acc[ch] + bias[ch]); `out` is then the fp32 strided (NCHW frame) tensor
  // of z and the C-channel intermediate never reaches memory (EF forecaster: last deconv + final 1x1 conv)
  const float* proj_w;    // device fp32 [proj_n][C]
  const float* proj_b;    // device fp32 [proj_n] or nullptr
  int proj_n;
  // optional statistics for a following GroupNorm, written by the tcgen05 halo kernel's BIAS_ACT epilogue:
  // gn_sums[b][slot][group][2] = (sum, sum of squares) of one warp's 32 positions; slot = gn_slot0 + (tile index inside
  // the image) * 4 + TMEM lane quadrant.  Every slot is written exactly once per launch (plain stores, no atomics), so
  // the apply pass adds them in a fixed order: results do not depend on the schedule.
  float* gn_sums;
  int gn_group_size;      // channels per group (0 = disabled); -1 = ONE group = the whole sample (LayerNorm over C, H, W):
                          //   gn_sums[b][slot][2] with slot = gn_slot0 + ((tile in image) * n_tiles + N tile) * 8 + quadrant * 2 + half
  int gn_slot0, gn_nslots;
  long long ps_row;       // EPI_SUBPIX: elements between two output rows (OW * C)
  // Causal LSTM / GHU (PredRNN++, causal.h) reuse the ST-LSTM epilogue instantiations with a warp-uniform variant:
  //   EPI_ST_C, variant 1: acc = (i', f', g', m_m)   m' = sig(f' + forget_bias) tanh(m_m) + sig(i') tanh(g'); stores m' to s0
  //                        (fp32, no read) and t0 (activation copy); s1 / t1 untouched
  //   EPI_ST_O, variant 1: acc = (o - o_part, last)  h' = tanh(o_part + acc0) * tanh(acc1)
  //   EPI_ST_O, variant 2: acc = (p, u), s0 = z      z' = sig(u) z + (1 - sig(u)) tanh(p); stores z' to s0 and `out`
  int variant;
};

struct ConvLaunch {
  // problem
  int B, H, W;            // output grid
  int nsrc;
  SrcView src[kMaxSrc];
  int nsteps;
  const ConvStep* steps;  // device pointer
  const void* wpacked;    // device, activation type, [N_pad][K_pad]
  int K_pad, N_pad;
  int G;                  // gates per channel; packed row n = ch*G + gate
  int Cn;                 // channels per N tile for the tensor-core kernel (tile N = Cn*G)
  EpiParams epi;
  // bookkeeping
  int op_f16;             // 16-bit operands are fp16 instead of bf16 (tensor-core kernels: instruction descriptor)
  double flops;           // 2*M*N*K of the real (unpadded) contraction
  int is_gate_gemm;       // counted in the gate-GEMM roofline figure
  // tcgen05 halo kernel, CTA pairs only: > 0 splits the G gate columns of a channel into two accumulator REGIONS, gates
  // [0, region_g0) and [region_g0, G), packed region-major inside each CTA's half of an N tile; a tap then issues only the
  // regions its weight tensor feeds (conv_halo.cu).  Multi-source gate convs whose sources feed different gates
  // (Causal LSTM: m_m is fed by m alone; ST-LSTM output: conv_last is 1 x 1) otherwise multiply zero weight rows.
  int region_g0;
};

// Environment switches.  Everything the shipping library reads with getenv() selects between code paths that compute
// the same frames (A/B runs; each alternative has a GPU test).  Switches that CHANGE results or skip work -- perf
// experiments (VPK_EXP_NO_PEEPHOLES, VPK_TC_DEBUG) -- exist only in developer builds (make EXTRA_FLAGS=-DVPK_DEV): in the
// product build dev_env() is constant nullptr, so a stray export on a box cannot silently void a number.
inline const char* dev_env(const char* name) {
#ifdef VPK_DEV
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// VPK_PDL=0 turns programmatic dependent launch off (A/B runs).
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* env = getenv("VPK_PDL");
    return env == nullptr || atoi(env) != 0;
  }();
  return on;
}

#ifdef __CUDACC__
// Launch of a kernel that starts with ptx::pdl_launch_dependents() / ptx::pdl_wait() (see ptx.cuh): with PDL enabled its
// CTAs may become resident (and run whatever precedes their pdl_wait) while the previous kernel of the stream drains.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  VPK_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
#endif

}  // namespace vpk
