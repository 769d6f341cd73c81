// Host-side lowering of conv / transposed-conv / gated-cell contractions to the generalised convolution launch,
// weight packing (reference OIHW fp32 -> [N_pad][K_pad] activation-type, gate-interleaved rows) and the device arena.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "conv_tc.h"

namespace vpk {

// A reference-layout weight tensor on the host.  conv: [O][I][KH][KW]; transposed conv: [I][O][KH][KW].
struct WeightRef {
  const float* w = nullptr;
  int O = 0, I = 0, KH = 0, KW = 0;
  bool transposed = false;
  // row block (of C rows) of this tensor that feeds packed gate g, or -1 if this tensor does not feed gate g
  int gate_block[4] = {0, -1, -1, -1};
};

struct BiasRef {
  const float* b = nullptr;
  int gate_block[4] = {0, -1, -1, -1};
};

struct HostStep {
  ConvStep s;
  int kw_valid;  // channels of this step that exist in the weight tensor (<= s.kc; the rest is zero padding)
  int wsplit;    // 0: weight value as is; 1 / 2: high / low 16-bit part of it (split operands)
  bool count_flops = true;   // false: an extra product of a split weight (not part of the algorithmic FLOP count)
  int wref;      // index into ConvSpec::wrefs
  int ky, kx;    // tap of that weight tensor
  int wc0;       // input-channel index of that weight tensor that corresponds to channel 0 of the source
  // sub-pixel transposed convs: the G = 4 "gates" are the four output parities, and each reads its OWN tap of the
  // weight tensor at this input offset (or none: -1)
  bool per_gate = false;
  signed char gky[4] = {-1, -1, -1, -1}, gkx[4] = {-1, -1, -1, -1};
};

struct PhaseSpec {
  int H = 0, W = 0;                 // output grid of this phase
  std::vector<HostStep> steps;
  EpiParams epi{};                  // per-phase (output offsets differ between transposed-conv parities)
};

struct ConvSpec {
  std::string name;
  int B = 0;
  int G = 1;                        // gates per channel
  int C = 0;                        // output channels per gate
  std::vector<SrcView> srcs;
  std::vector<WeightRef> wrefs;
  std::vector<BiasRef> biases;
  std::vector<PhaseSpec> phases;
  bool is_gate_gemm = false;
  int region_g0 = 0;                // request: accumulator regions [0, g0) / [g0, G) (ConvLaunch::region_g0); honoured when the
                                    // launch will run on CTA pairs of the tcgen05 halo kernel, ignored otherwise
};

struct ConvInput {                  // one logical input tensor of a conv (full-resolution NHWC view)
  SrcView view;
  int wref;                         // weight tensor it multiplies
  int wc0;                          // its channel offset inside that weight's input-channel dimension
  SrcView lo_view{nullptr, 0, 0, 0, 0, 0, 0};   // split-bf16 operand: `view` holds the high parts, `lo_view` the low parts
  int wc_count = -1;                // channels the weight really has for this input (-1: view.C); a view may carry
                                    // zero-padded extra channels (e.g. 49 -> 56 so that TMA strides are 16-byte)
  // weights only are split into a high and a low 16-bit part: A * W_hi + A * W_lo, two taps per weight tap over the SAME
  // activation tile (~22 weight mantissa bits; stride-1 convs; LayerNorm-fed ST-LSTM convs, see model_predrnn.cu)
  bool w_split = false;
  // split activations + weights used for PRECISION only (lo_view set): count the algorithmic FLOPs once, not three times
  bool extra_uncounted = false;
};

// Appends the K-steps of a k x k conv with the given stride (1 or 2) and padding over `inputs` to spec.phases[0]
// (creating it) and registers the (parity) source views.  Returns the output size through oh / ow.
void lower_conv(ConvSpec& spec, int k, int stride, int pad, const std::vector<ConvInput>& inputs, int in_h, int in_w,
                int esize, int* oh, int* ow);

// Transposed conv (stride 1 or 2): one phase per output parity.  `epi_for_phase(ry, rx, stride, OH, OW)` supplies the
// epilogue (strided output) of each phase.
void lower_conv_transpose(ConvSpec& spec, int k, int stride, int pad, int out_pad, const ConvInput& input, int in_h,
                          int in_w, int* oh, int* ow,
                          const std::function<EpiParams(int ry, int rx, int stride, int OH, int OW)>& epi_for_phase);

// Stride-2 transposed conv as ONE stride-1 conv over the input grid with four gate columns per channel = the four
// output parities ("sub-pixel" form): gate (ry, rx) of input position (q_y, q_x) is output pixel (2 q_y + ry, 2 q_x + rx).
// Every input offset of the union neighbourhood is one K-step whose weight rows are zero for the parities that do not
// use it, so the contraction is denser than the four per-parity launches (k4: 36 vs 16 tap-parity pairs, k3: 16 vs 9) --
// worth it only when the launches are latency-bound (small batches): one launch and one activation tile instead of four.
// Requires OH = 2 * in_h, OW = 2 * in_w (k4 p1, k3 p1 op1).
void lower_conv_transpose_subpixel(ConvSpec& spec, int k, int pad, int out_pad, const ConvInput& input, int in_h, int in_w,
                                   int* oh, int* ow);

// Simple bump allocator over a caller-provided (or library-owned) device range; base == nullptr only measures.
struct Arena {
  char* base = nullptr;
  size_t off = 0, cap = 0;
  void* alloc(size_t bytes, size_t align = 1024) {
    off = (off + align - 1) / align * align;
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (base && off > cap) VPK_THROW(4, "workspace too small");
    return p;
  }
  bool measuring() const { return base == nullptr; }
};

// Library-owned device buffers (packed weights, step tables, biases).
struct DeviceStore {
  std::vector<void*> ptrs;
  void* upload(const void* host, size_t bytes, cudaStream_t stream);
  void* zeros(size_t bytes, cudaStream_t stream);
  void release();
  ~DeviceStore() { release(); }
  // host staging kept alive until the upload stream has been synchronised
  std::vector<std::vector<char>> staging;
};

// A built phase: device-resident step table / packed weights / bias + the launch description.
struct BuiltConv {
  std::string name;
  ConvLaunch L;
  bool use_tc = false;
  bool use_direct = false;
  bool use_halo = false;
  TcPlan tc;
  HaloPlan halo;
};

// Packed-weight cache key -> device pointers, so the per-timestep launches of one layer share one packed copy.
struct PackedWeights {
  void* w = nullptr;
  float* bias = nullptr;
  ConvStep* steps = nullptr;
  int K_pad = 0, N_pad = 0, Cn = 0;
  // halo-kernel view of the same steps (block-major): device tables
  HaloBlock* blocks = nullptr;
  HaloTap* taps = nullptr;
  int nblocks = 0, ntaps = 0, radius = 0;
};

// N tiles the tensor-core kernels split `C` output channels with `G` gates per channel into (the tile rule of build_conv).
int conv_n_tiles(int C, int G);

// Packs the weights of every phase of `spec` for activation type `dtype` and returns one BuiltConv per phase (device
// pointers of the sources / outputs inside spec must already be final unless `measure_only`).
std::vector<BuiltConv> build_conv(const ConvSpec& spec, int dtype, int backend, DeviceStore& store,
                                  std::map<std::string, std::vector<PackedWeights>>& cache, cudaStream_t stream,
                                  int num_sms, bool measure_only);

}  // namespace vpk
