// tcgen05 / TMEM / TMA implicit-GEMM kernel for the generalised convolution launch: host-side plan.
#pragma once
#include "common.h"

namespace vpk {

struct alignas(64) TcPlan {
  CUtensorMap amap[kMaxSrc];   // 4-D (C, W, H, B) views of the activation sources, box (64, TW, TH, TB), 128B swizzle
  CUtensorMap bmap;            // 2-D (K_pad, N_pad) packed weights, box (64, tileN), 128B swizzle
  ConvLaunch L;
  int TW, TH, TB;              // output positions per tile: TW*TH*TB = 128
  int tiles_x, tiles_y, tiles_b;
  int n_tiles;                 // tiles along N
  int tileN;                   // Cn * G, multiple of 16, <= 256
  int stages;                  // smem ring depth
  int tmem_cols;               // power of two >= 2 * tileN
  unsigned smem_bytes;
  int grid;
  int debug;                   // VPK_TC_DEBUG bit mask (perf experiments only; results are wrong when set):
                               //   1 = skip epilogue math/IO, 2 = skip MMA issue, 4 = skip TMA loads
  int cta2;                    // 1: CTA-pair kernel (conv_tc2.cu): weight box holds tileN/2 rows, grid is even
};

// True when the launch can run on the tensor-core kernel (bf16, channel counts TMA-addressable, ...).
bool tc_eligible(const ConvLaunch& L, int dtype);
// Fills tensor maps and tiling for a launch whose device pointers are final.
void tc_make_plan(const ConvLaunch& L, TcPlan* plan, int num_sms);
void launch_conv_tc(const TcPlan& plan, cudaStream_t stream);   // dispatches on plan.cta2
void launch_conv_tc2(const TcPlan& plan, cudaStream_t stream);

void launch_conv_simt(const ConvLaunch& L, int dtype, cudaStream_t stream);

// Direct kernel for small, memory-bound convs (total K <= 128, N <= 32): conv_direct.cu
bool direct_eligible(const ConvLaunch& L);
void launch_conv_direct(const ConvLaunch& L, int dtype, int num_sms, cudaStream_t stream);

}  // namespace vpk
