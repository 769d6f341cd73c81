// Direct kernel for the memory-bound small convs of the path (image-channel stem convs, 1x1 frame heads, the last
// transposed conv): total K <= 128 and N <= 32, so one thread keeps a whole output position (all N channels) in
// registers, the packed weights sit in shared memory (every lane reads the same weight: broadcast), inputs are read
// straight from global/L1 (neighbouring lanes read neighbouring pixels) and outputs are written as one contiguous
// vector per thread.  Same K-step table, packed weights and fused epilogue as the two GEMM kernels.
#include "common.h"
#include "conv_tc.h"
#include "epilogue.cuh"

namespace vpk {

namespace {

constexpr int kDirectThreads = 256;
constexpr int kDirectMaxK = 256;
constexpr int kDirectMaxN = 32;
constexpr int kDirectMaxSteps = 64;

template <typename T, int NB>   // NB = N_pad / 8 blocks of 8 channels
__global__ void __launch_bounds__(kDirectThreads) conv_direct_kernel(const ConvLaunch L) {
  constexpr int N = NB * 8;
  __shared__ float s_w[kDirectMaxK][N];
  __shared__ ConvStep s_steps[kDirectMaxSteps];
  const T* wp = static_cast<const T*>(L.wpacked);
  for (int i = threadIdx.x; i < L.K_pad * N; i += kDirectThreads) {
    const int k = i / N, n = i - k * N;
    s_w[k][n] = (k < L.K_pad) ? to_f32(wp[static_cast<size_t>(n) * L.K_pad + k]) : 0.f;
  }
  for (int i = threadIdx.x; i < L.nsteps; i += kDirectThreads) s_steps[i] = L.steps[i];
  __syncthreads();

  const int HW = L.H * L.W;
  const long long M = static_cast<long long>(L.B) * HW;
  for (long long m = blockIdx.x * static_cast<long long>(kDirectThreads) + threadIdx.x; m < M;
       m += static_cast<long long>(gridDim.x) * kDirectThreads) {
    const int b = static_cast<int>(m / HW);
    const int r = static_cast<int>(m - static_cast<long long>(b) * HW);
    const int y = r / L.W, x = r - y * L.W;
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = 0.f;
    for (int s = 0; s < L.nsteps; ++s) {
      const ConvStep st = s_steps[s];
      const SrcView& sv = L.src[st.src];
      const int sy = y + st.dy, sx = x + st.dx;
      if (sy < 0 || sy >= sv.H || sx < 0 || sx >= sv.W) continue;
      const T* ap = static_cast<const T*>(sv.base) + b * sv.sB + sy * sv.sY + sx * sv.sX + st.c0;
      for (int j = 0; j < st.kc; ++j) {
        const float a = to_f32(ap[j]);
        const float* wr = s_w[st.wk + j];
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = fmaf(a, wr[n], acc[n]);
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      float g[1][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[0][j] = acc[q * 8 + j];
      epilogue_apply<T, 1, 8>(L.epi, b, y, x, L.H, L.W, q * 8, g);
    }
  }
}

template <typename T> void launch_n(const ConvLaunch& L, int grid, cudaStream_t stream) {
  switch (L.N_pad / 8) {
    case 1: conv_direct_kernel<T, 1><<<grid, kDirectThreads, 0, stream>>>(L); break;
    case 2: conv_direct_kernel<T, 2><<<grid, kDirectThreads, 0, stream>>>(L); break;
    case 3: conv_direct_kernel<T, 3><<<grid, kDirectThreads, 0, stream>>>(L); break;
    case 4: conv_direct_kernel<T, 4><<<grid, kDirectThreads, 0, stream>>>(L); break;
    default: VPK_THROW(1, "conv_direct: unsupported N");
  }
}

}  // namespace

bool direct_eligible(const ConvLaunch& L) {
  return L.G == 1 && L.epi.kind == EPI_BIAS_ACT && L.N_pad % 8 == 0 && L.N_pad <= kDirectMaxN &&
         L.K_pad <= kDirectMaxK && L.nsteps <= kDirectMaxSteps;
}

void launch_conv_direct(const ConvLaunch& L, int dtype, int num_sms, cudaStream_t stream) {
  const long long M = static_cast<long long>(L.B) * L.H * L.W;
  const long long blocks = (M + kDirectThreads - 1) / kDirectThreads;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(blocks, static_cast<long long>(num_sms) * 8)));
  if (dtype == DT_F32) launch_n<float>(L, grid, stream);
  else if (dtype == DT_F16) launch_n<__half>(L, grid, stream);
  else launch_n<__nv_bfloat16>(L, grid, stream);
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
