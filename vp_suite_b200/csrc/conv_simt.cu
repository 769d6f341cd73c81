// CUDA-core implicit-GEMM kernel for the generalised convolution launch (common.h).
//
// Role: (1) the fp32-operand mode (matches the reference to <= 1e-4), (2) the few layers whose K is too small or
// whose channel count is not TMA-addressable (image-channel convs), (3) an operand-identical cross-check of the
// tcgen05 kernel (same bf16 inputs and packed weights, fp32 accumulation).
//
// Tile: 128 output positions x (16 * TN) packed columns per 256-thread CTA, K consumed in 32-wide chunks through
// shared memory; each thread owns 8 positions x TN consecutive columns = TN/G whole channels with all their gates,
// so the fused epilogue (epilogue.cuh) runs on registers.
#include "common.h"
#include "epilogue.cuh"

namespace vpk {

namespace {

constexpr int kTM = 128;      // positions per CTA
constexpr int kBK = 32;       // K chunk
constexpr int kThreads = 256;

template <typename T, int G, int TN>
__global__ void __launch_bounds__(kThreads) conv_simt_kernel(const ConvLaunch L) {
  constexpr int NT = 16 * TN;          // columns per CTA
  constexpr int CPT = TN / G;          // channels per thread
  __shared__ float As[kBK][kTM + 4];
  __shared__ float Bs[kBK][NT + 4];
  __shared__ ConvStep s_steps[kMaxSteps];

  const int tid = threadIdx.x;
  for (int i = tid; i < L.nsteps; i += kThreads) s_steps[i] = L.steps[i];
  __syncthreads();

  const int HW = L.H * L.W;
  const long long M = static_cast<long long>(L.B) * HW;
  const long long m0 = static_cast<long long>(blockIdx.x) * kTM;
  const int n0 = blockIdx.y * NT;

  // loader roles: A: position tid>>1, 16 channels at (tid&1)*16;  B: rows r, r+128 (if NT > 128), same k half
  const int lp = tid >> 1;
  const int lk = (tid & 1) * 16;
  const long long lm = m0 + lp;
  int lb = 0, ly = 0, lx = 0;
  const bool lvalid = lm < M;
  if (lvalid) {
    lb = static_cast<int>(lm / HW);
    const int r = static_cast<int>(lm - static_cast<long long>(lb) * HW);
    ly = r / L.W;
    lx = r - ly * L.W;
  }

  const int tm = tid & 15;             // positions tm + 16*i
  const int tn = tid >> 4;             // columns tn*TN .. +TN
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const T* wp = static_cast<const T*>(L.wpacked);

  for (int s = 0; s < L.nsteps; ++s) {
    const ConvStep st = s_steps[s];
    const SrcView& sv = L.src[st.src];
    const int sy = ly + st.dy, sx = lx + st.dx;
    const bool inb = lvalid && sy >= 0 && sy < sv.H && sx >= 0 && sx < sv.W;
    const T* ap = static_cast<const T*>(sv.base) + lb * sv.sB + sy * sv.sY + sx * sv.sX + st.c0;
    for (int kk = 0; kk < st.kc; kk += kBK) {
      // ---- A tile: [kBK][kTM] (k-major) ----
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = kk + lk + j;
        float v = 0.f;
        if (inb && k < st.kc) v = to_f32(ap[k]);
        As[lk + j][lp] = v;
      }
      // ---- B tile: [kBK][NT] ----
      for (int r = lp; r < NT; r += kThreads / 2) {
        const int n = n0 + r;
        const T* bp = wp + static_cast<size_t>(n) * L.K_pad + st.wk + kk + lk;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = 0.f;
          if (n < L.N_pad && kk + lk + j < st.kc) v = to_f32(bp[j]);
          Bs[lk + j][r] = v;
        }
      }
      __syncthreads();
      const int kmax = min(kBK, st.kc - kk);
      for (int k = 0; k < kmax; ++k) {
        float a[8], bv[TN];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = As[k][tm + 16 * i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tn * TN + j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- fused epilogue ----
  const int ch0 = (n0 + tn * TN) / G;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + tm + 16 * i;
    if (m >= M) continue;
    const int b = static_cast<int>(m / HW);
    const int r = static_cast<int>(m - static_cast<long long>(b) * HW);
    const int y = r / L.W, x = r - y * L.W;
    float g[G][CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
      for (int q = 0; q < G; ++q) g[q][c] = acc[i][c * G + q];
    epilogue_apply<T, G, CPT>(L.epi, b, y, x, L.H, L.W, ch0, g);
  }
}

template <typename T, int G, int TN> void launch_t(const ConvLaunch& L, cudaStream_t stream) {
  constexpr int NT = 16 * TN;
  const long long M = static_cast<long long>(L.B) * L.H * L.W;
  dim3 grid(static_cast<unsigned>((M + kTM - 1) / kTM), static_cast<unsigned>((L.N_pad + NT - 1) / NT));
  conv_simt_kernel<T, G, TN><<<grid, kThreads, 0, stream>>>(L);
}

template <typename T> void launch_g(const ConvLaunch& L, cudaStream_t stream) {
  switch (L.G) {
    case 1: launch_t<T, 1, 8>(L, stream); break;
    case 2: launch_t<T, 2, 8>(L, stream); break;
    case 3: launch_t<T, 3, 6>(L, stream); break;
    case 4: launch_t<T, 4, 8>(L, stream); break;
    default: VPK_THROW(1, "conv_simt: unsupported gate count");
  }
}

}  // namespace

void launch_conv_simt(const ConvLaunch& L, int dtype, cudaStream_t stream) {
  VPK_REQUIRE(L.nsteps > 0 && L.nsteps <= kMaxSteps, "conv_simt: bad step count");
  if (dtype == DT_F32) launch_g<float>(L, stream);
  else if (dtype == DT_F16) launch_t<__half, 1, 8>(L, stream);     // fp16 operands exist for plain convs only (G = 1)
  else launch_g<__nv_bfloat16>(L, stream);
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
