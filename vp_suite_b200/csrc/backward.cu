// Backward kernels of the ndrplz ConvLSTM cell (backward.h).  fp32 CUDA-core kernels: this is the correctness-first
// differentiable entry (gradient-checked against the reference's autograd); the input gradient (a transposed conv over dz)
// reuses the generalised-conv launch and so runs on the tensor cores in 16-bit mode.
#include "backward.h"

#include "ptx.cuh"

namespace vpk {

namespace {

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + __expf(-v)); }

// one thread = one (position, channel)
template <typename T>
__global__ void __launch_bounds__(256) lstm_gate_backward_kernel(const float* __restrict__ z, const float* __restrict__ c,
                                                                 const float* __restrict__ dh, const float* __restrict__ dcn,
                                                                 float* __restrict__ dz, T* __restrict__ dz_act,
                                                                 float* __restrict__ dc_in, long long P, int C) {
  const long long total = P * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = idx / C;
    const int ch = static_cast<int>(idx - p * C);
    const float* zp = z + p * 4 * C;
    // forward (conv_lstm_ndrplz.py:34-41): split order (i, f, o, g)
    const float i = sigmoid_acc(zp[ch]), f = sigmoid_acc(zp[C + ch]), o = sigmoid_acc(zp[2 * C + ch]), g = tanhf(zp[3 * C + ch]);
    const float cp = c[idx];
    const float cn = f * cp + i * g;
    const float tc = tanhf(cn);
    const float gh = dh != nullptr ? dh[idx] : 0.f;
    const float dct = (dcn != nullptr ? dcn[idx] : 0.f) + gh * o * (1.f - tc * tc);      // dL/dc_next, all paths
    const float dzi = dct * g * i * (1.f - i);
    const float dzf = dct * cp * f * (1.f - f);
    const float dzo = gh * tc * o * (1.f - o);
    const float dzg = dct * i * (1.f - g * g);
    float* dp = dz + p * 4 * C;
    dp[ch] = dzi;
    dp[C + ch] = dzf;
    dp[2 * C + ch] = dzo;
    dp[3 * C + ch] = dzg;
    if (dz_act != nullptr) {
      T* ap = dz_act + p * 4 * C;
      ap[ch] = static_cast<T>(dzi);
      ap[C + ch] = static_cast<T>(dzf);
      ap[2 * C + ch] = static_cast<T>(dzo);
      ap[3 * C + ch] = static_cast<T>(dzg);
    }
    dc_in[idx] = dct * f;
  }
}

// Shi et al. ConvLSTM with peepholes (conv_lstm_hzzone.py:57-69): split order (i, f, g, o);
//   i = s(z_i + Wci c), f = s(z_f + Wcf c), c' = f c + i tanh(z_g), o = s(z_o + Wco c'), h = o tanh(c').
// One thread = one (position of the image, channel); it walks the batch, so the peephole gradients (sums over the
// batch) are accumulated in a fixed order without atomics.
template <typename T>
__global__ void __launch_bounds__(256) lstm_peep_gate_backward_kernel(const float* __restrict__ z, const float* __restrict__ c,
                                                                      const float* __restrict__ wci, const float* __restrict__ wcf,
                                                                      const float* __restrict__ wco, const float* __restrict__ dh,
                                                                      const float* __restrict__ dcn, float* __restrict__ dz,
                                                                      T* __restrict__ dz_act, float* __restrict__ dc_in,
                                                                      float* __restrict__ dwci, float* __restrict__ dwcf,
                                                                      float* __restrict__ dwco, int B, long long HW, int C) {
  const long long total = HW * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long hw = idx / C;
    const int ch = static_cast<int>(idx - hw * C);
    const float pi = wci ? wci[idx] : 0.f, pf = wcf ? wcf[idx] : 0.f, po = wco ? wco[idx] : 0.f;
    float gi = 0.f, gf = 0.f, go = 0.f;
    for (int b = 0; b < B; ++b) {
      const long long p = b * HW + hw;
      const float* zp = z + p * 4 * C;
      const float cp = c[p * C + ch];
      const float i = sigmoid_acc(zp[ch] + pi * cp), f = sigmoid_acc(zp[C + ch] + pf * cp), g = tanhf(zp[2 * C + ch]);
      const float cn = f * cp + i * g;
      const float o = sigmoid_acc(zp[3 * C + ch] + po * cn);
      const float tc = tanhf(cn);
      const float gh = dh != nullptr ? dh[p * C + ch] : 0.f;
      const float dzo = gh * tc * o * (1.f - o);
      const float dct = (dcn != nullptr ? dcn[p * C + ch] : 0.f) + gh * o * (1.f - tc * tc) + dzo * po;    // dL/dc', all paths
      const float dzi = dct * g * i * (1.f - i);
      const float dzf = dct * cp * f * (1.f - f);
      const float dzg = dct * i * (1.f - g * g);
      float* dp = dz + p * 4 * C;
      dp[ch] = dzi;
      dp[C + ch] = dzf;
      dp[2 * C + ch] = dzg;
      dp[3 * C + ch] = dzo;
      if (dz_act != nullptr) {
        T* ap = dz_act + p * 4 * C;
        ap[ch] = static_cast<T>(dzi);
        ap[C + ch] = static_cast<T>(dzf);
        ap[2 * C + ch] = static_cast<T>(dzg);
        ap[3 * C + ch] = static_cast<T>(dzo);
      }
      dc_in[p * C + ch] = dct * f + dzi * pi + dzf * pf;
      gi = fmaf(dzi, cp, gi);
      gf = fmaf(dzf, cp, gf);
      go = fmaf(dzo, cn, go);
    }
    if (dwci) dwci[idx] = gi;
    if (dwcf) dwcf[idx] = gf;
    if (dwco) dwco[idx] = go;
  }
}

// grid (ceil(Co / 64), ceil(Ci / 64), k * k); 256 threads = 16 x 16, each a 4 x 4 block of (o, i) for one tap
constexpr int kWgTile = 64, kWgPos = 16;
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz,
                                                         float* __restrict__ dw, int B, int H, int W, int Ci, int Co, int k,
                                                         int Ci_total, int i0) {
  __shared__ float s_dz[kWgPos][kWgTile + 1];
  __shared__ float s_in[kWgPos][kWgTile + 1];
  const int o0 = blockIdx.x * kWgTile, c0 = blockIdx.y * kWgTile;
  const int tap = blockIdx.z, ky = tap / k, kx = tap - ky * k, pad = k / 2;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long P = static_cast<long long>(B) * H * W;
  for (long long p0 = 0; p0 < P; p0 += kWgPos) {
    // stage 16 positions x 64 channels of dz and of the (shifted) input: 1024 values each, 4 per thread
    for (int e = threadIdx.x; e < kWgPos * kWgTile; e += 256) {
      const int pp = e / kWgTile, cc = e - pp * kWgTile;
      const long long p = p0 + pp;
      float vz = 0.f, vi = 0.f;
      if (p < P) {
        if (o0 + cc < Co) vz = dz[p * Co + o0 + cc];
        const int x = static_cast<int>(p % W), y = static_cast<int>((p / W) % H);
        const int ys = y + ky - pad, xs = x + kx - pad;
        if (c0 + cc < Ci && ys >= 0 && ys < H && xs >= 0 && xs < W)
          vi = in[(p + static_cast<long long>(ky - pad) * W + (kx - pad)) * Ci + c0 + cc];
      }
      s_dz[pp][cc] = vz;
      s_in[pp][cc] = vi;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kWgPos; ++pp) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = s_dz[pp][ty * 4 + q];
        b[q] = s_in[pp][tx * 4 + q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(a[q], b[r], acc[q][r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = o0 + ty * 4 + q, i = c0 + tx * 4 + r;
      if (o < Co && i < Ci) dw[((static_cast<long long>(o) * Ci_total + i0 + i) * k + ky) * k + kx] = acc[q][r];
    }
}

// grid ceil(Co / 32); 256 threads = 32 channels x 8 position lanes; fixed-order reduction
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dz, float* __restrict__ db, long long P, int Co) {
  __shared__ float s[8][33];
  const int ch = blockIdx.x * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
  float acc = 0.f;
  if (ch < Co)
    for (long long p = lane; p < P; p += 8) acc += dz[p * Co + ch];
  s[lane][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && ch < Co) {
    float t = 0.f;
    for (int l = 0; l < 8; ++l) t += s[l][threadIdx.x];
    db[ch] = t;
  }
}

}  // namespace

void launch_lstm_gate_backward(const float* z, const float* c, const float* dh_out, const float* dc_out, float* dz,
                               void* dz_act, int act_dtype, float* dc_in, long long P, int C, int num_sms,
                               cudaStream_t stream) {
  VPK_REQUIRE(z && c && dz && dc_in && P > 0 && C > 0, "lstm_gate_backward: bad arguments");
  const long long total = P * C;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 32ll * num_sms));
  if (dz_act == nullptr || act_dtype == DT_F32)
    lstm_gate_backward_kernel<float><<<grid, 256, 0, stream>>>(z, c, dh_out, dc_out, dz, nullptr, dc_in, P, C);
  else if (act_dtype == DT_F16)
    lstm_gate_backward_kernel<__half><<<grid, 256, 0, stream>>>(z, c, dh_out, dc_out, dz, static_cast<__half*>(dz_act), dc_in, P, C);
  else
    lstm_gate_backward_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(z, c, dh_out, dc_out, dz, static_cast<__nv_bfloat16*>(dz_act), dc_in, P, C);
  VPK_CUDA(cudaGetLastError());
}

void launch_lstm_peep_gate_backward(const float* z, const float* c, const float* wci, const float* wcf, const float* wco,
                                    const float* dh_out, const float* dc_out, float* dz, void* dz_act, int act_dtype,
                                    float* dc_in, float* dwci, float* dwcf, float* dwco, int B, long long HW, int C, int num_sms,
                                    cudaStream_t stream) {
  VPK_REQUIRE(z && c && dz && dc_in && B > 0 && HW > 0 && C > 0, "lstm_peep_gate_backward: bad arguments");
  const long long total = HW * C;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 32ll * num_sms));
  if (dz_act == nullptr || act_dtype == DT_F32)
    lstm_peep_gate_backward_kernel<float><<<grid, 256, 0, stream>>>(z, c, wci, wcf, wco, dh_out, dc_out, dz, nullptr, dc_in, dwci, dwcf, dwco, B, HW, C);
  else if (act_dtype == DT_F16)
    lstm_peep_gate_backward_kernel<__half><<<grid, 256, 0, stream>>>(z, c, wci, wcf, wco, dh_out, dc_out, dz, static_cast<__half*>(dz_act), dc_in, dwci, dwcf, dwco, B, HW, C);
  else
    lstm_peep_gate_backward_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(z, c, wci, wcf, wco, dh_out, dc_out, dz, static_cast<__nv_bfloat16*>(dz_act), dc_in, dwci, dwcf, dwco, B, HW, C);
  VPK_CUDA(cudaGetLastError());
}

void launch_conv_wgrad(const float* in, const float* dz, float* dw, int B, int H, int W, int Ci, int Co, int k, int Ci_total,
                       int i0, cudaStream_t stream) {
  VPK_REQUIRE(in && dz && dw && B > 0 && Ci > 0 && Co > 0 && k % 2 == 1, "conv_wgrad: bad arguments");
  const dim3 grid((Co + kWgTile - 1) / kWgTile, (Ci + kWgTile - 1) / kWgTile, k * k);
  conv_wgrad_kernel<<<grid, 256, 0, stream>>>(in, dz, dw, B, H, W, Ci, Co, k, Ci_total, i0);
  VPK_CUDA(cudaGetLastError());
}

void launch_bias_grad(const float* dz, float* db, long long P, int Co, cudaStream_t stream) {
  VPK_REQUIRE(dz && db && P > 0 && Co > 0, "bias_grad: bad arguments");
  bias_grad_kernel<<<(Co + 31) / 32, 256, 0, stream>>>(dz, db, P, Co);
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
