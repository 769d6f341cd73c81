// Memory-bound helper kernels: boundary layout conversion (fp32 NCHW frames <-> NHWC activations), PredRNN
// patchify / un-patchify, GroupNorm (+LeakyReLU), decouple-loss reduction.  Vectorised, coalesced on the side that
// dominates the traffic; grid-stride loops sized in multiples of the SM count.
#include "common.h"
#include "elementwise.h"
#include "epilogue.cuh"

namespace vpk {

namespace {

inline int grid_for(long long n, int threads, int num_sms) {
  long long blocks = (n + threads - 1) / threads;
  long long cap = static_cast<long long>(num_sms) * 16;
  return static_cast<int>(std::max<long long>(1, std::min(blocks, cap)));
}

// x fp32 [B, T, C, H, W] (one microbatch) -> out T [T][B][H][W][C]
template <typename T>
__global__ void frames_to_nhwc_kernel(const float* __restrict__ x, long long bstride, T* __restrict__ out, int B,
                                      int Tn, int C, int H, int W) {
  const long long HW = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(B) * Tn * HW;   // one thread per (b, t, pixel); loops channels
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % HW;
    const long long bt = i / HW;
    const int t = static_cast<int>(bt % Tn);
    const int b = static_cast<int>(bt / Tn);
    const float* src = x + static_cast<long long>(b) * bstride + static_cast<long long>(t) * C * HW + p;
    T* dst = out + ((static_cast<long long>(t) * B + b) * HW + p) * C;
    for (int c = 0; c < C; ++c) dst[c] = from_f32<T>(src[c * HW]);
  }
}

// PredRNN patchify (models/predrnn_v2.py:232-240): x fp32 [B, T, c, H, W] -> out T [T][B][H/p][W/p][p*p*c],
// patch-channel order (p_h, p_w, c).
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ x, long long bstride, T* __restrict__ out, int B, int Tn,
                                int C, int H, int W, int p) {
  const int hp = H / p, wp = W / p, cp = p * p * C;
  const long long total = static_cast<long long>(B) * Tn * hp * wp * cp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cp);
    long long r = i / cp;
    const int xx = static_cast<int>(r % wp); r /= wp;
    const int yy = static_cast<int>(r % hp); r /= hp;
    const int b = static_cast<int>(r % B);
    const int t = static_cast<int>(r / B);
    const int c = ch % C;
    const int pw = (ch / C) % p;
    const int ph = ch / (C * p);
    const float v = x[static_cast<long long>(b) * bstride +
                      ((static_cast<long long>(t) * C + c) * H + (yy * p + ph)) * W + (xx * p + pw)];
    out[i] = from_f32<T>(v);
  }
}

// inverse (models/predrnn_v2.py:242-250): in T [B][hp][wp][cp] (one frame) -> out fp32 frame t of [B, P, c, H, W]
template <typename T>
__global__ void unpatchify_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int P, int t, int C, int H,
                                  int W, int p) {
  const int hp = H / p, wp = W / p, cp = p * p * C;
  const long long total = static_cast<long long>(B) * C * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(i % W);
    long long r = i / W;
    const int yh = static_cast<int>(r % H); r /= H;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const int ch = ((yh % p) * p + (xw % p)) * C + c;
    const float v = to_f32(in[((static_cast<long long>(b) * hp + yh / p) * wp + xw / p) * cp + ch]);
    out[((static_cast<long long>(b) * P + t) * C + c) * H * W + static_cast<long long>(yh) * W + xw] = v;
  }
}

// in T [B][H][W][C] -> out fp32 [B][C][H][W]
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int B, int C, int H, int W) {
  const long long HW = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(B) * C * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % HW;
    const int c = static_cast<int>((i / HW) % C);
    const long long b = i / (HW * C);
    out[i] = to_f32(in[(b * HW + p) * C + c]);
  }
}

__global__ void cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long n2 = n / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(in)[i];
    reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(v.x, v.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = __float2bfloat16_rn(in[n - 1]);
}

// PredRNN-V2 decoupling loss term (models/predrnn_v2.py:197-211): ad = adapter(delta) as fp32 [2B][HW][C] with
// delta_c in the first B samples and delta_m in the last B.  Per (b, ch): |cos| between the two HW-vectors, each
// L2-normalised first (F.normalize eps 1e-12).  Adds sum over (b, ch) to *acc.
__global__ void decouple_reduce_kernel(const float* __restrict__ ad, int B, int HW, int C, double* acc) {
  const int b = blockIdx.x;
  const int ch = blockIdx.y * 128 + threadIdx.x;
  float v = 0.f;
  if (ch < C) {
    const float* pc = ad + static_cast<size_t>(b) * HW * C + ch;
    const float* pm = ad + (static_cast<size_t>(B) + b) * HW * C + ch;
    float dot = 0.f, nc = 0.f, nm = 0.f;
    for (int p = 0; p < HW; ++p) {
      const float a = pc[static_cast<size_t>(p) * C], m = pm[static_cast<size_t>(p) * C];
      dot = fmaf(a, m, dot);
      nc = fmaf(a, a, nc);
      nm = fmaf(m, m, nm);
    }
    v = fabsf(dot) / (fmaxf(sqrtf(nc), 1e-12f) * fmaxf(sqrtf(nm), 1e-12f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ float ws[4];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(acc, static_cast<double>(ws[0] + ws[1] + ws[2] + ws[3]));
}

// GroupNorm (+ optional LeakyReLU(0.2), + optional residual add) over one sample per CTA, NHWC.
//   in  [B][HW][Cs_in]  (first C channels are real), out [B][HW][Cs_out]; statistics per (sample, group) over
//   (C/groups) channels x HW positions, two-pass (mean, then centred variance) like ATen's kernel; eps inside the
//   sqrt; affine gamma/beta per channel.  Thread layout: tx = channel (coalesced), ty strides over positions.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) groupnorm_act_kernel(const TI* __restrict__ in, TO* __restrict__ out,
                                                            const TO* __restrict__ add, int HW, int C, int Cs_in,
                                                            int Cs_out, int groups, int TC,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int act) {
  __shared__ float s_acc[64], s_mean[64], s_rstd[64];
  const int b = blockIdx.x;
  const int tx = threadIdx.x % TC, ty = threadIdx.x / TC, rows = 256 / TC;
  const int cg = C / groups;
  const bool live = tx < C;
  const int g = live ? tx / cg : 0;
  const TI* ip = in + static_cast<size_t>(b) * HW * Cs_in + tx;
  const float n = static_cast<float>(cg) * HW;
  if (threadIdx.x < 64) s_acc[threadIdx.x] = 0.f;
  __syncthreads();
  float a = 0.f;
  if (live)
    for (int p = ty; p < HW; p += rows) a += to_f32(ip[static_cast<size_t>(p) * Cs_in]);
  if (live) atomicAdd(&s_acc[g], a);
  __syncthreads();
  if (threadIdx.x < groups) {
    s_mean[threadIdx.x] = s_acc[threadIdx.x] / n;
    s_acc[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const float mean = s_mean[g];
  a = 0.f;
  if (live)
    for (int p = ty; p < HW; p += rows) {
      const float d = to_f32(ip[static_cast<size_t>(p) * Cs_in]) - mean;
      a = fmaf(d, d, a);
    }
  if (live) atomicAdd(&s_acc[g], a);
  __syncthreads();
  if (threadIdx.x < groups) s_rstd[threadIdx.x] = rsqrtf(s_acc[threadIdx.x] / n + eps);
  __syncthreads();
  if (!live) return;
  const float sc = s_rstd[g] * gamma[tx];
  const float sh = beta[tx] - mean * sc;
  TO* op = out + static_cast<size_t>(b) * HW * Cs_out + tx;
  const TO* ap = add ? add + static_cast<size_t>(b) * HW * Cs_out + tx : nullptr;
  for (int p = ty; p < HW; p += rows) {
    float v = fmaf(to_f32(ip[static_cast<size_t>(p) * Cs_in]), sc, sh);
    if (act == ACT_LEAKY) v = v > 0.f ? v : 0.2f * v;
    if (ap) v += to_f32(ap[static_cast<size_t>(p) * Cs_out]);
    op[static_cast<size_t>(p) * Cs_out] = from_f32<TO>(v);
  }
}

__global__ void decouple_finalize_kernel(const double* acc, float* aux, double scale) {
  aux[0] = static_cast<float>(acc[0] * scale);
}

}  // namespace

void launch_nhwc_to_nchw(const void* in, int dtype, float* out, int B, int C, int H, int W, int num_sms,
                         cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32) nhwc_to_nchw_kernel<float><<<g, 256, 0, stream>>>(static_cast<const float*>(in), out, B, C, H, W);
  else nhwc_to_nchw_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, B, C, H, W);
  VPK_CUDA(cudaGetLastError());
}

void launch_frames_to_nhwc_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H,
                                   int W, int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * T * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    frames_to_nhwc_kernel<float><<<g, 256, 0, stream>>>(x, bstride, static_cast<float*>(out), B, T, C, H, W);
  else
    frames_to_nhwc_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(x, bstride, static_cast<__nv_bfloat16*>(out), B, T, C,
                                                               H, W);
  VPK_CUDA(cudaGetLastError());
}

void launch_frames_to_nhwc(const float* x, void* out, int dtype, int B, int T, int C, int H, int W, int num_sms,
                           cudaStream_t stream) {
  launch_frames_to_nhwc_strided(x, static_cast<long long>(T) * C * H * W, out, dtype, B, T, C, H, W, num_sms, stream);
}

void launch_cast_f32_to_bf16(const float* in, void* out, long long n, int num_sms, cudaStream_t stream) {
  cast_kernel<<<grid_for((n + 1) / 2, 256, num_sms), 256, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n);
  VPK_CUDA(cudaGetLastError());
}

void launch_decouple_reduce(const float* ad, int B, int HW, int C, double* acc, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((C + 127) / 128));
  decouple_reduce_kernel<<<grid, 128, 0, stream>>>(ad, B, HW, C, acc);
  VPK_CUDA(cudaGetLastError());
}

void launch_groupnorm_act(const void* in, int in_dtype, void* out, int out_dtype, const void* add, int B, int HW, int C,
                          int Cs_in, int Cs_out, int groups, const float* gamma, const float* beta, float eps,
                          int act, cudaStream_t stream) {
  VPK_REQUIRE(C <= 256 && groups <= 64 && C % groups == 0, "groupnorm: unsupported channel / group count");
  int TC = 1;
  while (TC < C) TC <<= 1;
#define VPK_GN(TI, TO)                                                                                          \
  groupnorm_act_kernel<TI, TO><<<B, 256, 0, stream>>>(static_cast<const TI*>(in), static_cast<TO*>(out),           \
                                                      static_cast<const TO*>(add), HW, C, Cs_in, Cs_out, groups, TC, \
                                                      gamma, beta, eps, act)
  if (in_dtype == DT_F32 && out_dtype == DT_F32) VPK_GN(float, float);
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16) VPK_GN(float, __nv_bfloat16);
  else if (in_dtype == DT_BF16 && out_dtype == DT_BF16) VPK_GN(__nv_bfloat16, __nv_bfloat16);
  else VPK_THROW(1, "groupnorm: unsupported dtype combination");
#undef VPK_GN
  VPK_CUDA(cudaGetLastError());
}

void launch_decouple_finalize(const double* acc, float* aux, double scale, cudaStream_t stream) {
  decouple_finalize_kernel<<<1, 1, 0, stream>>>(acc, aux, scale);
  VPK_CUDA(cudaGetLastError());
}

void launch_patchify_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H, int W,
                             int p, int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * T * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    patchify_kernel<float><<<g, 256, 0, stream>>>(x, bstride, static_cast<float*>(out), B, T, C, H, W, p);
  else
    patchify_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(x, bstride, static_cast<__nv_bfloat16*>(out), B, T, C, H, W,
                                                         p);
  VPK_CUDA(cudaGetLastError());
}

void launch_unpatchify(const void* in, float* out, int dtype, int B, int P, int t, int C, int H, int W, int p,
                       int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    unpatchify_kernel<float><<<g, 256, 0, stream>>>(static_cast<const float*>(in), out, B, P, t, C, H, W, p);
  else
    unpatchify_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, B, P, t, C, H,
                                                           W, p);
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
