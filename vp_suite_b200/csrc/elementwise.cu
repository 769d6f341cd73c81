// Memory-bound helper kernels: boundary layout conversion (fp32 NCHW frames <-> NHWC activations), PredRNN
// patchify / un-patchify, GroupNorm (+LeakyReLU), decouple-loss reduction.  Vectorised, coalesced on the side that
// dominates the traffic; grid-stride loops sized in multiples of the SM count.
#include <mutex>

#include "common.h"
#include "elementwise.h"
#include "epilogue.cuh"
#include "ptx.cuh"

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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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

// x fp32 [B, T, C, H, W] -> bf16 [T][B][H][W][8] with channels C..7 zero: one 16-byte store per pixel, and the frame
// becomes TMA-addressable (16-byte pixel stride), so the image-channel stem convs run on the tensor-core kernel.
// `lo` != nullptr: split-bf16, hi = bf16(v) and lo = bf16(v - hi) in two tensors of the same layout.
// F16: `hi` receives fp16 values instead (single tensor, no low parts).
template <bool F16>
__global__ void frames_to_nhwc8_kernel(const float* __restrict__ x, long long bstride, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, int B, int Tn, int C, int H, int W) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const long long HW = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(B) * Tn * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % HW;
    const long long bt = i / HW;
    const int t = static_cast<int>(bt % Tn);
    const int b = static_cast<int>(bt / Tn);
    const float* src = x + static_cast<long long>(b) * bstride + static_cast<long long>(t) * C * HW + p;
    __align__(16) __nv_bfloat16 vh[8], vl[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = (c < C) ? src[c * HW] : 0.f;
      vh[c] = __float2bfloat16_rn(v);
      vl[c] = __float2bfloat16_rn(v - __bfloat162float(vh[c]));
    }
    const long long o = ((static_cast<long long>(t) * B + b) * HW + p);
    if (F16) {
      __align__(16) __half vf[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) vf[c] = __float2half_rn((c < C) ? src[c * HW] : 0.f);
      reinterpret_cast<uint4*>(hi)[o] = *reinterpret_cast<const uint4*>(vf);
      continue;
    }
    reinterpret_cast<uint4*>(hi)[o] = *reinterpret_cast<const uint4*>(vh);
    if (lo != nullptr) reinterpret_cast<uint4*>(lo)[o] = *reinterpret_cast<const uint4*>(vl);
  }
}

// PredRNN patchify (models/predrnn_v2.py:232-240): x fp32 [B, T, c, H, W] -> out T [T][B][H/p][W/p][p*p*c],
// patch-channel order (p_h, p_w, c).
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ x, long long bstride, T* __restrict__ out, int B, int Tn,
                                int C, int H, int W, int p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const long long n2 = n / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(in)[i];
    reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(v.x, v.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = __float2bfloat16_rn(in[n - 1]);
}

// Evaluation metrics, stage 1: one block per (sample, frame) sums (p - y)^2 over the frame: fp32 products, fp64
// accumulation per thread, warp shuffles, then the 8 warp totals in order (deterministic).
__global__ void __launch_bounds__(256) metric_sse_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                         long long chw, double* __restrict__ sse) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ double s_w[8];
  const float* p = pred + static_cast<long long>(blockIdx.x) * chw;
  const float* y = tgt + static_cast<long long>(blockIdx.x) * chw;
  double acc = 0.0;
  if ((chw & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const long long n4 = chw >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 a = reinterpret_cast<const float4*>(p)[i], b = reinterpret_cast<const float4*>(y)[i];
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      acc += static_cast<double>(d0 * d0) + static_cast<double>(d1 * d1) + static_cast<double>(d2 * d2) +
             static_cast<double>(d3 * d3);
    }
  } else {
    for (long long i = threadIdx.x; i < chw; i += 256) {
      const float d = p[i] - y[i];
      acc += static_cast<double>(d * d);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    sse[blockIdx.x] = t;
  }
}
// stage 2: one block per frame adds the batch in a fixed order (thread-strided partials, then a tree in shared memory)
__global__ void __launch_bounds__(256) metric_frame_sums_kernel(const double* __restrict__ sse, int B, int P, long long chw,
                                                                double* __restrict__ out) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ double s_a[256], s_b[256];
  const int t = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) {
    const double v = sse[static_cast<long long>(i) * P + t];
    a += v;
    b += 10.0 * log10(v / static_cast<double>(chw));
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[t] = s_a[0];
    out[P + t] = s_b[0];
    if (t == 0) out[2 * P] = static_cast<double>(B);
  }
}

// SSIM (vp_suite/measure/image_wise.py:100-121 = piqa 1.1.7 SSIM with its defaults, restated: 11-tap Gaussian window,
// sigma 1.5, valid region only, K1 = 0.01, K2 = 0.03, value range 1; inputs mapped (v + 1) / 2 and clamped to [0, 1]
// as base_measure.py:59-75 does).  Stage 1: one block per (image plane, strip of R output rows).  The strip (+10 halo
// rows) of both images goes to shared memory once.  Pass 1 filters vertically: a thread owns one column and four
// output rows, so 14 loaded (x, y) pairs feed 4 x 5 filtered moments (x, y, xx, yy, xy); the moments are stored
// transposed ([moment][column][row], odd row pitch).  Pass 2 filters horizontally: a thread owns one row and four output
// columns (14 x 5 loads for four SSIM values; consecutive lanes = consecutive rows of the transposed store).  fp64
// block sum in a fixed order.  ~25 shared-memory loads and ~110 FMAs per output position: issue-bound, not HBM-bound.
struct SsimWin {
  float w[11];
};
__global__ void __launch_bounds__(256) metric_ssim_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                          int H, int W, int R, SsimWin win, double* __restrict__ part) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float ssim_sm[];
  __shared__ double s_w[8];
  const int Ho = H - 10, Wo = W - 10, RS = R + 1;
  const int r0 = blockIdx.y * R;
  const int rows_out = min(R, Ho - r0), rows_in = rows_out + 10;
  float* sx = ssim_sm;
  float* sy = sx + (R + 10) * W;
  float* vq = sy + (R + 10) * W;            // [5][W][RS]
  const int plane = W * RS;
  const long long base = (static_cast<long long>(blockIdx.x) * H + r0) * W;
  for (int i = threadIdx.x; i < rows_in * W; i += 256) {
    sx[i] = fminf(fmaxf((pred[base + i] + 1.f) * 0.5f, 0.f), 1.f);
    sy[i] = fminf(fmaxf((tgt[base + i] + 1.f) * 0.5f, 0.f), 1.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (R / 4) * W; i += 256) {
    const int rb = i / W, c = i - rb * W, rr = rb * 4;
    if (rr >= rows_out) continue;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, aa[4] = {0.f, 0.f, 0.f, 0.f},
          bb[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int row = rr + k;
      float xv = 0.f, yv = 0.f;
      if (row < rows_in) {
        xv = sx[row * W + c];
        yv = sy[row * W + c];
      }
      const float xx = xv * xv, yy = yv * yv, xy = xv * yv;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k - j;
        if (kk >= 0 && kk <= 10) {
          const float wk = win.w[kk];
          a[j] += wk * xv;
          b[j] += wk * yv;
          aa[j] += wk * xx;
          bb[j] += wk * yy;
          ab[j] += wk * xy;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rr + j < rows_out) {
        const int o = c * RS + rr + j;
        vq[o] = a[j];
        vq[plane + o] = b[j];
        vq[2 * plane + o] = aa[j];
        vq[3 * plane + o] = bb[j];
        vq[4 * plane + o] = ab[j];
      }
    }
  }
  __syncthreads();
  const float c1 = 1e-4f, c2 = 9e-4f;       // (K1 * range)^2, (K2 * range)^2
  const int ncb = (Wo + 3) / 4;
  double acc = 0.0;
  for (int i = threadIdx.x; i < R * ncb; i += 256) {
    const int r = i & (R - 1), c0 = (i / R) * 4;
    if (r >= rows_out) continue;
    float m[5][4];
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) m[q][j] = 0.f;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int col = c0 + k;
      float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (col < W) {
        const int o = col * RS + r;
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = vq[q * plane + o];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k - j;
        if (kk >= 0 && kk <= 10) {
          const float wk = win.w[kk];
#pragma unroll
          for (int q = 0; q < 5; ++q) m[q][j] += wk * v[q];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c0 + j < Wo) {
        const float mx = m[0][j], my = m[1][j];
        const float mxx = mx * mx, myy = my * my, mxy = mx * my;
        const float cs = (2.f * (m[4][j] - mxy) + c2) / ((m[2][j] - mxx) + (m[3][j] - myy) + c2);
        acc += static_cast<double>((2.f * mxy + c1) / (mxx + myy + c1) * cs);
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    part[static_cast<long long>(blockIdx.x) * gridDim.y + blockIdx.y] = t;
  }
}
// stage 2: one block per frame; out[t] = sum over the batch of the image's SSIM (mean of the map over channels and the
// valid region); fixed order: per sample the C * S strip sums in index order, thread-strided over the batch, smem tree
__global__ void __launch_bounds__(256) metric_ssim_sums_kernel(const double* __restrict__ part, int B, int P, int CS,
                                                               double inv_count, double* __restrict__ out) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ double s_a[256];
  const int t = blockIdx.x;
  double a = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) {
    const double* q = part + (static_cast<long long>(i) * P + t) * CS;
    double v = 0.0;
    for (int j = 0; j < CS; ++j) v += q[j];
    a += v * inv_count;
  }
  s_a[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) s_a[threadIdx.x] += s_a[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[t] = s_a[0];
}

// fp32 -> fp16, n % 4 == 0
__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n4) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01);
    hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    reinterpret_cast<uint2*>(out)[i] = hv;
  }
}

// PredRNN-V2 decoupling loss term (models/predrnn_v2.py:197-211): ad = adapter(delta) as fp32 [2B][HW][C] with
// delta_c in the first B samples and delta_m in the last B.  Per (b, ch): |cos| between the two HW-vectors, each
// L2-normalised first (F.normalize eps 1e-12).  Adds sum over (b, ch) to *acc.
// Thread layout: 32 channel quads (float4, coalesced 512-byte rows) x 8 position slices per block of 256 threads; the
// slices are combined through shared memory, then one warp finishes |cos| and the block adds one fp64 atomic.
__global__ void __launch_bounds__(256) decouple_reduce_kernel(const float* __restrict__ ad, int B, int HW, int C,
                                                              double* acc) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_part[8][3][128];
  const int b = blockIdx.x;
  const int q = threadIdx.x & 31;            // channel quad inside this block's 128-channel range
  const int sl = threadIdx.x >> 5;           // position slice
  const int ch = blockIdx.y * 128 + q * 4;
  float dot[4] = {0.f, 0.f, 0.f, 0.f}, nc[4] = {0.f, 0.f, 0.f, 0.f}, nm[4] = {0.f, 0.f, 0.f, 0.f};
  if (ch < C) {
    const float* pc = ad + static_cast<size_t>(b) * HW * C + ch;
    const float* pm = ad + (static_cast<size_t>(B) + b) * HW * C + ch;
    for (int p = sl; p < HW; p += 8) {
      const float4 a = *reinterpret_cast<const float4*>(pc + static_cast<size_t>(p) * C);
      const float4 m = *reinterpret_cast<const float4*>(pm + static_cast<size_t>(p) * C);
      dot[0] = fmaf(a.x, m.x, dot[0]); nc[0] = fmaf(a.x, a.x, nc[0]); nm[0] = fmaf(m.x, m.x, nm[0]);
      dot[1] = fmaf(a.y, m.y, dot[1]); nc[1] = fmaf(a.y, a.y, nc[1]); nm[1] = fmaf(m.y, m.y, nm[1]);
      dot[2] = fmaf(a.z, m.z, dot[2]); nc[2] = fmaf(a.z, a.z, nc[2]); nm[2] = fmaf(m.z, m.z, nm[2]);
      dot[3] = fmaf(a.w, m.w, dot[3]); nc[3] = fmaf(a.w, a.w, nc[3]); nm[3] = fmaf(m.w, m.w, nm[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_part[sl][0][q * 4 + j] = dot[j];
    s_part[sl][1][q * 4 + j] = nc[j];
    s_part[sl][2][q * 4 + j] = nm[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    float d = 0.f, a = 0.f, m = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      d += s_part[k][0][c];
      a += s_part[k][1][c];
      m += s_part[k][2][c];
    }
    float v = 0.f;
    if (blockIdx.y * 128 + c < C) v = fabsf(d) / (fmaxf(sqrtf(a), 1e-12f) * fmaxf(sqrtf(m), 1e-12f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ float ws[4];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncwarp();
    asm volatile("bar.sync 1, 128;");
    if (threadIdx.x == 0) atomicAdd(acc, static_cast<double>(ws[0] + ws[1] + ws[2] + ws[3]));
  }
}

// PhyCell F tail (model_blocks/phydnet.py:33-39, 60): h~ = h + conv2(GroupNorm(f1)) in ONE kernel.  f1 = F.conv1(h)
// is fp32 [B][HW][Cs] (hid real channels, channel-padded); GroupNorm(groups, hid) has no activation; conv2 is 1 x 1
// (hid -> C) with bias.  One block per sample: the sample is staged in shared memory (rows padded to an odd word count:
// conflict-free), statistics are two-pass over the staged copy with fixed-order sums, the normalised values overwrite it,
// and every thread then does the hid x C matrix-vector product of its positions against the shared-memory weights
// (broadcast LDS.128) in fp32.  Replaces a GroupNorm launch + a tensor-core 1x1 conv launch (K = 49: one MMA slice).
constexpr int kPhyTailThreads = 256;
__global__ void __launch_bounds__(kPhyTailThreads) phy_f_tail_kernel(const float* __restrict__ f1, const float* __restrict__ h,
                                                                    float* __restrict__ htilde, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, const float* __restrict__ w2,
                                                                    const float* __restrict__ b2, int HW, int hid, int Cs,
                                                                    int groups, int C, float eps) {
  ptx::pdl_launch_dependents();
  extern __shared__ __align__(16) float s_dyn[];
  const int rs = Cs | 1;                                   // row stride in words (odd)
  float* s_x = s_dyn;                                      // [HW][rs]
  float* s_w = s_x + ((HW * rs + 3) & ~3);                 // [hid][C]  (w2 transposed: input-channel major)
  float* s_b = s_w + hid * C;                              // [C]
  __shared__ float s_part[kPhyTailThreads / 32][64], s_mean[64], s_rstd[64];
  for (int i = threadIdx.x; i < hid * C; i += kPhyTailThreads) {
    const int k = i / C, o = i - k * C;
    s_w[i] = w2[o * hid + k];                              // reference layout [C][hid][1][1]
  }
  for (int i = threadIdx.x; i < C; i += kPhyTailThreads) s_b[i] = b2 ? b2[i] : 0.f;
  ptx::pdl_wait();
  const int b = blockIdx.x;
  const float* src = f1 + static_cast<size_t>(b) * HW * Cs;
  // 16-byte global loads (Cs % 4 == 0); padding channels of f1 (k >= hid) are never written by conv1: only whole quads
  // below hid are read, the ragged last quad element by element
  const int q4 = Cs >> 2;
  for (int i = threadIdx.x; i < HW * q4; i += kPhyTailThreads) {
    const int p = i / q4, k = (i - p * q4) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (k + 4 <= hid) {
      const float4 t = *reinterpret_cast<const float4*>(src + static_cast<size_t>(p) * Cs + k);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (k + j < hid) v[j] = src[static_cast<size_t>(p) * Cs + k + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s_x[p * rs + k + j] = v[j];
  }
  __syncthreads();
  const int cg = hid / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = kPhyTailThreads / 32;
  const float n = static_cast<float>(cg) * HW;
  // statistics: warp w sums positions w, w + nw, ...; lane = channel (two rounds cover hid <= 64)
  auto group_pass = [&](bool centred) {
    for (int c0 = 0; c0 < hid; c0 += 32) {
      const int c = c0 + lane;
      float a = 0.f;
      if (c < hid) {
        const float mu = centred ? s_mean[c / cg] : 0.f;
        for (int p = warp; p < HW; p += nw) {
          const float d = s_x[p * rs + c] - mu;
          a = centred ? fmaf(d, d, a) : a + d;
        }
      }
      s_part[warp][c0 + lane] = a;
    }
    __syncthreads();
    if (threadIdx.x < groups) {
      float t = 0.f;
      for (int w = 0; w < nw; ++w)
        for (int c = 0; c < cg; ++c) t += s_part[w][threadIdx.x * cg + c];
      if (centred) s_rstd[threadIdx.x] = rsqrtf(t / n + eps);
      else s_mean[threadIdx.x] = t / n;
    }
    __syncthreads();
  };
  group_pass(false);
  group_pass(true);
  for (int i = threadIdx.x; i < HW * hid; i += kPhyTailThreads) {
    const int p = i / hid, k = i - p * hid;
    const float sc = s_rstd[k / cg] * gamma[k];
    s_x[p * rs + k] = fmaf(s_x[p * rs + k] - s_mean[k / cg], sc, beta[k]);
  }
  __syncthreads();
  // 1x1 conv + bias + residual: one position per thread, 16 output channels at a time
  for (int p = threadIdx.x; p < HW; p += kPhyTailThreads) {
    const float* hp_ = h + (static_cast<size_t>(b) * HW + p) * C;
    float* op = htilde + (static_cast<size_t>(b) * HW + p) * C;
    for (int o0 = 0; o0 < C; o0 += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = s_b[o0 + j];
#pragma unroll 7
      for (int k = 0; k < hid; ++k) {
        const float xv = s_x[p * rs + k];
        const float4* wr = reinterpret_cast<const float4*>(s_w + k * C + o0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
          acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 hv = *reinterpret_cast<const float4*>(hp_ + o0 + 4 * q);
        *reinterpret_cast<float4*>(op + o0 + 4 * q) =
            make_float4(acc[4 * q] + hv.x, acc[4 * q + 1] + hv.y, acc[4 * q + 2] + hv.z, acc[4 * q + 3] + hv.w);
      }
    }
  }
}

// Decoupling-loss term from the partial sums the EPI_DECOUPLE conv epilogue left (slots[b][slot][C][3] = dot, |c|^2,
// |m|^2 over one warp's positions): per (b, ch) |cos| as above, summed over the channels of sample b in a fixed order
// into term[b].  One block per sample, one thread per channel.
__global__ void __launch_bounds__(256) decouple_cos_kernel(const float* __restrict__ slots, int nslots, int C,
                                                           float* __restrict__ term) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_v[256];
  const int b = blockIdx.x, c = threadIdx.x;
  float v = 0.f;
  if (c < C) {
    float d = 0.f, a = 0.f, m = 0.f;
    const float* p = slots + (static_cast<size_t>(b) * nslots * C + c) * 3;
    for (int s = 0; s < nslots; ++s) {
      d += p[0];
      a += p[1];
      m += p[2];
      p += static_cast<size_t>(C) * 3;
    }
    v = fabsf(d) / (fmaxf(sqrtf(a), 1e-12f) * fmaxf(sqrtf(m), 1e-12f));
  }
  s_v[c] = v;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (c < o) s_v[c] += s_v[c + o];
    __syncthreads();
  }
  if (c == 0) term[b] = s_v[0];
}
// *acc += sum of terms[0 .. n) in index order (one thread per 1/256th, fixed tree): the rollout's loss accumulator
__global__ void __launch_bounds__(256) decouple_sum_kernel(const float* __restrict__ terms, long long n, double* acc) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ double s_v[256];
  double v = 0.0;
  for (long long i = threadIdx.x; i < n; i += 256) v += static_cast<double>(terms[i]);
  s_v[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) s_v[threadIdx.x] += s_v[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *acc += s_v[0];
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  // per-thread partial sums go to s_part[thread] and are added in a fixed order: no atomics, results do not depend on
  // the warp schedule
  __shared__ float s_part[256], s_mean[64], s_rstd[64];
  const int b = blockIdx.x;
  const int tx = threadIdx.x % TC, ty = threadIdx.x / TC, rows = 256 / TC;
  const int cg = C / groups;
  const bool live = tx < C;
  const int g = live ? tx / cg : 0;
  const TI* ip = in + static_cast<size_t>(b) * HW * Cs_in + tx;
  const float n = static_cast<float>(cg) * HW;
  auto group_total = [&](int grp) {           // sum of the partials of group grp: rows x cg entries, fixed order
    float t = 0.f;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cg; ++c) t += s_part[r * TC + grp * cg + c];
    return t;
  };
  float a = 0.f;
  if (live)
    for (int p = ty; p < HW; p += rows) a += to_f32(ip[static_cast<size_t>(p) * Cs_in]);
  s_part[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < groups) s_mean[threadIdx.x] = group_total(threadIdx.x) / n;
  __syncthreads();
  const float mean = s_mean[g];
  a = 0.f;
  if (live)
    for (int p = ty; p < HW; p += rows) {
      const float d = to_f32(ip[static_cast<size_t>(p) * Cs_in]) - mean;
      a = fmaf(d, d, a);
    }
  s_part[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < groups) s_rstd[threadIdx.x] = rsqrtf(group_total(threadIdx.x) / n + eps);
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

// GroupNorm (+ LeakyReLU(0.2), + fp32 residual add) with the whole sample staged in shared memory: HBM sees one read
// and one write of the tensor (the generic kernel above reads it three times).  in fp32 dense [B][HW][C] with
// (4 * blockDim) % C == 0, so every thread owns the same four channels in each of its float4 pieces; statistics are
// two-pass (mean, then centred sum of squares) over the staged copy, like ATen.  OUT selects the output encoding:
// 0 fp32, 1 bf16, 2 split-bf16 (hi = bf16(v), lo = bf16(v - hi): together 16 mantissa bits, the operand format of the
// three-product tensor-core convs that replace fp32 convs).  STAGED = false re-reads global memory instead (samples
// larger than shared memory).
// lanes of a warp that own the same four channels (4 * lane distance is a multiple of C) are summed by shuffles first,
// then one lane per channel quad stores the warp's partial to s_part[warp][channel]; gn_total() adds the partials of a
// group in a fixed order (no atomics: results do not depend on the warp schedule).  s_part must be zero where a warp
// owns no lanes of a channel, so it is cleared before each pass.
constexpr int kGnMaxC = 256;
__device__ __forceinline__ void gn_commit(float (*s_part)[kGnMaxC], int C, int c0, float a0, float a1, float a2, float a3) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    if ((o * 4) % C == 0) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane * 4 < C) {
    s_part[warp][c0 + 0] = a0;
    s_part[warp][c0 + 1] = a1;
    s_part[warp][c0 + 2] = a2;
    s_part[warp][c0 + 3] = a3;
  }
}
__device__ __forceinline__ float gn_total(float (*s_part)[kGnMaxC], int nwarps, int grp, int cg) {
  float t = 0.f;
  for (int w = 0; w < nwarps; ++w)
    for (int c = 0; c < cg; ++c) t += s_part[w][grp * cg + c];
  return t;
}

constexpr int kGnThreads = 512;
template <int OUT, bool STAGED>
__global__ void __launch_bounds__(kGnThreads) groupnorm_smem_kernel(const float* __restrict__ in, void* __restrict__ out,
                                                                    void* __restrict__ out_lo,
                                                                    const float* __restrict__ add, int HW, int C,
                                                                    int groups, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float eps, int act) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float4 gn_stage[];
  __shared__ float s_part[kGnThreads / 32][kGnMaxC];
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.x;
  const int n4 = HW * C / 4;
  const float4* src = reinterpret_cast<const float4*>(in + static_cast<size_t>(b) * HW * C);
  const int c0 = (threadIdx.x * 4) % C;
  const int cg = C / groups;
  const float n = static_cast<float>(cg) * HW;
  auto clear_part = [&]() {
    for (int i = threadIdx.x; i < (kGnThreads / 32) * C; i += kGnThreads) s_part[i / C][i % C] = 0.f;
  };
  clear_part();
  __syncthreads();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int i = threadIdx.x; i < n4; i += kGnThreads) {
    const float4 v = src[i];
    if (STAGED) gn_stage[i] = v;
    a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
  }
  gn_commit(s_part, C, c0, a0, a1, a2, a3);
  __syncthreads();
  if (threadIdx.x < groups) s_mean[threadIdx.x] = gn_total(s_part, kGnThreads / 32, threadIdx.x, cg) / n;
  __syncthreads();
  clear_part();
  __syncthreads();
  const float m0 = s_mean[(c0 + 0) / cg], m1 = s_mean[(c0 + 1) / cg], m2 = s_mean[(c0 + 2) / cg],
              m3 = s_mean[(c0 + 3) / cg];
  a0 = a1 = a2 = a3 = 0.f;
  for (int i = threadIdx.x; i < n4; i += kGnThreads) {
    const float4 v = STAGED ? gn_stage[i] : src[i];
    a0 = fmaf(v.x - m0, v.x - m0, a0);
    a1 = fmaf(v.y - m1, v.y - m1, a1);
    a2 = fmaf(v.z - m2, v.z - m2, a2);
    a3 = fmaf(v.w - m3, v.w - m3, a3);
  }
  gn_commit(s_part, C, c0, a0, a1, a2, a3);
  __syncthreads();
  if (threadIdx.x < groups) s_rstd[threadIdx.x] = rsqrtf(gn_total(s_part, kGnThreads / 32, threadIdx.x, cg) / n + eps);
  __syncthreads();
  float sc[4], sh[4];
  const float mm[4] = {m0, m1, m2, m3};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = s_rstd[(c0 + j) / cg] * gamma[c0 + j];
    sh[j] = beta[c0 + j] - mm[j] * sc[j];
  }
  const float4* addp = add ? reinterpret_cast<const float4*>(add + static_cast<size_t>(b) * HW * C) : nullptr;
  const size_t obase = static_cast<size_t>(b) * n4;
  for (int i = threadIdx.x; i < n4; i += kGnThreads) {
    const float4 x = STAGED ? gn_stage[i] : src[i];
    float v[4] = {fmaf(x.x, sc[0], sh[0]), fmaf(x.y, sc[1], sh[1]), fmaf(x.z, sc[2], sh[2]), fmaf(x.w, sc[3], sh[3])};
    if (act == ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
    }
    if (addp) {
      const float4 r = addp[i];
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if (OUT == 0) {
      reinterpret_cast<float4*>(out)[obase + i] = make_float4(v[0], v[1], v[2], v[3]);
    } else if (OUT == 3) {
      const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01);
      hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      reinterpret_cast<uint2*>(out)[obase + i] = hv;
    } else {
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01);
      hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      reinterpret_cast<uint2*>(out)[obase + i] = hv;
      if (OUT == 2) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y),
                             l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
        uint2 lv;
        lv.x = *reinterpret_cast<const uint32_t*>(&l01);
        lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        reinterpret_cast<uint2*>(out_lo)[obase + i] = lv;
      }
    }
  }
}

// GroupNorm apply pass for convs whose epilogue already produced partial statistics (sums[b][slot][group][2] = sum, sum
// of squares of one warp's positions; nslots slots, added here in slot order -- deterministic):
// y = (x - mean) * rstd * gamma + beta, LeakyReLU, optional fp32 add.
// One block = one slice of ONE sample (blockIdx.y = sample), so scale / shift per channel are computed once per block.
// OUT as in groupnorm_smem_kernel (0 fp32, 1 bf16, 3 fp16).  var = E[x^2] - mean^2 in fp32 (n <= a few thousand).
template <int OUT, bool IN16>
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const void* __restrict__ in_, void* __restrict__ out,
                                                              const float* __restrict__ add,
                                                              const float* __restrict__ sums, int nslots, int HW, int C,
                                                              int groups,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int act) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_sc[256], s_sh[256];
  const int b = blockIdx.y;
  const int cg = C / groups;
  if (threadIdx.x < C) {
    const float n = static_cast<float>(cg) * HW;
    const float2* sp = reinterpret_cast<const float2*>(sums) + static_cast<size_t>(b) * nslots * groups + threadIdx.x / cg;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int sl = 0; sl < nslots; ++sl) {
      const float2 v = sp[static_cast<size_t>(sl) * groups];
      s0 += v.x;
      s1 += v.y;
    }
    const float mean = s0 / n;
    const float var = fmaxf(s1 / n - mean * mean, 0.f);
    const float sc = rsqrtf(var + eps) * gamma[threadIdx.x];
    s_sc[threadIdx.x] = sc;
    s_sh[threadIdx.x] = beta[threadIdx.x] - mean * sc;
  }
  __syncthreads();
  const int n4 = HW * C / 4;
  const size_t obase = static_cast<size_t>(b) * n4;
  // IN16: the producing conv stored its raw output as fp16 (its statistics came from the fp32 accumulators)
  const float4* src = reinterpret_cast<const float4*>(in_) + (IN16 ? 0 : obase);
  const uint2* src16 = reinterpret_cast<const uint2*>(in_) + (IN16 ? obase : 0);
  const float4* addp = add ? reinterpret_cast<const float4*>(add) + obase : nullptr;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) {
    const int c0 = (i * 4) % C;
    float4 x;
    if constexpr (IN16) {
      const uint2 q = src16[i];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
      x = make_float4(a.x, a.y, b.x, b.y);
    } else {
      x = src[i];
    }
    float v[4] = {fmaf(x.x, s_sc[c0], s_sh[c0]), fmaf(x.y, s_sc[c0 + 1], s_sh[c0 + 1]), fmaf(x.z, s_sc[c0 + 2], s_sh[c0 + 2]),
                  fmaf(x.w, s_sc[c0 + 3], s_sh[c0 + 3])};
    if (act == ACT_LEAKY) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
    }
    if (addp) {
      const float4 r = addp[i];
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if (OUT == 0) {
      reinterpret_cast<float4*>(out)[obase + i] = make_float4(v[0], v[1], v[2], v[3]);
    } else if (OUT == 3) {
      const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01);
      hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      reinterpret_cast<uint2*>(out)[obase + i] = hv;
    } else {
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
      uint2 hv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01);
      hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      reinterpret_cast<uint2*>(out)[obase + i] = hv;
    }
  }
}

// fp32 -> split-bf16 (hi, lo); n % 4 == 0
__global__ void split_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n4) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y),
                         l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01);
    hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01);
    lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi)[i] = hv;
    reinterpret_cast<uint2*>(lo)[i] = lv;
  }
}

// fp32 -> split-fp16 (hi, lo); n % 4 == 0
__global__ void split_f16_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, long long n4) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01);
    hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    lv.x = *reinterpret_cast<const uint32_t*>(&l01);
    lv.y = *reinterpret_cast<const uint32_t*>(&l23);
    reinterpret_cast<uint2*>(hi)[i] = hv;
    reinterpret_cast<uint2*>(lo)[i] = lv;
  }
}

// ---- action-conditional helpers ----
template <typename T> __device__ __forceinline__ T to_act(float v);
template <> __device__ __forceinline__ float to_act<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_act<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_act<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float from_act(float v) { return v; }
__device__ __forceinline__ float from_act(__half v) { return __half2float(v); }
__device__ __forceinline__ float from_act(__nv_bfloat16 v) { return __bfloat162float(v); }

// out[t][b][hw][ch] = ch < a ? actions[b * bstride + t * a + ch] : 0   (the action vector of step t, inflated spatially)
template <typename T>
__global__ void inflate_actions_kernel(const float* __restrict__ actions, long long bstride, int a, T* __restrict__ out, int B,
                                       int Tn, int HW, int a_pad) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const long long total = static_cast<long long>(Tn) * B * HW * a_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % a_pad);
    const long long tb = i / (static_cast<long long>(a_pad) * HW);
    const int b = static_cast<int>(tb % B), t = static_cast<int>(tb / B);
    out[i] = to_act<T>(ch < a ? actions[b * bstride + static_cast<long long>(t) * a + ch] : 0.f);
  }
}

// ST-Phy's action tensor (models/st_phy.py:48-56, 146-148): amap = Linear(action_t) viewed as [IA][H][W] (bias-free, weight wl
// [IA*H*W][a]), out = conv(5,1)(amap) + conv(1,5)(amap) (zero padding 2, weights wh [C][IA][5], ww [C][IA][5]), stored NHWC as
// out[t][b][y][x][c].  One block per (t, b): the map is built in shared memory, then one thread per output value.  fp32
// arithmetic in the reference's summation structure (Linear dot; each conv's taps; the sum of the two convs).
template <typename T>
__global__ void __launch_bounds__(256) stphy_action_tensor_kernel(const float* __restrict__ actions, long long bstride, int a,
                                                                  const float* __restrict__ wl, const float* __restrict__ wh,
                                                                  const float* __restrict__ ww, T* __restrict__ out, int B, int H,
                                                                  int W, int C, int IA) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float s_map[];                     // [IA][H][W]
  const int b = blockIdx.x % B, t = blockIdx.x / B;
  const int HW = H * W;
  const float* av = actions + b * bstride + static_cast<long long>(t) * a;
  for (int i = threadIdx.x; i < IA * HW; i += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < a; ++k) acc = fmaf(wl[static_cast<long long>(i) * a + k], av[k], acc);
    s_map[i] = acc;
  }
  __syncthreads();
  T* o = out + (static_cast<long long>(t) * B + b) * HW * C;
  for (int i = threadIdx.x; i < HW * C; i += blockDim.x) {
    const int c = i % C, p = i / C;
    const int y = p / W, x = p - y * W;
    float sh = 0.f, sw = 0.f;
    for (int ia = 0; ia < IA; ++ia) {
      const float* m = s_map + ia * HW;
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const int yy = y + d - 2, xx = x + d - 2;
        if (yy >= 0 && yy < H) sh = fmaf(wh[(c * IA + ia) * 5 + d], m[yy * W + x], sh);
        if (xx >= 0 && xx < W) sw = fmaf(ww[(c * IA + ia) * 5 + d], m[y * W + xx], sw);
      }
    }
    o[i] = to_act<T>(sh + sw);
  }
}

// out = T(x + y) elementwise; x of type TX, y fp32 (or absent)
template <typename TX, typename T>
__global__ void add_to_act_kernel(const TX* __restrict__ x, const float* __restrict__ y, T* __restrict__ out, long long n) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = to_act<T>(from_act(x[i]) + (y != nullptr ? y[i] : 0.f));
}

// fp32 [B][H][W][C] -> typed copies (hi / lo of type T1, cell of type T2, fp32), optionally L2-normalised along W first
// (F.normalize(x, p=2, dim=-1, eps): x / max(||x||_2 over W, eps), ST-Phy's encoder, model_blocks/enc.py:69).
// One thread = one (b, y, c) row of W values.
template <typename T1, typename T2, bool NORM>
__global__ void fanout_kernel(const float* __restrict__ in, T1* __restrict__ hi, T1* __restrict__ lo, T2* __restrict__ cell,
                              float* __restrict__ f32, long long rows, int W, int C, float eps) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(r % C);
    const long long by = r / C;
    const long long base = by * W * C + c;
    float scale = 1.f;
    if (NORM) {
      float ss = 0.f;
      for (int x = 0; x < W; ++x) {
        const float v = in[base + static_cast<long long>(x) * C];
        ss = fmaf(v, v, ss);
      }
      scale = 1.f / fmaxf(sqrtf(ss), eps);
    }
    for (int x = 0; x < W; ++x) {
      const long long i = base + static_cast<long long>(x) * C;
      const float v = in[i] * scale;
      if (hi != nullptr) {
        const T1 h = to_act<T1>(v);
        hi[i] = h;
        if (lo != nullptr) lo[i] = to_act<T1>(v - from_act(h));
      }
      if (cell != nullptr) cell[i] = to_act<T2>(v);
      if (f32 != nullptr) f32[i] = v;
    }
  }
}

// ---- TrajGRU (model_blocks/traj_gru.py:150-166, 198-206) ----
// warped[b, y, x, l * C + c] = bilinear sample of h[b, :, :, c] at (x - flow_x, y - flow_y) for flow pair l, zeros outside.
// The reference normalises with (W - 1) / (H - 1) and samples with grid_sample's default align_corners=False: pixel
// coordinate = ((2 s / (W - 1) - 1 + 1) * W - 1) / 2.  One thread = one (position, flow, 4 channels).
template <typename T>
__global__ void trajgru_warp_kernel(const float* __restrict__ h, const float* __restrict__ flows, int fpix, T* __restrict__ out,
                                    int B, int H, int W, int C, int L) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int cq = C >> 2;
  const long long total = static_cast<long long>(B) * H * W * L * cq;
  const float sx = static_cast<float>(W) / static_cast<float>(max(W - 1, 1)), sy = static_cast<float>(H) / static_cast<float>(max(H - 1, 1));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % cq);
    long long r = i / cq;
    const int l = static_cast<int>(r % L);
    const long long pos = r / L;
    const int x = static_cast<int>(pos % W), y = static_cast<int>((pos / W) % H);
    const long long b = pos / (static_cast<long long>(W) * H);
    const float* fp = flows + pos * fpix + 2 * l;
    // vgrid = grid + (-flow); normalised with (W - 1), sampled with align_corners = False
    const float xn = 2.0f * (static_cast<float>(x) - fp[0]) / static_cast<float>(max(W - 1, 1)) - 1.0f;
    const float yn = 2.0f * (static_cast<float>(y) - fp[1]) / static_cast<float>(max(H - 1, 1)) - 1.0f;
    const float ix = ((xn + 1.0f) * W - 1.0f) * 0.5f, iy = ((yn + 1.0f) * H - 1.0f) * 0.5f;
    (void)sx; (void)sy;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* hb = h + b * static_cast<long long>(H) * W * C + c4 * 4;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xs = x0 + dx, ys = y0 + dy;
        if (xs >= 0 && xs < W && ys >= 0 && ys < H) {
          const float wgt = (dx ? wx1 : wx0) * (dy ? wy1 : wy0);
          const float4 v = *reinterpret_cast<const float4*>(hb + (static_cast<long long>(ys) * W + xs) * C);
          acc.x = fmaf(wgt, v.x, acc.x);
          acc.y = fmaf(wgt, v.y, acc.y);
          acc.z = fmaf(wgt, v.z, acc.z);
          acc.w = fmaf(wgt, v.w, acc.w);
        }
      }
    T* o = out + pos * (static_cast<long long>(L) * C) + static_cast<long long>(l) * C + c4 * 4;
    o[0] = to_act<T>(acc.x);
    o[1] = to_act<T>(acc.y);
    o[2] = to_act<T>(acc.z);
    o[3] = to_act<T>(acc.w);
  }
}

// r = sig(i2h_0 + h2h_0), u = sig(i2h_1 + h2h_1), m = act(i2h_2 + r * h2h_2), h' = u * h + (1 - u) * m   (i2h may be absent)
template <typename T>
__global__ void trajgru_gates_kernel(const float* __restrict__ i2h, const float* __restrict__ h2h, const float* __restrict__ h,
                                     float* __restrict__ h_out, T* __restrict__ h_act, long long P, int C, int act) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const long long total = P * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / C;
    const int c = static_cast<int>(i - p * C);
    const float* a = h2h + p * 3 * C;
    float ir = 0.f, iu = 0.f, im = 0.f;
    if (i2h != nullptr) {
      const float* q = i2h + p * 3 * C;
      ir = q[c];
      iu = q[C + c];
      im = q[2 * C + c];
    }
    const float r = 1.f / (1.f + __expf(-(ir + a[c])));
    const float u = 1.f / (1.f + __expf(-(iu + a[C + c])));
    float m = im + r * a[2 * C + c];
    if (act == ACT_LEAKY) m = m > 0.f ? m : 0.2f * m;
    else if (act == ACT_RELU) m = fmaxf(m, 0.f);
    else if (act == ACT_SIGMOID) m = 1.f / (1.f + __expf(-m));
    const float hn = u * h[i] + (1.f - u) * m;
    h_out[i] = hn;
    if (h_act != nullptr) h_act[i] = to_act<T>(hn);
  }
}

__global__ void decouple_finalize_kernel(const double* acc, float* aux, double scale) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  aux[0] = static_cast<float>(acc[0] * scale);
}

}  // namespace

void launch_nhwc_to_nchw(const void* in, int dtype, float* out, int B, int C, int H, int W, int num_sms,
                         cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32) launch_pdl(nhwc_to_nchw_kernel<float>, dim3(g), dim3(256), 0, stream, static_cast<const float*>(in), out, B, C, H, W);
  else if (dtype == DT_F16) launch_pdl(nhwc_to_nchw_kernel<__half>, dim3(g), dim3(256), 0, stream, static_cast<const __half*>(in), out, B, C, H, W);
  else launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(in), out, B, C, H, W);
  VPK_CUDA(cudaGetLastError());
}

void launch_frames_to_nhwc_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H,
                                   int W, int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * T * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    launch_pdl(frames_to_nhwc_kernel<float>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<float*>(out), B, T, C, H, W);
  else if (dtype == DT_F16)
    launch_pdl(frames_to_nhwc_kernel<__half>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<__half*>(out), B, T, C, H, W);
  else
    launch_pdl(frames_to_nhwc_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<__nv_bfloat16*>(out), B, T, C,
                                                               H, W);
  VPK_CUDA(cudaGetLastError());
}

void launch_frames_to_nhwc8(const float* x, long long bstride, void* hi, void* lo, int out_dtype, int B, int T, int C,
                            int H, int W, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(C >= 1 && C <= 8, "frames_to_nhwc8: 1..8 image channels");
  const long long total = static_cast<long long>(B) * T * H * W;
  if (out_dtype == DT_F16)
    launch_pdl(frames_to_nhwc8_kernel<true>, dim3(grid_for(total, 256, num_sms)), dim3(256), 0, stream, 
        x, bstride, static_cast<__nv_bfloat16*>(hi), nullptr, B, T, C, H, W);
  else
    launch_pdl(frames_to_nhwc8_kernel<false>, dim3(grid_for(total, 256, num_sms)), dim3(256), 0, stream, 
        x, bstride, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), B, T, C, H, W);
  VPK_CUDA(cudaGetLastError());
}

void launch_frames_to_nhwc(const float* x, void* out, int dtype, int B, int T, int C, int H, int W, int num_sms,
                           cudaStream_t stream) {
  launch_frames_to_nhwc_strided(x, static_cast<long long>(T) * C * H * W, out, dtype, B, T, C, H, W, num_sms, stream);
}

void launch_cast_f32_to_bf16(const float* in, void* out, long long n, int num_sms, cudaStream_t stream) {
  launch_pdl(cast_kernel, dim3(grid_for((n + 1) / 2, 256, num_sms)), dim3(256), 0, stream, in, static_cast<__nv_bfloat16*>(out), n);
  VPK_CUDA(cudaGetLastError());
}

void launch_metric_partial_sums(const float* pred, const float* target, int B, int P, long long chw, double* scratch,
                                double* out, cudaStream_t stream) {
  VPK_REQUIRE(B > 0 && P > 0 && chw > 0, "metric_partial_sums: empty input");
  launch_pdl(metric_sse_kernel, dim3(static_cast<unsigned>(B) * P), dim3(256), 0, stream, pred, target, chw, scratch);
  launch_pdl(metric_frame_sums_kernel, dim3(P), dim3(256), 0, stream, static_cast<const double*>(scratch), B, P, chw, out);
}

namespace {
size_t ssim_smem_bytes(int R, int W) {     // both strips with halo rows + five transposed moment planes
  return (static_cast<size_t>(R + 10) * 2 * W + static_cast<size_t>(5) * W * (R + 1)) * sizeof(float);
}
}  // namespace

int metric_ssim_strip_rows(int H, int W) {
  if (H < 11 || W < 11) return 0;
  int R = 16;
  while (R > 4 && ssim_smem_bytes(R, W) > 200 * 1024) R /= 2;      // a power of two >= 4 (row blocks of 4)
  return ssim_smem_bytes(R, W) <= 200 * 1024 ? R : 0;
}

long long metric_ssim_scratch_elems(int B, int P, int C, int H, int W) {
  const int R = metric_ssim_strip_rows(H, W);
  if (R == 0 || B <= 0 || P <= 0 || C <= 0) return -1;
  return static_cast<long long>(B) * P * C * ((H - 10 + R - 1) / R);
}

void launch_metric_ssim_sums(const float* pred, const float* target, int B, int P, int C, int H, int W, double* scratch,
                             double* out, cudaStream_t stream) {
  const int R = metric_ssim_strip_rows(H, W);
  VPK_REQUIRE(B > 0 && P > 0 && C > 0 && R > 0, "metric_ssim_sums: images must be at least 11 x 11 (and at most ~4000 wide)");
  SsimWin win;                              // piqa's gaussian_kernel(11, 1.5), fp32 like torch
  float sum = 0.f;
  for (int k = 0; k < 11; ++k) {
    const float d = static_cast<float>(k) - 5.f;
    win.w[k] = expf(-(d * d) / (2.f * 1.5f * 1.5f));
    sum += win.w[k];
  }
  for (int k = 0; k < 11; ++k) win.w[k] /= sum;
  const int S = (H - 10 + R - 1) / R;
  const size_t smem = ssim_smem_bytes(R, W);
  VPK_CUDA(cudaFuncSetAttribute(metric_ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  launch_pdl(metric_ssim_kernel, dim3(static_cast<unsigned>(B) * P * C, S), dim3(256), smem, stream, pred, target, H, W, R, win,
             scratch);
  const double inv = 1.0 / (static_cast<double>(C) * (H - 10) * (W - 10));
  launch_pdl(metric_ssim_sums_kernel, dim3(P), dim3(256), 0, stream, static_cast<const double*>(scratch), B, P, C * S, inv, out);
}

void launch_cast_f32_to_f16(const float* in, void* out, long long n, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(n % 4 == 0, "cast_f32_to_f16: element count must be a multiple of 4");
  launch_pdl(cast_f16_kernel, dim3(grid_for(n / 4, 256, num_sms)), dim3(256), 0, stream, in, static_cast<__half*>(out), n / 4);
  VPK_CUDA(cudaGetLastError());
}

void launch_decouple_reduce(const float* ad, int B, int HW, int C, double* acc, cudaStream_t stream) {
  VPK_REQUIRE(C % 4 == 0, "decouple_reduce: channel count must be a multiple of 4");
  dim3 grid(static_cast<unsigned>(B), static_cast<unsigned>((C + 127) / 128));
  launch_pdl(decouple_reduce_kernel, dim3(grid), dim3(256), 0, stream, ad, B, HW, C, acc);
  VPK_CUDA(cudaGetLastError());
}

void launch_groupnorm_act(const void* in, int in_dtype, void* out, int out_dtype, const void* add, int B, int HW, int C,
                          int Cs_in, int Cs_out, int groups, const float* gamma, const float* beta, float eps,
                          int act, cudaStream_t stream) {
  VPK_REQUIRE(C <= 256 && groups <= 64 && C % groups == 0, "groupnorm: unsupported channel / group count");
  int TC = 1;
  while (TC < C) TC <<= 1;
#define VPK_GN(TI, TO)                                                                                          \
  launch_pdl(groupnorm_act_kernel<TI, TO>, dim3(B), dim3(256), 0, stream, static_cast<const TI*>(in), static_cast<TO*>(out),           \
                                                      static_cast<const TO*>(add), HW, C, Cs_in, Cs_out, groups, TC, \
                                                      gamma, beta, eps, act)
  if (in_dtype == DT_F32 && out_dtype == DT_F32) VPK_GN(float, float);
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16) VPK_GN(float, __nv_bfloat16);
  else if (in_dtype == DT_F32 && out_dtype == DT_F16) VPK_GN(float, __half);
  else if (in_dtype == DT_BF16 && out_dtype == DT_BF16) VPK_GN(__nv_bfloat16, __nv_bfloat16);
  else VPK_THROW(1, "groupnorm: unsupported dtype combination");
#undef VPK_GN
  VPK_CUDA(cudaGetLastError());
}

bool groupnorm_smem_supported(int HW, int C, int groups) {
  return C >= 4 && C <= kGnMaxC && (kGnThreads * 4) % C == 0 && groups <= 64 && C % groups == 0 && (HW * C) % 4 == 0;
}

void launch_groupnorm_smem(const float* in, void* out, void* out_lo, int out_kind, const float* add, int B, int HW,
                           int C, int groups, const float* gamma, const float* beta, float eps, int act,
                           cudaStream_t stream) {
  VPK_REQUIRE(groupnorm_smem_supported(HW, C, groups), "groupnorm_smem: unsupported shape");
  VPK_REQUIRE(out_kind >= 0 && out_kind <= 3 && (out_kind != 2 || out_lo != nullptr), "groupnorm_smem: bad output kind");
  const size_t bytes = static_cast<size_t>(HW) * C * sizeof(float);
  const bool staged = bytes <= 200 * 1024;
#define VPK_GNS(OUT, ST)                                                                                           \
  do {                                                                                                             \
    static std::once_flag once;                                                                                    \
    std::call_once(once, [] {                                                                                      \
      cudaFuncSetAttribute(groupnorm_smem_kernel<OUT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    });                                                                                                            \
    launch_pdl(groupnorm_smem_kernel<OUT, ST>, dim3(B), dim3(kGnThreads), (ST) ? bytes : 0, stream, in, out, out_lo, add, HW, C, groups, \
                                                                                gamma, beta, eps, act);            \
  } while (0)
  if (staged) {
    if (out_kind == 0) VPK_GNS(0, true);
    else if (out_kind == 1) VPK_GNS(1, true);
    else if (out_kind == 3) VPK_GNS(3, true);
    else VPK_GNS(2, true);
  } else {
    if (out_kind == 0) VPK_GNS(0, false);
    else if (out_kind == 1) VPK_GNS(1, false);
    else if (out_kind == 3) VPK_GNS(3, false);
    else VPK_GNS(2, false);
  }
#undef VPK_GNS
  VPK_CUDA(cudaGetLastError());
}

bool groupnorm_apply_supported(int HW, int C, int groups) {
  return C % 4 == 0 && C <= 256 && C % groups == 0 && (HW * C) % 4 == 0;
}

void launch_groupnorm_apply(const void* in, int in_f16, void* out, int out_kind, const float* add, const float* sums, int nslots,
                            int B, int HW, int C, int groups, const float* gamma, const float* beta, float eps, int act, int num_sms,
                            cudaStream_t stream) {
  VPK_REQUIRE(groupnorm_apply_supported(HW, C, groups), "groupnorm_apply: unsupported shape");
  VPK_REQUIRE(out_kind == 0 || out_kind == 1 || out_kind == 3, "groupnorm_apply: bad output kind");
  const int n4 = HW * C / 4;
  // ~2 waves of blocks over the whole batch, at least one block per sample
  const int per = std::max(1, std::min((n4 + 1023) / 1024, std::max(1, 2 * num_sms * 8 / std::max(1, B))));
  const dim3 grid(per, B);
#define VPK_GNA(OUT)                                                                                                          \
  do {                                                                                                                        \
    if (in_f16)                                                                                                               \
      launch_pdl(groupnorm_apply_kernel<OUT, true>, grid, dim3(256), 0, stream, in, out, add, sums, nslots, HW, C, groups, gamma, beta, eps, act); \
    else                                                                                                                      \
      launch_pdl(groupnorm_apply_kernel<OUT, false>, grid, dim3(256), 0, stream, in, out, add, sums, nslots, HW, C, groups, gamma, beta, eps, act); \
  } while (0)
  if (out_kind == 0) VPK_GNA(0);
  else if (out_kind == 1) VPK_GNA(1);
  else VPK_GNA(3);
#undef VPK_GNA
}

void launch_split_bf16(const float* in, void* hi, void* lo, long long n, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(n % 4 == 0, "split_bf16: element count must be a multiple of 4");
  launch_pdl(split_bf16_kernel, dim3(grid_for(n / 4, 256, num_sms)), dim3(256), 0, stream, in, static_cast<__nv_bfloat16*>(hi),
                                                                       static_cast<__nv_bfloat16*>(lo), n / 4);
  VPK_CUDA(cudaGetLastError());
}

void launch_split_f16(const float* in, void* hi, void* lo, long long n, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(n % 4 == 0, "split_f16: element count must be a multiple of 4");
  launch_pdl(split_f16_kernel, dim3(grid_for(n / 4, 256, num_sms)), dim3(256), 0, stream, in, static_cast<__half*>(hi),
             static_cast<__half*>(lo), n / 4);
  VPK_CUDA(cudaGetLastError());
}

void launch_stphy_action_tensor(const float* actions, long long bstride, int a, const float* wl, const float* wh, const float* ww,
                                void* out, int dtype, int B, int T, int H, int W, int C, int IA, cudaStream_t stream) {
  VPK_REQUIRE(a >= 1 && B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && IA >= 1, "stphy_action_tensor: bad shape");
  const size_t smem = static_cast<size_t>(IA) * H * W * sizeof(float);
  VPK_REQUIRE(smem <= 48 * 1024, "stphy_action_tensor: encoded map too large");
  const dim3 g(static_cast<unsigned>(B) * T);
  if (dtype == DT_F32) launch_pdl(stphy_action_tensor_kernel<float>, g, dim3(256), smem, stream, actions, bstride, a, wl, wh, ww, static_cast<float*>(out), B, H, W, C, IA);
  else if (dtype == DT_F16) launch_pdl(stphy_action_tensor_kernel<__half>, g, dim3(256), smem, stream, actions, bstride, a, wl, wh, ww, static_cast<__half*>(out), B, H, W, C, IA);
  else launch_pdl(stphy_action_tensor_kernel<__nv_bfloat16>, g, dim3(256), smem, stream, actions, bstride, a, wl, wh, ww, static_cast<__nv_bfloat16*>(out), B, H, W, C, IA);
}

void launch_inflate_actions(const float* actions, long long bstride, int a, void* out, int dtype, int B, int T, int HW,
                            int a_pad, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(a >= 1 && a <= a_pad && B > 0 && T > 0 && HW > 0, "inflate_actions: bad shape");
  const int g = grid_for(static_cast<long long>(T) * B * HW * a_pad, 256, num_sms);
  if (dtype == DT_F32) launch_pdl(inflate_actions_kernel<float>, dim3(g), dim3(256), 0, stream, actions, bstride, a, static_cast<float*>(out), B, T, HW, a_pad);
  else if (dtype == DT_F16) launch_pdl(inflate_actions_kernel<__half>, dim3(g), dim3(256), 0, stream, actions, bstride, a, static_cast<__half*>(out), B, T, HW, a_pad);
  else launch_pdl(inflate_actions_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, actions, bstride, a, static_cast<__nv_bfloat16*>(out), B, T, HW, a_pad);
  VPK_CUDA(cudaGetLastError());
}

void launch_trajgru_warp(const float* h, const float* flows, int fpix, void* out, int dtype, int B, int H, int W, int C, int L,
                         int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(C % 4 == 0 && L >= 1 && fpix >= 2 * L, "trajgru_warp: bad shape");
  const int g = grid_for(static_cast<long long>(B) * H * W * L * (C / 4), 256, num_sms);
  if (dtype == DT_F32) launch_pdl(trajgru_warp_kernel<float>, dim3(g), dim3(256), 0, stream, h, flows, fpix, static_cast<float*>(out), B, H, W, C, L);
  else if (dtype == DT_F16) launch_pdl(trajgru_warp_kernel<__half>, dim3(g), dim3(256), 0, stream, h, flows, fpix, static_cast<__half*>(out), B, H, W, C, L);
  else launch_pdl(trajgru_warp_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, h, flows, fpix, static_cast<__nv_bfloat16*>(out), B, H, W, C, L);
  VPK_CUDA(cudaGetLastError());
}

void launch_trajgru_gates(const float* i2h, const float* h2h, const float* h, float* h_out, void* h_act, int dtype, long long P,
                          int C, int act, int num_sms, cudaStream_t stream) {
  const int g = grid_for(P * C, 256, num_sms);
  if (dtype == DT_F32) launch_pdl(trajgru_gates_kernel<float>, dim3(g), dim3(256), 0, stream, i2h, h2h, h, h_out, static_cast<float*>(nullptr), P, C, act);
  else if (dtype == DT_F16) launch_pdl(trajgru_gates_kernel<__half>, dim3(g), dim3(256), 0, stream, i2h, h2h, h, h_out, static_cast<__half*>(h_act), P, C, act);
  else launch_pdl(trajgru_gates_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, i2h, h2h, h, h_out, static_cast<__nv_bfloat16*>(h_act), P, C, act);
  VPK_CUDA(cudaGetLastError());
}

void launch_fanout(const float* in, void* hi, void* lo, int hi_dtype, void* cell, int cell_dtype, float* f32, int B, int H,
                   int W, int C, bool norm_w, float eps, int num_sms, cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * H * C;
  const int g = grid_for(rows, 256, num_sms);
#define VPK_FAN(T1, T2)                                                                                                  \
  do {                                                                                                                  \
    if (norm_w) launch_pdl(fanout_kernel<T1, T2, true>, dim3(g), dim3(256), 0, stream, in, static_cast<T1*>(hi), static_cast<T1*>(lo), static_cast<T2*>(cell), f32, rows, W, C, eps); \
    else launch_pdl(fanout_kernel<T1, T2, false>, dim3(g), dim3(256), 0, stream, in, static_cast<T1*>(hi), static_cast<T1*>(lo), static_cast<T2*>(cell), f32, rows, W, C, eps); \
  } while (0)
  if (hi_dtype == DT_F16 && cell_dtype == DT_BF16) VPK_FAN(__half, __nv_bfloat16);
  else if (hi_dtype == DT_F16 && cell_dtype == DT_F16) VPK_FAN(__half, __half);
  else if (hi_dtype == DT_BF16 && cell_dtype == DT_BF16) VPK_FAN(__nv_bfloat16, __nv_bfloat16);
  else if (hi_dtype == DT_F32 && cell_dtype == DT_F32) VPK_FAN(float, float);
  else VPK_THROW(1, "fanout: unsupported dtype combination");
#undef VPK_FAN
  VPK_CUDA(cudaGetLastError());
}

void launch_add_to_act(const void* x, int x_dtype, const float* y, void* out, int out_dtype, long long n, int num_sms,
                       cudaStream_t stream) {
  const int g = grid_for(n, 256, num_sms);
#define VPK_ADD(TX, TO) launch_pdl(add_to_act_kernel<TX, TO>, dim3(g), dim3(256), 0, stream, static_cast<const TX*>(x), y, static_cast<TO*>(out), n)
  if (x_dtype == DT_F32 && out_dtype == DT_F32) VPK_ADD(float, float);
  else if (x_dtype == DT_F32 && out_dtype == DT_F16) VPK_ADD(float, __half);
  else if (x_dtype == DT_F32 && out_dtype == DT_BF16) VPK_ADD(float, __nv_bfloat16);
  else if (x_dtype == DT_F16 && out_dtype == DT_F16) VPK_ADD(__half, __half);
  else if (x_dtype == DT_BF16 && out_dtype == DT_BF16) VPK_ADD(__nv_bfloat16, __nv_bfloat16);
  else VPK_THROW(1, "add_to_act: unsupported dtype combination");
#undef VPK_ADD
  VPK_CUDA(cudaGetLastError());
}

size_t phy_f_tail_smem(int HW, int hid, int Cs, int C) {
  return (static_cast<size_t>((HW * (Cs | 1) + 3) & ~3) + static_cast<size_t>(hid) * C + C) * sizeof(float);
}
bool phy_f_tail_supported(int HW, int hid, int Cs, int groups, int C) {
  return hid <= 64 && groups <= 64 && hid % groups == 0 && C % 16 == 0 && Cs >= hid && Cs % 4 == 0 &&
         phy_f_tail_smem(HW, hid, Cs, C) <= 200 * 1024;
}
void launch_phy_f_tail(const float* f1, const float* h, float* htilde, const float* gamma, const float* beta,
                       const float* w2, const float* b2, int B, int HW, int hid, int Cs, int groups, int C, float eps,
                       cudaStream_t stream) {
  VPK_REQUIRE(phy_f_tail_supported(HW, hid, Cs, groups, C), "phy_f_tail: unsupported shape");
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(phy_f_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  launch_pdl(phy_f_tail_kernel, dim3(B), dim3(kPhyTailThreads), phy_f_tail_smem(HW, hid, Cs, C), stream, f1, h, htilde, gamma,
             beta, w2, b2, HW, hid, Cs, groups, C, eps);
}

void launch_decouple_cos(const float* slots, int nslots, int B, int C, float* term, cudaStream_t stream) {
  VPK_REQUIRE(C >= 1 && C <= 256, "decouple_cos: at most 256 channels");
  launch_pdl(decouple_cos_kernel, dim3(B), dim3(256), 0, stream, slots, nslots, C, term);
}
void launch_decouple_sum(const float* terms, long long n, double* acc, cudaStream_t stream) {
  launch_pdl(decouple_sum_kernel, dim3(1), dim3(256), 0, stream, terms, n, acc);
}

void launch_decouple_finalize(const double* acc, float* aux, double scale, cudaStream_t stream) {
  launch_pdl(decouple_finalize_kernel, dim3(1), dim3(1), 0, stream, acc, aux, scale);
  VPK_CUDA(cudaGetLastError());
}

void launch_patchify_strided(const float* x, long long bstride, void* out, int dtype, int B, int T, int C, int H, int W,
                             int p, int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * T * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    launch_pdl(patchify_kernel<float>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<float*>(out), B, T, C, H, W, p);
  else if (dtype == DT_F16)
    launch_pdl(patchify_kernel<__half>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<__half*>(out), B, T, C, H, W, p);
  else
    launch_pdl(patchify_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, x, bstride, static_cast<__nv_bfloat16*>(out), B, T, C, H, W,
                                                         p);
  VPK_CUDA(cudaGetLastError());
}

void launch_unpatchify(const void* in, float* out, int dtype, int B, int P, int t, int C, int H, int W, int p,
                       int num_sms, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * H * W;
  const int g = grid_for(total, 256, num_sms);
  if (dtype == DT_F32)
    launch_pdl(unpatchify_kernel<float>, dim3(g), dim3(256), 0, stream, static_cast<const float*>(in), out, B, P, t, C, H, W, p);
  else
    launch_pdl(unpatchify_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(in), out, B, P, t, C, H,
                                                           W, p);
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
