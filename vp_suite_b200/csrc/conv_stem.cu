// Direct CUDA-core kernels for the two memory-bound ends of the rollouts, where the contraction is far too small for a
// tensor-core tile (K = 9 * image channels, or N = image channels) and the bound is HBM:
//
//  * conv_stem_kernel:   Conv2d k3 p1 (stride 1 or 2) over image frames stored as 16-bit NHWC with 8 zero-padded channels
//                        (`frames_to_nhwc8`), <= 4 real channels, to 16 / 32 output channels + bias (+ LeakyReLU 0.2).
//                        EF encoder.stage1 (ef_blocks.py:15-49, c -> 16, stride 1) and PhyDNet encoder_E.c1's conv
//                        (model_blocks/conv.py:58-70, c -> 32, stride 2; its GroupNorm follows as a separate pass).
//  * deconv_tail_kernel: ConvTranspose2d k3 s2 p1 output_padding 1 from a 16-bit NHWC feature map to <= 4 image channels
//                        + bias + sigmoid, written as the fp32 NCHW predicted frame AND (optionally) as the 16-bit
//                        8-channel frame the next step's stem conv reads (PhyDNet decoder_D.upc3 + the sigmoid of
//                        models/phydnet.py:87-88 + the feedback of :121): all four output parities in one launch.
//
// Both keep the packed weights in shared memory (every lane reads the same weight: broadcast LDS.128), give each thread
// several output positions so that one weight read feeds 16 FMAs, read 16-byte pixels and write whole 16-byte vectors.
#include "common.h"
#include "conv_stem.h"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace vpk {

namespace {

constexpr int kStemThreads = 128;

template <bool F16> __device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (F16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    } else {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
}

// x: [B][H][W][8] 16-bit (fp16 if F16IN, else bf16); wsm layout [ky][kx][ci][N] fp32; out: [B][OH][OW][N] fp32 (F32OUT)
// or bf16.  One thread = PX consecutive output ROWS of one output column; consecutive threads = consecutive columns, so a
// warp's 16-byte pixel loads cover 512 contiguous bytes (4 lines; stride 2: 8) and its stores 8 lines.  (The first version
// gave a thread PX consecutive columns -- 16 lines per load, 32 per store; the row-wise form measured 125 -> 118 us on
// cfg 5's stem, 37 -> 31 us on PhyDNet's: the kernel is bound by FP32 issue / latency at ~0.3 of the FMA peak, not by the LSU.)
template <int CIN, int N, int STRIDE, bool F16IN, bool F32OUT, int PX>
__global__ void __launch_bounds__(kStemThreads) conv_stem_kernel(const uint4* __restrict__ x, const float* __restrict__ wpk,
                                                                const float* __restrict__ bias, void* __restrict__ out,
                                                                int B, int H, int W, int OH, int OW, int act) {
  __shared__ __align__(16) float s_w[9 * CIN * N];
  __shared__ float s_b[N];
  ptx::pdl_launch_dependents();
  for (int i = threadIdx.x; i < 9 * CIN * N; i += kStemThreads) s_w[i] = wpk[i];
  for (int i = threadIdx.x; i < N; i += kStemThreads) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  ptx::pdl_wait();

  constexpr int NROW = STRIDE * (PX - 1) + 3;             // input rows feeding PX output rows
  const int gy = (OH + PX - 1) / PX;
  const long long total = static_cast<long long>(B) * gy * OW;
  for (long long g = blockIdx.x * static_cast<long long>(kStemThreads) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * kStemThreads) {
    const int ox = static_cast<int>(g % OW);
    const long long r = g / OW;
    const int yg = static_cast<int>(r % gy);
    const int b = static_cast<int>(r / gy);
    const int oy0 = yg * PX;
    float acc[PX][N];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int n = 0; n < N; ++n) acc[p][n] = 0.f;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox * STRIDE - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ix0 + kx;
      const bool okx = ix >= 0 && ix < W;
      float v[NROW][CIN];
#pragma unroll
      for (int rr = 0; rr < NROW; ++rr) {
        const int iy = iy0 + rr;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (okx && iy >= 0 && iy < H) q = __ldg(x + (static_cast<long long>(b) * H + iy) * W + ix);
        float f[8];
        unpack8<F16IN>(q, f);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) v[rr][ci] = f[ci];
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * CIN + ci) * N);
#pragma unroll
          for (int n4 = 0; n4 < N / 4; ++n4) {
            const float4 wv = wr[n4];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float a = v[p * STRIDE + ky][ci];
              acc[p][4 * n4 + 0] = fmaf(a, wv.x, acc[p][4 * n4 + 0]);
              acc[p][4 * n4 + 1] = fmaf(a, wv.y, acc[p][4 * n4 + 1]);
              acc[p][4 * n4 + 2] = fmaf(a, wv.z, acc[p][4 * n4 + 2]);
              acc[p][4 * n4 + 3] = fmaf(a, wv.w, acc[p][4 * n4 + 3]);
            }
          }
        }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (oy0 + p >= OH) break;
      const long long obase = ((static_cast<long long>(b) * OH + oy0 + p) * OW + ox) * N;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        float y = acc[p][n] + s_b[n];
        if (act == ACT_LEAKY) y = y > 0.f ? y : 0.2f * y;
        else if (act == ACT_RELU) y = fmaxf(y, 0.f);
        acc[p][n] = y;
      }
      if (F32OUT) {
        float4* o = reinterpret_cast<float4*>(static_cast<float*>(out) + obase);
#pragma unroll
        for (int n4 = 0; n4 < N / 4; ++n4)
          o[n4] = make_float4(acc[p][4 * n4], acc[p][4 * n4 + 1], acc[p][4 * n4 + 2], acc[p][4 * n4 + 3]);
      } else {
        uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + obase);
#pragma unroll
        for (int n8 = 0; n8 < N / 8; ++n8) {
          uint4 t;
          __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
          for (int k = 0; k < 4; ++k) hh[k] = __floats2bfloat162_rn(acc[p][8 * n8 + 2 * k], acc[p][8 * n8 + 2 * k + 1]);
          o[n8] = t;
        }
      }
    }
  }
}

// x: [B][H][W][CI] fp16 (CI % 8 == 0); wsm layout [ci][ky*3+kx][4] fp32 (output channels padded to 4);
// out: fp32, element (b, co, Y, X) at out + b*oB + (co*2H + Y)*2W + X; fb (optional): fp16 [B][2H][2W][8].
// One thread = two neighbouring input positions (x0, x0+1) of one input row = a 2 x 4 block of output pixels.
__global__ void __launch_bounds__(kStemThreads) deconv_tail_kernel(const uint4* __restrict__ x, const float* __restrict__ wpk,
                                                                  const float* __restrict__ bias, float* __restrict__ out,
                                                                  long long oB, uint4* __restrict__ fb, int B, int H, int W,
                                                                  int CI, int CO, int act) {
  extern __shared__ __align__(16) float s_tw[];     // [CI][9][4]
  __shared__ float s_b[4];
  ptx::pdl_launch_dependents();
  for (int i = threadIdx.x; i < CI * 36; i += kStemThreads) s_tw[i] = wpk[i];
  if (threadIdx.x < 4) s_b[threadIdx.x] = (bias && threadIdx.x < CO) ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  ptx::pdl_wait();

  const int gx = W / 2;
  const int OH = 2 * H, OW = 2 * W;
  const int cq = CI / 8;                      // uint4 per pixel
  const long long total = static_cast<long long>(B) * H * gx;
  for (long long g = blockIdx.x * static_cast<long long>(kStemThreads) + threadIdx.x; g < total;
       g += static_cast<long long>(gridDim.x) * kStemThreads) {
    const int xg = static_cast<int>(g % gx);
    const long long r = g / gx;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    const int x0 = 2 * xg;
    float acc[2][4][4];                        // [ry][ox][co]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][o][c] = 0.f;
    const bool row1 = y + 1 < H;
    const bool col2 = x0 + 2 < W;
    const uint4* p0 = x + ((static_cast<long long>(b) * H + y) * W + x0) * cq;
    const uint4* p1 = p0 + static_cast<long long>(W) * cq;
    for (int q = 0; q < cq; ++q) {
      float v[2][3][8];                        // [iy][ix][channel of this chunk]
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      unpack8<true>(__ldg(p0 + q), v[0][0]);
      unpack8<true>(__ldg(p0 + cq + q), v[0][1]);
      unpack8<true>(col2 ? __ldg(p0 + 2 * cq + q) : z, v[0][2]);
      unpack8<true>(row1 ? __ldg(p1 + q) : z, v[1][0]);
      unpack8<true>(row1 ? __ldg(p1 + cq + q) : z, v[1][1]);
      unpack8<true>((row1 && col2) ? __ldg(p1 + 2 * cq + q) : z, v[1][2]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(s_tw + (q * 8 + j) * 36);
        float4 wt[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) wt[t] = wr[t];
        // output (2y + ry, 2x + rx) gets  sum over (ky, iy) in {ry=0: (1, y); ry=1: (2, y), (0, y+1)}  and likewise in x
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
          for (int xi = 0; xi < 2; ++xi)
#pragma unroll
            for (int rx = 0; rx < 2; ++rx) {
              float* a = acc[ry][2 * xi + rx];
#pragma unroll
              for (int ty = 0; ty < (ry ? 2 : 1); ++ty) {
                const int ky = ry ? (ty ? 0 : 2) : 1, iy = (ry && ty) ? 1 : 0;
#pragma unroll
                for (int tx = 0; tx < (rx ? 2 : 1); ++tx) {
                  const int kx = rx ? (tx ? 0 : 2) : 1, ix = xi + ((rx && tx) ? 1 : 0);
                  const float s = v[iy][ix][j];
                  const float4 wv = wt[ky * 3 + kx];
                  a[0] = fmaf(s, wv.x, a[0]);
                  a[1] = fmaf(s, wv.y, a[1]);
                  a[2] = fmaf(s, wv.z, a[2]);
                  a[3] = fmaf(s, wv.w, a[3]);
                }
              }
            }
      }
    }
#pragma unroll
    for (int ry = 0; ry < 2; ++ry)
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = acc[ry][o][c] + s_b[c];
          if (act == ACT_SIGMOID) t = sigmoid_f(t);
          else if (act == ACT_LEAKY) t = t > 0.f ? t : 0.2f * t;
          acc[ry][o][c] = t;
        }
    float* ob = out + static_cast<long long>(b) * oB;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= CO) break;
#pragma unroll
      for (int ry = 0; ry < 2; ++ry) {
        float4* o = reinterpret_cast<float4*>(ob + (static_cast<long long>(c) * OH + 2 * y + ry) * OW + 2 * x0);
        *o = make_float4(acc[ry][0][c], acc[ry][1][c], acc[ry][2][c], acc[ry][3][c]);
      }
    }
    if (fb != nullptr) {
#pragma unroll
      for (int ry = 0; ry < 2; ++ry)
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const __half2 h01 = __floats2half2_rn(acc[ry][o][0], CO > 1 ? acc[ry][o][1] : 0.f);
          const __half2 h23 = __floats2half2_rn(CO > 2 ? acc[ry][o][2] : 0.f, CO > 3 ? acc[ry][o][3] : 0.f);
          uint4 t = make_uint4(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23), 0u, 0u);
          fb[(static_cast<long long>(b) * OH + 2 * y + ry) * OW + 2 * x0 + o] = t;
        }
    }
  }
}

inline int grid_for_stem(long long n, int num_sms) {
  const long long blocks = (n + kStemThreads - 1) / kStemThreads;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(blocks, static_cast<long long>(num_sms) * 16)));
}

// positions per thread: 4 at N = 16 (64 accumulators); 2 at N = 32 (4 x 32 accumulators need 255 registers: 11 % occupancy)
template <int N> constexpr int stem_px() { return N >= 32 ? 2 : 4; }
template <int CIN, int N, int STRIDE> void launch_stem_t(const StemArgs& a, int OH, int OW, int num_sms, cudaStream_t s) {
  constexpr int PX = stem_px<N>();
  const long long groups = static_cast<long long>(a.B) * ((OH + PX - 1) / PX) * OW;
  const int grid = grid_for_stem(groups, num_sms);
  const uint4* x = static_cast<const uint4*>(a.x);
  if (a.x_f16)
    launch_pdl(conv_stem_kernel<CIN, N, STRIDE, true, true, PX>, dim3(grid), dim3(kStemThreads), 0, s, x, a.w, a.bias, a.out,
               a.B, a.H, a.W, OH, OW, a.act);
  else
    launch_pdl(conv_stem_kernel<CIN, N, STRIDE, false, false, PX>, dim3(grid), dim3(kStemThreads), 0, s, x, a.w, a.bias, a.out,
               a.B, a.H, a.W, OH, OW, a.act);
}
template <int CIN, int N> void launch_stem_s(const StemArgs& a, int OH, int OW, int num_sms, cudaStream_t s) {
  if (a.stride == 1) launch_stem_t<CIN, N, 1>(a, OH, OW, num_sms, s);
  else launch_stem_t<CIN, N, 2>(a, OH, OW, num_sms, s);
}
template <int CIN> void launch_stem_n(const StemArgs& a, int OH, int OW, int num_sms, cudaStream_t s) {
  if (a.N == 16) launch_stem_s<CIN, 16>(a, OH, OW, num_sms, s);
  else launch_stem_s<CIN, 32>(a, OH, OW, num_sms, s);
}

}  // namespace

int stem_cin_slots(int cin) { return cin == 1 ? 1 : (cin == 3 ? 3 : 4); }

bool conv_stem_supported(int k, int stride, int pad, int cin, int N, int H, int W) {
  if (k != 3 || pad != 1 || (stride != 1 && stride != 2) || cin < 1 || cin > 4 || (N != 16 && N != 32)) return false;
  const int OW = (W + 2 - 3) / stride + 1;
  return H >= 1 && W >= 1 && OW >= 1;
}

std::vector<float> conv_stem_pack(const float* w, int N, int cin, int round_to) {
  const int cs = stem_cin_slots(cin);
  std::vector<float> p(static_cast<size_t>(9) * cs * N, 0.f);
  for (int n = 0; n < N; ++n)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < 9; ++t) {
        float v = w[(static_cast<size_t>(n) * cin + ci) * 9 + t];
        // same operand values as the 16-bit tensor-core / CUDA-core GEMM kernels this replaces (their cross-check tests)
        if (round_to == DT_BF16) v = __bfloat162float(__float2bfloat16_rn(v));
        else if (round_to == DT_F16) v = __half2float(__float2half_rn(v));
        p[(static_cast<size_t>(t) * cs + ci) * N + n] = v;
      }
  return p;
}

void launch_conv_stem(const StemArgs& a, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(conv_stem_supported(3, a.stride, 1, a.cin, a.N, a.H, a.W), "conv_stem: unsupported shape");
  // the two instantiated type pairs: fp16 in -> fp32 out (PhyDNet), bf16 in -> bf16 out (EF)
  VPK_REQUIRE((a.x_f16 != 0) == (a.out_f32 != 0), "conv_stem: unsupported operand / output type pair");
  const int OH = (a.H + 2 - 3) / a.stride + 1, OW = (a.W + 2 - 3) / a.stride + 1;
  switch (stem_cin_slots(a.cin)) {
    case 1: launch_stem_n<1>(a, OH, OW, num_sms, stream); break;
    case 3: launch_stem_n<3>(a, OH, OW, num_sms, stream); break;
    default: launch_stem_n<4>(a, OH, OW, num_sms, stream); break;
  }
}

bool deconv_tail_supported(int k, int stride, int pad, int out_pad, int cin, int cout, int H, int W) {
  return k == 3 && stride == 2 && pad == 1 && out_pad == 1 && cin % 8 == 0 && cin >= 8 && cin <= 64 && cout >= 1 &&
         cout <= 4 && W % 2 == 0 && H >= 1;
}

std::vector<float> deconv_tail_pack(const float* w, int cin, int cout, int round_to) {   // w: [cin][cout][3][3]
  std::vector<float> p(static_cast<size_t>(cin) * 36, 0.f);
  for (int ci = 0; ci < cin; ++ci)
    for (int co = 0; co < cout; ++co)
      for (int t = 0; t < 9; ++t) {
        float v = w[(static_cast<size_t>(ci) * cout + co) * 9 + t];
        if (round_to == DT_F16) v = __half2float(__float2half_rn(v));
        p[(static_cast<size_t>(ci) * 9 + t) * 4 + co] = v;
      }
  return p;
}

void launch_deconv_tail(const TailArgs& a, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(deconv_tail_supported(3, 2, 1, 1, a.CI, a.CO, a.H, a.W), "deconv_tail: unsupported shape");
  const long long groups = static_cast<long long>(a.B) * a.H * (a.W / 2);
  launch_pdl(deconv_tail_kernel, dim3(grid_for_stem(groups, num_sms)), dim3(kStemThreads),
             static_cast<size_t>(a.CI) * 36 * sizeof(float), stream, static_cast<const uint4*>(a.x), a.w, a.bias, a.out, a.oB,
             static_cast<uint4*>(a.fb), a.B, a.H, a.W, a.CI, a.CO, a.act);
}

}  // namespace vpk
