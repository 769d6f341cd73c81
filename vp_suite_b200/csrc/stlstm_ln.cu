// Elementwise side of the LayerNorm ST-LSTM cell (stlstm_ln.h): per-sample statistics and the two fused gate kernels.
#include "stlstm_ln.h"

#include "epilogue.cuh"
#include "ptx.cuh"

namespace vpk {

namespace {

// grid (kLnSlices, B, ntens): slice s of sample b of tensor z -> (sum, sum of squares), fixed-order block reduction
__global__ void __launch_bounds__(256) ln_stats_kernel(const LnStatsArgs a) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_s[8], s_q[8];
  const int z = blockIdx.z, b = blockIdx.y, s = blockIdx.x;
  const long long n = a.n[z];
  const long long n4 = n >> 2;                                        // n % 4 == 0 (checked on the host)
  const long long chunk = (n4 + kLnSlices - 1) / kLnSlices;
  const long long lo = s * chunk, hi = min(n4, lo + chunk);
  const float4* p = reinterpret_cast<const float4*>(a.in[z] + static_cast<long long>(b) * n);
  float sum = 0.f, sq = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) {
    const float4 v = p[i];
    sum += (v.x + v.y) + (v.z + v.w);
    sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_s[threadIdx.x >> 5] = sum;
    s_q[threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int w = 0; w < 8; ++w) {
      ts += s_s[w];
      tq += s_q[w];
    }
    float* o = a.part + ((static_cast<long long>(z) * a.B + b) * kLnSlices + s) * 2;
    o[0] = ts;
    o[1] = tq;
  }
}

// mean / rstd of tensor z of sample b from its partial slots (double accumulation, fixed order)
__device__ __forceinline__ void ln_finalize(const float* part, int nslots, int b, double n, float* mean, float* rstd) {
  const float* p = part + static_cast<long long>(b) * nslots * 2;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < nslots; ++i) {
    s += static_cast<double>(p[2 * i]);
    q += static_cast<double>(p[2 * i + 1]);
  }
  const double mu = s / n;
  const double var = fmax(q / n - mu * mu, 0.0);
  *mean = static_cast<float>(mu);
  *rstd = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ln4(float4 v, float mean, float rstd, float4 g, float4 b) {
  return make_float4(fmaf((v.x - mean) * rstd, g.x, b.x), fmaf((v.y - mean) * rstd, g.y, b.y),
                     fmaf((v.z - mean) * rstd, g.z, b.z), fmaf((v.w - mean) * rstd, g.w, b.w));
}
template <typename T> __device__ __forceinline__ void st_act4(T* p, float4 v);
template <> __device__ __forceinline__ void st_act4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st_act4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&a);
  t.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

template <> __device__ __forceinline__ void st_act4<__half>(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 t;
  t.x = *reinterpret_cast<const uint32_t*>(&a);
  t.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// what the activation-type copy of v misses: v - float(T(v)), rounded to T (split operands: hi + lo ~ 22 mantissa bits)
template <typename T> __device__ __forceinline__ float4 lo_part4(float4 v);
template <> __device__ __forceinline__ float4 lo_part4<float>(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float4 lo_part4<__half>(float4 v) {
  return make_float4(v.x - __half2float(__float2half_rn(v.x)), v.y - __half2float(__float2half_rn(v.y)),
                     v.z - __half2float(__float2half_rn(v.z)), v.w - __half2float(__float2half_rn(v.w)));
}
template <> __device__ __forceinline__ float4 lo_part4<__nv_bfloat16>(float4 v) {
  return make_float4(v.x - __bfloat162float(__float2bfloat16_rn(v.x)), v.y - __bfloat162float(__float2bfloat16_rn(v.y)),
                     v.z - __bfloat162float(__float2bfloat16_rn(v.z)), v.w - __bfloat162float(__float2bfloat16_rn(v.w)));
}

// grid (item blocks, sample chunks): one thread = one (position, four channels) for kGateNB consecutive samples.  The
// LayerNorm affine parameters depend on (position, channel) only, so they are read once per gate group and reused for
// the whole chunk of samples (they would otherwise be two thirds of the kernel's L2 -> SM traffic: 28 parameter
// vectors against 14 data vectors per item).  Three passes over the chunk: c group, m group, output-gate part.
// HBM-bound (11 KB per position and sample): four resident blocks per SM (64 registers, a few spilled words) instead of two
// (107 registers) put enough loads in flight -- 261 -> 183 us per launch at cfg 3's shape (0.6 of the copy bandwidth).
constexpr int kGateNB = 8;
template <typename T>
__global__ void __launch_bounds__(256, 4) stlstm_ln_gates_kernel(const StLnGatesArgs a) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_mean[kGateNB][3], s_rstd[kGateNB][3];
  const int C = a.C, HW = a.HW, cq = C >> 2;
  const int b0 = blockIdx.y * kGateNB;
  const int nb = min(kGateNB, a.B - b0);
  if (threadIdx.x < 3 * nb) {
    const int j = threadIdx.x / 3, z = threadIdx.x - 3 * j;
    const double n = static_cast<double>(HW) * C * (z == 0 ? 7 : z == 1 ? 4 : 3);
    ln_finalize(a.part[z], a.nslots[z], b0 + j, n, &s_mean[j][z], &s_rstd[j][z]);
  }
  __syncthreads();
  const int it = blockIdx.x * 256 + threadIdx.x;
  if (it >= HW * cq) return;
  const int hw = it / cq, ch = (it - hw * cq) * 4;
  const long long ax = static_cast<long long>(hw) * 7 * C + ch, ah = static_cast<long long>(hw) * 4 * C + ch,
                  am = static_cast<long long>(hw) * 3 * C + ch;
  const float fb = a.forget_bias;
#define VPK_GATE3(f, i0, i1, f0, f1, g0, g1, d, s)                                                       \
  {                                                                                                      \
    const float i_ = sigmoid_f(i0.f + i1.f), f_ = sigmoid_f(f0.f + f1.f + fb), g_ = tanh_f(g0.f + g1.f); \
    d.f = i_ * g_;                                                                                       \
    s.f = fmaf(f_, s.f, d.f);                                                                            \
  }
  {   // ---- temporal memory: i, f, g from conv_x slices 0..2 and conv_h slices 0..2 ----
    const float4 g0 = ld4(a.gx + ax), g1 = ld4(a.gx + ax + C), g2 = ld4(a.gx + ax + 2 * C);
    const float4 e0 = ld4(a.bx + ax), e1 = ld4(a.bx + ax + C), e2 = ld4(a.bx + ax + 2 * C);
    const float4 p0 = ld4(a.gh + ah), p1 = ld4(a.gh + ah + C), p2 = ld4(a.gh + ah + 2 * C);
    const float4 q0 = ld4(a.bh + ah), q1 = ld4(a.bh + ah + C), q2 = ld4(a.bh + ah + 2 * C);
    for (int j = 0; j < nb; ++j) {
      const long long pos = static_cast<long long>(b0 + j) * HW + hw;
      const float* X = a.X + pos * 7 * C + ch;
      const float* H = a.H + pos * 4 * C + ch;
      const float mx = s_mean[j][0], rx = s_rstd[j][0], mh = s_mean[j][1], rh = s_rstd[j][1];
      const float4 ix = ln4(ld4(X), mx, rx, g0, e0), fx = ln4(ld4(X + C), mx, rx, g1, e1), gx = ln4(ld4(X + 2 * C), mx, rx, g2, e2);
      const float4 ih = ln4(ld4(H), mh, rh, p0, q0), fh = ln4(ld4(H + C), mh, rh, p1, q1), gh = ln4(ld4(H + 2 * C), mh, rh, p2, q2);
      float4 cv = ld4(a.c + pos * C + ch), dc;
      VPK_GATE3(x, ix, ih, fx, fh, gx, gh, dc, cv)
      VPK_GATE3(y, ix, ih, fx, fh, gx, gh, dc, cv)
      VPK_GATE3(z, ix, ih, fx, fh, gx, gh, dc, cv)
      VPK_GATE3(w, ix, ih, fx, fh, gx, gh, dc, cv)
      *reinterpret_cast<float4*>(a.c + pos * C + ch) = cv;
      st_act4<T>(static_cast<T*>(a.mem) + pos * 2 * C + ch, cv);
      st_act4<T>(static_cast<T*>(a.dc) + pos * C + ch, dc);
    }
  }
  {   // ---- spatio-temporal memory: i', f', g' from conv_x slices 3..5 and conv_m slices 0..2 ----
    const float4 g0 = ld4(a.gx + ax + 3 * C), g1 = ld4(a.gx + ax + 4 * C), g2 = ld4(a.gx + ax + 5 * C);
    const float4 e0 = ld4(a.bx + ax + 3 * C), e1 = ld4(a.bx + ax + 4 * C), e2 = ld4(a.bx + ax + 5 * C);
    const float4 p0 = ld4(a.gm + am), p1 = ld4(a.gm + am + C), p2 = ld4(a.gm + am + 2 * C);
    const float4 q0 = ld4(a.bm + am), q1 = ld4(a.bm + am + C), q2 = ld4(a.bm + am + 2 * C);
    for (int j = 0; j < nb; ++j) {
      const long long pos = static_cast<long long>(b0 + j) * HW + hw;
      const float* X = a.X + pos * 7 * C + ch;
      const float* M = a.M + pos * 3 * C + ch;
      const float mx = s_mean[j][0], rx = s_rstd[j][0], mm = s_mean[j][2], rm = s_rstd[j][2];
      const float4 ix = ln4(ld4(X + 3 * C), mx, rx, g0, e0), fx = ln4(ld4(X + 4 * C), mx, rx, g1, e1),
                   gx = ln4(ld4(X + 5 * C), mx, rx, g2, e2);
      const float4 im = ln4(ld4(M), mm, rm, p0, q0), fm = ln4(ld4(M + C), mm, rm, p1, q1), gm = ln4(ld4(M + 2 * C), mm, rm, p2, q2);
      float4 mv = ld4(a.m + pos * C + ch), dm;
      VPK_GATE3(x, ix, im, fx, fm, gx, gm, dm, mv)
      VPK_GATE3(y, ix, im, fx, fm, gx, gm, dm, mv)
      VPK_GATE3(z, ix, im, fx, fm, gx, gm, dm, mv)
      VPK_GATE3(w, ix, im, fx, fm, gx, gm, dm, mv)
      *reinterpret_cast<float4*>(a.m + pos * C + ch) = mv;
      st_act4<T>(static_cast<T*>(a.mem) + pos * 2 * C + C + ch, mv);
      st_act4<T>(static_cast<T*>(a.m_act) + pos * C + ch, mv);
      if (a.m_act_lo != nullptr) st_act4<T>(static_cast<T*>(a.m_act_lo) + pos * C + ch, lo_part4<T>(mv));
      st_act4<T>(static_cast<T*>(a.dm) + pos * C + ch, dm);
    }
  }
#undef VPK_GATE3
  {   // ---- output-gate part: conv_x slice 6 + conv_h slice 3 ----
    const float4 g0 = ld4(a.gx + ax + 6 * C), e0 = ld4(a.bx + ax + 6 * C), p0 = ld4(a.gh + ah + 3 * C), q0 = ld4(a.bh + ah + 3 * C);
    for (int j = 0; j < nb; ++j) {
      const long long pos = static_cast<long long>(b0 + j) * HW + hw;
      const float4 ox = ln4(ld4(a.X + pos * 7 * C + 6 * C + ch), s_mean[j][0], s_rstd[j][0], g0, e0);
      const float4 oh = ln4(ld4(a.H + pos * 4 * C + 3 * C + ch), s_mean[j][1], s_rstd[j][1], p0, q0);
      *reinterpret_cast<float4*>(a.opart + pos * C + ch) = make_float4(ox.x + oh.x, ox.y + oh.y, ox.z + oh.z, ox.w + oh.w);
    }
  }
}

// Action-conditional ST-LSTM gate kernel (model_blocks/predrnn.py:142-164), with or without LayerNorm: grid (item blocks,
// samples), one thread = one (sample, position, four channels).  Same outputs as stlstm_ln_gates_kernel; the i, f, g, o
// terms of conv_h are multiplied by those of conv_a (:149) after their (optional) LayerNorms.  These cells run on
// (patch_h / 4) x (patch_w / 4) latents -- a few thousand items -- so the kernel is kept simple.
template <typename T>
__global__ void __launch_bounds__(256) stlstm_ac_gates_kernel(const StLnGatesArgs a) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_mean[4], s_rstd[4];
  const int C = a.C, HW = a.HW, cq = C >> 2;
  const int b = blockIdx.y;
  const bool ln = a.use_ln != 0;
  if (threadIdx.x < 4) {
    const int z = threadIdx.x;
    if (ln) {
      const double n = static_cast<double>(HW) * C * (z == 0 ? 7 : z == 2 ? 3 : 4);
      ln_finalize(z == 3 ? a.part_a : a.part[z], z == 3 ? a.nslots_a : a.nslots[z], b, n, &s_mean[z], &s_rstd[z]);
    } else {
      s_mean[z] = 0.f;
      s_rstd[z] = 1.f;
    }
  }
  __syncthreads();
  const int it = blockIdx.x * 256 + threadIdx.x;
  if (it >= HW * cq) return;
  const int hw = it / cq, ch = (it - hw * cq) * 4;
  const long long pos = static_cast<long long>(b) * HW + hw;
  auto nrm = [&](const float* raw, const float* g, const float* be, int z, int k, int slice) {
    const float4 v = ld4(raw + pos * k * C + slice * C + ch);
    if (!ln) return v;
    const long long ap = static_cast<long long>(hw) * k * C + slice * C + ch;
    return ln4(v, s_mean[z], s_rstd[z], ld4(g + ap), ld4(be + ap));
  };
  auto mul4 = [](float4 p, float4 q) { return make_float4(p.x * q.x, p.y * q.y, p.z * q.z, p.w * q.w); };
  const float fb = a.forget_bias;
#define VPK_GATE3(f, i0, i1, f0, f1, g0, g1, d, s)                                                       \
  {                                                                                                      \
    const float i_ = sigmoid_f(i0.f + i1.f), f_ = sigmoid_f(f0.f + f1.f + fb), g_ = tanh_f(g0.f + g1.f); \
    d.f = i_ * g_;                                                                                       \
    s.f = fmaf(f_, s.f, d.f);                                                                            \
  }
  {   // temporal memory: conv_x slices 0..2, (conv_h * conv_a) slices 0..2
    const float4 ix = nrm(a.X, a.gx, a.bx, 0, 7, 0), fx = nrm(a.X, a.gx, a.bx, 0, 7, 1), gx = nrm(a.X, a.gx, a.bx, 0, 7, 2);
    const float4 ih = mul4(nrm(a.H, a.gh, a.bh, 1, 4, 0), nrm(a.A, a.ga, a.ba, 3, 4, 0));
    const float4 fh = mul4(nrm(a.H, a.gh, a.bh, 1, 4, 1), nrm(a.A, a.ga, a.ba, 3, 4, 1));
    const float4 gh = mul4(nrm(a.H, a.gh, a.bh, 1, 4, 2), nrm(a.A, a.ga, a.ba, 3, 4, 2));
    float4 cv = ld4(a.c + pos * C + ch), dc;
    VPK_GATE3(x, ix, ih, fx, fh, gx, gh, dc, cv)
    VPK_GATE3(y, ix, ih, fx, fh, gx, gh, dc, cv)
    VPK_GATE3(z, ix, ih, fx, fh, gx, gh, dc, cv)
    VPK_GATE3(w, ix, ih, fx, fh, gx, gh, dc, cv)
    *reinterpret_cast<float4*>(a.c + pos * C + ch) = cv;
    st_act4<T>(static_cast<T*>(a.mem) + pos * 2 * C + ch, cv);
    st_act4<T>(static_cast<T*>(a.dc) + pos * C + ch, dc);
  }
  {   // spatio-temporal memory: conv_x slices 3..5, conv_m slices 0..2
    const float4 ix = nrm(a.X, a.gx, a.bx, 0, 7, 3), fx = nrm(a.X, a.gx, a.bx, 0, 7, 4), gx = nrm(a.X, a.gx, a.bx, 0, 7, 5);
    const float4 im = nrm(a.M, a.gm, a.bm, 2, 3, 0), fm = nrm(a.M, a.gm, a.bm, 2, 3, 1), gm = nrm(a.M, a.gm, a.bm, 2, 3, 2);
    float4 mv = ld4(a.m + pos * C + ch), dm;
    VPK_GATE3(x, ix, im, fx, fm, gx, gm, dm, mv)
    VPK_GATE3(y, ix, im, fx, fm, gx, gm, dm, mv)
    VPK_GATE3(z, ix, im, fx, fm, gx, gm, dm, mv)
    VPK_GATE3(w, ix, im, fx, fm, gx, gm, dm, mv)
    *reinterpret_cast<float4*>(a.m + pos * C + ch) = mv;
    st_act4<T>(static_cast<T*>(a.mem) + pos * 2 * C + C + ch, mv);
    st_act4<T>(static_cast<T*>(a.m_act) + pos * C + ch, mv);
    st_act4<T>(static_cast<T*>(a.dm) + pos * C + ch, dm);
  }
#undef VPK_GATE3
  {   // output-gate part: conv_x slice 6 + (conv_h * conv_a) slice 3
    const float4 ox = nrm(a.X, a.gx, a.bx, 0, 7, 6);
    const float4 oh = mul4(nrm(a.H, a.gh, a.bh, 1, 4, 3), nrm(a.A, a.ga, a.ba, 3, 4, 3));
    *reinterpret_cast<float4*>(a.opart + pos * C + ch) = make_float4(ox.x + oh.x, ox.y + oh.y, ox.z + oh.z, ox.w + oh.w);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) stlstm_ln_out_kernel(const StLnOutArgs a) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  __shared__ float s_mean, s_rstd;
  const int b = blockIdx.y;
  const int C = a.C, HW = a.HW, cq = C >> 2;
  const bool ln = a.use_ln != 0;
  if (threadIdx.x == 0 && ln) ln_finalize(a.part, a.nslots, b, static_cast<double>(HW) * C, &s_mean, &s_rstd);
  __syncthreads();
  const float mo = ln ? s_mean : 0.f, ro = ln ? s_rstd : 1.f;
  const int items = HW * cq;
  for (int it = blockIdx.x * 256 + threadIdx.x; it < items; it += gridDim.x * 256) {
    const int hw = it / cq, ch = (it - hw * cq) * 4;
    const long long pos = static_cast<long long>(b) * HW + hw;
    const long long ao = static_cast<long long>(hw) * C + ch;
    const float4 o = ln ? ln4(ld4(a.O + pos * C + ch), mo, ro, ld4(a.go + ao), ld4(a.bo + ao)) : ld4(a.O + pos * C + ch);
    const float4 p = ld4(a.opart + pos * C + ch), l = ld4(a.Lraw + pos * C + ch);
    const float4 h = make_float4(sigmoid_f(p.x + o.x) * tanh_f(l.x), sigmoid_f(p.y + o.y) * tanh_f(l.y),
                                 sigmoid_f(p.z + o.z) * tanh_f(l.z), sigmoid_f(p.w + o.w) * tanh_f(l.w));
    st_act4<T>(static_cast<T*>(a.h) + pos * C + ch, h);
    if (a.h_lo != nullptr) st_act4<T>(static_cast<T*>(a.h_lo) + pos * C + ch, lo_part4<T>(h));
    if (a.h32 != nullptr) *reinterpret_cast<float4*>(a.h32 + pos * C + ch) = h;
  }
}

inline dim3 sample_grid(int items, int B, int num_sms) {
  const int per = std::max(1, std::min((items + 255) / 256, std::max(1, 4 * num_sms / std::max(1, B))));
  return dim3(static_cast<unsigned>(per), static_cast<unsigned>(B));
}

}  // namespace

void launch_ln_stats(const LnStatsArgs& a, cudaStream_t stream) {
  VPK_REQUIRE(a.ntens >= 1 && a.ntens <= 3 && a.B > 0, "ln_stats: bad arguments");
  for (int z = 0; z < a.ntens; ++z) VPK_REQUIRE(a.n[z] > 0 && a.n[z] % 4 == 0, "ln_stats: sample size must be a multiple of 4");
  launch_pdl(ln_stats_kernel, dim3(kLnSlices, a.B, a.ntens), dim3(256), 0, stream, a);
}

void launch_stlstm_ln_gates(const StLnGatesArgs& a, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(a.C % 4 == 0 && a.B > 0 && a.HW > 0, "stlstm_ln_gates: bad shape");
  if (a.A != nullptr) {      // action-conditional cell (with or without LayerNorm)
    const dim3 grid(static_cast<unsigned>((a.HW * (a.C / 4) + 255) / 256), static_cast<unsigned>(a.B));
    if (a.dtype == DT_F32) launch_pdl(stlstm_ac_gates_kernel<float>, grid, dim3(256), 0, stream, a);
    else if (a.dtype == DT_F16) launch_pdl(stlstm_ac_gates_kernel<__half>, grid, dim3(256), 0, stream, a);
    else launch_pdl(stlstm_ac_gates_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, a);
    return;
  }
  VPK_REQUIRE(a.use_ln != 0, "stlstm_ln_gates: the plain cell without LayerNorm runs through the fused conv epilogues");
  const dim3 grid(static_cast<unsigned>((a.HW * (a.C / 4) + 255) / 256), static_cast<unsigned>((a.B + kGateNB - 1) / kGateNB));
  if (a.dtype == DT_F32) launch_pdl(stlstm_ln_gates_kernel<float>, grid, dim3(256), 0, stream, a);
  else if (a.dtype == DT_F16) launch_pdl(stlstm_ln_gates_kernel<__half>, grid, dim3(256), 0, stream, a);
  else launch_pdl(stlstm_ln_gates_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, a);
}

void launch_stlstm_ln_out(const StLnOutArgs& a, int num_sms, cudaStream_t stream) {
  VPK_REQUIRE(a.C % 4 == 0 && a.B > 0 && a.HW > 0, "stlstm_ln_out: bad shape");
  const dim3 grid = sample_grid(a.HW * (a.C / 4), a.B, num_sms);
  if (a.dtype == DT_F32) launch_pdl(stlstm_ln_out_kernel<float>, grid, dim3(256), 0, stream, a);
  else if (a.dtype == DT_F16) launch_pdl(stlstm_ln_out_kernel<__half>, grid, dim3(256), 0, stream, a);
  else launch_pdl(stlstm_ln_out_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, a);
}

}  // namespace vpk
