// Lean epilogue for the tcgen05 kernels: same arithmetic as epilogue.cuh (which stays the CUDA-core kernels' epilogue
// and the cross-check), specialised at COMPILE time on the epilogue kind, with everything that does not depend on the
// channel chunk hoisted to once per (thread, tile): position decode, 64-bit offsets, layout selection.  The packed
// bias lives in shared memory (every lane reads the same float4: one broadcast wavefront).  One call = 8 channels.
#pragma once
#include "ptx.cuh"
#include "epilogue.cuh"

namespace vpk {

struct EpiTile {            // per thread, per output tile
  long long out_off;        // b*oB + y*oY + x*oX of the primary output
  long long pix_c;          // dense pixel index * C  (NHWC tensors: t1, res, q0, h32, PHY state)
  long long pix_t0;         // dense pixel index * t0_pix (ST-LSTM mem buffer)
  long long st_off;         // fp32 state tensors: offset of channel 0 at this position ...
  long long st_g;           // ... and the stride between channel quads (0: NHWC, channels contiguous)
  long long pp_off;         // peepholes (no batch dimension), same layout rule
  int pos, hw;              // y * W + x and H * W (packed bf16 peepholes)
};

__device__ __forceinline__ EpiTile epi_tile(const EpiParams& E, int b, int y, int x, int H, int W) {
  EpiTile t;
  const long long pix = (static_cast<long long>(b) * H + y) * W + x;
  t.out_off = b * E.oB + y * E.oY + x * E.oX;
  t.pix_c = pix * E.C;
  t.pix_t0 = pix * E.t0_pix;
  t.pos = y * W + x;
  t.hw = H * W;
  if (E.state_c4) {
    const long long hw = static_cast<long long>(H) * W;
    const long long p = static_cast<long long>(y) * W + x;
    t.st_off = (static_cast<long long>(b) * (E.C >> 2) * hw + p) * 4;
    t.pp_off = p * 4;
    t.st_g = hw * 4;
  } else {
    t.st_off = pix * E.C;
    t.pp_off = (static_cast<long long>(y) * W + x) * E.C;
    t.st_g = 0;
  }
  return t;
}

// 8 consecutive channels of an fp32 state tensor (full chunks only: the tcgen05 path requires C % 8 == 0 for these)
__device__ __forceinline__ void ld_state8(const float* base, long long off, long long g, int ch, float (&v)[8]) {
  const float* p = base + off + (g ? (ch >> 2) * g : ch);
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + (g ? g : 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st_state8(float* base, long long off, long long g, int ch, const float (&v)[8]) {
  float* p = base + off + (g ? (ch >> 2) * g : ch);
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + (g ? g : 4)) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void ld_bf16x8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    v[2 * k] = f.x; v[2 * k + 1] = f.y;
  }
}
__device__ __forceinline__ void st_f32x8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ld_f32x8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---- GroupNorm statistics of a conv output, accumulated by the conv's own epilogue ---------------------------------
// A thread owns up to four 8-channel chunks (k = 0..3) of one output position; s[(k * (8/GS) + j) * 2 + {0, 1}] holds the
// sum / sum of squares of group j of chunk k (GS = channels per group, 8 % GS == 0): at most 16 values per thread.
template <int GS>
__device__ __forceinline__ void gn_accum(float (&s)[16], int k, const float (&v)[8]) {
  constexpr int GPC = 8 / GS;
#pragma unroll
  for (int j = 0; j < GPC; ++j) {
    const int idx = (k * GPC + j) * 2;
    if (idx < 16) {
#pragma unroll
      for (int c = 0; c < GS; ++c) {
        const float x = v[j * GS + c];
        s[idx] += x;
        s[idx + 1] = fmaf(x, x, s[idx + 1]);
      }
    }
  }
}
// Sums the 16 per-thread values over the 32 lanes of a warp with 8 + 4 + 2 + 1 + 1 shuffles (each step halves the values a
// lane keeps).  Returns the total of value gn_lane_value(lane) -- the same in lanes 2i and 2i + 1.
__device__ __forceinline__ float gn_warp_reduce16(float (&s)[16], int lane) {
#pragma unroll
  for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? s[i] : s[i + w];
      const float keep = up ? s[i + w] : s[i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return s[0] + __shfl_xor_sync(0xffffffffu, s[0], 1);
}
__device__ __forceinline__ int gn_lane_value(int lane) { return lane >> 1; }   // bits 4..1 of the lane = value index
// The same for 32 values (16 + 8 + 4 + 2 + 1 shuffles): lane i ends up with the warp total of value i.
__device__ __forceinline__ float warp_reduce32(float (&s)[32], int lane) {
#pragma unroll
  for (int w = 16, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? s[i] : s[i + w];
      const float keep = up ? s[i + w] : s[i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return s[0];
}

// Whether the fast path may be used for this launch: whole 8-channel chunks, 16-byte aligned rows.
__host__ __device__ inline bool epi_tc_fast_ok(const EpiParams& E) {
  if (E.C % 8 != 0) return false;
  if (E.kind == EPI_BIAS_ACT) {
    if (E.oC != 1) return false;                             // channel-strided (NCHW) outputs use the generic path
    const long long m = E.out_f32 ? 4 : 8;
    if ((E.oB % m) || (E.oY % m) || (E.oX % m)) return false;
  }
  return true;
}

// ---- prefetch: the global operands of channels [ch, ch+8) ---------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void epi_tc_prefetch(const EpiParams& E, const EpiTile& t, int ch, EpiOperands<8>& o) {
  using bf16 = __nv_bfloat16;
  if constexpr (KIND == EPI_DECOUPLE || KIND == EPI_SUBPIX) return;      // handled inside the halo kernel itself
  if constexpr (KIND == EPI_LSTM) {
    ld_state8(E.s0, t.st_off, t.st_g, ch, o.a);
    if (E.p0 != nullptr) {
      ld_state8(E.p0, t.pp_off, t.st_g, ch, o.b);
      ld_state8(E.p1, t.pp_off, t.st_g, ch, o.c);
      ld_state8(E.p2, t.pp_off, t.st_g, ch, o.d);
    }
  } else if constexpr (KIND == EPI_ST_C || KIND == EPI_ST_M || KIND == EPI_ST_O) {
    ld_state8(E.s0, t.st_off, t.st_g, ch, o.a);
  } else if constexpr (KIND == EPI_ST_O1) {
    ld_state8(E.s0, t.st_off, t.st_g, ch, o.a);
    ld_f32x8(E.res + t.pix_c + ch, o.b);
  } else if constexpr (KIND == EPI_PHY_GATE) {
    ld_bf16x8(static_cast<const bf16*>(E.q0) + t.pix_c + ch, o.a);
    ld_f32x8(E.res + t.pix_c + ch, o.b);
  } else {   // EPI_BIAS_ACT
    if (E.res != nullptr) ld_f32x8(E.res + t.pix_c + ch, o.a);
  }
}

// ---- finish: acc[g][j] = gate g of channel ch + j; s_bias = packed bias of this N tile in shared memory ----------
template <int KIND, int G>
__device__ __forceinline__ void epi_tc_finish(const EpiParams& E, const EpiTile& t, int ch, const float* s_bias,
                                              float (&acc)[G][8], EpiOperands<8>& o) {
  using bf16 = __nv_bfloat16;
  if constexpr (KIND == EPI_DECOUPLE || KIND == EPI_SUBPIX) return;      // handled inside the halo kernel itself
  if (s_bias != nullptr) {   // packed order (ch + j) * G + g: 8*G consecutive floats
    const float4* bp = reinterpret_cast<const float4*>(s_bias + ch * G);
#pragma unroll
    for (int q = 0; q < 2 * G; ++q) {
      const float4 bv = bp[q];
      const float e[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int idx = 4 * q + r;      // = j * G + g
        acc[idx % G][idx / G] += e[r];
      }
    }
  }
  if constexpr (KIND == EPI_BIAS_ACT) {
    float v[8];
    switch (E.act) {
      case ACT_LEAKY:
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc[0][j] > 0.f ? acc[0][j] : 0.2f * acc[0][j];
        break;
      case ACT_SIGMOID:
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = sigmoid_fast(acc[0][j]);
        break;
      case ACT_RELU:
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[0][j], 0.f);
        break;
      default:
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc[0][j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = v[j];      // the activated values, for callers that accumulate statistics
    if (E.res != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += o.a[j];
    }
    if (E.out_f32) st_f32x8(static_cast<float*>(E.out) + t.out_off + ch, v);
    else st_bf16x8(static_cast<bf16*>(E.out) + t.out_off + ch, v);
  } else if constexpr (KIND == EPI_PHY_GATE) {   // h' = h~ + sigmoid(acc) * (x - h~)
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = o.b[j] + sigmoid_fast(acc[0][j]) * (o.a[j] - o.b[j]);
    st_f32x8(E.s0 + t.pix_c + ch, v);
    st_bf16x8(static_cast<bf16*>(E.out) + t.pix_c + ch, v);
  } else if constexpr (KIND == EPI_LSTM) {
    float h[8];
    if (E.p0 != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float ig = sigmoid_fast(fmaf(o.b[j], o.a[j], acc[0][j]));
        const float fg = sigmoid_fast(fmaf(o.c[j], o.a[j], acc[1][j]));
        const float cn = fmaf(fg, o.a[j], ig * tanh_fast(acc[2][j]));
        const float og = sigmoid_fast(fmaf(o.d[j], cn, acc[3][j]));
        o.a[j] = cn;
        h[j] = og * tanh_fast(cn);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float cn = fmaf(sigmoid_fast(acc[1][j]), o.a[j], sigmoid_fast(acc[0][j]) * tanh_fast(acc[2][j]));
        o.a[j] = cn;
        h[j] = sigmoid_fast(acc[3][j]) * tanh_fast(cn);
      }
    }
    st_state8(E.s0, t.st_off, t.st_g, ch, o.a);
    st_bf16x8(static_cast<bf16*>(E.out) + t.out_off + ch, h);
    if (E.h32 != nullptr) st_f32x8(E.h32 + t.pix_c + ch, h);
  } else if constexpr (KIND == EPI_ST_C) {   // acc = (i, f, g, o_x + o_h)
    if (E.variant == 1) {   // Causal LSTM spatial memory (causal.h): acc = (i', f', g', m_m)
      float mn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        mn[j] = fmaf(sigmoid_fast(acc[1][j] + E.forget_bias), tanh_fast(acc[3][j]), sigmoid_fast(acc[0][j]) * tanh_fast(acc[2][j]));
      st_state8(E.s0, t.st_off, t.st_g, ch, mn);
      st_bf16x8(static_cast<bf16*>(E.t0) + t.pix_t0 + ch, mn);
      return;
    }
    float dc[8], op[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = sigmoid_fast(acc[0][j]);
      const float fg = sigmoid_fast(acc[1][j] + E.forget_bias);
      dc[j] = ig * tanh_fast(acc[2][j]);
      o.a[j] = fmaf(fg, o.a[j], dc[j]);
      op[j] = acc[3][j];
    }
    st_state8(E.s0, t.st_off, t.st_g, ch, o.a);
    st_state8(E.s1, t.st_off, t.st_g, ch, op);
    st_bf16x8(static_cast<bf16*>(E.t0) + t.pix_t0 + ch, o.a);
    if (E.t1 != nullptr) st_bf16x8(static_cast<bf16*>(E.t1) + t.pix_c + ch, dc);
  } else if constexpr (KIND == EPI_ST_M) {   // acc = (i', f', g')
    float dm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = sigmoid_fast(acc[0][j]);
      const float fg = sigmoid_fast(acc[1][j] + E.forget_bias);
      dm[j] = ig * tanh_fast(acc[2][j]);
      o.a[j] = fmaf(fg, o.a[j], dm[j]);
    }
    st_state8(E.s0, t.st_off, t.st_g, ch, o.a);
    st_bf16x8(static_cast<bf16*>(E.t0) + t.pix_t0 + ch, o.a);
    if (E.t1 != nullptr) st_bf16x8(static_cast<bf16*>(E.t1) + t.pix_c + ch, dm);
  } else if constexpr (KIND == EPI_ST_O1) {   // (o.a = o_part)  variant & 2: acc = conv_o(mem), o.b = conv_last(mem); else swapped
    float h[8];
    const bool swapped = (E.variant & 2) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gate = o.a[j] + (swapped ? acc[0][j] : o.b[j]);
      const float last = swapped ? o.b[j] : acc[0][j];
      h[j] = ((E.variant & 1) ? tanh_fast(gate) : sigmoid_fast(gate)) * tanh_fast(last);
    }
    st_bf16x8(static_cast<bf16*>(E.out) + t.out_off + ch, h);
  } else {   // EPI_ST_O: acc = (conv_o(mem), conv_last(mem))
    float h[8];
    if (E.variant == 2) {   // gradient highway unit (causal.h): acc = (p, u), o.a = z
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = sigmoid_fast(acc[1][j]);
        h[j] = fmaf(u, o.a[j] - tanh_fast(acc[0][j]), tanh_fast(acc[0][j]));      // u z + (1 - u) tanh(p)
      }
      st_state8(E.s0, t.st_off, t.st_g, ch, h);
    } else if (E.variant == 1) {   // Causal LSTM output: tanh output gate
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = tanh_fast(o.a[j] + acc[0][j]) * tanh_fast(acc[1][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = sigmoid_fast(o.a[j] + acc[0][j]) * tanh_fast(acc[1][j]);
    }
    st_bf16x8(static_cast<bf16*>(E.out) + t.out_off + ch, h);
  }
}

// ---- ConvLSTM epilogue with the operands of a WHOLE tile in flight ----------------------------------------------
// The gate GEMMs of the Shi et al. ConvLSTM have a short K (720 .. 1728): their epilogue (c in/out, three peepholes,
// five transcendentals per channel) is as long as the MMA main loop, and with operands requested only one chunk ahead
// every chunk exposed a global-memory round trip.  Here a thread keeps the cell state of all its (up to four) chunks of
// the NEXT tile in registers: chunk k of tile i+1 is requested right after chunk k of tile i has been consumed, a
// full tile epilogue before it is needed (HBM latency).  Peepholes (L2-resident, shared by the whole batch) come as
// packed bf16, 12 registers per chunk, requested two chunks ahead when the chunk count is even, else a tile ahead.
struct LstmOps {            // cell state of one 8-channel chunk ...
  float c[8];
};
struct LstmPeep {           // ... and its three peepholes (bf16 x 8 each)
  uint4 p[3];
};
struct LstmTile {           // per thread and tile; everything else is warp-uniform and re-derived from EpiParams
  long long out_off;        // h' (dense NHWC, so also the offset into the optional fp32 copy h32)
  long long st_off;         // c
  int pos;                  // y * W + x
};
__device__ __forceinline__ LstmTile lstm_tile(const EpiParams& E, int b, int y, int x, int H, int W) {
  LstmTile t;
  t.pos = y * W + x;
  t.out_off = b * E.oB + y * E.oY + x * E.oX;
  const long long hw = static_cast<long long>(H) * W;
  t.st_off = E.state_c4 ? (static_cast<long long>(b) * (E.C >> 2) * hw + t.pos) * 4 : (b * hw + t.pos) * E.C;
  return t;
}

__device__ __forceinline__ void lstm_c_load(const EpiParams& E, const LstmTile& t, long long hw, int ch, LstmOps& o) {
  ld_state8(E.s0, t.st_off, E.state_c4 ? hw * 4 : 0, ch, o.c);
}
__device__ __forceinline__ void lstm_peep_load(const EpiParams& E, const LstmTile& t, long long hw, int ch, LstmPeep& o) {
  const uint4* p = static_cast<const uint4*>(E.pp16) + (static_cast<long long>(ch >> 3) * hw + t.pos) * 3;
  o.p[0] = __ldg(p);
  o.p[1] = __ldg(p + 1);
  o.p[2] = __ldg(p + 2);
}

__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __bfloat1622float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}

// STORE_C = false: the new cell state only stays in o.c (sequence mode keeps it in shared memory)
// `sb` = SHARED-space address of this chunk's 32 staged bias values in the HALVED form of the kernels that use this
// function (conv_halo.cu stages 0.5 * b for the three sigmoid gates), and the packed peepholes are halved at pack time:
// sigmoid(z + b + w c) = 0.5 tanh(0.5 z + 0.5 b + (0.5 w) c) + 0.5 then costs FFMA, FFMA, MUFU, FFMA instead of FADD, FFMA,
// FMUL, MUFU, FFMA -- and, scaling by 0.5 being exact, rounds exactly as the plain form does (bit-identical results).
template <bool PEEP, bool STORE_C = true, bool PTR_OUT = false>
__device__ __forceinline__ void lstm_finish(const EpiParams& E, const LstmTile& t, long long hw, int ch,
                                            uint32_t sb, float (&acc)[4][8], LstmOps& o, const LstmPeep& pp,
                                            __nv_bfloat16* hout = nullptr, float* h32out = nullptr) {
  using bf16 = __nv_bfloat16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 bv = ptx::lds_f4(sb + static_cast<uint32_t>(16 * j));     // (0.5 b_i, 0.5 b_f, b_g, 0.5 b_o)
    acc[0][j] = fmaf(acc[0][j], 0.5f, bv.x);
    acc[1][j] = fmaf(acc[1][j], 0.5f, bv.y);
    acc[2][j] += bv.z;
    acc[3][j] = fmaf(acc[3][j], 0.5f, bv.w);
  }
  float h[8];
  if constexpr (PEEP) {
    // peepholes (pre-halved) are decoded two channels at a time (one 32-bit word of each gate): 6 transient registers, not 24
    const uint32_t* pi = reinterpret_cast<const uint32_t*>(&pp.p[0]);
    const uint32_t* pf = reinterpret_cast<const uint32_t*>(&pp.p[1]);
    const uint32_t* po = reinterpret_cast<const uint32_t*>(&pp.p[2]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float wi[2] = {__uint_as_float(pi[q] << 16), __uint_as_float(pi[q] & 0xFFFF0000u)};
      const float wf[2] = {__uint_as_float(pf[q] << 16), __uint_as_float(pf[q] & 0xFFFF0000u)};
      const float wo[2] = {__uint_as_float(po[q] << 16), __uint_as_float(po[q] & 0xFFFF0000u)};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * q + e;
        const float ig = fmaf(0.5f, tanh_fast(fmaf(wi[e], o.c[j], acc[0][j])), 0.5f);
        const float fg = fmaf(0.5f, tanh_fast(fmaf(wf[e], o.c[j], acc[1][j])), 0.5f);
        const float cn = fmaf(fg, o.c[j], ig * tanh_fast(acc[2][j]));
        const float og = fmaf(0.5f, tanh_fast(fmaf(wo[e], cn, acc[3][j])), 0.5f);
        o.c[j] = cn;
        h[j] = og * tanh_fast(cn);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = fmaf(0.5f, tanh_fast(acc[0][j]), 0.5f), fg = fmaf(0.5f, tanh_fast(acc[1][j]), 0.5f);
      const float cn = fmaf(fg, o.c[j], ig * tanh_fast(acc[2][j]));
      o.c[j] = cn;
      h[j] = fmaf(0.5f, tanh_fast(acc[3][j]), 0.5f) * tanh_fast(cn);
    }
  }
  if constexpr (STORE_C) st_state8(E.s0, t.st_off, E.state_c4 ? hw * 4 : 0, ch, o.c);
  if constexpr (PTR_OUT) {      // sequence mode: the chunk's h' address was prepared before the accumulator wait
    st_bf16x8(hout, h);
    if (h32out != nullptr) st_f32x8(h32out, h);
  } else {
    st_bf16x8(static_cast<bf16*>(E.out) + t.out_off + ch, h);
    if (E.h32 != nullptr) st_f32x8(E.h32 + t.out_off + ch, h);
  }
}

}  // namespace vpk
