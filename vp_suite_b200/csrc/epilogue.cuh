// Fused epilogues shared by the CUDA-core and the tcgen05 kernels: gate non-linearities, peepholes, cell-state update
// and the stores of h / c / m, applied to accumulators while they are still on chip, so that gate pre-activations
// never reach HBM.  One call handles NCH consecutive channels (all G gates of each) of ONE output position.
#pragma once
#include "common.h"

namespace vpk {

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// sigmoid / tanh through ex2: abs error ~1e-7, far inside the 1e-4 fp32-mode bound
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
  float e = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
  return __fdividef(e - 1.f, e + 1.f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == ACT_SIGMOID) return sigmoid_f(v);
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// ---- masked / vectorised row accessors ------------------------------------------------------------------------
template <int N> __device__ __forceinline__ void load_f32(const float* p, float (&v)[N], int nvalid) {
  if (nvalid >= N && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      float4 t = reinterpret_cast<const float4*>(p)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i < nvalid) ? p[i] : 0.f;
  }
}
template <int N> __device__ __forceinline__ void store_f32(float* p, const float (&v)[N], int nvalid) {
  if (nvalid >= N && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < nvalid) p[i] = v[i];
  }
}
template <int N> __device__ __forceinline__ void load_act(const float* p, float (&v)[N], int nvalid) {
  load_f32<N>(p, v, nvalid);
}
template <int N> __device__ __forceinline__ void load_act(const __nv_bfloat16* p, float (&v)[N], int nvalid) {
  if (nvalid >= N && (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t = reinterpret_cast<const uint4*>(p)[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(h[k]);
        v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i < nvalid) ? __bfloat162float(p[i]) : 0.f;
  }
}
template <int N> __device__ __forceinline__ void store_act(float* p, const float (&v)[N], int nvalid) {
  store_f32<N>(p, v, nvalid);
}
template <int N> __device__ __forceinline__ void store_act(__nv_bfloat16* p, const float (&v)[N], int nvalid) {
  if (nvalid >= N && (N % 8 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < N / 8; ++i) {
      uint4 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  } else if (nvalid >= N && (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
      reinterpret_cast<__nv_bfloat162*>(p)[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// acc[g][j]: gate g of channel ch0 + j at output position (b, y, x) of an (H, W) grid.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int G, int NCH>
__device__ __forceinline__ void epilogue_apply(const EpiParams& E, int b, int y, int x, int H, int W, int ch0,
                                               float (&acc)[G][NCH]) {
  const int C = E.C;
  const int nvalid = min(NCH, C - ch0);
  if (nvalid <= 0) return;
  const size_t pix = (static_cast<size_t>(b) * H + y) * W + x;

  if (E.bias != nullptr) {   // packed order: (ch0 + j) * G + g
    const float* bp = E.bias + static_cast<size_t>(ch0) * G;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
      if (j < nvalid) {
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g][j] += __ldg(bp + j * G + g);
      }
  }

  if constexpr (G == 1) {
    if (E.kind == EPI_BIAS_ACT) {
      float v[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) v[j] = apply_act(acc[0][j], E.act);
      if (E.res != nullptr) {
        float r[NCH];
        load_f32<NCH>(E.res + pix * C + ch0, r, nvalid);
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] += r[j];
      }
      const long long off = b * E.oB + y * E.oY + x * E.oX;
      if (E.oC == 1) {
        if (E.out_f32) store_f32<NCH>(static_cast<float*>(E.out) + off + ch0, v, nvalid);
        else store_act<NCH>(static_cast<T*>(E.out) + off + ch0, v, nvalid);
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
          if (j < nvalid) {
            const long long o = off + static_cast<long long>(ch0 + j) * E.oC;
            if (E.out_f32) static_cast<float*>(E.out)[o] = v[j];
            else static_cast<T*>(E.out)[o] = from_f32<T>(v[j]);
          }
      }
    } else {   // EPI_PHY_GATE: h' = h~ + sigmoid(acc) * (x - h~)      (model_blocks/phydnet.py:58-61)
      float xf[NCH], ht[NCH], v[NCH];
      load_act<NCH>(static_cast<const T*>(E.q0) + pix * C + ch0, xf, nvalid);
      load_f32<NCH>(E.res + pix * C + ch0, ht, nvalid);
#pragma unroll
      for (int j = 0; j < NCH; ++j) v[j] = ht[j] + sigmoid_f(acc[0][j]) * (xf[j] - ht[j]);
      store_f32<NCH>(E.s0 + pix * C + ch0, v, nvalid);                       // fp32 master of the hidden state
      store_act<NCH>(static_cast<T*>(E.out) + pix * C + ch0, v, nvalid);     // conv-operand copy
    }
  } else if constexpr (G == 4) {
    float c[NCH];
    float* cp = E.s0 + pix * C + ch0;
    load_f32<NCH>(cp, c, nvalid);
    if (E.kind == EPI_LSTM) {
      // conv_lstm_hzzone.py:62-68 (with peepholes) and conv_lstm_ndrplz.py:34-41 (without); rows packed as i,f,g,o
      float h[NCH];
      if (E.p0 != nullptr) {
        const size_t pp = (static_cast<size_t>(y) * W + x) * C + ch0;
        float wi[NCH], wf[NCH], wo[NCH];
        load_f32<NCH>(E.p0 + pp, wi, nvalid);
        load_f32<NCH>(E.p1 + pp, wf, nvalid);
        load_f32<NCH>(E.p2 + pp, wo, nvalid);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float ig = sigmoid_f(acc[0][j] + wi[j] * c[j]);
          const float fg = sigmoid_f(acc[1][j] + wf[j] * c[j]);
          const float cn = fg * c[j] + ig * tanh_f(acc[2][j]);
          const float og = sigmoid_f(acc[3][j] + wo[j] * cn);
          c[j] = cn;
          h[j] = og * tanh_f(cn);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float cn = sigmoid_f(acc[1][j]) * c[j] + sigmoid_f(acc[0][j]) * tanh_f(acc[2][j]);
          c[j] = cn;
          h[j] = sigmoid_f(acc[3][j]) * tanh_f(cn);
        }
      }
      store_f32<NCH>(cp, c, nvalid);
      store_act<NCH>(static_cast<T*>(E.out) + b * E.oB + y * E.oY + x * E.oX + ch0, h, nvalid);
      if (E.h32 != nullptr) store_f32<NCH>(E.h32 + pix * C + ch0, h, nvalid);
    } else {   // EPI_ST_C: predrnn.py:65-70; acc = (i, f, g, o_x + o_h)
      float dc[NCH], op[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const float ig = sigmoid_f(acc[0][j]);
        const float fg = sigmoid_f(acc[1][j] + E.forget_bias);
        dc[j] = ig * tanh_f(acc[2][j]);
        c[j] = fg * c[j] + dc[j];
        op[j] = acc[3][j];
      }
      store_f32<NCH>(cp, c, nvalid);
      store_f32<NCH>(E.s1 + pix * C + ch0, op, nvalid);
      store_act<NCH>(static_cast<T*>(E.t0) + pix * E.t0_pix + ch0, c, nvalid);
      store_act<NCH>(static_cast<T*>(E.t1) + pix * C + ch0, dc, nvalid);
    }
  } else if constexpr (G == 3) {   // EPI_ST_M: predrnn.py:72-77; acc = (i', f', g')
    float m[NCH], dm[NCH];
    float* mp = E.s0 + pix * C + ch0;
    load_f32<NCH>(mp, m, nvalid);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const float ig = sigmoid_f(acc[0][j]);
      const float fg = sigmoid_f(acc[1][j] + E.forget_bias);
      dm[j] = ig * tanh_f(acc[2][j]);
      m[j] = fg * m[j] + dm[j];
    }
    store_f32<NCH>(mp, m, nvalid);
    store_act<NCH>(static_cast<T*>(E.t0) + pix * E.t0_pix + ch0, m, nvalid);
    store_act<NCH>(static_cast<T*>(E.t1) + pix * C + ch0, dm, nvalid);
  } else if constexpr (G == 2) {   // EPI_ST_O: predrnn.py:79-80; acc = (conv_o(mem), conv_last(mem))
    float op[NCH], h[NCH];
    load_f32<NCH>(E.s0 + pix * C + ch0, op, nvalid);
#pragma unroll
    for (int j = 0; j < NCH; ++j) h[j] = sigmoid_f(op[j] + acc[0][j]) * tanh_f(acc[1][j]);
    store_act<NCH>(static_cast<T*>(E.out) + b * E.oB + y * E.oY + x * E.oX + ch0, h, nvalid);
  }
}

}  // namespace vpk
