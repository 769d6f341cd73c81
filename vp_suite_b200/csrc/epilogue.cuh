// Fused epilogues shared by the CUDA-core and the tcgen05 kernels: gate non-linearities, peepholes, cell-state update
// and the stores of h / c / m, applied to accumulators while they are still on chip, so that gate pre-activations
// never reach HBM.  One call handles NCH consecutive channels (all G gates of each) of ONE output position.
#pragma once
#include "common.h"

namespace vpk {

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// sigmoid / tanh through ex2: abs error ~1e-7, far inside the 1e-4 fp32-mode bound
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
  float e = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
  return __fdividef(e - 1.f, e + 1.f);
}
// Fast variants for the bf16 tensor-core path: one MUFU op each (tanh.approx.f32, max rel. error 2^-11 -- an order of
// magnitude below the bf16 operand rounding of that path; measured effect on a 10+10 rollout: ~2e-4).
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
template <bool FAST> __device__ __forceinline__ float sigmoid_t(float x) { return FAST ? sigmoid_fast(x) : sigmoid_f(x); }
template <bool FAST> __device__ __forceinline__ float tanh_t(float x) { return FAST ? tanh_fast(x) : tanh_f(x); }

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
template <int N> __device__ __forceinline__ void load_act(const __half* p, float (&v)[N], int nvalid) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = (i < nvalid) ? __half2float(p[i]) : 0.f;
}
template <int N> __device__ __forceinline__ void store_act(__half* p, const float (&v)[N], int nvalid) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < nvalid) p[i] = __float2half_rn(v[i]);
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

// ---- epilogue-private fp32 state tensors: NHWC or channel-quad layout [B][C/4][H][W][4] -----------------------------
struct StateAddr {
  long long off;      // element offset of channel ch0 at this position
  long long gstride;  // distance between consecutive channel quads (0 = NHWC: channels contiguous)
};
__device__ __forceinline__ StateAddr state_addr(const EpiParams& E, int b, int y, int x, int H, int W, int ch0,
                                                bool with_batch) {
  StateAddr a;
  if (E.state_c4) {
    const long long hw = static_cast<long long>(H) * W;
    const long long q = (with_batch ? static_cast<long long>(b) * (E.C >> 2) : 0) + (ch0 >> 2);
    a.off = (q * hw + static_cast<long long>(y) * W + x) * 4 + (ch0 & 3);
    a.gstride = hw * 4;
  } else {
    a.off = ((with_batch ? static_cast<long long>(b) * H : 0) + y) * W * static_cast<long long>(E.C) +
            static_cast<long long>(x) * E.C + ch0;
    a.gstride = 0;
  }
  return a;
}
template <int N> __device__ __forceinline__ void load_state(const float* base, const StateAddr& a, float (&v)[N],
                                                            int nvalid) {
  if (a.gstride == 0) {
    load_f32<N>(base + a.off, v, nvalid);
  } else if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      if (4 * i < nvalid) {
        const float4 t = *reinterpret_cast<const float4*>(base + a.off + i * a.gstride);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      } else {
        v[4 * i] = v[4 * i + 1] = v[4 * i + 2] = v[4 * i + 3] = 0.f;
      }
    }
  } else {   // short runs (N = 2) stay inside one quad
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i < nvalid) ? base[a.off + i] : 0.f;
  }
}
template <int N> __device__ __forceinline__ void store_state(float* base, const StateAddr& a, const float (&v)[N],
                                                             int nvalid) {
  if (a.gstride == 0) {
    store_f32<N>(base + a.off, v, nvalid);
  } else if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i)
      if (4 * i < nvalid)
        *reinterpret_cast<float4*>(base + a.off + i * a.gstride) =
            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (i < nvalid) base[a.off + i] = v[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The epilogue of one (position, NCH-channel chunk) is split in two so that the tcgen05 kernel can issue the global
// loads of chunk k+1 before it does the math of chunk k (the loads' L2 latency is then hidden behind the MUFU work):
//   epilogue_prefetch : loads the fp32 / activation operands the kind needs (state, peepholes, residual, ...)
//   epilogue_finish   : bias, gate math, state update, stores
// acc[g][j]: gate g of channel ch0 + j at output position (b, y, x) of an (H, W) grid.
// ---------------------------------------------------------------------------------------------------------------
template <int NCH> struct EpiOperands {
  float a[NCH], b[NCH], c[NCH], d[NCH];
};

template <typename T, int G, int NCH>
__device__ __forceinline__ void epilogue_prefetch(const EpiParams& E, int b, int y, int x, int H, int W, int ch0,
                                                  EpiOperands<NCH>& o) {
  const int C = E.C;
  const int nvalid = min(NCH, C - ch0);
  if (nvalid <= 0) return;
  const size_t pix = (static_cast<size_t>(b) * H + y) * W + x;
  if constexpr (G == 1) {
    if (E.kind == EPI_PHY_GATE) {
      load_act<NCH>(static_cast<const T*>(E.q0) + pix * C + ch0, o.a, nvalid);
      load_f32<NCH>(E.res + pix * C + ch0, o.b, nvalid);
    } else if (E.kind == EPI_ST_O1) {
      load_state<NCH>(E.s0, state_addr(E, b, y, x, H, W, ch0, true), o.a, nvalid);
      load_f32<NCH>(E.res + pix * C + ch0, o.b, nvalid);
    } else if (E.res != nullptr) {
      load_f32<NCH>(E.res + pix * C + ch0, o.a, nvalid);
    }
  } else {
    load_state<NCH>(E.s0, state_addr(E, b, y, x, H, W, ch0, true), o.a, nvalid);       // c / m / o_part
    if (G == 4 && E.kind == EPI_LSTM && E.p0 != nullptr) {
      const StateAddr pa = state_addr(E, b, y, x, H, W, ch0, false);
      load_state<NCH>(E.p0, pa, o.b, nvalid);
      load_state<NCH>(E.p1, pa, o.c, nvalid);
      load_state<NCH>(E.p2, pa, o.d, nvalid);
    }
  }
}

template <typename T, int G, int NCH, bool FAST = false>
__device__ __forceinline__ void epilogue_finish(const EpiParams& E, int b, int y, int x, int H, int W, int ch0,
                                                float (&acc)[G][NCH], EpiOperands<NCH>& o) {
  const int C = E.C;
  const int nvalid = min(NCH, C - ch0);
  if (nvalid <= 0) return;
  const size_t pix = (static_cast<size_t>(b) * H + y) * W + x;

  if (E.bias != nullptr && !(E.debug & 16)) {   // packed order: (ch0 + j) * G + g
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
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] += o.a[j];
      }
      const long long off = b * E.oB + y * E.oY + x * E.oX;
      if (E.debug & 8) {
        if (v[0] == 123.456f) static_cast<float*>(E.out)[0] = v[1];   // keep the math alive
      } else if (E.oC == 1) {
        if (E.out_f32) store_f32<NCH>(static_cast<float*>(E.out) + off + ch0, v, nvalid);
        else store_act<NCH>(static_cast<T*>(E.out) + off + ch0, v, nvalid);
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
          if (j < nvalid) {
            const long long oo = off + static_cast<long long>(ch0 + j) * E.oC;
            if (E.out_f32) static_cast<float*>(E.out)[oo] = v[j];
            else static_cast<T*>(E.out)[oo] = from_f32<T>(v[j]);
          }
      }
    } else if (E.kind == EPI_ST_O1) {   // (o.a = o_part)  variant & 2: acc = conv_o(mem), o.b = conv_last(mem); else swapped
      float h[NCH];
      const bool swapped = (E.variant & 2) != 0;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const float gate = o.a[j] + (swapped ? acc[0][j] : o.b[j]);
        const float last = swapped ? o.b[j] : acc[0][j];
        h[j] = ((E.variant & 1) ? tanh_t<FAST>(gate) : sigmoid_t<FAST>(gate)) * tanh_t<FAST>(last);
      }
      store_act<NCH>(static_cast<T*>(E.out) + b * E.oB + y * E.oY + x * E.oX + ch0, h, nvalid);
    } else {   // EPI_PHY_GATE: h' = h~ + sigmoid(acc) * (x - h~)      (model_blocks/phydnet.py:58-61)
      float v[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) v[j] = o.b[j] + sigmoid_t<FAST>(acc[0][j]) * (o.a[j] - o.b[j]);
      store_f32<NCH>(E.s0 + pix * C + ch0, v, nvalid);                       // fp32 master of the hidden state
      store_act<NCH>(static_cast<T*>(E.out) + pix * C + ch0, v, nvalid);     // conv-operand copy
    }
  } else if constexpr (G == 4) {
    const StateAddr sa = state_addr(E, b, y, x, H, W, ch0, true);
    if (E.kind == EPI_LSTM) {
      // conv_lstm_hzzone.py:62-68 (with peepholes) and conv_lstm_ndrplz.py:34-41 (without); rows packed as i,f,g,o
      float h[NCH];
      if (E.p0 != nullptr) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float ig = sigmoid_t<FAST>(acc[0][j] + o.b[j] * o.a[j]);
          const float fg = sigmoid_t<FAST>(acc[1][j] + o.c[j] * o.a[j]);
          const float cn = fg * o.a[j] + ig * tanh_t<FAST>(acc[2][j]);
          const float og = sigmoid_t<FAST>(acc[3][j] + o.d[j] * cn);
          o.a[j] = cn;
          h[j] = og * tanh_t<FAST>(cn);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float cn = sigmoid_t<FAST>(acc[1][j]) * o.a[j] + sigmoid_t<FAST>(acc[0][j]) * tanh_t<FAST>(acc[2][j]);
          o.a[j] = cn;
          h[j] = sigmoid_t<FAST>(acc[3][j]) * tanh_t<FAST>(cn);
        }
      }
      store_state<NCH>(E.s0, sa, o.a, nvalid);
      store_act<NCH>(static_cast<T*>(E.out) + b * E.oB + y * E.oY + x * E.oX + ch0, h, nvalid);
      if (E.h32 != nullptr) store_f32<NCH>(E.h32 + pix * C + ch0, h, nvalid);
    } else if (E.variant == 1) {   // Causal LSTM spatial memory (causal.h): acc = (i', f', g', m_m)
      float mn[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        mn[j] = sigmoid_t<FAST>(acc[1][j] + E.forget_bias) * tanh_t<FAST>(acc[3][j]) +
                sigmoid_t<FAST>(acc[0][j]) * tanh_t<FAST>(acc[2][j]);
      store_state<NCH>(E.s0, sa, mn, nvalid);
      store_act<NCH>(static_cast<T*>(E.t0) + pix * E.t0_pix + ch0, mn, nvalid);
    } else {   // EPI_ST_C: predrnn.py:65-70; acc = (i, f, g, o_x + o_h)
      float dc[NCH], op[NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const float ig = sigmoid_t<FAST>(acc[0][j]);
        const float fg = sigmoid_t<FAST>(acc[1][j] + E.forget_bias);
        dc[j] = ig * tanh_t<FAST>(acc[2][j]);
        o.a[j] = fg * o.a[j] + dc[j];
        op[j] = acc[3][j];
      }
      store_state<NCH>(E.s0, sa, o.a, nvalid);
      store_state<NCH>(E.s1, sa, op, nvalid);
      store_act<NCH>(static_cast<T*>(E.t0) + pix * E.t0_pix + ch0, o.a, nvalid);
      if (E.t1 != nullptr) store_act<NCH>(static_cast<T*>(E.t1) + pix * C + ch0, dc, nvalid);
    }
  } else if constexpr (G == 3) {   // EPI_ST_M: predrnn.py:72-77; acc = (i', f', g')
    float dm[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const float ig = sigmoid_t<FAST>(acc[0][j]);
      const float fg = sigmoid_t<FAST>(acc[1][j] + E.forget_bias);
      dm[j] = ig * tanh_t<FAST>(acc[2][j]);
      o.a[j] = fg * o.a[j] + dm[j];
    }
    store_state<NCH>(E.s0, state_addr(E, b, y, x, H, W, ch0, true), o.a, nvalid);
    store_act<NCH>(static_cast<T*>(E.t0) + pix * E.t0_pix + ch0, o.a, nvalid);
    if (E.t1 != nullptr) store_act<NCH>(static_cast<T*>(E.t1) + pix * C + ch0, dm, nvalid);
  } else if constexpr (G == 2) {   // EPI_ST_O: predrnn.py:79-80; acc = (conv_o(mem), conv_last(mem))
    float h[NCH];
    if (E.variant == 2) {   // gradient highway unit (causal.h): acc = (p, u), o.a = z
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const float u = sigmoid_t<FAST>(acc[1][j]);
        h[j] = u * o.a[j] + (1.f - u) * tanh_t<FAST>(acc[0][j]);
      }
      store_state<NCH>(E.s0, state_addr(E, b, y, x, H, W, ch0, true), h, nvalid);
    } else if (E.variant == 1) {   // Causal LSTM output: tanh output gate
#pragma unroll
      for (int j = 0; j < NCH; ++j) h[j] = tanh_t<FAST>(o.a[j] + acc[0][j]) * tanh_t<FAST>(acc[1][j]);
    } else {
#pragma unroll
      for (int j = 0; j < NCH; ++j) h[j] = sigmoid_t<FAST>(o.a[j] + acc[0][j]) * tanh_t<FAST>(acc[1][j]);
    }
    store_act<NCH>(static_cast<T*>(E.out) + b * E.oB + y * E.oY + x * E.oX + ch0, h, nvalid);
  }
}

template <typename T, int G, int NCH>
__device__ __forceinline__ void epilogue_apply(const EpiParams& E, int b, int y, int x, int H, int W, int ch0,
                                               float (&acc)[G][NCH]) {
  EpiOperands<NCH> o;
  epilogue_prefetch<T, G, NCH>(E, b, y, x, H, W, ch0, o);
  epilogue_finish<T, G, NCH>(E, b, y, x, H, W, ch0, acc, o);
}

}  // namespace vpk
