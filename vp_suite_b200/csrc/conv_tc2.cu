// 2-CTA (cta_group::2) variant of the tcgen05 / TMEM / TMA implicit-GEMM kernel (see conv_tc.cu for the single-CTA one).
//
// A CTA pair (cluster of 2, one TPC) computes a 256-position x tileN tile: each CTA TMA-loads its own 128-position
// activation box and HALF of the weight tile (tileN/2 rows), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256)
// which reads both CTAs' shared memory and writes both CTAs' TMEM.  Per CTA and K-step this moves 16 KB + tileN*64 B
// instead of 16 KB + tileN*128 B: the kernel is bound by operand delivery (L2 -> smem), so halving the weight traffic
// per FLOP is the point.  Barrier protocol:
//   full[s]   (leader only, 2 arrivals + tx bytes of both CTAs)   TMA of both CTAs -> leader's MMA warp
//   empty[s]  (both CTAs, 1 arrival via multicast tcgen05.commit) leader's MMA     -> both producers
//   tfull[a]  (both CTAs, multicast commit)                       leader's MMA     -> both epilogues
//   tempty[a] (leader only, 16 arrivals)                          both epilogues   -> leader's MMA
#include <cstdio>
#include <mutex>

#include "common.h"
#include "conv_tc.h"
#include "epilogue_tc.cuh"
#include "ptx.cuh"

namespace vpk {

namespace {

constexpr int kTcThreads = 320;          // warps 0,1: TMA / MMA;  warps 2..9: epilogue (two per TMEM lane quadrant)
constexpr int kEpiThreads = 256;
constexpr uint32_t kAStageBytes = 128 * 128;
constexpr unsigned kMaxSmem = 232448;

__host__ __device__ constexpr int gates_of2(int kind) {
  return (kind == EPI_LSTM || kind == EPI_ST_C) ? 4 : (kind == EPI_ST_M) ? 3 : (kind == EPI_ST_O) ? 2 : 1;
}

template <int KIND, bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    conv_tc2_kernel(const __grid_constant__ TcPlan P) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  using bf16 = __nv_bfloat16;
  constexpr int G = gates_of2(KIND);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int stages = P.stages;
  const int tileN = P.tileN;
  const int halfN = tileN / 2;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(halfN) * 128u;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + stages * b_stage_bytes);
  const uint32_t full_bar = ptx::smem_u32(bars);
  const uint32_t empty_bar = ptx::smem_u32(bars + stages);
  const uint32_t tfull_bar = ptx::smem_u32(bars + 2 * stages);
  const uint32_t tempty_bar = ptx::smem_u32(bars + 2 * stages + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);
  ConvStep* s_steps = reinterpret_cast<ConvStep*>(tmem_slot + 4);
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_steps + P.L.nsteps) + 15) & ~uintptr_t(15));
  int* s_nk = reinterpret_cast<int*>(s_bias + P.L.N_pad);     // K=16 slices per step, for the MMA thread

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nsteps = P.L.nsteps;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  for (int i = threadIdx.x; i < nsteps; i += kTcThreads) {
    const ConvStep st = P.L.steps[i];
    s_steps[i] = st;
    s_nk[i] = (st.kc + 15) >> 4;
  }
  for (int i = threadIdx.x; i < P.L.N_pad; i += kTcThreads) s_bias[i] = P.L.epi.bias ? P.L.epi.bias[i] : 0.f;

  if (warp == 0 && ptx::elect_one()) {
    for (int i = 0; i < P.L.nsrc; ++i) ptx::prefetch_tensormap(&P.amap[i]);
    ptx::prefetch_tensormap(&P.bmap);
  } else if (warp == 1 && ptx::elect_one()) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full_bar + 8 * i, 2);
      ptx::mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar + 8 * i, 1);
      ptx::mbar_init(tempty_bar + 8 * i, 2 * (kEpiThreads / 32));   // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  ptx::cluster_sync_all();          // both CTAs' barriers exist before anything may signal them
  if (warp == 2) {
    ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), static_cast<uint32_t>(P.tmem_cols));
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tiles: a pair owns M tiles (2p, 2p+1); CTA `rank` loads / finalises tile 2p + rank
  const int m_tiles = P.tiles_b * P.tiles_y * P.tiles_x;
  const int m_pairs = (m_tiles + 1) / 2;
  const int total = m_pairs * P.n_tiles;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =========================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < total; t += npairs) {
        const int nt = t % P.n_tiles;
        const int mt = (t / P.n_tiles) * 2 + static_cast<int>(rank);
        const int x0 = (mt % P.tiles_x) * P.TW;
        const int y0 = ((mt / P.tiles_x) % P.tiles_y) * P.TH;
        const int b0 = (mt / (P.tiles_x * P.tiles_y)) * P.TB;      // mt == m_tiles (odd tail) -> b0 >= B: all zero fill
        const int n0 = nt * tileN + static_cast<int>(rank) * halfN;
        for (int s = 0; s < nsteps; ++s) {
          const ConvStep st = s_steps[s];
          ptx::mbar_wait_fast(empty_bar + 8 * stage, phase ^ 1u);
          const uint32_t fb = full_bar + 8 * stage;
          if (P.debug & 4) {
            if (leader) ptx::mbar_arrive(fb);
            else ptx::mbar_arrive_cluster(fb, 0);
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(fb, 2u * (kAStageBytes + b_stage_bytes));
            else ptx::mbar_arrive_cluster(fb, 0);
            ptx::tma_load_4d_pair(&P.amap[st.src], fb, ptx::smem_u32(smem_a + stage * kAStageBytes), st.c0,
                                  x0 + st.dx, y0 + st.dy, b0);
            ptx::tma_load_2d_pair(&P.bmap, fb, ptx::smem_u32(smem_b + stage * b_stage_bytes), st.wk, n0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) =====================================
    // ONE thread runs the whole loop; descriptors advance arithmetically (16-byte units, no carry out of the field).
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(256, tileN, P.L.op_f16 != 0);
      const uint64_t adesc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem_a));
      const uint64_t bdesc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem_b));
      const uint32_t a_u = kAStageBytes >> 4, b_u = b_stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = pair; t < total; t += npairs, ++iter) {
        const int acc = iter & 1;
        ptx::mbar_wait_fast(tempty_bar + 8 * acc, ((iter >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * tileN);
        uint32_t accum = 0;
        for (int s = 0; s < nsteps; ++s) {
          const int nk = (P.debug & 2) ? 0 : s_nk[s];
          ptx::mbar_wait_fast(full_bar + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * a_u);
          const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_u);
          if (nk > 0) ptx::mma_bf16_ss_pair(tmem_d, ad, bd, idesc, accum);
          if (nk > 1) ptx::mma_bf16_ss_pair(tmem_d, ad + 2, bd + 2, idesc, 1u);
          if (nk > 2) ptx::mma_bf16_ss_pair(tmem_d, ad + 4, bd + 4, idesc, 1u);
          if (nk > 3) ptx::mma_bf16_ss_pair(tmem_d, ad + 6, bd + 6, idesc, 1u);
          accum = 1u;
          ptx::mma_commit_pair(empty_bar + 8 * stage, 3);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        ptx::mma_commit_pair(tfull_bar + 8 * acc, 3);
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue (warps 2..5, both CTAs) =================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // the two warps of a quadrant take alternate 8-channel chunks
    const int row = quad * 32 + lane;          // accumulator row = output position inside the tile
    const int rx = row % P.TW;
    const int ry = (row / P.TW) % P.TH;
    const int rb = row / (P.TW * P.TH);
    const int Cn = tileN / G;
    int iter = 0;
    for (int t = pair; t < total; t += npairs, ++iter) {
      const int nt = t % P.n_tiles;
      const int mt = (t / P.n_tiles) * 2 + static_cast<int>(rank);
      const int x = (mt % P.tiles_x) * P.TW + rx;
      const int y = ((mt / P.tiles_x) % P.tiles_y) * P.TH + ry;
      const int b = (mt / (P.tiles_x * P.tiles_y)) * P.TB + rb;
      const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      EpiOperands<8> ops0, ops1;
      const int C = P.L.epi.C;
      const int ch_base = nt * Cn;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
      auto tmem_chunk = [&](int ch, uint32_t (&r)[8 * G]) {
        const uint32_t ta = taddr + static_cast<uint32_t>(ch * G);
        if constexpr (G == 4) ptx::tmem_ld32(ta, r);
        else if constexpr (G == 2) ptx::tmem_ld16(ta, r);
        else if constexpr (G == 1) ptx::tmem_ld8(ta, r);
        else { ptx::tmem_ld8(ta, r); ptx::tmem_ld8(ta + 8, r + 8); ptx::tmem_ld8(ta + 16, r + 16); }
      };
      if constexpr (FAST) {
        EpiTile et;
        if (valid) et = epi_tile(P.L.epi, b, y, x, P.L.H, P.L.W);
        const float* bias = s_bias;
        if (valid && ch_base + half * 8 < C) epi_tc_prefetch<KIND>(P.L.epi, et, ch_base + half * 8, ops0);
        ptx::mbar_wait_fast(tfull_bar + 8 * acc, acc_phase);
        ptx::tc_fence_after();
        auto do_chunk = [&](int ch, EpiOperands<8>& cur, EpiOperands<8>& nxt) {
          uint32_t r[8 * G];
          tmem_chunk(ch, r);
          if (valid && ch + 16 < Cn && ch_base + ch + 16 < C) epi_tc_prefetch<KIND>(P.L.epi, et, ch_base + ch + 16, nxt);
          ptx::tmem_ld_wait();
          if (valid && ch_base + ch < C) {
            float a[G][8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
              for (int g = 0; g < G; ++g) a[g][j] = __uint_as_float(r[j * G + g]);
            epi_tc_finish<KIND, G>(P.L.epi, et, ch_base + ch, bias, a, cur);
          }
        };
        for (int ch = half * 8; ch < Cn; ch += 32) {
          do_chunk(ch, ops0, ops1);
          if (ch + 16 < Cn) do_chunk(ch + 16, ops1, ops0);
        }
      } else {
        if (valid && half * 8 < Cn) epilogue_prefetch<bf16, G, 8>(P.L.epi, b, y, x, P.L.H, P.L.W, ch_base + half * 8, ops0);
        ptx::mbar_wait_fast(tfull_bar + 8 * acc, acc_phase);
        ptx::tc_fence_after();
        auto do_chunk = [&](int ch, EpiOperands<8>& cur, EpiOperands<8>& nxt) {
          uint32_t r[8 * G];
          tmem_chunk(ch, r);
          if (valid && ch + 16 < Cn)
            epilogue_prefetch<bf16, G, 8>(P.L.epi, b, y, x, P.L.H, P.L.W, ch_base + ch + 16, nxt);
          ptx::tmem_ld_wait();
          if (valid) {
            float a[G][8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
              for (int g = 0; g < G; ++g) a[g][j] = __uint_as_float(r[j * G + g]);
            epilogue_finish<bf16, G, 8, true>(P.L.epi, b, y, x, P.L.H, P.L.W, ch_base + ch, a, cur);
          }
        };
        for (int ch = half * 8; ch < Cn; ch += 32) {
          do_chunk(ch, ops0, ops1);
          if (ch + 16 < Cn) do_chunk(ch + 16, ops1, ops0);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(tempty_bar + 8 * acc, 0);   // the leader's MMA warp owns the accumulator hand-off
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == 2) ptx::tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(P.tmem_cols));
#endif
}

template <int KIND, bool FAST> void launch_kind_f(const TcPlan& P, cudaStream_t stream) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(conv_tc2_kernel<KIND, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(kMaxSmem));
  });
  conv_tc2_kernel<KIND, FAST><<<P.grid, kTcThreads, P.smem_bytes, stream>>>(P);
}
template <int KIND> void launch_kind(const TcPlan& P, cudaStream_t stream) {
  if (P.fast_epi) launch_kind_f<KIND, true>(P, stream);
  else launch_kind_f<KIND, false>(P, stream);
}

}  // namespace

void launch_conv_tc2(const TcPlan& P, cudaStream_t stream) {
  switch (P.L.epi.kind) {
    case EPI_BIAS_ACT: launch_kind<EPI_BIAS_ACT>(P, stream); break;
    case EPI_LSTM: launch_kind<EPI_LSTM>(P, stream); break;
    case EPI_ST_C: launch_kind<EPI_ST_C>(P, stream); break;
    case EPI_ST_M: launch_kind<EPI_ST_M>(P, stream); break;
    case EPI_ST_O: launch_kind<EPI_ST_O>(P, stream); break;
    case EPI_ST_O1: launch_kind<EPI_ST_O1>(P, stream); break;
    case EPI_PHY_GATE: launch_kind<EPI_PHY_GATE>(P, stream); break;
    default: VPK_THROW(1, "conv_tc2: unsupported epilogue kind");
  }
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
