// tcgen05 / TMEM / TMA implicit-GEMM kernel for the generalised convolution launch (sm_100a only).
//
//   D[128 positions, tileN] = sum over K-steps  A_step[128, <=64 ch] * W_step[tileN, <=64 ch]^T      (bf16 -> fp32)
//
// * A_step is one TMA tiled load of a (64 ch, TW, TH, TB) box from the NHWC activation view at the tap's spatial
//   offset; out-of-image coordinates are zero-filled by TMA, which is the conv padding (halo load, no im2col buffer,
//   no torch.cat: x, h, m are separate tensor maps feeding consecutive K-steps).  The box lands in shared memory as
//   128 rows x 128 B with the 128-byte swizzle = the canonical K-major UMMA operand layout.
// * W_step is a (64, tileN) box of the packed weights, same layout.
// * One elected thread issues tcgen05.mma (M=128, N=tileN, K=16) ceil(kc/16) times per step into a TMEM
//   accumulator; two accumulators (2 * tileN <= 512 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// * 4 epilogue warps read the accumulator with tcgen05.ld (thread = output position, registers = gate columns) and
//   run the fused gate / state-update epilogue (epilogue.cuh): pre-activations never leave the SM.
// * Persistent grid (<= #SMs CTAs), static round-robin over (m-tile, n-tile) with the n-tiles of one m-tile adjacent.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (warp 2 also owns the TMEM allocation).
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.h"
#include "conv_tc.h"
#include "epilogue_tc.cuh"
#include "ptx.cuh"

namespace vpk {

namespace {

constexpr int kTcThreads = 320;          // warps 0,1: TMA / MMA;  warps 2..9: epilogue (two per TMEM lane quadrant)
constexpr int kEpiThreads = 256;
constexpr uint32_t kAStageBytes = 128 * 128;   // 128 rows x 64 bf16
constexpr unsigned kMaxSmem = 232448;          // 227 KB

template <int G>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcPlan P) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  using bf16 = __nv_bfloat16;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // 1024-B alignment for the 128B swizzle

  const int stages = P.stages;
  const int tileN = P.tileN;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(tileN) * 128u;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * kAStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + stages * b_stage_bytes);
  const uint32_t full_bar = ptx::smem_u32(bars);                    // [stages]  TMA -> MMA
  const uint32_t empty_bar = ptx::smem_u32(bars + stages);          // [stages]  MMA -> TMA
  const uint32_t tfull_bar = ptx::smem_u32(bars + 2 * stages);      // [2]       MMA -> epilogue
  const uint32_t tempty_bar = ptx::smem_u32(bars + 2 * stages + 2); // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);
  ConvStep* s_steps = reinterpret_cast<ConvStep*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nsteps = P.L.nsteps;

  for (int i = threadIdx.x; i < nsteps; i += kTcThreads) s_steps[i] = P.L.steps[i];

  if (warp == 0 && ptx::elect_one()) {
    for (int i = 0; i < P.L.nsrc; ++i) ptx::prefetch_tensormap(&P.amap[i]);
    ptx::prefetch_tensormap(&P.bmap);
  } else if (warp == 1 && ptx::elect_one()) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full_bar + 8 * i, 1);
      ptx::mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull_bar + 8 * i, 1);
      ptx::mbar_init(tempty_bar + 8 * i, kEpiThreads);
    }
    ptx::fence_barrier_init();
  } else if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), static_cast<uint32_t>(P.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = P.tiles_b * P.tiles_y * P.tiles_x;
  const int total = m_tiles * P.n_tiles;

  if (warp == 0) {
    // ===================================== TMA producer =====================================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int nt = t % P.n_tiles;
        const int mt = t / P.n_tiles;
        const int x0 = (mt % P.tiles_x) * P.TW;
        const int y0 = ((mt / P.tiles_x) % P.tiles_y) * P.TH;
        const int b0 = (mt / (P.tiles_x * P.tiles_y)) * P.TB;
        const int n0 = nt * tileN;
        for (int s = 0; s < nsteps; ++s) {
          const ConvStep st = s_steps[s];
          ptx::mbar_wait(empty_bar + 8 * stage, phase ^ 1u);
          const uint32_t fb = full_bar + 8 * stage;
          if (P.debug & 4) {
            ptx::mbar_arrive(fb);
          } else {
            ptx::mbar_arrive_expect_tx(fb, kAStageBytes + b_stage_bytes);
            ptx::tma_load_4d(&P.amap[st.src], fb, ptx::smem_u32(smem_a + stage * kAStageBytes), st.c0, x0 + st.dx,
                             y0 + st.dy, b0);
            ptx::tma_load_2d(&P.bmap, fb, ptx::smem_u32(smem_b + stage * b_stage_bytes), st.wk, n0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================================
    const uint32_t idesc = ptx::idesc_bf16_f32(128, tileN, P.L.op_f16 != 0);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      ptx::mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * tileN);
      for (int s = 0; s < nsteps; ++s) {
        const int nk = (s_steps[s].kc + 15) >> 4;
        ptx::mbar_wait(full_bar + 8 * stage, phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t adesc = ptx::smem_desc_sw128(ptx::smem_u32(smem_a + stage * kAStageBytes));
          const uint64_t bdesc = ptx::smem_desc_sw128(ptx::smem_u32(smem_b + stage * b_stage_bytes));
          for (int k = 0; k < ((P.debug & 2) ? 0 : nk); ++k)   // +32 B per K=16 slice inside the 128-B swizzle row
            ptx::mma_bf16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (s > 0 || k > 0) ? 1u : 0u);
          ptx::mma_commit(empty_bar + 8 * stage);              // frees the smem slot when these MMAs retire
          if (s == nsteps - 1) ptx::mma_commit(tfull_bar + 8 * acc);   // accumulator complete
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue (warps 2..5) ============================================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // the two warps of a quadrant take alternate 8-channel chunks
    const int row = quad * 32 + lane;          // accumulator row = output position inside the tile
    const int rx = row % P.TW;
    const int ry = (row / P.TW) % P.TH;
    const int rb = row / (P.TW * P.TH);
    const int Cn = tileN / G;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const int nt = t % P.n_tiles;
      const int mt = t / P.n_tiles;
      const int x = (mt % P.tiles_x) * P.TW + rx;
      const int y = ((mt / P.tiles_x) % P.tiles_y) * P.TH + ry;
      const int b = (mt / (P.tiles_x * P.tiles_y)) * P.TB + rb;
      const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      EpiOperands<8> ops0, ops1;
      // operands of this warp's first chunk are requested before the accumulator is even complete
      if (valid && half * 8 < Cn) epilogue_prefetch<bf16, G, 8>(P.L.epi, b, y, x, P.L.H, P.L.W, nt * Cn + half * 8, ops0);
      ptx::mbar_wait(tfull_bar + 8 * acc, acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
      auto do_chunk = [&](int ch, EpiOperands<8>& cur, EpiOperands<8>& nxt) {
        uint32_t r[8 * G];
        const uint32_t ta = taddr + static_cast<uint32_t>(ch * G);
        if constexpr (G == 4) ptx::tmem_ld32(ta, r);
        else if constexpr (G == 2) ptx::tmem_ld16(ta, r);
        else if constexpr (G == 1) ptx::tmem_ld8(ta, r);
        else { ptx::tmem_ld8(ta, r); ptx::tmem_ld8(ta + 8, r + 8); ptx::tmem_ld8(ta + 16, r + 16); }
        if (valid && ch + 16 < Cn)   // next chunk's global operands fly while this chunk is computed
          epilogue_prefetch<bf16, G, 8>(P.L.epi, b, y, x, P.L.H, P.L.W, nt * Cn + ch + 16, nxt);
        ptx::tmem_ld_wait();
        if (valid) {
          float a[G][8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int g = 0; g < G; ++g) a[g][j] = __uint_as_float(r[j * G + g]);
          epilogue_finish<bf16, G, 8, true>(P.L.epi, b, y, x, P.L.H, P.L.W, nt * Cn + ch, a, cur);
        }
      };
      for (int ch = half * 8; ch < Cn; ch += 32) {
        do_chunk(ch, ops0, ops1);
        if (ch + 16 < Cn) do_chunk(ch + 16, ops1, ops0);
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar + 8 * acc);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(P.tmem_cols));
#endif
}

// ---- driver entry point for tensor-map encoding (no link-time dependency on libcuda) ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (fn == nullptr) VPK_THROW(2, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

void encode(CUtensorMap* map, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                               const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, static_cast<int>(r));
    VPK_THROW(2, buf);
  }
}

int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int G> void set_smem_attr() {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(conv_tc_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxSmem));
  });
}

}  // namespace

bool tc_eligible(const ConvLaunch& L, int dtype) {
  if (dtype != DT_BF16 && dtype != DT_F16) return false;
  // fp16 operands: the tensor-core epilogues store activation-type outputs as bf16, so fp16 launches are fp32-out only
  // (or fp16-out through the halo kernel's lean epilogue: out_f16, checked again by halo_make_plan)
  if (dtype == DT_F16 && !((L.epi.kind == EPI_BIAS_ACT && (L.epi.out_f32 || L.epi.out_f16)) || L.epi.kind == EPI_DECOUPLE)) return false;
  if (L.G < 1 || L.G > 4) return false;
  const int tileN = L.Cn * L.G;
  if (L.Cn % 8 != 0 || tileN % 16 != 0 || tileN > 256 || L.N_pad % tileN != 0) return false;
  if (L.K_pad % 8 != 0) return false;
  for (int i = 0; i < L.nsrc; ++i) {
    const SrcView& s = L.src[i];
    if (s.C % 8 != 0) return false;                                   // 16-byte global strides / base
    if ((s.sX % 8) || (s.sY % 8) || (s.sB % 8)) return false;
    if (reinterpret_cast<uintptr_t>(s.base) % 16 != 0) return false;
  }
  return true;
}

void tc_make_plan(const ConvLaunch& L, TcPlan* plan, int num_sms) {
  TcPlan& P = *plan;
  P.L = L;
  P.tileN = L.Cn * L.G;
  P.n_tiles = L.N_pad / P.tileN;
  P.TW = std::min(pow2_at_least(L.W), 16);
  P.TH = std::min(pow2_at_least(L.H), 128 / P.TW);
  P.TB = 128 / (P.TW * P.TH);
  P.tiles_x = (L.W + P.TW - 1) / P.TW;
  P.tiles_y = (L.H + P.TH - 1) / P.TH;
  P.tiles_b = (L.B + P.TB - 1) / P.TB;
  P.tmem_cols = std::max(32, pow2_at_least(2 * P.tileN));
  VPK_REQUIRE(P.tmem_cols <= 512, "tcgen05 plan: accumulators exceed TMEM");
  const long long m_tiles_total = static_cast<long long>(P.tiles_x) * P.tiles_y * P.tiles_b;
  // CTA pairs pay off once every SM pair has at least one 256-position tile; VPK_TC_PAIR=0/1 overrides (testing)
  P.cta2 = (P.tileN % 16 == 0 && m_tiles_total * P.n_tiles >= num_sms) ? 1 : 0;
  P.debug = 0;
  if (const char* env = dev_env("VPK_TC_DEBUG")) P.debug = atoi(env);
  P.L.epi.debug = P.debug;
  const int gk = (L.epi.kind == EPI_LSTM || L.epi.kind == EPI_ST_C) ? 4 : (L.epi.kind == EPI_ST_M) ? 3 : (L.epi.kind == EPI_ST_O) ? 2 : 1;
  P.fast_epi = (epi_tc_fast_ok(L.epi) && gk == L.G) ? 1 : 0;
  if (const char* env = getenv("VPK_TC_PAIR")) P.cta2 = (atoi(env) != 0 && P.tileN % 16 == 0) ? 1 : 0;
  const unsigned b_rows = static_cast<unsigned>(P.cta2 ? P.tileN / 2 : P.tileN);
  const unsigned stage_bytes = kAStageBytes + b_rows * 128u;
  const unsigned fixed = 1024 /*alignment slack*/ + 512 /*barriers, tmem slot*/ +
                         static_cast<unsigned>(L.nsteps) * (sizeof(ConvStep) + 4) + static_cast<unsigned>(L.N_pad) * 4 + 160;
  int stages = static_cast<int>((kMaxSmem - fixed) / stage_bytes);
  stages = std::max(2, std::min(stages, 8));
  stages = std::min(stages, std::max(2, L.nsteps));
  P.stages = stages;
  P.smem_bytes = fixed + static_cast<unsigned>(stages) * stage_bytes;
  VPK_REQUIRE(P.smem_bytes <= kMaxSmem, "tcgen05 plan: shared memory budget exceeded");
  if (P.cta2) {
    const long long pairs = (m_tiles_total + 1) / 2 * P.n_tiles;
    P.grid = 2 * static_cast<int>(std::min<long long>(pairs, num_sms / 2));
  } else {
    const long long total = m_tiles_total * P.n_tiles;
    P.grid = static_cast<int>(std::min<long long>(total, num_sms));
  }

  for (int i = 0; i < L.nsrc; ++i) {
    const SrcView& s = L.src[i];
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(s.C), static_cast<cuuint64_t>(s.W), static_cast<cuuint64_t>(s.H),
                          static_cast<cuuint64_t>(L.B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(s.sX) * 2, static_cast<cuuint64_t>(s.sY) * 2,
                             static_cast<cuuint64_t>(s.sB) * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(P.TW), static_cast<cuuint32_t>(P.TH),
                         static_cast<cuuint32_t>(P.TB)};
    encode(&P.amap[i], 4, s.base, dims, strides, box, "activation view");
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(L.K_pad), static_cast<cuuint64_t>(L.N_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(L.K_pad) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(b_rows)};
    encode(&P.bmap, 2, L.wpacked, dims, strides, box, "packed weights");
  }
}

void launch_conv_tc(const TcPlan& P, cudaStream_t stream) {
  if (P.cta2) {
    launch_conv_tc2(P, stream);
    return;
  }
  switch (P.L.G) {
    case 1: set_smem_attr<1>(); conv_tc_kernel<1><<<P.grid, kTcThreads, P.smem_bytes, stream>>>(P); break;
    case 2: set_smem_attr<2>(); conv_tc_kernel<2><<<P.grid, kTcThreads, P.smem_bytes, stream>>>(P); break;
    case 3: set_smem_attr<3>(); conv_tc_kernel<3><<<P.grid, kTcThreads, P.smem_bytes, stream>>>(P); break;
    case 4: set_smem_attr<4>(); conv_tc_kernel<4><<<P.grid, kTcThreads, P.smem_bytes, stream>>>(P); break;
    default: VPK_THROW(1, "conv_tc: unsupported gate count");
  }
  VPK_CUDA(cudaGetLastError());
}

}  // namespace vpk
