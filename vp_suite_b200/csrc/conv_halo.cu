// Halo-reuse tcgen05 / TMEM / TMA implicit-GEMM kernel (sm_100a): the main tensor-core kernel of libvpk.
//
// Difference to conv_tc.cu / conv_tc2.cu (one activation load per tap): here the activation tile of one (source,
// 64-channel block) is loaded ONCE, with its k x k halo, as a (64 ch, 8+2P, 16+2P) TMA box, and every tap reads it
// through a SHIFTED UMMA shared-memory descriptor: start address = slot + ((P+dy)*(8+2P) + (P+dx)) * 128 B, stride
// between 8-row groups (SBO) = (8+2P) * 128 B.  This relies on the hardware applying the 128-byte swizzle from the
// absolute shared-memory address bits (so that a view starting at any 128-byte row of what TMA wrote stays
// consistent); tools/umma_probe.cu verifies exactly that on the B200 (start rows 1..26, SBO 1280/1536/2048, exact).
// Activation traffic L2 -> smem drops from k*k * 16 KB to one ~23 KB (3x3) / 30 KB (5x5) / 39 KB (7x7) box per block;
// only the weight tiles still stream per tap.
//
// Output tile: 8 (x) x 16 (y) positions of one image = 128 accumulator rows, row r = ty*8 + tx, so each 8-row core
// group is one image row segment.  PAIR = true: cta_group::2, M = 256 across a cluster of two CTAs (each CTA its own
// spatial tile and half of the weight rows), as in conv_tc2.cu.
//
// Warps: 0 activation-TMA producer, 1 weight-TMA producer, 2 MMA issuer (one elected thread; the warp also owns the
// TMEM allocation), 3..10 epilogue (two warps per TMEM lane quadrant = warp % 4, operands prefetched one 8-channel
// chunk ahead).  352 threads leave 184 registers per thread: the LSTM epilogue holds two operand chunks in flight.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.h"
#include "conv_tc.h"
#include "epilogue_tc.cuh"
#include "ptx.cuh"

namespace vpk {

namespace {

constexpr int kHaloThreads = 352;   // warps 0..2: activation TMA, weight TMA, MMA issue (+ TMEM alloc); warps 3..10: epilogue
constexpr int kEpiThreads = 256;
constexpr int kTW = 8, kTH = 16;
#ifdef VPK_TRACE
constexpr unsigned kMaxSmem = 232448 - 1024;   // room for the static trace slots
#else
constexpr unsigned kMaxSmem = 232448;
#endif
constexpr uint32_t kTapFirstOfBlock = 1, kTapLastOfBlock = 2, kTapFirstOfGroup = 4, kTapLastOfGroup = 8;
constexpr uint32_t kTapRegionA = 32, kTapRegionB = 64;   // accumulator regions the tap feeds (ConvLaunch::region_g0)
constexpr uint32_t kTapFuseNext = 16;   // the next tap needs no hand-off in between and both are full (4 K-slices): one asm block
// taps of this block share streamed weight tiles four at a time (see the tap-table build in the kernel); VPK_HALO_PACK16=0 off
__device__ __forceinline__ bool halo_block_packed(const HaloPlan& P, const HaloBlock& blk) {
  return P.pack16 != 0 && !P.resident && blk.kc <= 16 && blk.ntaps > 1;
}

__device__ __forceinline__ uint64_t smem_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;     // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;     // SWIZZLE_128B; base_offset stays 0: the swizzle follows absolute addresses
  return d;
}

__host__ __device__ constexpr int gates_of(int kind) {
  return (kind == EPI_LSTM || kind == EPI_ST_C || kind == EPI_SUBPIX) ? 4
         : (kind == EPI_ST_M) ? 3 : (kind == EPI_ST_O || kind == EPI_DECOUPLE) ? 2 : 1;
}

// MODE: 0 generic epilogue, 1 lean compile-time epilogue, 2 ConvLSTM epilogue with a whole tile of operands in flight,
//       3 bias + activation + fused 1x1 projection to <= 4 channels (fp32 strided output),
//       4 SEQUENCE mode (EPI_LSTM only): one launch runs P.seq_T timesteps of the layer.  Every role wraps its tile loop
//         in a timestep loop (the same (M unit, N tile) list per CTA every step), the cell state c of the CTA's tiles
//         lives in shared memory from the first to the last step, h'_t goes to slot t of the output sequence and is
//         read back as step t+1's recurrent input through TMA after a grid-wide barrier (release / acquire on a global
//         counter + fence.proxy.async: the stores are generic-proxy, the loads async-proxy).  All CTAs must be
//         co-resident (grid <= what cudaOccupancyMaxActiveClusters reports; one CTA per SM).
template <int KIND, bool PAIR, int MODE>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ HaloPlan P) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  using bf16 = __nv_bfloat16;
  constexpr int G = gates_of(KIND);
  constexpr bool FAST = MODE == 1 || MODE == 2 || MODE == 4;
  constexpr bool SEQ = MODE == 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  float* s_cstate = reinterpret_cast<float*>(smem);          // SEQ: cell state of this CTA's tiles, in front of the rings
  if constexpr (SEQ) smem += P.seq_c_bytes;
  const uint32_t s_stage = ptx::smem_u32(smem);              // lean_tma: two staged output tiles
  smem += P.stage_bytes;
  const int T_steps = SEQ ? P.seq_T : 1;
  // PDL: the next kernel of the stream may become resident as soon as every CTA of this grid is (it then waits in its
  // own pdl_wait); everything up to our pdl_wait below reads only plan tables, biases and packed weights' descriptors
  ptx::pdl_launch_dependents();
  // Built with -DVPK_TRACE and run with VPK_TC_DEBUG & 128: CTA 0 prints a timeline of its phases (globaltimer, ns) --
  // developer aid for the fixed cost of short launches; compiled out of the product build
#ifdef VPK_TRACE
  __shared__ unsigned long long s_trace[8];
  const bool trace = (P.debug & 128) && blockIdx.x == 0;
  auto stamp = [&](int i) {
    if (trace) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      s_trace[i] = t;
    }
  };
#else
  auto stamp = [](int) {};
#endif
  if (threadIdx.x == 0) stamp(0);

  const int SA = P.SA, SB = P.SB;
  const int tileN = P.tileN;
  const int rowsB = PAIR ? tileN / 2 : tileN;
  // kernels that may run with accumulator regions (ConvLaunch::region_g0): CTA pairs, general fast epilogue, ST-LSTM kinds
  constexpr bool kRegionKind = PAIR && MODE == 1 && (KIND == EPI_ST_C || KIND == EPI_ST_O);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SA * P.a_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + SB * P.b_slot_bytes);
  const uint32_t afull = ptx::smem_u32(bars);
  const uint32_t aempty = ptx::smem_u32(bars + SA);
  const uint32_t bfull = ptx::smem_u32(bars + 2 * SA);
  const uint32_t bempty = ptx::smem_u32(bars + 2 * SA + SB);
  const uint32_t tfull = ptx::smem_u32(bars + 2 * SA + 2 * SB);
  const uint32_t tempty = ptx::smem_u32(bars + 2 * SA + 2 * SB + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);
  HaloBlock* s_blocks = reinterpret_cast<HaloBlock*>(tmem_slot + 4);
  HaloTap* s_taps = reinterpret_cast<HaloTap*>(s_blocks + P.nblocks);
  // per tap, for the MMA thread: x = offset of the shifted activation view in descriptor units (16 B), y = K=16 slices
  uint2* s_tapmma = reinterpret_cast<uint2*>((reinterpret_cast<uintptr_t>(s_taps + P.ntaps) + 15) & ~uintptr_t(15));
  float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_tapmma + P.ntaps + 2) + 15) & ~uintptr_t(15));

  float* s_proj = s_bias + P.L.N_pad;      // MODE 3: [4][N_pad] projection weights, then 4 projection biases
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // PAIR: clusters of 2 (one CTA pair) or, P.mc, of 4 = two pairs (ranks 0,1 and 2,3) that walk the SAME sequence of
  // (N tile, tap) steps on different M tiles and share every weight tile through TMA multicast
  const bool MC = PAIR && P.mc != 0;
  const uint32_t crank = PAIR ? ptx::cluster_ctarank() : 0u;     // rank in the cluster
  const uint32_t rank = crank & 1u;                               // rank in the CTA pair
  const uint32_t prank = crank >> 1;                              // pair in the cluster (0 unless MC)
  const uint32_t lead = crank & ~1u;                              // cluster rank of this pair's leader CTA
  const bool leader = rank == 0;
  const int csize = MC ? 4 : (PAIR ? 2 : 1);
  const int unit0 = blockIdx.x / csize;
  const int nunits = gridDim.x / csize;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << lead);  // both CTAs of this pair
  // M tile of iteration t for this CTA (t / n_tiles = M unit of the cluster)
  auto m_tile_of = [&](int t) {
    const int mu = t / P.n_tiles;
    return MC ? ((mu * 2 + static_cast<int>(prank)) * 2 + static_cast<int>(rank)) : (mu * (PAIR ? 2 : 1) + static_cast<int>(rank));
  };
  const int nblocks = P.nblocks;

  for (int i = threadIdx.x; i < nblocks; i += kHaloThreads) s_blocks[i] = P.blocks[i];
  for (int i = threadIdx.x; i < P.ntaps; i += kHaloThreads) {
    const HaloTap tp = P.taps[i];
    s_taps[i] = tp;
    // MMA-thread table, one entry per tap in issue order:
    //   x = offset of the shifted activation view (16-byte units) | flags << 16 | K=16 slices << 24
    //   y = offset of this tap's weight tile inside its ring slot (16-byte units)
    int bi = 0;
    while (bi + 1 < nblocks && P.blocks[bi + 1].first_tap <= i) ++bi;
    const HaloBlock blk = P.blocks[bi];
    // Blocks of <= 16 channels (one K = 16 slice per tap; patch frames, EF stage-1 features): the packed weights give
    // every tap 16 consecutive columns, so the 64-column weight tile fetched for tap 4j already holds taps 4j .. 4j+3.
    // Four such taps SHARE one streamed tile (K slice q & 3 of it) instead of fetching four overlapping ones: the weight
    // ring, not the MMA, bounds these layers (a 16-channel source cost a quarter of the FLOPs but a full tile per tap).
    const bool packed = halo_block_packed(P, blk);
    const int q = i - blk.first_tap, wt = packed ? q >> 2 : q, sub = packed ? q & 3 : 0, g = wt % P.bgroup;
    uint32_t flags = 0;
    if (q == 0) flags |= kTapFirstOfBlock;
    if (q == blk.ntaps - 1) flags |= kTapLastOfBlock;
    if (!P.resident) {
      if (g == 0 && sub == 0) flags |= kTapFirstOfGroup;
      if ((g == P.bgroup - 1 && (!packed || sub == 3)) || q == blk.ntaps - 1) flags |= kTapLastOfGroup;
    }
    const int hwp = kTW + 2 * P.P;
    const uint32_t off = (P.debug & 32) ? 0u : static_cast<uint32_t>(((P.P + tp.dy) * hwp + (P.P + tp.dx)) * 128);
    const uint32_t nk = (P.debug & 2) ? 0u : static_cast<uint32_t>(tp.nk & 0xFF);
    const uint32_t regions = static_cast<uint32_t>(tp.nk >> 8) & 3u;
    if constexpr (kRegionKind) {
      if (regions & 1u) flags |= kTapRegionA;
      if (regions & 2u) flags |= kTapRegionB;
    }
    // fuse with the next tap: same block, no group boundary in between, both with 4 K-slices, feeding the same accumulator regions
    if (q + 1 < blk.ntaps && !(flags & (kTapLastOfGroup | kTapLastOfBlock)) && nk == 4 && (P.taps[i + 1].nk & 0xFF) == 4 &&
        (P.resident || (g + 1) % P.bgroup != 0) && (P.taps[i + 1].nk >> 8) == (tp.nk >> 8))
      flags |= kTapFuseNext;
    s_tapmma[i] = make_uint2((off >> 4) | (flags << 16) | (nk << 24),
                             static_cast<uint32_t>(P.resident ? i : g) * (P.b_tap_stride >> 4) + static_cast<uint32_t>(2 * sub));
  }
  if constexpr (MODE == 3) {
    const int np = P.L.N_pad;
    for (int i = threadIdx.x; i < 4 * np + 4; i += kHaloThreads) {
      float v = 0.f;
      if (i < 4 * np) {
        const int pr = i / np, ch = i - pr * np;
        if (pr < P.L.epi.proj_n && ch < P.L.epi.C) v = P.L.epi.proj_w[pr * P.L.epi.C + ch];
      } else if (P.L.epi.proj_b != nullptr && i - 4 * np < P.L.epi.proj_n) {
        v = P.L.epi.proj_b[i - 4 * np];
      }
      s_proj[i] = v;
    }
  }
  if (threadIdx.x < 2) s_tapmma[P.ntaps + threadIdx.x] = make_uint2(0u, 0u);
  for (int i = threadIdx.x; i < P.L.N_pad; i += kHaloThreads) {
    float v = P.L.epi.bias ? P.L.epi.bias[i] : 0.f;
    // lstm_finish (the rolled and the sequence-mode ConvLSTM epilogues) takes the sigmoid gates' biases halved: packed order
    // ch * 4 + (i, f, g, o)
    if constexpr (KIND == EPI_LSTM && (MODE == 2 || MODE == 4))
      if ((i & 3) != 2) v *= 0.5f;
    s_bias[i] = v;
  }

  if (warp == 0 && ptx::elect_one()) {
    for (int i = 0; i < P.L.nsrc; ++i) ptx::prefetch_tensormap(&P.amap[i]);
    ptx::prefetch_tensormap(&P.bmap);
  } else if (warp == 1 && ptx::elect_one()) {
    const uint32_t prod = PAIR ? 2u : 1u;
    for (int i = 0; i < SA; ++i) {
      ptx::mbar_init(afull + 8 * i, prod);
      ptx::mbar_init(aempty + 8 * i, 1);
    }
    for (int i = 0; i < SB; ++i) {
      ptx::mbar_init(bfull + 8 * i, prod);
      ptx::mbar_init(bempty + 8 * i, MC ? 2 : 1);      // MC: a slot is free once BOTH pairs' MMAs have consumed it
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull + 8 * i, 1);
      ptx::mbar_init(tempty + 8 * i, prod * (kEpiThreads / 32));   // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if constexpr (PAIR) ptx::cluster_sync_all();
  if (warp == 2) {
    if constexpr (PAIR) {
      ptx::tmem_alloc_pair(ptx::smem_u32(tmem_slot), static_cast<uint32_t>(P.tmem_cols));
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(tmem_slot), static_cast<uint32_t>(P.tmem_cols));
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(1);
  ptx::pdl_wait();           // the previous kernel has completed: activations / cell state may be read from here on
  if (threadIdx.x == 0) stamp(2);

  const int m_tiles = P.L.B * P.tiles_y * P.tiles_x;
  const int units = MC ? ((m_tiles + 1) / 2 + 1) / 2 : (PAIR ? (m_tiles + 1) / 2 : m_tiles);   // M units per cluster
  const int total = units * P.n_tiles;
  const int rad = P.P;
  const int HWp = kTW + 2 * rad;
  const uint32_t a_box_bytes = static_cast<uint32_t>((kTH + 2 * rad) * HWp * 128);
  const uint32_t b_box_bytes = static_cast<uint32_t>(rowsB) * 128u;
  const uint32_t ncta = PAIR ? 2u : 1u;

  if (warp == 0) {
    // ===================================== activation (halo tile) producer ==================================
    if (ptx::elect_one()) {
      int sa = 0;
      uint32_t ph = 0;
#ifdef VPK_TRACE
      long long p_bar = 0;
      const long long p_begin = clock64();
#endif
      for (int ts = 0; ts < T_steps; ++ts) {
      bool synced = false;      // SEQ: this step's recurrent input has been waited for
      for (int t = unit0; t < total; t += nunits) {
        const int mt = m_tile_of(t);
        const int x0 = (mt % P.tiles_x) * kTW;
        const int y0 = ((mt / P.tiles_x) % P.tiles_y) * kTH;
        const int b0 = mt / (P.tiles_x * P.tiles_y);            // odd tail of a pair: b0 == B -> zero fill
        for (int bi = 0; bi < nblocks; ++bi) {
          const HaloBlock blk = s_blocks[bi];
          int src = blk.src, samp = b0;
          if constexpr (SEQ) {
            if (src == P.seq_recur_src) {
              if (ts == 0) {
                src = P.seq_h0_src;
              } else if (!synced) {
                // step ts reads what every CTA's epilogue wrote in step ts - 1: wait for all arrivals, then order the
                // async-proxy (TMA) reads behind the acquire
                const unsigned target = static_cast<unsigned>(ts) * gridDim.x;
                unsigned seen;
#ifdef VPK_TRACE
                const long long b_t0 = clock64();
#endif
                do {
                  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.seq_barrier) : "memory");
                } while (seen < target);
                asm volatile("fence.proxy.async;" ::: "memory");
#ifdef VPK_TRACE
                p_bar += clock64() - b_t0;
#endif
                synced = true;
              }
            }
            samp = (b0 < P.L.B) ? b0 * P.seq_sb[src] + ts * P.seq_st[src] + P.seq_off[src] : P.seq_oob;
          }
          ptx::mbar_wait_spin(aempty + 8 * sa, ph ^ 1u);
          const uint32_t fb = afull + 8 * sa;
          const uint32_t dst = ptx::smem_u32(smem_a + sa * P.a_slot_bytes);
          if (P.debug & 4) {
            if (leader) ptx::mbar_arrive(fb); else ptx::mbar_arrive_cluster(fb, lead);
          } else if constexpr (PAIR) {
            if (leader) ptx::mbar_arrive_expect_tx(fb, ncta * a_box_bytes); else ptx::mbar_arrive_cluster(fb, lead);
            ptx::tma_load_4d_pair(&P.amap[src], fb, dst, blk.c0, x0 - rad, y0 - rad, samp);
          } else {
            ptx::mbar_arrive_expect_tx(fb, a_box_bytes);
            ptx::tma_load_4d(&P.amap[src], fb, dst, blk.c0, x0 - rad, y0 - rad, samp);
          }
          if (++sa == SA) { sa = 0; ph ^= 1u; }
        }
      }
      }
#ifdef VPK_TRACE
      if (SEQ && trace)
        printf("halo seq trace N=%d taps=%d T=%d: activation producer %lld cycles, of which waiting at the grid barrier %lld\n",
               tileN, P.ntaps, T_steps, clock64() - p_begin, p_bar);
#endif
    }
  } else if (warp == 1) {
    // ===================================== weight producer (groups of taps per slot) ========================
    if (P.resident) {
      // small layers (one N tile, all taps fit): the whole packed weight matrix is loaded once and stays resident
      if (ptx::elect_one() && unit0 < total) {
        const int n0 = static_cast<int>(rank) * (PAIR ? rowsB : 0);
        const uint32_t dst = ptx::smem_u32(smem_b);
        if (P.debug & 4) {
          if (leader) ptx::mbar_arrive(bfull); else ptx::mbar_arrive_cluster(bfull, lead);
        } else {
          if constexpr (PAIR) {
            if (leader) ptx::mbar_arrive_expect_tx(bfull, ncta * b_box_bytes * P.ntaps); else ptx::mbar_arrive_cluster(bfull, lead);
          } else {
            ptx::mbar_arrive_expect_tx(bfull, b_box_bytes * P.ntaps);
          }
          for (int q = 0; q < P.ntaps; ++q) {
            if constexpr (PAIR) ptx::tma_load_2d_pair(&P.bmap, bfull, dst + q * P.b_tap_stride, s_taps[q].wk, n0);
            else ptx::tma_load_2d(&P.bmap, bfull, dst + q * P.b_tap_stride, s_taps[q].wk, n0);
          }
        }
      }
    } else if (ptx::elect_one()) {
      int sb = 0;
      uint32_t ph = 0;
      for (int ts = 0; ts < T_steps; ++ts)
      for (int t = unit0; t < total; t += nunits) {
        const int n0 = (t % P.n_tiles) * tileN + static_cast<int>(rank) * (PAIR ? rowsB : 0);
        for (int bi = 0; bi < nblocks; ++bi) {
          const HaloBlock blk = s_blocks[bi];
          const int tstep = halo_block_packed(P, blk) ? 4 : 1;             // taps per streamed weight tile
          const int ntile = (blk.ntaps + tstep - 1) / tstep;
          for (int q0 = 0; q0 < ntile; q0 += P.bgroup) {
            const int gn = min(P.bgroup, ntile - q0);
            ptx::mbar_wait_spin(bempty + 8 * sb, ph ^ 1u);
            const uint32_t fb = bfull + 8 * sb;
            const uint32_t dst = ptx::smem_u32(smem_b + sb * P.b_slot_bytes);
            if (P.debug & 4) {
              if (leader) ptx::mbar_arrive(fb); else ptx::mbar_arrive_cluster(fb, lead);
            } else {
              if constexpr (PAIR) {
                if (leader) ptx::mbar_arrive_expect_tx(fb, ncta * b_box_bytes * gn); else ptx::mbar_arrive_cluster(fb, lead);
              } else {
                ptx::mbar_arrive_expect_tx(fb, b_box_bytes * gn);
              }
              for (int q = 0; q < gn; ++q) {
                const int wk = s_taps[blk.first_tap + (q0 + q) * tstep].wk;
                if constexpr (PAIR) {
                  if (MC) {    // taps alternate between the two pairs' producers; each load feeds the same-half CTA of both
                    if ((q & 1) == static_cast<int>(prank))
                      ptx::tma_load_2d_pair_mc(&P.bmap, fb, dst + q * P.b_tap_stride, wk, n0,
                                               static_cast<uint16_t>((1u << rank) | (4u << rank)));
                  } else {
                    ptx::tma_load_2d_pair(&P.bmap, fb, dst + q * P.b_tap_stride, wk, n0);
                  }
                } else {
                  ptx::tma_load_2d(&P.bmap, fb, dst + q * P.b_tap_stride, wk, n0);
                }
              }
            }
            if (++sb == SB) { sb = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================== MMA issuer (leader CTA of a pair) ================================
    // ONE thread runs a flat, table-driven loop over the taps of a tile (s_tapmma): per tap one 8-byte table read, at
    // most two barrier waits, one asm block with the (up to four) tcgen05.mma of the tap, at most two commits.  Every
    // instruction here competes for issue slots with two epilogue warps on the same scheduler, and the tensor pipe
    // idles whenever this thread falls behind (ncu: 73 % tensor-active with the previous ~60-instruction tap).
    // Descriptors move by their low word only (start address in 16-byte units; shared memory < 256 KB: no carry).
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(PAIR ? 256 : 128, tileN, P.L.op_f16 != 0);
      // accumulator regions (pairs only): widths over the pair, first column of region B, first weight row of region B in a half
      const int rg_a = P.L.region_g0, rg_cn = tileN / G;
      const uint32_t idesc_ra = ptx::idesc_bf16_f32(256, rg_a > 0 ? rg_cn * rg_a : tileN, P.L.op_f16 != 0);
      const uint32_t idesc_rb = ptx::idesc_bf16_f32(256, rg_a > 0 ? rg_cn * (G - rg_a) : tileN, P.L.op_f16 != 0);
      const uint32_t region_col_b = static_cast<uint32_t>(rg_cn * rg_a);
      const uint32_t region_row_b = static_cast<uint32_t>((rg_cn / 2) * rg_a * 128) >> 4;
      const uint32_t sbo = (P.debug & 64) ? 1024u : static_cast<uint32_t>(HWp * 128);
      const uint64_t adesc0 = smem_desc_sw128_sbo(ptx::smem_u32(smem_a), sbo);
      const uint64_t bdesc0 = ptx::smem_desc_sw128(ptx::smem_u32(smem_b));
      const uint32_t a_slot_u = P.a_slot_bytes >> 4, b_slot_u = P.b_slot_bytes >> 4;
      const int ntaps = P.ntaps;
      const uint32_t tab0 = ptx::smem_u32(s_tapmma);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      uint64_t a_d = adesc0, b_d = bdesc0;
      int iter = 0;
#ifdef VPK_TRACE
      long long w_acc = 0, w_a = 0, w_b = 0, t_begin = clock64();     // cycles the MMA thread spent waiting, by cause
#define VPK_TIMED(counter, stmt) do { const long long c0_ = clock64(); stmt; counter += clock64() - c0_; } while (0)
#else
#define VPK_TIMED(counter, stmt) stmt
#endif
      if (P.resident && unit0 < total) ptx::mbar_wait_spin(bfull, 0);
      for (int ts = 0; ts < T_steps; ++ts)
      for (int t = unit0; t < total; t += nunits, ++iter) {
        const int acc = iter & 1;
        VPK_TIMED(w_acc, ptx::mbar_wait_spin(tempty + 8 * acc, ((iter >> 1) & 1u) ^ 1u));
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * tileN);
        uint32_t accum = 0, accum_b = 0;
        uint32_t tab = tab0;
        uint2 ti = ptx::lds_u2(tab);
        for (int i = 0; i < ntaps; ++i) {
          const uint2 cur = ti;
          tab += 8;
          ti = ptx::lds_u2(tab);            // the table has one padding entry: no bounds test on the prefetch
          const uint32_t flags = cur.x >> 16;
          if (flags & (kTapFirstOfBlock | kTapFirstOfGroup)) {
            if (flags & kTapFirstOfBlock) {
              VPK_TIMED(w_a, ptx::mbar_wait_spin(afull + 8 * sa, pha));
              a_d = adesc0 + static_cast<uint32_t>(sa) * a_slot_u;
              if (iter == 0 && i == 0) stamp(3);
            }
            if (flags & kTapFirstOfGroup) {
              VPK_TIMED(w_b, ptx::mbar_wait_spin(bfull + 8 * sb, phb));
              b_d = bdesc0 + static_cast<uint32_t>(sb) * b_slot_u;
            }
          }
          uint32_t lflags = flags;
          // accumulator regions exist only in the CTA-pair instantiations of the ST-LSTM kinds (kRegionKind): every other
          // kernel keeps the plain issue loop -- the small-N layers are bound by this thread's instruction count
          bool region_tap = false;
          if constexpr (kRegionKind) region_tap = (flags & (kTapRegionA | kTapRegionB)) != 0;
          if (flags & kTapFuseNext) {
            const uint2 nx = ti;              // the table entry after `cur` is consumed by the same asm block
            tab += 8;
            ti = ptx::lds_u2(tab);
            ++i;
            if constexpr (PAIR) {
              if (kRegionKind && region_tap) {      // accumulator regions (see the single-tap form below)
                if (flags & kTapRegionA) {
                  ptx::mma_bf16_ss_tap2_pair(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, a_d + (nx.x & 0xFFFFu), b_d + nx.y, idesc_ra, accum);
                  accum = 1u;
                }
                if (flags & kTapRegionB) {
                  ptx::mma_bf16_ss_tap2_pair(tmem_d + region_col_b, a_d + (cur.x & 0xFFFFu), b_d + cur.y + region_row_b,
                                             a_d + (nx.x & 0xFFFFu), b_d + nx.y + region_row_b, idesc_rb, accum_b);
                  accum_b = 1u;
                }
              } else {
                ptx::mma_bf16_ss_tap2_pair(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, a_d + (nx.x & 0xFFFFu), b_d + nx.y, idesc, accum);
                accum = 1u;
              }
            } else {
              ptx::mma_bf16_ss_tap2(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, a_d + (nx.x & 0xFFFFu), b_d + nx.y, idesc, accum);
              accum = 1u;
            }
            lflags = nx.x >> 16;
          } else if (kRegionKind && region_tap) {
            // accumulator regions: region A = columns [0, N_A) <- the first rows of each CTA's weight half, region B behind it;
            // a tap multiplies only what its weight tensor feeds (N is constant per region, as cta_group::2's column split needs)
            if constexpr (PAIR) {
              if (flags & kTapRegionA) {
                ptx::mma_bf16_ss_tap_pair(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, idesc_ra, accum, (cur.x >> 24) & 0xFFu);
                accum = 1u;
              }
              if (flags & kTapRegionB) {
                ptx::mma_bf16_ss_tap_pair(tmem_d + region_col_b, a_d + (cur.x & 0xFFFFu), b_d + cur.y + region_row_b, idesc_rb, accum_b,
                                          (cur.x >> 24) & 0xFFu);
                accum_b = 1u;
              }
            }
          } else {
            if constexpr (PAIR) ptx::mma_bf16_ss_tap_pair(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, idesc, accum, (cur.x >> 24) & 0xFFu);
            else ptx::mma_bf16_ss_tap(tmem_d, a_d + (cur.x & 0xFFFFu), b_d + cur.y, idesc, accum, (cur.x >> 24) & 0xFFu);
            accum = 1u;
          }
          if (lflags & (kTapLastOfGroup | kTapLastOfBlock)) {
            const uint32_t flags = lflags;
            if (flags & kTapLastOfGroup) {
              if constexpr (PAIR) ptx::mma_commit_pair(bempty + 8 * sb, MC ? static_cast<uint16_t>(0xF) : pair_mask);
              else ptx::mma_commit(bempty + 8 * sb);
              if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
            if (flags & kTapLastOfBlock) {
              if constexpr (PAIR) ptx::mma_commit_pair(aempty + 8 * sa, pair_mask);
              else ptx::mma_commit(aempty + 8 * sa);
              if (++sa == SA) { sa = 0; pha ^= 1u; }
            }
          }
        }
        if constexpr (PAIR) ptx::mma_commit_pair(tfull + 8 * acc, pair_mask);
        else ptx::mma_commit(tfull + 8 * acc);
        if (iter == 0) stamp(4);
      }
      stamp(5);
#ifdef VPK_TRACE
      if (trace)
        printf("halo trace N=%d taps=%d: MMA thread %lld cycles for %d tiles: waiting for accumulator %lld, activations %lld, "
               "weights %lld\n", tileN, ntaps, clock64() - t_begin, iter, w_acc, w_a, w_b);
#endif
#undef VPK_TIMED
    }
    __syncwarp();
  } else if (warp >= 3) {
    // ===================================== epilogue (warps 4..11) ===========================================
    const int quad = warp & 3;
    const int half = (warp - 3) >> 2;
    const int row = quad * 32 + lane;
    const int rx = row % kTW;
    const int ry = row / kTW;
    const int Cn = tileN / G;
    int iter = 0;
    const uint32_t sb_lstm = ptx::smem_u32(s_bias);        // lstm_finish reads the staged biases through ld.shared
#ifdef VPK_TRACE
    long long e_wait = 0, e_body = 0, e_t1 = 0, e_ld = 0;
#endif
    constexpr bool rolled = (MODE == 2 && KIND == EPI_LSTM);
    if constexpr (rolled) {
      {
        // ---- whole-tile operand prefetch (epilogue_tc.cuh: LstmOps) ----
        auto run = [&](auto peep_c, auto nch_c) {
          constexpr bool PEEP = decltype(peep_c)::value;
          constexpr int NCH = decltype(nch_c)::value;      // this warp's chunks: channels (half + 2k) * 8 of the tile
          constexpr int PPD = (NCH % 2 == 0) ? 2 : NCH;    // peephole prefetch distance in chunks (= ring size)
          const EpiParams& E = P.L.epi;
          const int C = E.C;
          const long long hw = static_cast<long long>(P.L.H) * P.L.W;
          LstmOps ops[NCH];
          LstmPeep pps[PPD];
          auto locate = [&](int t, LstmTile& et, int& chb) -> bool {
            const int nt = t % P.n_tiles;
            const int mt = m_tile_of(t);
            const int x = (mt % P.tiles_x) * kTW + rx;
            const int y = ((mt / P.tiles_x) % P.tiles_y) * kTH + ry;
            const int b = mt / (P.tiles_x * P.tiles_y);
            chb = nt * Cn;
            const bool ok = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
            if (ok) et = lstm_tile(E, b, y, x, P.L.H, P.L.W);
            return ok;
          };
          LstmTile et{};
          int chb = 0;
          bool valid = false;
          int t = unit0;
          if (t < total) {
            valid = locate(t, et, chb);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              const int chl = (half + 2 * k) * 8;
              if (valid && chb + chl < C) {
                lstm_c_load(E, et, hw, chb + chl, ops[k]);
                if constexpr (PEEP)
                  if (k < PPD) lstm_peep_load(E, et, hw, chb + chl, pps[k]);
              }
            }
          }
          for (; t < total; t += nunits, ++iter) {
            LstmTile etn{};
            int chbn = 0;
            bool validn = false;
            if (t + nunits < total) validn = locate(t + nunits, etn, chbn);
            const int acc = iter & 1;
            ptx::mbar_wait_fast(tfull + 8 * acc, (iter >> 1) & 1u);
            ptx::tc_fence_after();
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              const int chl = (half + 2 * k) * 8;
              uint32_t r[32];
              ptx::tmem_ld32(taddr + static_cast<uint32_t>(chl * 4), r);
              ptx::tmem_ld_wait();
              if (valid && chb + chl < C) {
                float a[4][8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                  for (int g = 0; g < 4; ++g) a[g][j] = __uint_as_float(r[j * 4 + g]);
                lstm_finish<PEEP>(E, et, hw, chb + chl, sb_lstm + static_cast<uint32_t>((chb + chl) * 16), a, ops[k], pps[k % PPD]);
              }
              if (validn && chbn + chl < C) lstm_c_load(E, etn, hw, chbn + chl, ops[k]);
              if constexpr (PEEP) {
                const int kk = k + PPD;                    // chunk whose peepholes take this ring slot next
                if (kk < NCH) {
                  const int chl2 = (half + 2 * kk) * 8;
                  if (valid && chb + chl2 < C) lstm_peep_load(E, et, hw, chb + chl2, pps[k % PPD]);
                } else {
                  const int chl2 = (half + 2 * (kk - NCH)) * 8;
                  if (validn && chbn + chl2 < C) lstm_peep_load(E, etn, hw, chbn + chl2, pps[k % PPD]);
                }
              }
            }
            ptx::tc_fence_before();
            __syncwarp();            // every lane's tcgen05.ld has completed (wait::ld) and is fenced: one arrival per warp
            if (lane == 0) {
              if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
              else ptx::mbar_arrive(tempty + 8 * acc);
            }
            et = etn;
            chb = chbn;
            valid = validn;
          }
        };
        auto run_n = [&](auto peep_c) {
          switch ((Cn / 8 - half + 1) / 2) {
            case 1: run(peep_c, std::integral_constant<int, 1>{}); break;
            case 2: run(peep_c, std::integral_constant<int, 2>{}); break;
            case 3: run(peep_c, std::integral_constant<int, 3>{}); break;
            case 4: run(peep_c, std::integral_constant<int, 4>{}); break;
            default: __trap();      // the plan only selects this kernel for 8 <= Cn <= 64
          }
        };
        if (P.L.epi.pp16 != nullptr) run_n(std::true_type{});
        else run_n(std::false_type{});
      }
    }
    if constexpr (SEQ && KIND == EPI_LSTM) {
      // ---- sequence mode: T_steps timesteps, cell state resident in shared memory ----
      auto run = [&](auto peep_c, auto nch_c) {
        constexpr bool PEEP = decltype(peep_c)::value;
        constexpr int NCH = decltype(nch_c)::value;        // this warp's chunks: channels (half + 2k) * 8 of the tile
        const EpiParams& E = P.L.epi;
        const int C = E.C;
        const long long hw = static_cast<long long>(P.L.H) * P.L.W;
        const int chunks = Cn >> 3;
        for (int ts = 0; ts < T_steps; ++ts) {
          int slot = 0;
          for (int t = unit0; t < total; t += nunits, ++iter, ++slot) {
            const int nt = t % P.n_tiles;
            const int mt = m_tile_of(t);
            const int x = (mt % P.tiles_x) * kTW + rx;
            const int y = ((mt / P.tiles_x) % P.tiles_y) * kTH + ry;
            const int b = mt / (P.tiles_x * P.tiles_y);
            const int chb = nt * Cn;
            const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B);
            LstmTile et{};
            // everything the chunks address is prepared here, BEFORE the accumulator wait (this part overlaps the tile's
            // MMAs; what follows the wait is exposed once per timestep): h' / peephole / shared-state pointers of chunk 0,
            // chunk k = + k * constant
            bf16* hout0 = nullptr;
            float* h32out0 = nullptr;
            const uint4* pp0 = nullptr;
            if (valid) {
              et = lstm_tile(E, b, y, x, P.L.H, P.L.W);      // st_off: c_0 / c_T in global memory, indexed by the sequence b
              const long long bo = static_cast<long long>(b) * P.seq_out_sb + static_cast<long long>(ts) * P.seq_out_st + P.seq_out_off;
              et.out_off = bo * E.oB + y * E.oY + x * E.oX;  // h'_t: slot (b, ts) of the output sequence
              hout0 = static_cast<bf16*>(E.out) + et.out_off + chb + half * 8;
              if (E.h32 != nullptr) h32out0 = E.h32 + et.out_off + chb + half * 8;
              if constexpr (PEEP) pp0 = static_cast<const uint4*>(E.pp16) + (static_cast<long long>((chb >> 3) + half) * hw + et.pos) * 3;
            }
            const long long pp_step = 2 * hw * 3;             // uint4 units between this warp's consecutive chunks (16 channels)
            float4* const cp0 = reinterpret_cast<float4*>(s_cstate + ((static_cast<size_t>(slot) * chunks + half) * 128 + row) * 8);
            const int acc = iter & 1;
#ifdef VPK_TRACE
            const long long q_t0 = clock64();
#endif
            ptx::mbar_wait_fast(tfull + 8 * acc, (iter >> 1) & 1u);
            ptx::tc_fence_after();
#ifdef VPK_TRACE
            const long long q_t1 = clock64();
            e_wait += q_t1 - q_t0;
#endif
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
            LstmPeep pp[2];
            auto peep_ld = [&](int k, LstmPeep& dst) {
              const uint4* q = pp0 + k * pp_step;
              dst.p[0] = __ldg(q);
              dst.p[1] = __ldg(q + 1);
              dst.p[2] = __ldg(q + 2);
            };
            if constexpr (PEEP)
              if (valid && chb + half * 8 < C) peep_ld(0, pp[0]);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              const int chl = (half + 2 * k) * 8;
              uint32_t r[32];
              ptx::tmem_ld32(taddr + static_cast<uint32_t>(chl * 4), r);
              if constexpr (PEEP)
                if (k + 1 < NCH && valid && chb + chl + 16 < C) peep_ld(k + 1, pp[(k + 1) & 1]);
              ptx::tmem_ld_wait();
              if (valid && chb + chl < C) {
                float4* cp = cp0 + k * (2 * 128 * 2);          // chunk (half + 2k): 2 chunks x 128 rows x 2 float4 further
                LstmOps o;
                if (ts == 0) {
                  if (P.seq_c_zero) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o.c[j] = 0.f;
                  } else {
                    lstm_c_load(E, et, hw, chb + chl, o);
                  }
                } else {
                  const float4 c0 = cp[0], c1 = cp[1];
                  o.c[0] = c0.x; o.c[1] = c0.y; o.c[2] = c0.z; o.c[3] = c0.w;
                  o.c[4] = c1.x; o.c[5] = c1.y; o.c[6] = c1.z; o.c[7] = c1.w;
                }
                float a[4][8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                  for (int g = 0; g < 4; ++g) a[g][j] = __uint_as_float(r[j * 4 + g]);
                lstm_finish<PEEP, false, true>(E, et, hw, chb + chl, sb_lstm + static_cast<uint32_t>((chb + chl) * 16), a, o, pp[k & 1],
                                               hout0 + 16 * k, h32out0 != nullptr ? h32out0 + 16 * k : nullptr);
                cp[0] = make_float4(o.c[0], o.c[1], o.c[2], o.c[3]);
                cp[1] = make_float4(o.c[4], o.c[5], o.c[6], o.c[7]);
                if (ts == T_steps - 1 && E.s0 != nullptr) st_state8(E.s0, et.st_off, E.state_c4 ? hw * 4 : 0, chb + chl, o.c);
              }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
              else ptx::mbar_arrive(tempty + 8 * acc);
            }
#ifdef VPK_TRACE
            e_body += clock64() - q_t1;
#endif
          }
          if (ts + 1 < T_steps) {
            // this CTA's h'_ts is complete: publish it (generic-proxy stores -> other CTAs' TMA reads) and arrive once
#ifdef VPK_TRACE
            const long long q_t2 = clock64();
#endif
            // (every thread orders its own generic-proxy stores for the async proxy; the CTA barrier orders them before the
            // elected thread, whose gpu-scope fence + release are cumulative over them -- one fence per CTA, not 256:
            // the publish phase was 2.3 us of every 17 us step)
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (warp == 3 && lane == 0) {
              __threadfence();
              asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(P.seq_barrier), "r"(1u) : "memory");
            }
#ifdef VPK_TRACE
            e_ld += clock64() - q_t2;
#endif
          }
        }
#ifdef VPK_TRACE
        if (trace && warp == 3 && lane == 0)
          printf("halo seq trace N=%d taps=%d T=%d: epilogue warp, %d tiles: waiting for the accumulator %lld cycles, fused update %lld, "
                 "publish (fences + CTA barrier + release) %lld\n", tileN, P.ntaps, T_steps, iter, e_wait, e_body, e_ld);
#endif
      };
      auto run_n = [&](auto peep_c) {
        switch ((Cn / 8 - half + 1) / 2) {
          case 1: run(peep_c, std::integral_constant<int, 1>{}); break;
          case 2: run(peep_c, std::integral_constant<int, 2>{}); break;
          case 3: run(peep_c, std::integral_constant<int, 3>{}); break;
          case 4: run(peep_c, std::integral_constant<int, 4>{}); break;
          default: __trap();
        }
      };
      if (P.L.epi.pp16 != nullptr) run_n(std::true_type{});
      else run_n(std::false_type{});
    }
    if constexpr (MODE == 3) {
      // ---- bias + activation + 1x1 projection: one warp per TMEM quadrant reads all channels of its positions ----
      const EpiParams& E = P.L.epi;
      const int np = P.L.N_pad;
      for (int t = unit0; t < total; t += nunits, ++iter) {
        const int mt = m_tile_of(t);                                         // n_tiles == 1
        const int x = (mt % P.tiles_x) * kTW + rx;
        const int y = ((mt / P.tiles_x) % P.tiles_y) * kTH + ry;
        const int b = mt / (P.tiles_x * P.tiles_y);
        const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
        const int acc = iter & 1;
        ptx::mbar_wait_fast(tfull + 8 * acc, (iter >> 1) & 1u);
        ptx::tc_fence_after();
        if (half == 0) {
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
          float z[4] = {s_proj[4 * np], s_proj[4 * np + 1], s_proj[4 * np + 2], s_proj[4 * np + 3]};
          for (int ch0 = 0; ch0 < Cn; ch0 += 32) {          // up to four 8-channel chunks per tcgen05.wait::ld
            uint32_t r4[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (ch0 + 8 * q < Cn) ptx::tmem_ld8(taddr + static_cast<uint32_t>(ch0 + 8 * q), r4[q]);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int ch = ch0 + 8 * q;
              if (ch < Cn) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float v = __uint_as_float(r4[q][j]) + s_bias[ch + j];
                  if (E.act == ACT_LEAKY) v = v > 0.f ? v : 0.2f * v;
                  else if (E.act == ACT_RELU) v = fmaxf(v, 0.f);
                  else if (E.act == ACT_SIGMOID) v = sigmoid_fast(v);
#pragma unroll
                  for (int pr = 0; pr < 4; ++pr) z[pr] = fmaf(s_proj[pr * np + ch + j], v, z[pr]);
                }
              }
            }
          }
          if (valid) {
            float* o = static_cast<float*>(E.out) + b * E.oB + y * E.oY + x * E.oX;
#pragma unroll
            for (int pr = 0; pr < 4; ++pr)
              if (pr < E.proj_n) o[pr * E.oC] = z[pr];
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
          else ptx::mbar_arrive(tempty + 8 * acc);
        }
      }
    }
    bool lean_done = false;
    if constexpr (KIND == EPI_BIAS_ACT && MODE == 1) {
      if (P.lean) {
        // ---- bias + slope activation + 16-bit store (+ GroupNorm partial sums), nothing else: stage convs, sub-pixel deconv
        // parities, the DCGAN convs.  An epilogue warp is ONE instruction stream: at ~7 cycles per dependent instruction the
        // ~340 instructions per tile of the general loop (tile decode by division, per-chunk predicates, activation switch,
        // 64-bit address chains) took 2.6 k cycles -- longer than the 24 MMAs of an N = 96, K = 384 tile -- and its
        // per-thread st.global (32 lines per instruction) stalled the warps another 2.3 k.  Here: tile decode by
        // multiply-high, chunk count and activation compile-time, all TMEM reads of a tile behind one wait, and (TMA) the
        // tile staged in shared memory and written by bulk tensor stores. ----
        lean_done = true;
        const EpiParams& E = P.L.epi;
        const uint32_t tpi = static_cast<uint32_t>(P.tiles_x * P.tiles_y);
        const uint32_t sb0 = ptx::smem_u32(s_bias) + static_cast<uint32_t>(half * 32);
        bf16* const outp = static_cast<bf16*>(E.out) + half * 8;
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(half * 8);
        const float slope = E.act == ACT_LEAKY ? 0.2f : 1.f;
        const bool f16o = E.out_f16 != 0;                    // 16-bit output as fp16 (fp16-operand launches), else bf16
        auto run = [&](auto nj_c, auto relu_c, auto tma_c, auto stats_c) {
          constexpr int NJ = decltype(nj_c)::value;          // this warp's 8-channel chunks: channels half * 8 + 16 j
          constexpr bool RELU = decltype(relu_c)::value;
          constexpr bool TMA = decltype(tma_c)::value;
          constexpr bool STATS = decltype(stats_c)::value;   // per-group sum / sum of squares for a following GroupNorm
          // TMA: the tile is staged as NJ / 2 sub-tiles of [128 positions][32 channels] (64 B rows, 64B swizzle; one
          // [128][16] sub-tile with 32 B rows and the 32B swizzle when NJ == 1) -- conflict-free 16 B st.shared
          constexpr int SUBS = NJ == 1 ? 1 : NJ / 2;
          constexpr uint32_t kSubBytes = NJ == 1 ? 128u * 32u : 128u * 64u;
          const bool issuer = TMA && warp == 3 && lane == 0;
          uint32_t srow = 0;
          if constexpr (TMA) srow = NJ == 1 ? static_cast<uint32_t>(row * 32) : static_cast<uint32_t>(row * 64);
          const uint32_t sxor = NJ == 1 ? static_cast<uint32_t>((row >> 2) & 1) : static_cast<uint32_t>((row >> 1) & 3);
          for (int t = unit0; t < total; t += nunits, ++iter) {
            // (M unit, N tile) -> (sequence, tile row, tile column, first channel): exact multiply-high divisions (plan)
            uint32_t mu = static_cast<uint32_t>(t), nt = 0;
            if (P.n_tiles != 1) {
              mu = __umulhi(static_cast<uint32_t>(t), P.magic_nt);
              nt = static_cast<uint32_t>(t) - mu * static_cast<uint32_t>(P.n_tiles);
            }
            const uint32_t mt = MC ? ((mu * 2 + prank) * 2 + rank) : (mu * (PAIR ? 2u : 1u) + rank);
            const int b = static_cast<int>(tpi == 1 ? mt : __umulhi(mt, P.magic_tpi));
            const uint32_t rem = mt - static_cast<uint32_t>(b) * tpi;
            const int ty = static_cast<int>(P.tiles_x == 1 ? rem : __umulhi(rem, P.magic_tx));
            const int tx = static_cast<int>(rem) - ty * P.tiles_x;
            const int ch_base = static_cast<int>(nt) * Cn;
            const int x = tx * kTW + rx, y = ty * kTH + ry;
            const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
            bf16* o = outp + (b * E.oB + y * E.oY + x * E.oX) + ch_base;
            const uint32_t sb = sb0 + static_cast<uint32_t>(ch_base * 4);
            const int acc = iter & 1;
#ifdef VPK_TRACE
            const long long l_t0 = clock64();
#endif
            ptx::mbar_wait_fast(tfull + 8 * acc, (iter >> 1) & 1u);
            ptx::tc_fence_after();
#ifdef VPK_TRACE
            const long long l_t1 = clock64();
            e_wait += l_t1 - l_t0;
#endif
            const uint32_t ta = tq + static_cast<uint32_t>(acc * tileN);
            uint32_t r[NJ][8];
#pragma unroll
            for (int j = 0; j < NJ; ++j) ptx::tmem_ld8(ta + static_cast<uint32_t>(16 * j), r[j]);
            float4 bv[NJ][2];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              bv[j][0] = ptx::lds_f4(sb + static_cast<uint32_t>(64 * j));
              bv[j][1] = ptx::lds_f4(sb + static_cast<uint32_t>(64 * j + 16));
            }
            ptx::tmem_ld_wait();
#ifdef VPK_TRACE
            const long long l_t2 = clock64();
            e_ld += l_t2 - l_t1;
#endif
            if constexpr (TMA) {      // the accumulator is in registers: hand the TMEM buffer back before the stores
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
                else ptx::mbar_arrive(tempty + 8 * acc);
              }
            }
            const uint32_t sbuf = s_stage + static_cast<uint32_t>(iter & 1) * (SUBS * kSubBytes) + srow;
            float gs16[STATS ? 16 : 1];
            if constexpr (STATS) {
#pragma unroll
              for (int q = 0; q < 16; ++q) gs16[q] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              const float bj[8] = {bv[j][0].x, bv[j][0].y, bv[j][0].z, bv[j][0].w, bv[j][1].x, bv[j][1].y, bv[j][1].z, bv[j][1].w};
              float v[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float a = __uint_as_float(r[j][k]) + bj[k];
                v[k] = RELU ? fmaxf(a, 0.f) : (a > 0.f ? a : slope * a);
              }
              if constexpr (TMA) {
                uint4 pk;
                if (f16o) {
                  __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                  for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                } else {
                  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                  for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                }
                // channel half * 8 + 16 j: sub-tile j / 2, 16-byte chunk (half + 2 j) % 4 of its row
                const uint32_t chunk = NJ == 1 ? static_cast<uint32_t>(half) : static_cast<uint32_t>((half + 2 * j) & 3);
                ptx::sts_u4(sbuf + static_cast<uint32_t>(j >> 1) * kSubBytes + ((chunk ^ sxor) << 4), pk);
              } else {
                if (valid) {
                  if (f16o) {
                    uint4 pk;
                    __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) h2[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                    *reinterpret_cast<uint4*>(o + 16 * j) = pk;
                  } else {
                    st_bf16x8(o + 16 * j, v);
                  }
                }
              }
              if constexpr (STATS) {       // positions outside the image contribute zeros
                if (!valid) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) v[k] = 0.f;
                }
                const int gsz = E.gn_group_size;
                if (gsz == 2) gn_accum<2>(gs16, j, v);
                else if (gsz == 4) gn_accum<4>(gs16, j, v);
                else gn_accum<8>(gs16, j, v);
              }
            }
            if constexpr (TMA) {
              ptx::fence_proxy_async_smem();               // this thread's st.shared -> visible to the bulk store
              if (issuer) ptx::bulk_wait_read0();          // the previous tile's store has read its buffer (the one the
                                                           // NEXT tile writes, after the barrier below)
              asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
              if (issuer && b < P.L.B && !(P.debug & 1)) {
#pragma unroll
                for (int u = 0; u < SUBS; ++u)
                  ptx::tma_store_4d(&P.omap, s_stage + static_cast<uint32_t>(iter & 1) * (SUBS * kSubBytes) + u * kSubBytes,
                                    ch_base + u * 32, tx * kTW, ty * kTH, b);
                ptx::bulk_commit();
              }
            } else {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
                else ptx::mbar_arrive(tempty + 8 * acc);
              }
            }
            if constexpr (STATS) {
              const int gsz = E.gn_group_size;
              const float tot = gn_warp_reduce16(gs16, lane);
              const int vi = gn_lane_value(lane), pairi = vi >> 1, gpc = 8 / gsz;
              const int kk = pairi / gpc, jj = pairi - kk * gpc;
              const int chg = ch_base + half * 8 + 16 * kk;       // first channel of that chunk
              if ((lane & 1) == 0 && b < P.L.B && kk < NJ) {
                const int slot = E.gn_slot0 + static_cast<int>(rem) * 4 + quad;
                E.gn_sums[((static_cast<long long>(b) * E.gn_nslots + slot) * (E.C / gsz) + chg / gsz + jj) * 2 + (vi & 1)] = tot;
              }
            }
#ifdef VPK_TRACE
            e_body += clock64() - l_t2;
#endif
          }
          if (issuer) ptx::bulk_wait0();                   // all staged tiles are in global memory before the CTA exits
        };
        auto run_r = [&](auto nj_c, auto tma_c) {
          if (E.act == ACT_RELU) run(nj_c, std::true_type{}, tma_c, std::false_type{});
          else run(nj_c, std::false_type{}, tma_c, std::false_type{});
        };
        const bool stats = E.gn_sums != nullptr;             // (the plan admits group statistics with <= 64 channels per tile)
        if (stats && P.lean_tma) {
          switch (Cn >> 4) {
            case 1: run(std::integral_constant<int, 1>{}, std::false_type{}, std::true_type{}, std::true_type{}); break;
            case 2: run(std::integral_constant<int, 2>{}, std::false_type{}, std::true_type{}, std::true_type{}); break;
            case 4: run(std::integral_constant<int, 4>{}, std::false_type{}, std::true_type{}, std::true_type{}); break;
            default: __trap();
          }
        } else if (stats) {
          switch (Cn >> 4) {
            case 1: run(std::integral_constant<int, 1>{}, std::false_type{}, std::false_type{}, std::true_type{}); break;
            case 2: run(std::integral_constant<int, 2>{}, std::false_type{}, std::false_type{}, std::true_type{}); break;
            case 3: run(std::integral_constant<int, 3>{}, std::false_type{}, std::false_type{}, std::true_type{}); break;
            case 4: run(std::integral_constant<int, 4>{}, std::false_type{}, std::false_type{}, std::true_type{}); break;
            default: __trap();
          }
        } else if (P.lean_tma) {
          switch (Cn >> 4) {
            case 1: run_r(std::integral_constant<int, 1>{}, std::true_type{}); break;
            case 2: run_r(std::integral_constant<int, 2>{}, std::true_type{}); break;
            case 4: run_r(std::integral_constant<int, 4>{}, std::true_type{}); break;
            case 6: run_r(std::integral_constant<int, 6>{}, std::true_type{}); break;
            default: __trap();    // the plan sets `lean_tma` for 16 / 32 / 64 / 96 channels only
          }
        } else
        switch (Cn >> 4) {
          case 1: run_r(std::integral_constant<int, 1>{}, std::false_type{}); break;
          case 2: run_r(std::integral_constant<int, 2>{}, std::false_type{}); break;
          case 3: run_r(std::integral_constant<int, 3>{}, std::false_type{}); break;
          case 4: run_r(std::integral_constant<int, 4>{}, std::false_type{}); break;
          case 6: run_r(std::integral_constant<int, 6>{}, std::false_type{}); break;
          default: __trap();      // the plan sets `lean` for these chunk counts only
        }
      }
    }
    if constexpr (!rolled && MODE != 3 && !SEQ)
    if (!lean_done)
    for (int t = unit0; t < total; t += nunits, ++iter) {
      const int nt = t % P.n_tiles;
      const int mt = m_tile_of(t);
      const int x = (mt % P.tiles_x) * kTW + rx;
      const int y = ((mt / P.tiles_x) % P.tiles_y) * kTH + ry;
      const int b = mt / (P.tiles_x * P.tiles_y);
      const bool valid = (x < P.L.W) && (y < P.L.H) && (b < P.L.B) && !(P.debug & 1);
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      EpiOperands<8> ops0, ops1;
      const int C = P.L.epi.C;
      const int ch_base = nt * Cn;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * tileN);
      auto tmem_chunk = [&](int ch, uint32_t (&r)[8 * G]) {
        if constexpr (kRegionKind) {
          if (P.L.region_g0 > 0) {
            // accumulator regions: columns [0, N_A) hold gates [0, g0), CTA halves side by side inside each region
            // (cta_group::2 takes columns [0, N/2) of an MMA from the leader's weight rows and [N/2, N) from its peer's)
            constexpr int GA = (G == 4) ? 3 : 1, GB = 1;
            const int hc = Cn >> 1, hf = ch >= hc ? 1 : 0, cj = ch - hf * hc;
            const uint32_t ca = taddr + static_cast<uint32_t>(hf * hc * GA + cj * GA);
            const uint32_t cb = taddr + static_cast<uint32_t>(Cn * GA + hf * hc * GB + cj * GB);
            uint32_t ra[8 * GA], rb[8 * GB];
            if constexpr (GA == 3) { ptx::tmem_ld8(ca, ra); ptx::tmem_ld8(ca + 8, ra + 8); ptx::tmem_ld8(ca + 16, ra + 16); }
            else ptx::tmem_ld8(ca, ra);
            ptx::tmem_ld8(cb, rb);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
              for (int g = 0; g < GA; ++g) r[j * G + g] = ra[j * GA + g];
              r[j * G + GA] = rb[j];
            }
            return;
          }
        }
        const uint32_t ta = taddr + static_cast<uint32_t>(ch * G);
        if constexpr (G == 4) ptx::tmem_ld32(ta, r);
        else if constexpr (G == 2) ptx::tmem_ld16(ta, r);
        else if constexpr (G == 1) ptx::tmem_ld8(ta, r);
        else { ptx::tmem_ld8(ta, r); ptx::tmem_ld8(ta + 8, r + 8); ptx::tmem_ld8(ta + 16, r + 16); }
      };
      if constexpr (FAST) {
        // ---- lean path: kind known at compile time, offsets hoisted, bias from shared memory ----
        EpiTile et;
        if (valid) et = epi_tile(P.L.epi, b, y, x, P.L.H, P.L.W);
        const float* bias = s_bias;
        if (valid && ch_base + half * 8 < C) epi_tc_prefetch<KIND>(P.L.epi, et, ch_base + half * 8, ops0);
#ifdef VPK_TRACE
        const long long e_t0 = clock64();
#endif
        ptx::mbar_wait_fast(tfull + 8 * acc, acc_phase);
        ptx::tc_fence_after();
#ifdef VPK_TRACE
        e_t1 = clock64();
        e_wait += e_t1 - e_t0;
#endif
        const bool ln_stats = (KIND == EPI_BIAS_ACT) && P.L.epi.gn_sums != nullptr && P.L.epi.gn_group_size < 0;
        float ln_s = 0.f, ln_q = 0.f;
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
            if constexpr (KIND == EPI_BIAS_ACT) {
              if (ln_stats) {          // whole-sample statistics (a following LayerNorm): the values just stored
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  ln_s += a[0][j];
                  ln_q = fmaf(a[0][j], a[0][j], ln_q);
                }
              }
            }
          }
        };
        bool stats = false;
        if constexpr (KIND == EPI_BIAS_ACT) stats = P.L.epi.gn_sums != nullptr && P.L.epi.gn_group_size > 0;
        if constexpr (KIND == EPI_SUBPIX) {
          // sub-pixel transposed conv: the four gates of a channel are the four output pixels of this input position
          stats = false;
          const EpiParams& E = P.L.epi;
          for (int ch = half * 8; ch < Cn; ch += 16) {
            uint32_t r[8 * G];
            tmem_chunk(ch, r);
            ptx::tmem_ld_wait();
            if (valid && ch_base + ch < C) {
              const float4* bp = reinterpret_cast<const float4*>(bias + (ch_base + ch) * 4);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 bv = bp[j];
                  const float bj = g == 0 ? bv.x : g == 1 ? bv.y : g == 2 ? bv.z : bv.w;
                  float t = __uint_as_float(r[j * 4 + g]) + bj;
                  if (E.act == ACT_LEAKY) t = t > 0.f ? t : 0.2f * t;
                  else if (E.act == ACT_RELU) t = fmaxf(t, 0.f);
                  v[j] = t;
                }
                st_bf16x8(static_cast<bf16*>(E.out) + et.out_off + (g >> 1) * E.ps_row + (g & 1) * C + ch_base + ch, v);
              }
            }
          }
        } else
        if constexpr (KIND == EPI_DECOUPLE) {
          // PredRNN-V2 decoupling loss: acc = (adapter(delta_c), adapter(delta_m)) of 8 channels at this position; the
          // warp's 32 positions are reduced to dot / |c|^2 / |m|^2 per channel (32-value butterfly) and stored to this
          // warp's slot -- the adapter outputs never reach memory
          stats = false;
          const EpiParams& E = P.L.epi;
          const int slot = E.gn_slot0 + (mt % (P.tiles_x * P.tiles_y)) * 4 + quad;
          for (int ch = half * 8; ch < Cn; ch += 16) {
            uint32_t r[8 * G];
            tmem_chunk(ch, r);
            ptx::tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float ac = valid ? __uint_as_float(r[2 * j]) : 0.f, am = valid ? __uint_as_float(r[2 * j + 1]) : 0.f;
              v[j] = ac * am;
              v[8 + j] = ac * ac;
              v[16 + j] = am * am;
              v[24 + j] = 0.f;
            }
            const float tot = warp_reduce32(v, lane);
            if (lane < 24 && b < P.L.B && ch_base + ch < C)
              E.s1[((static_cast<long long>(b) * E.gn_nslots + slot) * C + ch_base + ch + (lane & 7)) * 3 + (lane >> 3)] = tot;
          }
        } else
        if (stats) {
          // conv feeding a GroupNorm: per-group sum / sum of squares of the stored values, reduced over the 32 positions
          // of the warp and stored to this warp's slot of gn_sums (the plan guarantees Cn <= 64, Cn / gs <= 16, no
          // residual); positions outside the image contribute zeros
          if constexpr (KIND == EPI_BIAS_ACT) {
            float gs16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) gs16[i] = 0.f;
            const int gsz = P.L.epi.gn_group_size;
            uint32_t r4[4][8 * G];       // all (up to four) chunks of this warp are read behind ONE tcgen05.wait::ld
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (half * 8 + 16 * k < Cn) tmem_chunk(half * 8 + 16 * k, r4[k]);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int ch = half * 8 + 16 * k;
              if (ch < Cn) {
                const uint32_t* r = r4[k];
                float a[G][8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[0][j] = 0.f;
                if (valid && ch_base + ch < C) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) a[0][j] = __uint_as_float(r[j]);
                  epi_tc_finish<KIND, G>(P.L.epi, et, ch_base + ch, bias, a, ops0);
                }
                if (gsz == 2) gn_accum<2>(gs16, k, a[0]);
                else if (gsz == 4) gn_accum<4>(gs16, k, a[0]);
                else gn_accum<8>(gs16, k, a[0]);
              }
            }
            const float tot = gn_warp_reduce16(gs16, lane);
            const int vi = gn_lane_value(lane), pairi = vi >> 1, gpc = 8 / gsz;
            const int kk = pairi / gpc, jj = pairi - kk * gpc;
            const int chg = ch_base + half * 8 + 16 * kk;       // first channel of that chunk
            if ((lane & 1) == 0 && b < P.L.B && half * 8 + 16 * kk < Cn && chg < C) {
              const int slot = P.L.epi.gn_slot0 + (mt % (P.tiles_x * P.tiles_y)) * 4 + quad;
              P.L.epi.gn_sums[((static_cast<long long>(b) * P.L.epi.gn_nslots + slot) * (C / gsz) + chg / gsz + jj) * 2 +
                              (vi & 1)] = tot;
            }
          }
        } else {
          if constexpr (KIND == EPI_BIAS_ACT) {
            // bias + activation (+ residual / whole-sample statistics): nothing has to be fetched from memory first, so the
            // only latency per chunk is the TMEM read itself (~0.5 k cycles; six serial reads made the N = 96 deconv
            // epilogue slower than its MMAs).  Up to four chunks are read per tcgen05.wait::ld.
            const int nj = (Cn - half * 8 + 15) / 16;          // this warp's chunks: channels half * 8 + 16 j
            for (int j0 = 0; j0 < nj; j0 += 4) {
              uint32_t r4[4][8];
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (j0 + q < nj) ptx::tmem_ld8(taddr + static_cast<uint32_t>(half * 8 + 16 * (j0 + q)), r4[q]);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int ch = half * 8 + 16 * (j0 + q);
                if (j0 + q < nj && valid && ch_base + ch < C) {
                  if (P.L.epi.res != nullptr) epi_tc_prefetch<KIND>(P.L.epi, et, ch_base + ch, ops0);
                  float a[G][8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) a[0][j] = __uint_as_float(r4[q][j]);
                  epi_tc_finish<KIND, G>(P.L.epi, et, ch_base + ch, bias, a, ops0);
                  if (ln_stats) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      ln_s += a[0][j];
                      ln_q = fmaf(a[0][j], a[0][j], ln_q);
                    }
                  }
                }
              }
            }
          } else {
          for (int ch = half * 8; ch < Cn; ch += 32) {
            do_chunk(ch, ops0, ops1);
            if (ch + 16 < Cn) do_chunk(ch + 16, ops1, ops0);
          }
          }
          if (ln_stats) {      // one (sum, sum of squares) pair per warp and tile, in this warp's own slot: deterministic
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
              ln_s += __shfl_xor_sync(0xffffffffu, ln_s, o);
              ln_q += __shfl_xor_sync(0xffffffffu, ln_q, o);
            }
            if (lane == 0 && b < P.L.B) {
              const int slot = P.L.epi.gn_slot0 + ((mt % (P.tiles_x * P.tiles_y)) * P.n_tiles + nt) * 8 + quad * 2 + half;
              float* o = P.L.epi.gn_sums + (static_cast<long long>(b) * P.L.epi.gn_nslots + slot) * 2;
              o[0] = ln_s;
              o[1] = ln_q;
            }
          }
        }
      } else {
        // ---- generic path (ragged channel counts, channel-strided outputs) ----
        if (valid && half * 8 < Cn) epilogue_prefetch<bf16, G, 8>(P.L.epi, b, y, x, P.L.H, P.L.W, ch_base + half * 8, ops0);
        ptx::mbar_wait_fast(tfull + 8 * acc, acc_phase);
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
      __syncwarp();              // one arrival per warp (512 per-thread remote arrivals per tile serialised on the barrier)
      if (lane == 0) {
        if constexpr (PAIR) ptx::mbar_arrive_cluster(tempty + 8 * acc, lead);
        else ptx::mbar_arrive(tempty + 8 * acc);
      }
#ifdef VPK_TRACE
      if (e_t1) e_body += clock64() - e_t1;
#endif
    }
#ifdef VPK_TRACE
    if (trace && warp == 3 && lane == 0 && e_body)
      printf("halo trace N=%d taps=%d: epilogue warp, %d tiles: waiting for the accumulator %lld cycles, TMEM reads %lld, working %lld "
             "cycles (lean %d)\n", tileN, P.ntaps, iter, e_wait, e_ld, e_body, P.lean);
#endif
  }

  if (threadIdx.x == kHaloThreads - 1) stamp(6);      // last epilogue warp has finished its tiles
  ptx::tc_fence_before();
  if constexpr (PAIR) ptx::cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(P.tmem_cols));
    else ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(P.tmem_cols));
  }
#ifdef VPK_TRACE
  if (trace && threadIdx.x == 0) {
    stamp(7);
    const unsigned long long t0 = s_trace[0];
    printf("halo trace N=%d taps=%d tiles/cta=%d: setup %llu  pdl_wait %llu  first_A %llu  tile0_issued %llu  all_issued %llu  "
           "epi_done %llu  exit %llu (ns since entry)\n",
           P.tileN, P.ntaps, (total + nunits - 1) / nunits, s_trace[1] - t0, s_trace[2] - t0, s_trace[3] - t0,
           s_trace[4] - t0, s_trace[5] - t0, s_trace[6] - t0, s_trace[7] - t0);
  }
#endif
#endif
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) VPK_THROW(2, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}
void encode(CUtensorMap* map, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
            const cuuint32_t* box, const char* what, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                           const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, static_cast<int>(r));
    VPK_THROW(2, buf);
  }
}
int pow2_at_least(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int KIND, bool PAIR, int MODE> void launch_one(const HaloPlan& P, cudaStream_t stream) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(conv_halo_kernel<KIND, PAIR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(kMaxSmem));
  });
  cudaLaunchConfig_t cfg{};
  int grid = P.grid;
  if (PAIR && P.mc) {
    // four-CTA clusters must sit inside one GPC: ask how many fit at once and keep the persistent grid to that
    static int max_clusters = -1;
    static unsigned max_for_smem = 0;
    if (max_clusters < 0 || P.smem_bytes > max_for_smem) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(static_cast<unsigned>(P.grid));
      q.blockDim = dim3(kHaloThreads);
      q.dynamicSmemBytes = P.smem_bytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 4;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      int n = 0;
      VPK_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_halo_kernel<KIND, PAIR, MODE>, &q));
      max_clusters = std::max(1, n);
      max_for_smem = P.smem_bytes;
      if (getenv("VPK_VERBOSE")) fprintf(stderr, "conv_halo: %d clusters of 4 CTAs fit (%u bytes of shared memory)\n", n, P.smem_bytes);
    }
    grid = std::min(grid, 4 * max_clusters);
  }
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = P.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? (P.mc ? 4 : 2) : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VPK_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<KIND, PAIR, MODE>, P));
}

}  // namespace

bool halo_eligible(const ConvLaunch& L, int dtype, int radius, int nblocks, int ntaps) {
  if (!tc_eligible(L, dtype)) return false;
  if (L.epi.proj_n > 0 && (L.epi.kind != EPI_BIAS_ACT || L.epi.proj_n > 4 || L.N_pad != L.Cn || L.Cn > 256)) return false;
  return radius >= 0 && radius <= 3 && nblocks >= 1 && nblocks <= 64 && ntaps <= kMaxSteps;
}

bool halo_will_pair(int B, int H, int W, int tileN, int n_tiles, int num_sms) {
  const long long m_tiles = static_cast<long long>(B) * ((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH);
  bool pair = tileN % 16 == 0 && m_tiles * n_tiles >= num_sms;
  if (const char* env = getenv("VPK_TC_PAIR")) pair = atoi(env) != 0 && tileN % 16 == 0;
  return pair;
}

void halo_make_plan(const ConvLaunch& L, const HaloBlock* d_blocks, const HaloTap* d_taps, int nblocks, int ntaps,
                    int radius, HaloPlan* plan, int num_sms, unsigned reserve_smem) {
  HaloPlan& P = *plan;
  P.L = L;
  P.seq_T = 0;
  P.seq_c_bytes = 0;
  P.seq_barrier = nullptr;
  P.blocks = d_blocks;
  P.taps = d_taps;
  P.nblocks = nblocks;
  P.ntaps = ntaps;
  P.P = radius;
  P.tileN = L.Cn * L.G;
  P.n_tiles = L.N_pad / P.tileN;
  P.tiles_x = (L.W + kTW - 1) / kTW;
  P.tiles_y = (L.H + kTH - 1) / kTH;
  const long long m_tiles = static_cast<long long>(L.B) * P.tiles_x * P.tiles_y;
  P.pair = halo_will_pair(L.B, L.H, L.W, P.tileN, P.n_tiles, num_sms) ? 1 : 0;
  VPK_REQUIRE(L.region_g0 == 0 || P.pair, "conv_halo: accumulator regions need a CTA-pair launch");
  P.debug = 0;
  if (const char* env = dev_env("VPK_TC_DEBUG")) P.debug = atoi(env);
  P.L.epi.debug = P.debug;
  P.pack16 = 1;
  if (const char* env = getenv("VPK_HALO_PACK16")) P.pack16 = atoi(env) != 0 ? 1 : 0;
  P.fast_epi = (epi_tc_fast_ok(L.epi) && gates_of(L.epi.kind) == L.G) ? 1 : 0;
  if (const char* env = getenv("VPK_TC_FAST_EPI")) P.fast_epi = P.fast_epi && atoi(env) != 0;
  VPK_REQUIRE(L.region_g0 == 0 || (P.fast_epi && (L.epi.kind == EPI_ST_C || L.epi.kind == EPI_ST_O)),
              "conv_halo: accumulator regions need the specialised epilogue of an ST-LSTM kind");
  P.roll = (P.fast_epi && L.epi.kind == EPI_LSTM && (L.epi.pp16 != nullptr || L.epi.p0 == nullptr)) ? 1 : 0;
  if (const char* env = getenv("VPK_EPI_ROLL")) P.roll = P.roll && atoi(env) != 0;
  {
    const int cn = P.tileN / L.G, nj = cn / 16;
    const bool gn = L.epi.gn_sums != nullptr;
    const int gs = L.epi.gn_group_size;
    const unsigned long long tpi = static_cast<unsigned long long>(P.tiles_x) * P.tiles_y;
    // exact floor(n / d) = umulhi(n, ceil(2^32 / d)) needs n * d < 2^32 for every n the epilogue divides
    const unsigned long long nmax = static_cast<unsigned long long>(m_tiles + 4) * P.n_tiles;
    const bool magic_ok = nmax * std::max<unsigned long long>(tpi, static_cast<unsigned long long>(P.n_tiles)) < (1ull << 32);
    auto magic = [](unsigned long long d) { return static_cast<unsigned>(((1ull << 32) + d - 1) / d); };
    P.magic_nt = P.n_tiles > 1 ? magic(static_cast<unsigned long long>(P.n_tiles)) : 0u;
    P.magic_tpi = tpi > 1 ? magic(tpi) : 0u;
    P.magic_tx = P.tiles_x > 1 ? magic(static_cast<unsigned long long>(P.tiles_x)) : 0u;
    P.lean = (P.fast_epi && L.epi.kind == EPI_BIAS_ACT && L.G == 1 && L.epi.proj_n == 0 && L.epi.res == nullptr &&
              !L.epi.out_f32 && cn % 16 == 0 && L.N_pad == L.epi.C && magic_ok &&
              (gn ? (gs > 0 && nj <= 4 && L.epi.act != ACT_RELU) : (nj == 1 || nj == 2 || nj == 3 || nj == 4 || nj == 6)) &&
              (L.epi.act == ACT_NONE || L.epi.act == ACT_LEAKY || L.epi.act == ACT_RELU)) ? 1 : 0;
    if (const char* env = getenv("VPK_EPI_LEAN")) P.lean = P.lean && (atoi(env) != 0 || L.epi.out_f16);
    VPK_REQUIRE(!L.epi.out_f16 || P.lean, "halo plan: fp16 outputs need the lean BIAS_ACT epilogue");
    // staged bulk-tensor stores: 16 / 32 / 64 / 96 channels (whole 32-channel sub-tiles), 16-byte aligned strides
    P.lean_tma = (P.lean && reserve_smem == 0 && (nj == 1 || nj == 2 || nj == 4 || nj == 6) && L.epi.oC == 1 &&
                  L.epi.oX % 8 == 0 && L.epi.oY % 8 == 0 && L.epi.oB % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(L.epi.out) % 16 == 0) ? 1 : 0;
    if (const char* env = getenv("VPK_EPI_TMA")) P.lean_tma = P.lean_tma && atoi(env) != 0;
    P.stage_bytes = P.lean_tma ? 2u * 128u * static_cast<unsigned>(cn) * 2u : 0u;
    reserve_smem += P.stage_bytes;
  }
  if (L.epi.gn_sums != nullptr && L.epi.gn_group_size < 0) {
    VPK_REQUIRE(P.fast_epi && L.epi.kind == EPI_BIAS_ACT && L.epi.proj_n == 0 && L.epi.res == nullptr,
                "halo plan: fused LayerNorm statistics need the lean BIAS_ACT epilogue");
  } else if (L.epi.gn_sums != nullptr) {
    const int gs = L.epi.gn_group_size;
    VPK_REQUIRE(P.fast_epi && L.epi.kind == EPI_BIAS_ACT && L.epi.proj_n == 0 && L.epi.res == nullptr &&
                    (gs == 2 || gs == 4 || gs == 8) && L.Cn <= 64 && L.Cn / gs <= 16 && L.epi.C % gs == 0,
                "halo plan: fused GroupNorm statistics need the lean BIAS_ACT epilogue, <= 64 channels per tile and groups of 2/4/8");
  }
  P.tmem_cols = std::max(32, pow2_at_least(2 * P.tileN));
  VPK_REQUIRE(P.tmem_cols <= 512, "halo plan: accumulators exceed TMEM");
  const int HWp = kTW + 2 * radius, HHp = kTH + 2 * radius;
  const unsigned a_box = static_cast<unsigned>(HWp * HHp * 128);
  P.a_slot_bytes = (a_box + 1023u) / 1024u * 1024u;
  const unsigned b_rows = static_cast<unsigned>(P.pair ? P.tileN / 2 : P.tileN);
  P.b_tap_stride = (b_rows * 128u + 1023u) / 1024u * 1024u;
  // several taps share one weight slot (one barrier round trip) when the tiles are small: per-tap barrier latency,
  // not bytes, bounds the small-N layers
  // (measured, cfg 5: two taps per 16 KB-tile slot instead of one = -8 % on the N = 256 gate GEMM, -12 % at N = 192)
  // (40 KB slots: three 12 KB taps at N = 192 -- forecaster.rnn2 760 -> 719 us, encoder.rnn2 589 -> 573 us -- two 16 KB taps at 256)
  P.bgroup = std::max(1, std::min<int>(9, static_cast<int>(40960u / P.b_tap_stride)));
  if (const char* env = getenv("VPK_HALO_BGROUP")) P.bgroup = std::max(1, atoi(env));
  P.b_slot_bytes = P.bgroup * P.b_tap_stride;
  // Resident weights: with one N tile and <= 96 KB of weight tiles the ring (and its per-group wait + commit in the
  // MMA thread, which bounds these small-N layers: ~250 ns per tap whatever N is) disappears.
  P.resident = (P.n_tiles == 1 && static_cast<unsigned>(ntaps) * P.b_tap_stride <= 96u * 1024u) ? 1 : 0;
  if (const char* env = getenv("VPK_HALO_RESIDENT")) P.resident = P.resident && atoi(env) != 0;
  if (P.resident) {
    P.bgroup = ntaps;
    P.b_slot_bytes = static_cast<unsigned>(ntaps) * P.b_tap_stride;
  }
  const unsigned fixed = 1024 + 1024 + static_cast<unsigned>(nblocks) * sizeof(HaloBlock) +
                         static_cast<unsigned>(ntaps) * (sizeof(HaloTap) + 8) + static_cast<unsigned>(L.N_pad) * 4 + 176 +
                         (L.epi.proj_n > 0 ? static_cast<unsigned>(L.N_pad) * 16 + 16 : 0u);
  // Ring depths by bytes: ~60 % of shared memory for weight tiles (4..24 slots), the rest for activation halo tiles
  // (2..8).  What matters is the number of TMA operations in flight against their ~2 us latency: small-N layers have
  // small weight tiles and get deep rings, the N = 256 gate GEMMs get 7-8 x 16 KB.
  // Experiment kept behind VPK_HALO_SMEM_CAP=<bytes> (default off): capping short low-register launches at half of the
  // SM's shared memory lets the next kernel's CTAs become resident under PDL and overlap their set-up -- measured on
  // cfg 4 / cfg 2 it is SLOWER (10.37 vs 9.70 ms, 10.45 vs 10.02 ms): the shallower activation ring costs more than the
  // hidden set-up saves.
  if (P.lean_tma && !P.resident) {
    // the staging buffers come out of the ring budget: streamed-weight layers that are bound by their activation ring
    // (stride-2 stage convs: one TMA box per tap and output row) keep the ring instead -- an activation ring of two
    // slots cost more than the stores saved (cfg 5 encoder.stage3: 83 -> 100 us)
    auto depth_a = [&](unsigned res) {
      const unsigned av = kMaxSmem - res - fixed;
      int b = std::max(4, std::min(24, static_cast<int>(av * 6 / 10 / P.b_slot_bytes)));
      while (b > 2 && b * P.b_slot_bytes + 2 * P.a_slot_bytes > av) --b;
      return static_cast<int>((av - b * P.b_slot_bytes) / P.a_slot_bytes);
    };
    if (depth_a(reserve_smem) < 3 && depth_a(reserve_smem - P.stage_bytes) >= 3) {
      reserve_smem -= P.stage_bytes;
      P.stage_bytes = 0;
      P.lean_tma = 0;
    }
  }
  VPK_REQUIRE(reserve_smem % 1024 == 0 && reserve_smem + 96 * 1024 <= kMaxSmem, "halo plan: shared-memory reserve too large");
  unsigned budget = kMaxSmem - reserve_smem;
  if (const char* env = getenv("VPK_HALO_SMEM_CAP")) {
    const unsigned cap = static_cast<unsigned>(atoi(env));
    const bool low_reg = (L.epi.kind == EPI_BIAS_ACT && L.epi.proj_n == 0) || L.epi.kind == EPI_PHY_GATE;
    if (cap > 0 && cap <= kMaxSmem && low_reg && P.fast_epi && m_tiles * P.n_tiles <= 8ll * num_sms &&
        fixed + (P.resident ? 1u : 2u) * P.b_slot_bytes + 2u * P.a_slot_bytes <= cap)
      budget = cap;
  }
  const unsigned avail = budget - fixed;
  int sb = static_cast<int>(avail * 6 / 10 / P.b_slot_bytes);
  sb = std::max(4, std::min(24, sb));
  while (sb > 2 && sb * P.b_slot_bytes + 2 * P.a_slot_bytes > avail) --sb;
  if (P.resident) sb = 1;
  if (reserve_smem > P.stage_bytes && !((sb >= 2 || P.resident) && sb * P.b_slot_bytes + 2 * P.a_slot_bytes <= avail)) {
    P.SA = P.SB = 0;          // sequence plan that does not fit: the caller falls back to per-step launches
    P.smem_bytes = 0;
    return;
  }
  VPK_REQUIRE((sb >= 2 || P.resident) && sb * P.b_slot_bytes + 2 * P.a_slot_bytes <= avail,
              "halo plan: shared memory budget exceeded");
  P.SB = sb;
  P.SA = std::max(2, std::min<int>(8, static_cast<int>((avail - P.SB * P.b_slot_bytes) / P.a_slot_bytes)));
  if (const char* env = getenv("VPK_HALO_SA")) {
    const int sa = atoi(env);
    if (!P.resident && sa >= 2 && fixed + sa * P.a_slot_bytes + 2 * P.b_slot_bytes <= budget) {
      P.SA = sa;
      P.SB = std::min<int>(24, static_cast<int>((avail - P.SA * P.a_slot_bytes) / P.b_slot_bytes));
    }
  }
  P.smem_bytes = fixed + P.SA * P.a_slot_bytes + P.SB * P.b_slot_bytes + reserve_smem;
  // Streamed weights are the largest L2 -> SM stream of the gate GEMMs (27 taps x 32 KB per 256 x 256 tile against
  // 69 KB of activations; the MMA thread waits for weight tiles 25-34 % of its time, phase timeline in profiles/).
  // VPK_HALO_MC=1 (size rule) / 2 (always) runs clusters of four CTAs that fetch every weight tile once for two CTA
  // pairs (TMA multicast).  Measured on the B200: +6.7 % per SM on the N = 256 gate GEMM, but only 33 four-CTA
  // clusters are co-resident (132 of 148 SMs; a cluster must sit inside one GPC), so the whole launch is 4.5 % SLOWER
  // (cfg 5: 285.7 vs 277.0 ms) -- off by default; the pair kernel keeps all 148 SMs busy.
  P.mc = 0;
  if (const char* env = getenv("VPK_HALO_MC")) {
    const int v = atoi(env);
    P.mc = (P.pair && !P.resident && (v > 1 || (v == 1 && (m_tiles + 1) / 2 * P.n_tiles >= 4ll * num_sms))) ? 1 : 0;
  }
  if (P.mc) {
    const long long cl_units = ((m_tiles + 1) / 2 + 1) / 2 * P.n_tiles;
    P.grid = 4 * static_cast<int>(std::min<long long>(cl_units, num_sms / 4));
  } else
  if (P.pair) {
    const long long units = (m_tiles + 1) / 2 * P.n_tiles;
    P.grid = 2 * static_cast<int>(std::min<long long>(units, num_sms / 2));
  } else {
    P.grid = static_cast<int>(std::min<long long>(m_tiles * P.n_tiles, num_sms));
  }
  for (int i = 0; i < L.nsrc; ++i) {
    const SrcView& s = L.src[i];
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(s.C), static_cast<cuuint64_t>(s.W), static_cast<cuuint64_t>(s.H),
                          static_cast<cuuint64_t>(L.B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(s.sX) * 2, static_cast<cuuint64_t>(s.sY) * 2,
                             static_cast<cuuint64_t>(s.sB) * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(HWp), static_cast<cuuint32_t>(HHp), 1};
    encode(&P.amap[i], 4, s.base, dims, strides, box, "activation halo view");
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(L.K_pad), static_cast<cuuint64_t>(L.N_pad)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(L.K_pad) * 2};
  cuuint32_t box[2] = {64, b_rows};
  encode(&P.bmap, 2, L.wpacked, dims, strides, box, "packed weights");
  if (P.lean_tma) {
    const int cn = P.tileN / L.G;
    cuuint64_t od[4] = {static_cast<cuuint64_t>(L.epi.C), static_cast<cuuint64_t>(L.W), static_cast<cuuint64_t>(L.H),
                        static_cast<cuuint64_t>(L.B)};
    cuuint64_t os[3] = {static_cast<cuuint64_t>(L.epi.oX) * 2, static_cast<cuuint64_t>(L.epi.oY) * 2,
                        static_cast<cuuint64_t>(L.epi.oB) * 2};
    cuuint32_t ob[4] = {static_cast<cuuint32_t>(std::min(cn, 32)), static_cast<cuuint32_t>(kTW), static_cast<cuuint32_t>(kTH), 1};
    encode(&P.omap, 4, L.epi.out, od, os, ob, "staged output tile",
           cn == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
  }
}

namespace {
template <bool PAIR> int max_coresident_ctas(unsigned smem_bytes, int grid) {
  cudaFuncSetAttribute(conv_halo_kernel<EPI_LSTM, PAIR, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMaxSmem));
  cudaLaunchConfig_t q{};
  q.gridDim = dim3(static_cast<unsigned>(grid));
  q.blockDim = dim3(kHaloThreads);
  q.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute qa[1];
  qa[0].id = cudaLaunchAttributeClusterDimension;
  qa[0].val.clusterDim.x = PAIR ? 2 : 1;
  qa[0].val.clusterDim.y = 1;
  qa[0].val.clusterDim.z = 1;
  q.attrs = qa;
  q.numAttrs = 1;
  int n = 0;
  VPK_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_halo_kernel<EPI_LSTM, PAIR, 4>, &q));
  return n * (PAIR ? 2 : 1);
}
}  // namespace

bool halo_make_seq_plan(const ConvLaunch& L, const HaloBlock* d_blocks, const HaloTap* d_taps, int nblocks, int ntaps,
                        int radius, const HaloSeqSpec& seq, HaloPlan* plan, int num_sms) {
  HaloPlan& P = *plan;
  VPK_REQUIRE(L.epi.kind == EPI_LSTM && L.G == 4 && seq.T >= 1 && seq.barrier != nullptr, "sequence plan: ConvLSTM layers only");
  // pass 1: tiling / pairing / grid of one timestep (all of shared memory); pass 2: the same with the cell-state reserve
  halo_make_plan(L, d_blocks, d_taps, nblocks, ntaps, radius, &P, num_sms);
  if (!(P.fast_epi && (L.epi.pp16 != nullptr || L.epi.p0 == nullptr)) || P.mc) return false;
  const long long m_tiles = static_cast<long long>(L.B) * P.tiles_x * P.tiles_y;
  const long long total = (P.pair ? (m_tiles + 1) / 2 : m_tiles) * P.n_tiles;
  auto reserve_for = [&](int grid) {
    const int nunits = grid / (P.pair ? 2 : 1);
    const int slots = static_cast<int>((total + nunits - 1) / nunits);
    const unsigned bytes = static_cast<unsigned>(slots) * (L.Cn / 8) * 128 * 8 * sizeof(float);
    return std::make_pair(slots, (bytes + 1023u) / 1024u * 1024u);
  };
  int grid = P.grid;
  for (int attempt = 0; attempt < 3; ++attempt) {
    const auto rs = reserve_for(grid);
    if (rs.second + 96u * 1024u > kMaxSmem) return false;       // leave at least ~96 KB to the operand rings
    const int pair = P.pair;
    halo_make_plan(L, d_blocks, d_taps, nblocks, ntaps, radius, &P, num_sms, rs.second);
    if (P.smem_bytes == 0 || P.pair != pair) return false;
    const int fit = pair ? max_coresident_ctas<true>(P.smem_bytes, grid) : max_coresident_ctas<false>(P.smem_bytes, grid);
    if (fit <= 0) return false;
    if (fit >= grid) {
      P.grid = grid;
      P.seq_slots = rs.first;
      P.seq_c_bytes = rs.second;
      break;
    }
    grid = fit;                 // fewer CTAs can be co-resident than the grid wants: more tiles (state slots) per CTA
    if (attempt == 2) return false;
  }
  P.roll = 0;
  P.seq_T = seq.T;
  for (int i = 0; i < kMaxSrc; ++i) {
    P.seq_sb[i] = seq.sb[i];
    P.seq_st[i] = seq.st[i];
    P.seq_off[i] = seq.off[i];
  }
  P.seq_recur_src = seq.recur_src;
  P.seq_h0_src = seq.h0_src;
  P.seq_out_sb = seq.out_sb;
  P.seq_out_st = seq.out_st;
  P.seq_out_off = seq.out_off;
  P.seq_c_zero = seq.c_zero;
  P.seq_barrier = seq.barrier;
  int oob = 0;
  const int HWp = kTW + 2 * radius, HHp = kTH + 2 * radius;
  for (int i = 0; i < L.nsrc; ++i) {       // tensor maps over the WHOLE sequence buffers (4th dimension = samples)
    const SrcView& sv = L.src[i];
    oob = std::max(oob, seq.samples[i]);
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(sv.C), static_cast<cuuint64_t>(sv.W), static_cast<cuuint64_t>(sv.H),
                          static_cast<cuuint64_t>(seq.samples[i])};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(sv.sX) * 2, static_cast<cuuint64_t>(sv.sY) * 2,
                             static_cast<cuuint64_t>(sv.sB) * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(HWp), static_cast<cuuint32_t>(HHp), 1};
    encode(&P.amap[i], 4, sv.base, dims, strides, box, "activation sequence view");
  }
  P.seq_oob = oob + 1;
  return true;
}

void launch_conv_halo(const HaloPlan& P, cudaStream_t stream) {
  if (P.seq_T > 0) {
    VPK_REQUIRE(P.L.epi.kind == EPI_LSTM, "conv_halo: sequence mode is a ConvLSTM mode");
    if (P.pair) launch_one<EPI_LSTM, true, 4>(P, stream);
    else launch_one<EPI_LSTM, false, 4>(P, stream);
    return;
  }
#define VPK_HALO(KIND)                                                                   \
  if (P.pair) {                                                                          \
    if (P.fast_epi) launch_one<KIND, true, 1>(P, stream);                                \
    else launch_one<KIND, true, 0>(P, stream);                                           \
  } else {                                                                               \
    if (P.fast_epi) launch_one<KIND, false, 1>(P, stream);                               \
    else launch_one<KIND, false, 0>(P, stream);                                          \
  }                                                                                      \
  break
  if (P.L.epi.kind == EPI_BIAS_ACT && P.L.epi.proj_n > 0) {
    if (P.pair) launch_one<EPI_BIAS_ACT, true, 3>(P, stream);
    else launch_one<EPI_BIAS_ACT, false, 3>(P, stream);
    return;
  }
  if (P.L.epi.kind == EPI_LSTM && P.roll) {
    if (P.pair) launch_one<EPI_LSTM, true, 2>(P, stream);
    else launch_one<EPI_LSTM, false, 2>(P, stream);
    return;
  }
  switch (P.L.epi.kind) {
    case EPI_BIAS_ACT: VPK_HALO(EPI_BIAS_ACT);
    case EPI_LSTM: VPK_HALO(EPI_LSTM);
    case EPI_ST_C: VPK_HALO(EPI_ST_C);
    case EPI_ST_M: VPK_HALO(EPI_ST_M);
    case EPI_ST_O: VPK_HALO(EPI_ST_O);
    case EPI_ST_O1: VPK_HALO(EPI_ST_O1);
    case EPI_PHY_GATE: VPK_HALO(EPI_PHY_GATE);
    case EPI_SUBPIX:
      VPK_REQUIRE(P.fast_epi, "conv_halo: the sub-pixel epilogue needs whole 8-channel chunks");
      if (P.pair) launch_one<EPI_SUBPIX, true, 1>(P, stream);
      else launch_one<EPI_SUBPIX, false, 1>(P, stream);
      break;
    case EPI_DECOUPLE:
      VPK_REQUIRE(P.fast_epi, "conv_halo: the decoupling-loss epilogue needs whole 8-channel chunks");
      if (P.pair) launch_one<EPI_DECOUPLE, true, 1>(P, stream);
      else launch_one<EPI_DECOUPLE, false, 1>(P, stream);
      break;
    default: VPK_THROW(1, "conv_halo: unsupported epilogue kind");
  }
#undef VPK_HALO
}

}  // namespace vpk
