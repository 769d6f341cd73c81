// tcgen05 / TMEM / TMA implicit-GEMM kernel for the generalised convolution launch: host-side plan.
#pragma once
#include "common.h"

namespace vpk {

struct alignas(64) TcPlan {
  CUtensorMap amap[kMaxSrc];   // 4-D (C, W, H, B) views of the activation sources, box (64, TW, TH, TB), 128B swizzle
  CUtensorMap bmap;            // 2-D (K_pad, N_pad) packed weights, box (64, tileN), 128B swizzle
  ConvLaunch L;
  int TW, TH, TB;              // output positions per tile: TW*TH*TB = 128
  int tiles_x, tiles_y, tiles_b;
  int n_tiles;                 // tiles along N
  int tileN;                   // Cn * G, multiple of 16, <= 256
  int stages;                  // smem ring depth
  int tmem_cols;               // power of two >= 2 * tileN
  unsigned smem_bytes;
  int grid;
  int debug;                   // VPK_TC_DEBUG bit mask (perf experiments only; results are wrong when set):
                               //   1 = skip epilogue math/IO, 2 = skip MMA issue, 4 = skip TMA loads
  int fast_epi;                // lean compile-time-specialised epilogue usable (CTA-pair kernel)
  int cta2;                    // 1: CTA-pair kernel (conv_tc2.cu): weight box holds tileN/2 rows, grid is even
};

// ---- halo-reuse kernel (conv_halo.cu) ----------------------------------------------------------------------------
struct HaloBlock {             // one (source, 64-channel block): its activation halo tile is loaded once
  short src, c0, kc, ntaps;
  int first_tap;
};
struct HaloTap {               // one tap of a block: shifted view of the halo tile x one weight tile
  signed char dy, dx;
  short nk;                    // low byte: K=16 MMA slices (ceil(kc / 16)); high byte: accumulator regions this tap feeds
                               // (1 = first, 2 = second, 0 = no regions in use; ConvLaunch::region_g0)
  int wk;                      // column offset in the packed weights
};
struct alignas(64) HaloPlan {
  CUtensorMap amap[kMaxSrc];   // 4-D (C, W, H, B) views, box (64, 8+2P, 16+2P, 1), 128B swizzle
  CUtensorMap bmap;            // 2-D packed weights, box (64, tileN or tileN/2)
  ConvLaunch L;
  const HaloBlock* blocks;     // device
  const HaloTap* taps;         // device
  int nblocks, ntaps;
  int P;                       // halo radius = max |dy|, |dx|
  int pack16;                  // taps of <= 16-channel blocks share streamed weight tiles four at a time (conv_halo.cu)
  int tiles_x, tiles_y;        // 8 x 16 output tiles per image
  int n_tiles, tileN;
  int SA, SB;                  // activation / weight ring depths
  unsigned a_slot_bytes, b_slot_bytes;
  unsigned b_tap_stride;       // bytes between the weight tiles of the taps sharing one slot
  int bgroup;                  // taps per weight slot
  int resident;                // 1: every tap's weight tile is loaded ONCE per CTA and stays in shared memory (SB = 1)
  int tmem_cols;
  unsigned smem_bytes;
  int grid;
  int pair;                    // cta_group::2
  int mc;                      // 1: clusters of FOUR CTAs = two CTA pairs that run the same weight-tile sequence on different
                               //    M tiles; every weight tile is fetched from L2 once per cluster (TMA multicast)
  int fast_epi;                // lean compile-time-specialised epilogue (epilogue_tc.cuh) usable for this launch
  int roll;                    // ConvLSTM epilogue with a whole tile of operands in flight (lstm_ops_load / lstm_finish)
  CUtensorMap omap;            // lean_tma: (C, W, H, B) view of the output, box (min(C, 32), 8, 16, 1), 64B / 32B swizzle
  int lean_tma;                // lean epilogue that stages the 16-bit tile in shared memory and writes it with bulk tensor
                               // stores (one thread, full lines) instead of 32-line st.global instructions
  unsigned stage_bytes;        // shared memory in front of the rings for two staged output tiles
  unsigned magic_nt, magic_tpi, magic_tx;   // ceil(2^32 / d) for d = n_tiles, tiles_x * tiles_y, tiles_x (lean epilogue)
  int lean;                    // bias + (leaky) ReLU + 16-bit store only: the short straight-line epilogue loop (the per-tile
                               // instruction stream of ONE warp, not bandwidth, bounds the epilogue of narrow layers)
  int debug;
  // ---- sequence mode (conv_halo_kernel MODE 4): ONE launch runs seq_T timesteps of a ConvLSTM layer, persistent CTAs,
  //      the cell state c resident in shared memory across the timesteps, a grid-wide barrier between timesteps ----
  int seq_T;                   // 0: ordinary single-step launch
  int seq_sb[kMaxSrc], seq_st[kMaxSrc], seq_off[kMaxSrc];   // sample index source s reads at (sequence b, step t): b*sb + t*st + off
  int seq_oob;                 // a sample index outside every source (zero fill for the odd tail of a CTA pair)
  int seq_recur_src;           // source that is the layer's own output sequence (step t reads what step t-1 wrote: waits for
                               // the grid barrier), or -1
  int seq_h0_src;              // source read INSTEAD of seq_recur_src at t = 0 (the initial hidden state)
  int seq_out_sb, seq_out_st, seq_out_off;   // sample index of h'_t in the output tensor (L.epi.out, sample stride epi.oB)
  int seq_c_zero;              // 1: c starts at zero; 0: c_0 is read from L.epi.s0.  c_T is always stored to L.epi.s0
  int seq_slots;               // (M unit, N tile) iterations per CTA and timestep = cell-state slots in shared memory
  unsigned seq_c_bytes;        // shared memory reserved for the cell state (in front of the rings)
  unsigned* seq_barrier;       // device counter, zeroed before every launch
};
// Sequence-mode plan of a ConvLSTM layer (see HaloPlan::seq_*).  `L` describes ONE timestep over L.B sequences; src_samples[s]
// = number of samples (4th tensor-map dimension) of source s.  Returns false when the layer does not fit (cell state of a
// CTA's tiles above the shared-memory reserve, or fewer co-resident CTAs than the grid needs).
struct HaloSeqSpec {
  int T;
  int sb[kMaxSrc], st[kMaxSrc], off[kMaxSrc], samples[kMaxSrc];
  int recur_src, h0_src;
  int out_sb, out_st, out_off;
  int c_zero;
  unsigned* barrier;
};
bool halo_make_seq_plan(const ConvLaunch& L, const HaloBlock* d_blocks, const HaloTap* d_taps, int nblocks, int ntaps,
                        int radius, const HaloSeqSpec& seq, HaloPlan* plan, int num_sms);
bool halo_eligible(const ConvLaunch& L, int dtype, int radius, int nblocks, int ntaps);
// the CTA-pair rule of halo_make_plan (shared with build_conv, which packs region-major weights for pair launches only)
bool halo_will_pair(int B, int H, int W, int tileN, int n_tiles, int num_sms);
void halo_make_plan(const ConvLaunch& L, const HaloBlock* d_blocks, const HaloTap* d_taps, int nblocks, int ntaps,
                    int radius, HaloPlan* plan, int num_sms, unsigned reserve_smem = 0);
void launch_conv_halo(const HaloPlan& plan, cudaStream_t stream);

// True when the launch can run on the tensor-core kernel (bf16, channel counts TMA-addressable, ...).
bool tc_eligible(const ConvLaunch& L, int dtype);
// Fills tensor maps and tiling for a launch whose device pointers are final.
void tc_make_plan(const ConvLaunch& L, TcPlan* plan, int num_sms);
void launch_conv_tc(const TcPlan& plan, cudaStream_t stream);   // dispatches on plan.cta2
void launch_conv_tc2(const TcPlan& plan, cudaStream_t stream);

void launch_conv_simt(const ConvLaunch& L, int dtype, cudaStream_t stream);

// Direct kernel for small, memory-bound convs (total K <= 128, N <= 32): conv_direct.cu
bool direct_eligible(const ConvLaunch& L);
void launch_conv_direct(const ConvLaunch& L, int dtype, int num_sms, cudaStream_t stream);

}  // namespace vpk
