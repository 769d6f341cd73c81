#include "lowering.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace vpk {

namespace {

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// One step per (tap, 64-channel block).  Steps are emitted BLOCK-MAJOR (all taps of one (source, channel block) are
// consecutive) so that the halo kernel can load that block's activation tile once and reuse it for every tap.
void add_step(std::vector<HostStep>& steps, int src, int dy, int dx, int C, int c0, int wref, int ky, int kx, int wc0,
              int wc_count, int wsplit = 0) {
  if (wc_count < 0) wc_count = C;
  {
    HostStep h{};
    h.s.src = static_cast<short>(src);
    h.s.dy = static_cast<signed char>(dy);
    h.s.dx = static_cast<signed char>(dx);
    h.s.c0 = static_cast<short>(c0);
    h.s.kc = static_cast<short>(std::min(64, C - c0));
    h.s.wk = 0;
    h.kw_valid = std::max(0, std::min<int>(h.s.kc, wc_count - c0));
    h.wref = wref;
    h.ky = ky;
    h.kx = kx;
    h.wc0 = wc0;
    h.wsplit = wsplit;
    steps.push_back(h);
  }
}

}  // namespace

void lower_conv(ConvSpec& spec, int k, int stride, int pad, const std::vector<ConvInput>& inputs, int in_h, int in_w,
                int esize, int* oh, int* ow) {
  VPK_REQUIRE(stride == 1 || stride == 2, "conv stride must be 1 or 2");
  const int OH = (in_h + 2 * pad - k) / stride + 1;
  const int OW = (in_w + 2 * pad - k) / stride + 1;
  if (spec.phases.empty()) spec.phases.emplace_back();
  PhaseSpec& ph = spec.phases[0];
  ph.H = OH;
  ph.W = OW;
  if (stride == 1) {
    // Split-bf16 inputs (view = high parts, lo_view = low parts) contribute three products per tap:
    // A_hi*W_hi + A_hi*W_lo (same activation block: one halo tile, 2 k*k taps) and A_lo*W_hi (second block).
    for (const ConvInput& in : inputs) {
      VPK_REQUIRE(in.view.H == in_h && in.view.W == in_w, "conv input size mismatch");
      const bool split = in.lo_view.base != nullptr || in.lo_view.C > 0;
      VPK_REQUIRE(!(split && in.w_split), "conv input: either split activations + weights or split weights only");
      const bool parts = split || in.w_split;
      const int hi = static_cast<int>(spec.srcs.size());
      spec.srcs.push_back(in.view);
      int lo = -1;
      if (split) {
        lo = static_cast<int>(spec.srcs.size());
        spec.srcs.push_back(in.lo_view);
      }
      for (int c0 = 0; c0 < in.view.C; c0 += 64) {
        for (int part = parts ? 1 : 0; part <= (parts ? 2 : 0); ++part)
          for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
              add_step(ph.steps, hi, ky - pad, kx - pad, in.view.C, c0, in.wref, ky, kx, in.wc0, in.wc_count, part);
              if ((in.w_split || in.extra_uncounted) && part == 2) ph.steps.back().count_flops = false;
            }
        if (split)
          for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx)
            {
              add_step(ph.steps, lo, ky - pad, kx - pad, in.view.C, c0, in.wref, ky, kx, in.wc0, in.wc_count, 1);
              if (in.extra_uncounted) ph.steps.back().count_flops = false;
            }
      }
    }
  } else {
    VPK_REQUIRE(inputs.size() == 1, "stride-2 conv takes a single input");
    VPK_REQUIRE(in_h % 2 == 0 && in_w % 2 == 0, "stride-2 conv needs even input size");
    const ConvInput& in = inputs[0];
    const bool split = in.lo_view.base != nullptr || in.lo_view.C > 0;
    auto parity_views = [&](const SrcView& full) {
      const int first = static_cast<int>(spec.srcs.size());
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {   // parity view: rows py, py+2, ...; columns px, px+2, ...
          SrcView v = full;
          v.H = in_h / 2;
          v.W = in_w / 2;
          v.sY = full.sY * 2;
          v.sX = full.sX * 2;
          v.base = static_cast<const char*>(full.base) + (py * full.sY + px * full.sX) * esize;
          spec.srcs.push_back(v);
        }
      return first;
    };
    const int hi_idx = parity_views(in.view);
    const int lo_idx = split ? parity_views(in.lo_view) : -1;
    auto emit = [&](int base_idx, int wsplit, int par, int c0) {
      for (int ky = 0; ky < k; ++ky) {
        const int ty = ky - pad;
        const int py = ((ty % 2) + 2) % 2;
        const int dy = floordiv(ty - py, 2);
        for (int kx = 0; kx < k; ++kx) {
          const int tx = kx - pad;
          const int px = ((tx % 2) + 2) % 2;
          const int dx = floordiv(tx - px, 2);
          if (py * 2 + px != par) continue;
          add_step(ph.steps, base_idx + par, dy, dx, in.view.C, c0, in.wref, ky, kx, in.wc0, in.wc_count, wsplit);
        }
      }
    };
    for (int par = 0; par < 4; ++par)          // block-major: parity view, channel block, then its taps
      for (int c0 = 0; c0 < in.view.C; c0 += 64) {
        if (!split) {
          emit(hi_idx, 0, par, c0);
        } else {
          emit(hi_idx, 1, par, c0);
          emit(hi_idx, 2, par, c0);
        }
      }
    if (split)
      for (int par = 0; par < 4; ++par)
        for (int c0 = 0; c0 < in.view.C; c0 += 64) emit(lo_idx, 1, par, c0);
  }
  VPK_REQUIRE(spec.srcs.size() <= static_cast<size_t>(kMaxSrc), "too many conv sources");
  *oh = OH;
  *ow = OW;
}

void lower_conv_transpose(ConvSpec& spec, int k, int stride, int pad, int out_pad, const ConvInput& input, int in_h,
                          int in_w, int* oh, int* ow,
                          const std::function<EpiParams(int, int, int, int, int)>& epi_for_phase) {
  VPK_REQUIRE(stride == 1 || stride == 2, "transposed-conv stride must be 1 or 2");
  const int OH = (in_h - 1) * stride - 2 * pad + k + out_pad;
  const int OW = (in_w - 1) * stride - 2 * pad + k + out_pad;
  const bool split = input.lo_view.base != nullptr || input.lo_view.C > 0;
  const int src = static_cast<int>(spec.srcs.size());
  spec.srcs.push_back(input.view);
  int src_lo = -1;
  if (split) {
    src_lo = static_cast<int>(spec.srcs.size());
    spec.srcs.push_back(input.lo_view);
  }
  for (int ry = 0; ry < stride; ++ry)
    for (int rx = 0; rx < stride; ++rx) {
      PhaseSpec ph;
      ph.H = (OH - ry + stride - 1) / stride;
      ph.W = (OW - rx + stride - 1) / stride;
      // y = stride*iy - pad + ky  with  y = stride*q + ry   =>   iy = q + (ry + pad - ky) / stride
      // taps in ascending (dy, dx) order = descending (ky, kx): the same order in which the sub-pixel form
      // (lower_conv_transpose_subpixel) meets them, so that both forms add a pixel's products in the same sequence and
      // agree bit for bit (which form runs depends on the problem size)
      auto emit = [&](int s_idx, int wsplit, int c0) {
        for (int ky = k - 1; ky >= 0; --ky) {
          if (((ry + pad - ky) % stride + stride) % stride != 0) continue;
          const int dy = floordiv(ry + pad - ky, stride);
          for (int kx = k - 1; kx >= 0; --kx) {
            if (((rx + pad - kx) % stride + stride) % stride != 0) continue;
            const int dx = floordiv(rx + pad - kx, stride);
            add_step(ph.steps, s_idx, dy, dx, input.view.C, c0, input.wref, ky, kx, input.wc0, input.wc_count, wsplit);
          }
        }
      };
      for (int c0 = 0; c0 < input.view.C; c0 += 64) {
        if (!split) {
          emit(src, 0, c0);
        } else {
          emit(src, 1, c0);
          emit(src, 2, c0);
        }
      }
      if (split)
        for (int c0 = 0; c0 < input.view.C; c0 += 64) emit(src_lo, 1, c0);
      ph.epi = epi_for_phase(ry, rx, stride, OH, OW);
      spec.phases.push_back(ph);
    }
  *oh = OH;
  *ow = OW;
}

void lower_conv_transpose_subpixel(ConvSpec& spec, int k, int pad, int out_pad, const ConvInput& input, int in_h, int in_w,
                                   int* oh, int* ow) {
  const int OH = (in_h - 1) * 2 - 2 * pad + k + out_pad;
  const int OW = (in_w - 1) * 2 - 2 * pad + k + out_pad;
  VPK_REQUIRE(OH == 2 * in_h && OW == 2 * in_w, "sub-pixel transposed conv needs an output of exactly twice the input size");
  VPK_REQUIRE(input.lo_view.base == nullptr && input.lo_view.C == 0, "sub-pixel transposed conv: no split operands");
  const int src = static_cast<int>(spec.srcs.size());
  spec.srcs.push_back(input.view);
  // per axis and parity r: y = 2 i - pad + kk with y = 2 q + r  =>  i = q + (r + pad - kk) / 2 for kk = r + pad (mod 2)
  auto tap_of = [&](int r, int d) {          // kernel index that parity r reads at input offset d, or -1
    const int kk = r + pad - 2 * d;
    return (kk >= 0 && kk < k) ? kk : -1;
  };
  int dmin = 0, dmax = 0;
  for (int r = 0; r < 2; ++r)
    for (int kk = 0; kk < k; ++kk)
      if (((r + pad - kk) % 2 + 2) % 2 == 0) {
        const int d = floordiv(r + pad - kk, 2);
        dmin = std::min(dmin, d);
        dmax = std::max(dmax, d);
      }
  PhaseSpec ph;
  ph.H = in_h;
  ph.W = in_w;
  for (int c0 = 0; c0 < input.view.C; c0 += 64)
    for (int dy = dmin; dy <= dmax; ++dy)
      for (int dx = dmin; dx <= dmax; ++dx) {
        add_step(ph.steps, src, dy, dx, input.view.C, c0, input.wref, 0, 0, input.wc0, input.wc_count, 0);
        HostStep& h = ph.steps.back();
        h.per_gate = true;
        bool any = false;
        for (int g = 0; g < 4; ++g) {
          const int ky = tap_of(g >> 1, dy), kx = tap_of(g & 1, dx);
          if (ky >= 0 && kx >= 0) {
            h.gky[g] = static_cast<signed char>(ky);
            h.gkx[g] = static_cast<signed char>(kx);
            any = true;
          }
        }
        if (!any) ph.steps.pop_back();
      }
  spec.phases.push_back(ph);
  *oh = OH;
  *ow = OW;
}

// ---------------------------------------------------------------------------------------------------------------
void* DeviceStore::upload(const void* host, size_t bytes, cudaStream_t stream) {
  void* d = nullptr;
  VPK_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  ptrs.push_back(d);
  staging.emplace_back(static_cast<const char*>(host), static_cast<const char*>(host) + bytes);
  VPK_CUDA(cudaMemcpyAsync(d, staging.back().data(), bytes, cudaMemcpyHostToDevice, stream));
  return d;
}
void* DeviceStore::zeros(size_t bytes, cudaStream_t stream) {
  void* d = nullptr;
  VPK_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  ptrs.push_back(d);
  VPK_CUDA(cudaMemsetAsync(d, 0, bytes, stream));
  return d;
}
void DeviceStore::release() {
  for (void* p : ptrs) cudaFree(p);
  ptrs.clear();
  staging.clear();
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

int choose_cn(int C, int G) {
  const int step = (G % 2 == 0) ? 8 : 16;
  int ntiles = std::max(1, (C * G + 255) / 256);
  for (;; ++ntiles) {
    const int cn = round_up((C + ntiles - 1) / ntiles, step);
    if (cn * G <= 256) return cn;
  }
}

}  // namespace

int conv_n_tiles(int C, int G) {
  const int cn = choose_cn(C, G);
  return (C + cn - 1) / cn;
}

namespace {

float weight_at(const WeightRef& w, int oc, int ic, int ky, int kx) {
  if (!w.transposed) return w.w[((static_cast<size_t>(oc) * w.I + ic) * w.KH + ky) * w.KW + kx];
  return w.w[((static_cast<size_t>(ic) * w.O + oc) * w.KH + ky) * w.KW + kx];
}

}  // namespace

std::vector<BuiltConv> build_conv(const ConvSpec& spec, int dtype, int backend, DeviceStore& store,
                                  std::map<std::string, std::vector<PackedWeights>>& cache, cudaStream_t stream,
                                  int num_sms, bool measure_only) {
  std::vector<BuiltConv> out;
  const int G = spec.G, C = spec.C;
  VPK_REQUIRE(G >= 1 && G <= 4 && C > 0, "bad conv spec");
  const int Cn = choose_cn(C, G);
  const int N_pad = (C + Cn - 1) / Cn * Cn * G;

  // Accumulator regions (ConvLaunch::region_g0): only for CTA-pair launches of the tcgen05 halo kernel with whole 8-channel
  // chunks per CTA half, no bias, and one of the two gate splits the epilogue reads ((3, 1) of G = 4, (1, 1) of G = 2).
  // The packed copy is region-major then, so it gets its own cache entry (a small batch of the same layer runs without pairs)
  int rg0 = 0;
  if (spec.region_g0 > 0 && backend == 0 && dtype != DT_F32 && spec.biases.empty() && spec.phases.size() == 1 &&
      Cn % 16 == 0 && C % Cn == 0 && ((G == 4 && spec.region_g0 == 3) || (G == 2 && spec.region_g0 == 1)) &&
      (spec.phases[0].epi.kind == EPI_ST_C || spec.phases[0].epi.kind == EPI_ST_O) && getenv("VPK_NO_REGIONS") == nullptr &&
      getenv("VPK_TC_FAST_EPI") == nullptr) {
    const char* halo_env = getenv("VPK_TC_HALO");
    if ((halo_env == nullptr || atoi(halo_env) != 0) &&
        halo_will_pair(spec.B, spec.phases[0].H, spec.phases[0].W, Cn * G, N_pad / (Cn * G), num_sms))
      rg0 = spec.region_g0;
  }
  const std::string cache_key = spec.name + (rg0 ? "#regions" : "");
  // packed row of (channel, gate): gate-interleaved, or region-major inside each CTA half of the N tile
  auto row_of = [&](int ch, int g) {
    if (rg0 == 0) return ch * G + g;
    const int tile = ch / Cn, cl = ch % Cn, hc = Cn / 2, half = cl / hc, cj = cl % hc, ga = rg0, gb = G - rg0;
    return tile * Cn * G + half * hc * G + (g < ga ? cj * ga + g : hc * ga + cj * gb + (g - ga));
  };

  std::vector<PackedWeights>* packed = nullptr;
  if (!measure_only) {
    auto it = cache.find(cache_key);
    if (it == cache.end()) {
      std::vector<PackedWeights> pw(spec.phases.size());
      // bias (shared by the phases)
      float* d_bias = nullptr;
      if (!spec.biases.empty()) {
        std::vector<float> hb(N_pad, 0.f);
        for (const BiasRef& br : spec.biases)
          for (int ch = 0; ch < C; ++ch)
            for (int g = 0; g < G; ++g)
              if (br.gate_block[g] >= 0) hb[ch * G + g] += br.b[br.gate_block[g] * C + ch];
        d_bias = static_cast<float*>(store.upload(hb.data(), hb.size() * sizeof(float), stream));
      }
      for (size_t p = 0; p < spec.phases.size(); ++p) {
        std::vector<HostStep> steps = spec.phases[p].steps;
        VPK_REQUIRE(!steps.empty() && steps.size() <= static_cast<size_t>(kMaxSteps), "conv step count out of range");
        int k = 0;
        for (HostStep& h : steps) {
          h.s.wk = k;
          k += round_up(h.s.kc, 16);
        }
        const int K_pad = round_up(k, 64);
        std::vector<float> wf(static_cast<size_t>(N_pad) * K_pad, 0.f);
        for (const HostStep& h : steps) {
          const WeightRef& w = spec.wrefs[h.wref];
          for (int ch = 0; ch < C; ++ch)
            for (int g = 0; g < G; ++g) {
              if (w.gate_block[g] < 0) continue;
              const int oc = w.gate_block[g] * C + ch;
              float* row = wf.data() + static_cast<size_t>(row_of(ch, g)) * K_pad + h.s.wk;
              if (h.per_gate && h.gky[g] < 0) continue;       // this parity does not read this input offset
              const int ky = h.per_gate ? h.gky[g] : h.ky, kx = h.per_gate ? h.gkx[g] : h.kx;
              for (int j = 0; j < h.kw_valid; ++j) {
                float v = weight_at(w, oc, h.wc0 + h.s.c0 + j, ky, kx);
                if (h.wsplit != 0) {   // split operand: high 16-bit part, or what the high part misses
                  const float hi = (dtype == DT_F16) ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
                  v = (h.wsplit == 1) ? hi : v - hi;
                }
                row[j] = v;
              }
            }
        }
        PackedWeights& q = pw[p];
        q.K_pad = K_pad;
        q.N_pad = N_pad;
        q.Cn = Cn;
        q.bias = d_bias;
        if (dtype == DT_F32) {
          q.w = store.upload(wf.data(), wf.size() * sizeof(float), stream);
        } else if (dtype == DT_F16) {
          std::vector<__half> wh(wf.size());
          for (size_t i = 0; i < wf.size(); ++i) wh[i] = __float2half_rn(wf[i]);
          q.w = store.upload(wh.data(), wh.size() * sizeof(__half), stream);
        } else {
          std::vector<__nv_bfloat16> wb(wf.size());
          for (size_t i = 0; i < wf.size(); ++i) wb[i] = __float2bfloat16_rn(wf[i]);
          q.w = store.upload(wb.data(), wb.size() * sizeof(__nv_bfloat16), stream);
        }
        std::vector<ConvStep> cs(steps.size());
        for (size_t i = 0; i < steps.size(); ++i) cs[i] = steps[i].s;
        std::vector<int> region_mask(steps.size(), 0);
        if (rg0)
          for (size_t i = 0; i < steps.size(); ++i) {
            const WeightRef& w = spec.wrefs[steps[i].wref];
            for (int g = 0; g < G; ++g)
              if (w.gate_block[g] >= 0) region_mask[i] |= (g < rg0) ? 1 : 2;
            VPK_REQUIRE(region_mask[i] != 0, "conv step feeds no gate: " + spec.name);
          }
        q.steps = static_cast<ConvStep*>(store.upload(cs.data(), cs.size() * sizeof(ConvStep), stream));
        // halo-kernel tables: consecutive steps of one (source, channel block) form a block
        std::vector<HaloBlock> hb;
        std::vector<HaloTap> ht;
        int radius = 0;
        for (size_t ci = 0; ci < cs.size(); ++ci) {
          const ConvStep& c = cs[ci];
          if (hb.empty() || hb.back().src != c.src || hb.back().c0 != c.c0) {
            HaloBlock b{};
            b.src = c.src;
            b.c0 = c.c0;
            b.kc = c.kc;
            b.ntaps = 0;
            b.first_tap = static_cast<int>(ht.size());
            hb.push_back(b);
          }
          HaloTap t{};
          t.dy = c.dy;
          t.dx = c.dx;
          t.nk = static_cast<short>(((c.kc + 15) / 16) | (region_mask[ci] << 8));
          t.wk = c.wk;
          ht.push_back(t);
          hb.back().ntaps++;
          radius = std::max(radius, std::max(std::abs(static_cast<int>(c.dy)), std::abs(static_cast<int>(c.dx))));
        }
        q.blocks = static_cast<HaloBlock*>(store.upload(hb.data(), hb.size() * sizeof(HaloBlock), stream));
        q.taps = static_cast<HaloTap*>(store.upload(ht.data(), ht.size() * sizeof(HaloTap), stream));
        q.nblocks = static_cast<int>(hb.size());
        q.ntaps = static_cast<int>(ht.size());
        q.radius = radius;
      }
      it = cache.emplace(cache_key, std::move(pw)).first;
    }
    packed = &it->second;
    VPK_REQUIRE(packed->size() == spec.phases.size(), "packed-weight cache mismatch for " + spec.name);
  }

  for (size_t p = 0; p < spec.phases.size(); ++p) {
    const PhaseSpec& ph = spec.phases[p];
    BuiltConv bc;
    bc.name = spec.name + (spec.phases.size() > 1 ? "#" + std::to_string(p) : "");
    ConvLaunch& L = bc.L;
    std::memset(&L, 0, sizeof L);
    L.B = spec.B;
    L.H = ph.H;
    L.W = ph.W;
    L.nsrc = static_cast<int>(spec.srcs.size());
    for (int i = 0; i < L.nsrc; ++i) L.src[i] = spec.srcs[i];
    L.nsteps = static_cast<int>(ph.steps.size());
    L.G = G;
    L.Cn = Cn;
    L.N_pad = N_pad;
    L.epi = ph.epi;
    L.epi.C = C;
    L.is_gate_gemm = spec.is_gate_gemm ? 1 : 0;
    L.region_g0 = rg0;
    L.op_f16 = (dtype == DT_F16) ? 1 : 0;
    VPK_REQUIRE(dtype != DT_F16 || G == 1 || ph.epi.kind == EPI_DECOUPLE, "fp16 operands are for plain convs only: " + spec.name);
    double kreal = 0;
    for (const HostStep& h : ph.steps) {                      // split-bf16 layers count their three products
      double used = 1.0;
      if (h.per_gate) {                                       // sub-pixel form: only the parities that read this offset
        int nv = 0;
        for (int g = 0; g < G; ++g) nv += h.gky[g] >= 0 ? 1 : 0;
        used = static_cast<double>(nv) / G;
      } else if (h.wref >= 0 && h.wref < static_cast<int>(spec.wrefs.size())) {
        // multi-source gate convs: a weight tensor that does not feed gate g (gate_block -1) leaves zero rows in the packed
        // matrix -- executed by the MMA, but not algorithmic FLOPs (ST-LSTM / Causal LSTM output launches, the decoupling adapter)
        int nv = 0;
        for (int g = 0; g < G && g < 4; ++g) nv += spec.wrefs[h.wref].gate_block[g] >= 0 ? 1 : 0;
        used = static_cast<double>(nv) / std::min(G, 4);
      }
      if (h.count_flops) kreal += h.kw_valid * used;
    }
    L.flops = 2.0 * static_cast<double>(spec.B) * ph.H * ph.W * (static_cast<double>(G) * C) * kreal;
    if (!measure_only) {
      const PackedWeights& q = (*packed)[p];
      L.steps = q.steps;
      L.wpacked = q.w;
      L.K_pad = q.K_pad;
      L.epi.bias = q.bias;
      // the CUDA-core direct kernel is for what the tensor-core path cannot address (image-channel stems) and for
      // single-K-chunk heads; anything larger runs ~4x faster through tcgen05 even at N = 16 (profiles/)
      // (a one-tap transposed-conv parity with 32 output channels: 39 us direct vs 15 us through the halo kernel)
      bc.use_direct = direct_eligible(L) && !(backend == 0 && tc_eligible(L, dtype) &&
                                              (L.K_pad > 64 || L.N_pad >= 32 || L.epi.gn_sums != nullptr));
      // The halo-reuse kernel is the default for everything it can address.  (Early in the round the per-tap CTA-pair
      // kernel was ~25 % faster at tile N >= 192; with the flat single-thread issue loop, two taps per weight slot and
      // the whole-tile operand prefetch the halo kernel wins everywhere: cfg 3 gate GEMMs 1.51 -> 1.63 PFLOP/s, cfg 5
      // N = 192 layers 1.02 -> 1.16.)  VPK_TC_HALO=0 forces the per-tap kernels (tests, A/B runs).
      bool prefer_halo = true;
      if (const char* env = getenv("VPK_TC_HALO")) prefer_halo = atoi(env) != 0;
      bc.use_halo = !bc.use_direct && (backend == 0) && prefer_halo &&
                    halo_eligible(L, dtype, q.radius, q.nblocks, q.ntaps);
      bc.use_tc = !bc.use_direct && !bc.use_halo && (backend == 0) && tc_eligible(L, dtype);
      if (L.epi.proj_n > 0) {
        bc.use_direct = false;
        bc.use_halo = (backend == 0) && halo_eligible(L, dtype, q.radius, q.nblocks, q.ntaps);
        bc.use_tc = false;
        VPK_REQUIRE(bc.use_halo, "fused projection epilogue needs the tcgen05 halo kernel for " + spec.name);
      }
      VPK_REQUIRE(L.epi.kind != EPI_DECOUPLE || bc.use_halo, "the decoupling-loss epilogue needs the tcgen05 halo kernel");
      VPK_REQUIRE(L.epi.kind != EPI_SUBPIX || bc.use_halo, "the sub-pixel epilogue needs the tcgen05 halo kernel");
      VPK_REQUIRE(L.epi.gn_sums == nullptr || bc.use_halo,
                  "fused GroupNorm statistics need the tcgen05 halo kernel for " + spec.name);
      VPK_REQUIRE(rg0 == 0 || bc.use_halo, "accumulator regions need the tcgen05 halo kernel for " + spec.name);
      if (bc.use_halo) halo_make_plan(L, q.blocks, q.taps, q.nblocks, q.ntaps, q.radius, &bc.halo, num_sms);
      if (bc.use_tc) tc_make_plan(L, &bc.tc, num_sms);
    }
    out.push_back(std::move(bc));
  }
  return out;
}

}  // namespace vpk
