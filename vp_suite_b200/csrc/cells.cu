// Single-step cells (VPModelBlock boundary): NCHW fp32 in/out, internally the same generalised-conv launches as the
// rollouts.  The handle owns its scratch (activation-layout copies of the operands) per batch size.
#include "cells.h"

#include <map>
#include <memory>
#include <vector>

#include "../../include/vpk.h"
#include "backward.h"
#include "builders.h"
#include "causal.h"
#include "elementwise.h"
#include "phycell.h"
#include "stlstm.h"
#include "stlstm_ln.h"

namespace vpk {

namespace {

class CellBase : public Cell {
 public:
  CellBase(int precision, int backend_) : backend(backend_) {
    dtype = (precision == VPK_PREC_BF16) ? DT_BF16 : DT_F32;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      VPK_THROW(2, "no CUDA device: libvpk has no CPU fallback");
    }
    int dev = 0, sms = 0;
    VPK_CUDA(cudaGetDevice(&dev));
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) num_sms = sms;
  }
  ~CellBase() override {
    for (auto& kv : scratch) cudaFree(kv.second.first);
  }

 protected:
  int dtype, backend, num_sms = 148;
  DeviceStore store;
  std::map<std::string, std::vector<PackedWeights>> cache;
  std::map<std::string, std::pair<void*, size_t>> scratch;
  int built_batch = -1;
  std::vector<BuiltConv> convs;

  int esize() const { return static_cast<int>(dtype_size(dtype)); }
  ActInfo act() const { return ActInfo{dtype, esize()}; }

  void* buf(const std::string& name, size_t bytes) {
    auto it = scratch.find(name);
    if (it != scratch.end() && it->second.second >= bytes) return it->second.first;
    if (it != scratch.end()) {
      cudaFree(it->second.first);
      scratch.erase(it);
      built_batch = -1;
    }
    void* p = nullptr;
    VPK_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
    scratch[name] = {p, bytes};
    return p;
  }
  void to_nhwc(const float* src, void* dst, int out_dtype, int B, int C, int H, int W, cudaStream_t s) {
    launch_frames_to_nhwc(src, dst, out_dtype, B, 1, C, H, W, num_sms, s);
  }
  void run(const BuiltConv& bc, cudaStream_t s) {
    if (bc.use_halo) launch_conv_halo(bc.halo, s);
    else if (bc.use_tc) launch_conv_tc(bc.tc, s);
    else if (bc.use_direct) launch_conv_direct(bc.L, dtype, num_sms, s);
    else launch_conv_simt(bc.L, dtype, s);
  }
  void add(const ConvSpec& spec, cudaStream_t s) {
    for (BuiltConv& bc : build_conv(spec, dtype, backend, store, cache, s, num_sms, false)) convs.push_back(bc);
  }
  void finish_build(cudaStream_t s) {
    VPK_CUDA(cudaStreamSynchronize(s));
    store.staging.clear();
  }
};

// ------------------------------------------------------------------------------------------------------------------
class ConvLstmCell : public CellBase {
 public:
  ConvLstmCell(int precision, int backend_, int cin_, int ch_, int h_, int w_, int k_, int order, const float* weight,
               const float* bias)
      : CellBase(precision, backend_), cin(cin_), ch(ch_), h(h_), w(w_), k(k_), ifog(order == 1) {
    VPK_REQUIRE(cin >= 0 && ch > 0 && h > 0 && w > 0 && k % 2 == 1, "bad ConvLSTM cell shape");
    hw.assign(weight, weight + static_cast<size_t>(4) * ch * (cin + ch) * k * k);
    if (bias) hb.assign(bias, bias + 4 * ch);
  }
  // in: x, h, c, wci, wcf, wco   out: h', c'
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    const float* x = in[0];
    const bool peep = in[3] != nullptr;
    VPK_REQUIRE(!peep || (in[4] && in[5]), "peepholes must be given together");
    const size_t px = static_cast<size_t>(B) * h * w;
    void* xb = buf("x", px * std::max(cin, 1) * esize());
    void* hi = buf("h_in", px * ch * esize());
    void* ho = buf("h_out", px * ch * esize());
    float* cb = static_cast<float*>(buf("c", px * ch * sizeof(float)));
    float* pw[3] = {nullptr, nullptr, nullptr};
    if (peep)
      for (int i = 0; i < 3; ++i) pw[i] = static_cast<float*>(buf("peep" + std::to_string(i), sizeof(float) * ch * h * w));
    const int variant = (x ? 1 : 0) | (peep ? 2 : 0);
    if (built_batch != B || built_variant != variant) {
      convs.clear();
      LstmArgs a{std::string("cell.") + (x ? "xh" : "h"), B, h, w, cin, ch, k, x ? xb : nullptr, hi, ho, cb, hw.data(),
                 hb.empty() ? nullptr : hb.data(), ifog, pw[0], pw[1], pw[2]};
      add(lstm_spec(a, act()), s);
      finish_build(s);
      built_batch = B;
      built_variant = variant;
    }
    if (x) to_nhwc(x, xb, dtype, B, cin, h, w, s);
    to_nhwc(in[1], hi, dtype, B, ch, h, w, s);
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    if (peep)
      for (int i = 0; i < 3; ++i) to_nhwc(in[3 + i], pw[i], DT_F32, 1, ch, h, w, s);
    for (const BuiltConv& bc : convs) run(bc, s);
    launch_nhwc_to_nchw(ho, dtype, out[0], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(cb, DT_F32, out[1], B, ch, h, w, num_sms, s);
  }

  // Backward of one ConvLSTMCell step (conv_lstm_ndrplz.py:28-43; training through base_model.py:148-179): recomputes the
  // gate pre-activations z (one forward conv), differentiates the gate math elementwise, then
  //   d[x; h] = conv_transpose(dz, W)      -- the generalised conv launch (tensor cores in 16-bit mode)
  //   dW      = sum over positions of dz (x) shifted cat(x, h), db = sum of dz      -- fp32 CUDA-core kernels
  // in: x, h, c, dh_out, dc_out [, Wci, Wcf, Wco]   out: dx, dh, dc, dw, db [, dWci, dWcf, dWco]
  // The Shi et al. cell (gate order i, f, g, o; conv_lstm_hzzone.py:57-69) takes its peepholes in in[5..7] (NCHW [1, C, H, W],
  // nullptr = zero), returns their gradients in out[5..7], and accepts x == nullptr (the forecaster's all-zero input:
  // no dx then).
  void backward(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    VPK_REQUIRE(cin > 0 && in[1] && in[2] && out[1] && out[2] && out[3], "convlstm backward: null argument");
    VPK_REQUIRE(!ifog || (in[0] && out[0]), "convlstm backward: null argument");
    const size_t px = static_cast<size_t>(B) * h * w;
    const int cio = cin + ch;
    void* xb = buf("bw_x", px * cin * esize());
    void* hi = buf("bw_h", px * ch * esize());
    float* x32 = dtype == DT_F32 ? static_cast<float*>(xb) : static_cast<float*>(buf("bw_x32", px * cin * 4));
    float* h32 = dtype == DT_F32 ? static_cast<float*>(hi) : static_cast<float*>(buf("bw_h32", px * ch * 4));
    float* cb = static_cast<float*>(buf("bw_c", px * ch * 4));
    float* gh = in[3] ? static_cast<float*>(buf("bw_dh", px * ch * 4)) : nullptr;
    float* gc = in[4] ? static_cast<float*>(buf("bw_dc", px * ch * 4)) : nullptr;
    float* z = static_cast<float*>(buf("bw_z", px * 4 * ch * 4));
    float* dz = static_cast<float*>(buf("bw_dz", px * 4 * ch * 4));
    void* dza = dtype == DT_F32 ? static_cast<void*>(dz) : buf("bw_dz_act", px * 4 * ch * esize());
    float* dci = static_cast<float*>(buf("bw_dc_in", px * ch * 4));
    float* dxo = static_cast<float*>(buf("bw_dx", px * cin * 4));
    float* dho = static_cast<float*>(buf("bw_dh_in", px * ch * 4));
    if (bw_batch != B) {
      bw_convs.clear();
      if (wx_t.empty()) {      // dgrad weights: the x / h input-channel slices of W, kept in W's own [4C][.][k][k] layout,
        const size_t kk = static_cast<size_t>(k) * k;      // which IS ConvTranspose2d's [Cin_t = 4C][Cout_t][k][k]
        wx_t.resize(static_cast<size_t>(4) * ch * cin * kk);
        wh_t.resize(static_cast<size_t>(4) * ch * ch * kk);
        for (int o = 0; o < 4 * ch; ++o) {
          std::copy(hw.begin() + (static_cast<size_t>(o) * cio) * kk, hw.begin() + (static_cast<size_t>(o) * cio + cin) * kk,
                    wx_t.begin() + static_cast<size_t>(o) * cin * kk);
          std::copy(hw.begin() + (static_cast<size_t>(o) * cio + cin) * kk, hw.begin() + (static_cast<size_t>(o) * cio + cio) * kk,
                    wh_t.begin() + static_cast<size_t>(o) * ch * kk);
        }
      }
      int oh, ow;
      {   // z = conv(cat(x, h)) + b, reference row order (ndrplz: i | f | o | g; Shi et al.: i | f | g | o), dense fp32
        ConvSpec sp;
        sp.name = "cell.bw.z";
        sp.B = B;
        sp.G = 1;
        sp.C = 4 * ch;
        WeightRef wr;
        wr.w = hw.data();
        wr.O = 4 * ch;
        wr.I = cio;
        wr.KH = wr.KW = k;
        sp.wrefs.push_back(wr);
        if (!hb.empty()) {
          BiasRef br;
          br.b = hb.data();
          sp.biases.push_back(br);
        }
        lower_conv(sp, k, 1, k / 2, {ConvInput{make_view(xb, h, w, cin), 0, 0}, ConvInput{make_view(hi, h, w, ch), 0, cin}}, h, w,
                   esize(), &oh, &ow);
        EpiParams& e = sp.phases[0].epi;
        e.kind = EPI_BIAS_ACT;
        e.act = ACT_NONE;
        e.out_f32 = 1;
        dense_out(e, z, h, w, 4 * ch);
        for (BuiltConv& bc : build_conv(sp, dtype, backend, store, cache, s, num_sms, false)) bw_convs.push_back(bc);
      }
      n_z = static_cast<int>(bw_convs.size());
      auto dgrad = [&](const char* name, const std::vector<float>& wt, int co, float* dst) {
        DeconvArgs a{std::string("cell.bw.") + name, B, h, w, 4 * ch, co, k, 1, k / 2, 0, dza, wt.data(), nullptr, ACT_NONE, dst};
        a.out_f32 = true;
        for (BuiltConv& bc : build_conv(deconv_spec(a, act(), &oh, &ow), dtype, backend, store, cache, s, num_sms, false))
          bw_convs.push_back(bc);
      };
      dgrad("dx", wx_t, cin, dxo);
      dgrad("dh", wh_t, ch, dho);
      finish_build(s);
      bw_batch = B;
    }
    if (in[0]) {
      to_nhwc(in[0], xb, dtype, B, cin, h, w, s);
    } else {
      VPK_CUDA(cudaMemsetAsync(xb, 0, px * cin * esize(), s));
    }
    to_nhwc(in[1], hi, dtype, B, ch, h, w, s);
    if (dtype != DT_F32) {
      if (in[0]) to_nhwc(in[0], x32, DT_F32, B, cin, h, w, s);
      else VPK_CUDA(cudaMemsetAsync(x32, 0, px * cin * 4, s));
      to_nhwc(in[1], h32, DT_F32, B, ch, h, w, s);
    }
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    if (gh) to_nhwc(in[3], gh, DT_F32, B, ch, h, w, s);
    if (gc) to_nhwc(in[4], gc, DT_F32, B, ch, h, w, s);
    for (int i = 0; i < n_z; ++i) run(bw_convs[i], s);
    if (ifog) {
      launch_lstm_gate_backward(z, cb, gh, gc, dz, dtype == DT_F32 ? nullptr : dza, dtype, dci, static_cast<long long>(px), ch, num_sms, s);
    } else {
      const size_t pp = static_cast<size_t>(h) * w * ch;
      float* pk[3] = {nullptr, nullptr, nullptr};
      float* dpk[3] = {nullptr, nullptr, nullptr};
      for (int q = 0; q < 3; ++q) {
        if (in[5 + q]) {
          pk[q] = static_cast<float*>(buf(q == 0 ? "bw_wci" : q == 1 ? "bw_wcf" : "bw_wco", pp * 4));
          to_nhwc(in[5 + q], pk[q], DT_F32, 1, ch, h, w, s);
        }
        if (out[5 + q]) dpk[q] = static_cast<float*>(buf(q == 0 ? "bw_dwci" : q == 1 ? "bw_dwcf" : "bw_dwco", pp * 4));
      }
      launch_lstm_peep_gate_backward(z, cb, pk[0], pk[1], pk[2], gh, gc, dz, dtype == DT_F32 ? nullptr : dza, dtype, dci, dpk[0],
                                     dpk[1], dpk[2], B, static_cast<long long>(h) * w, ch, num_sms, s);
      for (int q = 0; q < 3; ++q)
        if (out[5 + q]) launch_nhwc_to_nchw(dpk[q], DT_F32, out[5 + q], 1, ch, h, w, num_sms, s);
    }
    for (size_t i = n_z; i < bw_convs.size(); ++i) run(bw_convs[i], s);
    if (out[0]) launch_nhwc_to_nchw(dxo, DT_F32, out[0], B, cin, h, w, num_sms, s);
    launch_nhwc_to_nchw(dho, DT_F32, out[1], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(dci, DT_F32, out[2], B, ch, h, w, num_sms, s);
    launch_conv_wgrad(x32, dz, out[3], B, h, w, cin, 4 * ch, k, cio, 0, s);
    launch_conv_wgrad(h32, dz, out[3], B, h, w, ch, 4 * ch, k, cio, cin, s);
    if (out[4]) launch_bias_grad(dz, out[4], static_cast<long long>(px), 4 * ch, s);
  }

 private:
  int cin, ch, h, w, k;
  bool ifog;
  int built_variant = -1;
  std::vector<float> hw, hb;
  std::vector<float> wx_t, wh_t;
  std::vector<BuiltConv> bw_convs;
  int bw_batch = -1, n_z = 0;
};

// ------------------------------------------------------------------------------------------------------------------
class StLstmCell : public CellBase {
 public:
  StLstmCell(int precision, int backend_, int cin_, int ch_, int h_, int w_, int k_, const float* w_x, const float* w_h,
             const float* w_m, const float* w_o, const float* w_last)
      : CellBase(precision, backend_), cin(cin_), ch(ch_), h(h_), w(w_), k(k_) {
    VPK_REQUIRE(cin > 0 && ch > 0 && h > 0 && w > 0 && k % 2 == 1, "bad ST-LSTM cell shape");
    const size_t kk = static_cast<size_t>(k) * k;
    wx.assign(w_x, w_x + 7 * ch * cin * kk);
    wh.assign(w_h, w_h + 4 * ch * ch * kk);
    wm.assign(w_m, w_m + 3 * ch * ch * kk);
    wo.assign(w_o, w_o + static_cast<size_t>(ch) * 2 * ch * kk);
    wl.assign(w_last, w_last + static_cast<size_t>(ch) * 2 * ch);
  }
  void set_layer_norm(const float* const* params) override {
    const int mult[4] = {7, 4, 3, 1};
    const int HW = h * w;
    for (int i = 0; i < 8; ++i) {          // reference [kC, H, W] -> NHWC [HW][kC], the order of the raw conv outputs
      const int kc = mult[i / 2] * ch;
      ln[i].resize(static_cast<size_t>(kc) * HW);
      for (int c = 0; c < kc; ++c)
        for (int q = 0; q < HW; ++q) ln[i][static_cast<size_t>(q) * kc + c] = params[i][static_cast<size_t>(c) * HW + q];
    }
    has_ln = true;
    built_batch = -1;
  }
  // layer_norm=True: the pipeline of stlstm_ln.h with a separate statistics launch (this boundary converts NCHW <-> NHWC
  // around every call anyway; the fused-statistics form lives in the rollout)
  void step_ln(int B, const float* const* in, float* const* out, cudaStream_t s) {
    // fp16 operands in 16-bit mode, as in the rollout (LayerNorm amplifies bf16 operand rounding past the tolerance)
    const int adt = (dtype == DT_BF16) ? DT_F16 : dtype;
    const ActInfo a16{adt, esize()};
    const size_t px = static_cast<size_t>(B) * h * w;
    const int HW = h * w;
    void* xb = buf("x", px * cin * esize());
    void* hi = buf("h_in", px * ch * esize());
    void* mi = buf("m_in", px * ch * esize());
    void* ho = buf("h_out", px * ch * esize());
    void* mem = buf("mem", px * 2 * ch * esize());
    void* mact = buf("m_act", px * ch * esize());
    void* dc = buf("dc", px * ch * esize());
    void* dm = buf("dm", px * ch * esize());
    float* cb = static_cast<float*>(buf("c", px * ch * sizeof(float)));
    float* mb = static_cast<float*>(buf("m", px * ch * sizeof(float)));
    float* op = static_cast<float*>(buf("o_part", px * ch * sizeof(float)));
    float* xr = static_cast<float*>(buf("x_raw", px * 7 * ch * sizeof(float)));
    float* hr = static_cast<float*>(buf("h_raw", px * 4 * ch * sizeof(float)));
    float* mr = static_cast<float*>(buf("m_raw", px * 3 * ch * sizeof(float)));
    float* orw = static_cast<float*>(buf("o_raw", px * ch * sizeof(float)));
    float* lr = static_cast<float*>(buf("l_raw", px * ch * sizeof(float)));
    float* part = static_cast<float*>(buf("ln_part", static_cast<size_t>(3) * B * kLnSlices * 2 * sizeof(float)));
    if (d_ln[0] == nullptr)
      for (int i = 0; i < 8; ++i) d_ln[i] = static_cast<float*>(store.upload(ln[i].data(), ln[i].size() * sizeof(float), s));
    if (built_batch != B) {
      convs.clear();
      int oh, ow;
      // 16-bit mode: split fp16 weights for conv_x / conv_h / conv_m (the two-product form of the rollout's
      // VPK_LN_PRODUCTS, model_predrnn.cu; a single step does not need the split activations: 7e-4 against the golden block)
      const char* ws_env = getenv("VPK_LN_PRODUCTS");
      const bool w_split_on = adt == DT_F16 && (ws_env == nullptr || atoi(ws_env) >= 2);
      auto raw_conv = [&](const char* name, const void* src, int ci, int co, int kk, const float* wt, float* dst,
                          bool wsplit = false) {
        ConvArgs a{std::string("cell.ln.") + name, B, h, w, ci, co, kk, 1, kk / 2, src, wt, nullptr, ACT_NONE, dst};
        a.out_f32_dense = true;
        a.w_split = wsplit && w_split_on && ((ci + 63) / 64) * 2 * kk * kk <= kMaxSteps;
        for (BuiltConv& bc : build_conv(conv_spec(a, a16, &oh, &ow), adt, backend, store, cache, s, num_sms, false))
          convs.push_back(bc);
      };
      raw_conv("x", xb, cin, 7 * ch, k, wx.data(), xr, true);
      raw_conv("h", hi, ch, 4 * ch, k, wh.data(), hr, true);
      raw_conv("m", mi, ch, 3 * ch, k, wm.data(), mr, true);
      raw_conv("o", mem, 2 * ch, ch, k, wo.data(), orw);
      raw_conv("last", mem, 2 * ch, ch, 1, wl.data(), lr);
      finish_build(s);
      built_batch = B;
    }
    to_nhwc(in[0], xb, adt, B, cin, h, w, s);
    to_nhwc(in[1], hi, adt, B, ch, h, w, s);
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    to_nhwc(in[3], mi, adt, B, ch, h, w, s);
    to_nhwc(in[3], mb, DT_F32, B, ch, h, w, s);
    auto run16 = [&](const BuiltConv& bc) {
      if (bc.use_halo) launch_conv_halo(bc.halo, s);
      else if (bc.use_tc) launch_conv_tc(bc.tc, s);
      else if (bc.use_direct) launch_conv_direct(bc.L, adt, num_sms, s);
      else launch_conv_simt(bc.L, adt, s);
    };
    for (int i = 0; i < 3; ++i) run16(convs[i]);
    LnStatsArgs sa{{xr, hr, mr}, {7ll * ch * HW, 4ll * ch * HW, 3ll * ch * HW}, 3, B, part};
    launch_ln_stats(sa, s);
    float* px_ = part;
    float* ph_ = part + static_cast<size_t>(B) * kLnSlices * 2;
    float* pm_ = ph_ + static_cast<size_t>(B) * kLnSlices * 2;
    StLnGatesArgs ga{xr, hr, mr, {px_, ph_, pm_}, {kLnSlices, kLnSlices, kLnSlices},
                     d_ln[0], d_ln[1], d_ln[2], d_ln[3], d_ln[4], d_ln[5], cb, mb, mem, mact, dc, dm, op, B, HW, ch, adt, 1.0f};
    launch_stlstm_ln_gates(ga, num_sms, s);
    run16(convs[3]);
    run16(convs[4]);
    LnStatsArgs so{{orw, nullptr, nullptr}, {1ll * ch * HW, 0, 0}, 1, B, part};
    launch_ln_stats(so, s);
    StLnOutArgs oa{orw, lr, part, kLnSlices, d_ln[6], d_ln[7], op, ho, B, HW, ch, adt};
    launch_stlstm_ln_out(oa, num_sms, s);
    launch_nhwc_to_nchw(ho, adt, out[0], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(cb, DT_F32, out[1], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(mb, DT_F32, out[2], B, ch, h, w, num_sms, s);
    if (out[3]) launch_nhwc_to_nchw(dc, adt, out[3], B, ch, h, w, num_sms, s);
    if (out[4]) launch_nhwc_to_nchw(dm, adt, out[4], B, ch, h, w, num_sms, s);
  }
  // in: x, h, c, m    out: h', c', m', delta_c, delta_m
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    if (has_ln) return step_ln(B, in, out, s);
    const size_t px = static_cast<size_t>(B) * h * w;
    void* xb = buf("x", px * cin * esize());
    void* hi = buf("h_in", px * ch * esize());
    void* mi = buf("m_in", px * ch * esize());
    void* ho = buf("h_out", px * ch * esize());
    void* mem = buf("mem", px * 2 * ch * esize());
    void* dc = buf("dc", px * ch * esize());
    void* dm = buf("dm", px * ch * esize());
    float* cb = static_cast<float*>(buf("c", px * ch * sizeof(float)));
    float* mb = static_cast<float*>(buf("m", px * ch * sizeof(float)));
    float* op = static_cast<float*>(buf("o_part", px * ch * sizeof(float)));
    if (built_batch != B) {
      convs.clear();
      StLstmArgs a{"cell.", B, h, w, cin, ch, k, xb, hi, make_view(mi, h, w, ch), ho, cb, mb, op, mem, dc, dm,
                   wx.data(), wh.data(), wm.data(), wo.data(), wl.data()};
      for (const ConvSpec& sp : stlstm_specs(a, act())) add(sp, s);
      finish_build(s);
      built_batch = B;
    }
    to_nhwc(in[0], xb, dtype, B, cin, h, w, s);
    to_nhwc(in[1], hi, dtype, B, ch, h, w, s);
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    to_nhwc(in[3], mi, dtype, B, ch, h, w, s);
    to_nhwc(in[3], mb, DT_F32, B, ch, h, w, s);
    for (const BuiltConv& bc : convs) run(bc, s);
    launch_nhwc_to_nchw(ho, dtype, out[0], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(cb, DT_F32, out[1], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(mb, DT_F32, out[2], B, ch, h, w, num_sms, s);
    if (out[3]) launch_nhwc_to_nchw(dc, dtype, out[3], B, ch, h, w, num_sms, s);
    if (out[4]) launch_nhwc_to_nchw(dm, dtype, out[4], B, ch, h, w, num_sms, s);
  }

 private:
  int cin, ch, h, w, k;
  std::vector<float> wx, wh, wm, wo, wl;
  bool has_ln = false;
  std::vector<float> ln[8];
  float* d_ln[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};


// ------------------------------------------------------------------------------------------------------------------
// Causal LSTM cell and gradient highway unit of PredRNN++ (causal.h; absent from the reference checkout: parity unpinned,
// checked against oracle/causal.py) behind the NCHW block boundary.
class CausalLstmCell : public CellBase {
 public:
  CausalLstmCell(int precision, int backend_, int cin_, int cm_, int ch_, int h_, int w_, int k_, const float* const* weights)
      : CellBase(precision, backend_), cin(cin_), cm(cm_), ch(ch_), h(h_), w(w_), k(k_) {
    VPK_REQUIRE(cin > 0 && cm > 0 && ch > 0 && h > 0 && w > 0 && k % 2 == 1, "bad Causal LSTM cell shape");
    const size_t kk = static_cast<size_t>(k) * k, cc = static_cast<size_t>(ch) * ch;
    const size_t n[7] = {7u * ch * cin * kk, 4 * cc * kk, 3 * cc * kk, 3u * ch * cm * kk, 4 * cc * kk, cc * kk, 2 * cc};
    for (int i = 0; i < 7; ++i) {
      VPK_REQUIRE(weights[i] != nullptr, "null Causal LSTM weight");
      wt[i].assign(weights[i], weights[i] + n[i]);
    }
  }
  // in: x, h, c, m    out: h', c', m'
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    const size_t px = static_cast<size_t>(B) * h * w;
    void* xb = buf("x", px * cin * esize());
    void* hi = buf("h_in", px * ch * esize());
    void* ci = buf("c_in", px * ch * esize());
    void* mi = buf("m_in", px * cm * esize());
    void* ho = buf("h_out", px * ch * esize());
    void* mem = buf("mem", px * 2 * ch * esize());
    float* cb = static_cast<float*>(buf("c", px * ch * sizeof(float)));
    float* mb = static_cast<float*>(buf("m", px * ch * sizeof(float)));
    float* op = static_cast<float*>(buf("o_part", px * ch * sizeof(float)));
    if (built_batch != B) {
      convs.clear();
      CausalArgs a{"cell.", B, h, w, cin, ch, k, xb, hi, make_view(ci, h, w, ch), make_view(mi, h, w, cm), ho, cb, mb, op, mem,
                   wt[0].data(), wt[1].data(), wt[2].data(), wt[3].data(), wt[4].data(), wt[5].data(), wt[6].data()};
      a.Cm = cm;
      for (const ConvSpec& sp : causal_lstm_specs(a, act())) add(sp, s);
      finish_build(s);
      built_batch = B;
    }
    to_nhwc(in[0], xb, dtype, B, cin, h, w, s);
    to_nhwc(in[1], hi, dtype, B, ch, h, w, s);
    to_nhwc(in[2], ci, dtype, B, ch, h, w, s);
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    to_nhwc(in[3], mi, dtype, B, cm, h, w, s);
    VPK_CUDA(cudaMemsetAsync(mb, 0, px * ch * sizeof(float), s));        // write-only state (the prefetch reads it)
    for (const BuiltConv& bc : convs) run(bc, s);
    launch_nhwc_to_nchw(ho, dtype, out[0], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(cb, DT_F32, out[1], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(mb, DT_F32, out[2], B, ch, h, w, num_sms, s);
  }

 private:
  int cin, cm, ch, h, w, k;
  std::vector<float> wt[7];
};

class GhuCell : public CellBase {
 public:
  GhuCell(int precision, int backend_, int ch_, int h_, int w_, int k_, const float* w_x, const float* w_z)
      : CellBase(precision, backend_), ch(ch_), h(h_), w(w_), k(k_) {
    VPK_REQUIRE(ch > 0 && h > 0 && w > 0 && k % 2 == 1, "bad GHU shape");
    const size_t n = 2 * static_cast<size_t>(ch) * ch * k * k;
    wx.assign(w_x, w_x + n);
    wz.assign(w_z, w_z + n);
  }
  // in: x, z    out: z'
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    const size_t px = static_cast<size_t>(B) * h * w;
    void* xb = buf("x", px * ch * esize());
    void* zi = buf("z_in", px * ch * esize());
    void* zo = buf("z_out", px * ch * esize());
    float* zb = static_cast<float*>(buf("z", px * ch * sizeof(float)));
    if (built_batch != B) {
      convs.clear();
      GhuArgs g{"ghu.", B, h, w, ch, k, xb, zi, zo, zb, wx.data(), wz.data()};
      add(ghu_spec(g, act()), s);
      finish_build(s);
      built_batch = B;
    }
    to_nhwc(in[0], xb, dtype, B, ch, h, w, s);
    to_nhwc(in[1], zi, dtype, B, ch, h, w, s);
    to_nhwc(in[1], zb, DT_F32, B, ch, h, w, s);
    for (const BuiltConv& bc : convs) run(bc, s);
    launch_nhwc_to_nchw(zb, DT_F32, out[0], B, ch, h, w, num_sms, s);
  }

 private:
  int ch, h, w, k;
  std::vector<float> wx, wz;
};

// ------------------------------------------------------------------------------------------------------------------
// ActionConditionalSpatioTemporalLSTMCell.forward (model_blocks/predrnn.py:142-169), layer_norm off or on: the rollout's
// pipeline (model_predrnn.cu: add_ac_cell) behind the NCHW block boundary -- raw convs with bias (x, h, a, m), per-sample
// statistics when layer_norm, the action-conditional gate kernel, conv_o / conv_last, the output kernel.
class StLstmAcCell : public CellBase {
 public:
  StLstmAcCell(int precision, int backend_, int cin_, int ch_, int h_, int w_, int k_, const float* const* weights,
               const float* const* biases)
      : CellBase(precision, backend_), cin(cin_), ch(ch_), h(h_), w(w_), k(k_) {
    VPK_REQUIRE(cin > 0 && ch > 0 && ch % 4 == 0 && h > 0 && w > 0 && k % 2 == 1, "bad action-conditional ST-LSTM cell shape");
    const size_t kk = static_cast<size_t>(k) * k;
    const size_t wn[6] = {7 * ch * cin * kk, 4 * ch * ch * kk, 4 * ch * ch * kk, 3 * ch * ch * kk,
                          static_cast<size_t>(ch) * 2 * ch * kk, static_cast<size_t>(ch) * 2 * ch};
    const size_t bn[6] = {7u * ch, 4u * ch, 4u * ch, 3u * ch, 1u * ch, 1u * ch};
    for (int i = 0; i < 6; ++i) {
      wts[i].assign(weights[i], weights[i] + wn[i]);
      bs[i].assign(biases[i], biases[i] + bn[i]);
    }
  }
  // params: (gamma, beta) x (conv_x, conv_h, conv_a, conv_m, conv_o), reference layout [k*C, H, W]
  void set_layer_norm(const float* const* params) override {
    const int mult[5] = {7, 4, 4, 3, 1};
    const int HW = h * w;
    for (int i = 0; i < 10; ++i) {
      const int kc = mult[i / 2] * ch;
      ln[i].resize(static_cast<size_t>(kc) * HW);
      for (int c = 0; c < kc; ++c)
        for (int q = 0; q < HW; ++q) ln[i][static_cast<size_t>(q) * kc + c] = params[i][static_cast<size_t>(c) * HW + q];
    }
    has_ln = true;
  }
  // in: x, h, c, m, a    out: h', c', m', delta_c, delta_m
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    const int adt = (dtype == DT_BF16) ? DT_F16 : dtype;      // fp16 operands in 16-bit mode, as in the rollout
    const ActInfo a16{adt, esize()};
    const size_t px = static_cast<size_t>(B) * h * w;
    const int HW = h * w;
    void* xb = buf("x", px * cin * esize());
    void* hi = buf("h_in", px * ch * esize());
    void* ai = buf("a_in", px * ch * esize());
    void* mi = buf("m_in", px * ch * esize());
    void* ho = buf("h_out", px * ch * esize());
    void* mem = buf("mem", px * 2 * ch * esize());
    void* mact = buf("m_act", px * ch * esize());
    void* dc = buf("dc", px * ch * esize());
    void* dm = buf("dm", px * ch * esize());
    float* cb = static_cast<float*>(buf("c", px * ch * sizeof(float)));
    float* mb = static_cast<float*>(buf("m", px * ch * sizeof(float)));
    float* op = static_cast<float*>(buf("o_part", px * ch * sizeof(float)));
    float* xr = static_cast<float*>(buf("x_raw", px * 7 * ch * sizeof(float)));
    float* hr = static_cast<float*>(buf("h_raw", px * 4 * ch * sizeof(float)));
    float* ar = static_cast<float*>(buf("a_raw", px * 4 * ch * sizeof(float)));
    float* mr = static_cast<float*>(buf("m_raw", px * 3 * ch * sizeof(float)));
    float* orw = static_cast<float*>(buf("o_raw", px * ch * sizeof(float)));
    float* lr = static_cast<float*>(buf("l_raw", px * ch * sizeof(float)));
    float* part = static_cast<float*>(buf("ln_part", static_cast<size_t>(4) * B * kLnSlices * 2 * sizeof(float)));
    if (has_ln && d_ln[0] == nullptr)
      for (int i = 0; i < 10; ++i) d_ln[i] = static_cast<float*>(store.upload(ln[i].data(), ln[i].size() * sizeof(float), s));
    if (built_batch != B) {
      convs.clear();
      int oh, ow;
      const char* ws_env = getenv("VPK_LN_PRODUCTS");
      const bool w_split_on = has_ln && adt == DT_F16 && (ws_env == nullptr || atoi(ws_env) >= 2);
      auto raw_conv = [&](const char* name, const void* src, int ci, int co, int kk, int idx, float* dst, bool precise) {
        ConvArgs a{std::string("cell.ac.") + name, B, h, w, ci, co, kk, 1, kk / 2, src, wts[idx].data(), bs[idx].data(),
                   ACT_NONE, dst};
        a.out_f32_dense = true;
        a.w_split = precise && w_split_on && ((ci + 63) / 64) * 2 * kk * kk <= kMaxSteps;
        for (BuiltConv& bc : build_conv(conv_spec(a, a16, &oh, &ow), adt, backend, store, cache, s, num_sms, false))
          convs.push_back(bc);
      };
      raw_conv("x", xb, cin, 7 * ch, k, 0, xr, true);
      raw_conv("h", hi, ch, 4 * ch, k, 1, hr, true);
      raw_conv("a", ai, ch, 4 * ch, k, 2, ar, true);
      raw_conv("m", mi, ch, 3 * ch, k, 3, mr, true);
      raw_conv("o", mem, 2 * ch, ch, k, 4, orw, false);
      raw_conv("last", mem, 2 * ch, ch, 1, 5, lr, false);
      finish_build(s);
      built_batch = B;
    }
    to_nhwc(in[0], xb, adt, B, cin, h, w, s);
    to_nhwc(in[1], hi, adt, B, ch, h, w, s);
    to_nhwc(in[2], cb, DT_F32, B, ch, h, w, s);
    to_nhwc(in[3], mi, adt, B, ch, h, w, s);
    to_nhwc(in[3], mb, DT_F32, B, ch, h, w, s);
    to_nhwc(in[4], ai, adt, B, ch, h, w, s);
    auto run16 = [&](const BuiltConv& bc) {
      if (bc.use_halo) launch_conv_halo(bc.halo, s);
      else if (bc.use_tc) launch_conv_tc(bc.tc, s);
      else if (bc.use_direct) launch_conv_direct(bc.L, adt, num_sms, s);
      else launch_conv_simt(bc.L, adt, s);
    };
    for (int i = 0; i < 4; ++i) run16(convs[i]);
    const size_t reg = static_cast<size_t>(B) * kLnSlices * 2;
    float *px_ = part, *ph_ = part + reg, *pm_ = part + 2 * reg, *pa_ = part + 3 * reg;
    if (has_ln) {
      LnStatsArgs sa{{xr, hr, mr}, {7ll * ch * HW, 4ll * ch * HW, 3ll * ch * HW}, 3, B, part};
      LnStatsArgs sb{{ar, nullptr, nullptr}, {4ll * ch * HW, 0, 0}, 1, B, pa_};
      launch_ln_stats(sa, s);
      launch_ln_stats(sb, s);
    }
    StLnGatesArgs ga{xr, hr, mr, {px_, ph_, pm_}, {kLnSlices, kLnSlices, kLnSlices},
                     d_ln[0], d_ln[1], d_ln[2], d_ln[3], d_ln[6], d_ln[7], cb, mb, mem, mact, dc, dm, op, B, HW, ch, adt, 1.0f};
    ga.A = ar;
    ga.part_a = pa_;
    ga.nslots_a = kLnSlices;
    ga.ga = d_ln[4];
    ga.ba = d_ln[5];
    ga.use_ln = has_ln ? 1 : 0;
    launch_stlstm_ln_gates(ga, num_sms, s);
    run16(convs[4]);
    run16(convs[5]);
    if (has_ln) {
      LnStatsArgs so{{orw, nullptr, nullptr}, {1ll * ch * HW, 0, 0}, 1, B, part};
      launch_ln_stats(so, s);
    }
    StLnOutArgs oa{orw, lr, part, kLnSlices, d_ln[8], d_ln[9], op, ho, B, HW, ch, adt};
    oa.use_ln = has_ln ? 1 : 0;
    launch_stlstm_ln_out(oa, num_sms, s);
    launch_nhwc_to_nchw(ho, adt, out[0], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(cb, DT_F32, out[1], B, ch, h, w, num_sms, s);
    launch_nhwc_to_nchw(mb, DT_F32, out[2], B, ch, h, w, num_sms, s);
    if (out[3]) launch_nhwc_to_nchw(dc, adt, out[3], B, ch, h, w, num_sms, s);
    if (out[4]) launch_nhwc_to_nchw(dm, adt, out[4], B, ch, h, w, num_sms, s);
  }

 private:
  int cin, ch, h, w, k;
  std::vector<float> wts[6], bs[6];
  bool has_ln = false;
  std::vector<float> ln[10];
  float* d_ln[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// ------------------------------------------------------------------------------------------------------------------
class PhyCellCell : public CellBase {
 public:
  PhyCellCell(int precision, int backend_, int ch_, int hid_, int h_, int w_, int k_, const float* c1w, const float* c1b,
              const float* gnw, const float* gnb, const float* c2w, const float* c2b, const float* gw, const float* gb)
      : CellBase(precision, backend_), ch(ch_), hid(hid_), h(h_), w(w_), k(k_) {
    VPK_REQUIRE(ch > 0 && hid > 0 && h > 0 && w > 0 && k % 2 == 1, "bad PhyCell shape");
    w1.assign(c1w, c1w + static_cast<size_t>(hid) * ch * k * k);
    b1.assign(c1b, c1b + hid);
    gw_.assign(gnw, gnw + hid);
    gb_.assign(gnb, gnb + hid);
    w2.assign(c2w, c2w + static_cast<size_t>(ch) * hid);
    b2.assign(c2b, c2b + ch);
    wg.assign(gw, gw + static_cast<size_t>(ch) * 2 * ch * 9);
    bg.assign(gb, gb + ch);
    int sq = 1;
    while ((sq + 1) * (sq + 1) <= hid) ++sq;          // floor(sqrt(hid)); model_blocks/phydnet.py:348-362
    while (hid % sq != 0) --sq;
    groups = hid / sq;
  }
  // action_conditional=True (model_blocks/phydnet.py:44-55): frame / hidden first go through their 1x1 action convs
  void set_action_convs(int action_size, const float* fw_, const float* fb_, const float* hw_, const float* hb_) override {
    VPK_REQUIRE(action_size > 0 && action_size <= 8, "PhyCell action convs: action_size must be in 1..8");
    a_sz = action_size;
    const size_t n = static_cast<size_t>(ch) * (ch + a_sz);
    afw.assign(fw_, fw_ + n);
    afb.assign(fb_, fb_ + ch);
    ahw.assign(hw_, hw_ + n);
    ahb.assign(hb_, hb_ + ch);
    built_batch = -1;
  }
  // in: x, h [, action [b, a]]    out: h'
  // 16-bit mode: F.conv1 / F.conv2 run on FP16 operands (h and the GroupNorm output are O(1); the GroupNorm between them
  // amplifies operand rounding: with bf16 operands the golden block is 6.8e-3 off after ONE step, with fp16 3e-3), the
  // gate conv and its blend epilogue on bf16.
  void step(int B, const float* const* in, float* const* out, cudaStream_t s) override {
    const size_t px = static_cast<size_t>(B) * h * w;
    const int Cp = phycell_padded_channels(hid);
    const int fdt = (dtype == DT_BF16) ? DT_F16 : dtype;
    void* xb = buf("x", px * ch * esize());
    void* hi = buf("h_act", px * ch * esize());
    void* hf = (fdt != dtype) ? buf("h_f16", px * ch * esize()) : hi;
    void* ho = buf("h_act_out", px * ch * esize());
    float* hm = static_cast<float*>(buf("h_master", px * ch * sizeof(float)));
    float* ht = static_cast<float*>(buf("htilde", px * ch * sizeof(float)));
    float* f1 = static_cast<float*>(buf("f1raw", px * Cp * sizeof(float)));
    void* f1n = buf("f1n", px * Cp * esize());
    const bool ac = a_sz > 0;
    VPK_REQUIRE(!ac || in[2] != nullptr, "Given actions are None or of the wrong size!");
    // action-conditional: fp32 frame / hidden, the inflated action, and the convolved frame / hidden (fp32; `hidden` is what
    // the prediction h~ = hidden + F(hidden) starts from)
    float *x32 = nullptr, *h32 = nullptr, *a32 = nullptr, *fa32 = nullptr, *ha32 = nullptr;
    if (ac) {
      x32 = static_cast<float*>(buf("x32", px * ch * 4));
      h32 = static_cast<float*>(buf("h32", px * ch * 4));
      a32 = static_cast<float*>(buf("a32", px * 8 * 4));
      fa32 = static_cast<float*>(buf("fa32", px * ch * 4));
      ha32 = static_cast<float*>(buf("ha32", px * ch * 4));
    }
    if (built_batch != B) {
      convs.clear();
      PhyCellArgs a{"cell.", B, h, w, ch, hid, k, xb, hi, ho, hm, ht, f1, f1n, w1.data(), b1.data(), w2.data(),
                    b2.data(), wg.data(), bg.data()};
      a.h_f = hf;
      if (ac) a.h_res = ha32;
      const std::vector<ConvSpec> specs = phycell_specs(a, act(), ActInfo{fdt, esize()});
      for (int i = 0; i < 3; ++i)
        for (BuiltConv& bc : build_conv(specs[i], i < 2 ? fdt : dtype, backend, store, cache, s, num_sms, false))
          convs.push_back(bc);
      if (ac) {       // two fp32-operand 1x1 convs over (tensor, inflated action): convs[3] frame, convs[4] hidden
        auto action_conv = [&](const char* name, const float* src, const std::vector<float>& wt, const std::vector<float>& bi,
                               float* dst) {
          ConvSpec s1;
          s1.name = std::string("cell.") + name;
          s1.B = B;
          s1.G = 1;
          s1.C = ch;
          WeightRef wr;
          wr.w = wt.data();
          wr.O = ch;
          wr.I = ch + a_sz;
          wr.KH = wr.KW = 1;
          s1.wrefs.push_back(wr);
          BiasRef br;
          br.b = bi.data();
          s1.biases.push_back(br);
          ConvInput i0{make_view(src, h, w, ch), 0, 0};
          ConvInput i1{make_view(a32, h, w, 8), 0, ch};
          i1.wc_count = a_sz;
          int oh_, ow_;
          lower_conv(s1, 1, 1, 0, {i0, i1}, h, w, 4, &oh_, &ow_);
          EpiParams& e = s1.phases[0].epi;
          e.kind = EPI_BIAS_ACT;
          e.act = ACT_NONE;
          e.out_f32 = 1;
          dense_out(e, dst, h, w, ch);
          for (BuiltConv& bc : build_conv(s1, DT_F32, backend, store, cache, s, num_sms, false)) convs.push_back(bc);
        };
        action_conv("frame_action_conv", x32, afw, afb, fa32);
        action_conv("hidden_action_conv", h32, ahw, ahb, ha32);
      }
      d_gamma = static_cast<float*>(store.upload(gw_.data(), gw_.size() * sizeof(float), s));
      d_beta = static_cast<float*>(store.upload(gb_.data(), gb_.size() * sizeof(float), s));
      finish_build(s);
      VPK_CUDA(cudaMemsetAsync(f1n, 0, px * Cp * esize(), s));   // padded channels stay zero
      built_batch = B;
    }
    auto run_dt = [&](const BuiltConv& bc, int dt) {
      if (bc.use_halo) launch_conv_halo(bc.halo, s);
      else if (bc.use_tc) launch_conv_tc(bc.tc, s);
      else if (bc.use_direct) launch_conv_direct(bc.L, dt, num_sms, s);
      else launch_conv_simt(bc.L, dt, s);
    };
    if (ac) {
      const long long n = static_cast<long long>(px) * ch;
      to_nhwc(in[0], x32, DT_F32, B, ch, h, w, s);
      to_nhwc(in[1], h32, DT_F32, B, ch, h, w, s);
      launch_inflate_actions(in[2], a_sz, a_sz, a32, DT_F32, B, 1, h * w, 8, num_sms, s);
      run_dt(convs[3], DT_F32);
      run_dt(convs[4], DT_F32);
      launch_add_to_act(fa32, DT_F32, nullptr, xb, dtype, n, num_sms, s);       // frame' as the gate conv's / blend's operand
      launch_add_to_act(ha32, DT_F32, nullptr, hi, dtype, n, num_sms, s);       // hidden' for the gate conv
      if (hf != hi) launch_add_to_act(ha32, DT_F32, nullptr, hf, fdt, n, num_sms, s);
    } else {
      to_nhwc(in[0], xb, dtype, B, ch, h, w, s);
      to_nhwc(in[1], hi, dtype, B, ch, h, w, s);
      if (hf != hi) to_nhwc(in[1], hf, fdt, B, ch, h, w, s);
      to_nhwc(in[1], hm, DT_F32, B, ch, h, w, s);
    }
    run_dt(convs[0], fdt);
    launch_groupnorm_act(f1, DT_F32, f1n, fdt, nullptr, B, h * w, hid, Cp, Cp, groups, d_gamma, d_beta, 1e-5f,
                         ACT_NONE, s);
    run_dt(convs[1], fdt);
    run_dt(convs[2], dtype);
    launch_nhwc_to_nchw(hm, DT_F32, out[0], B, ch, h, w, num_sms, s);
  }

 private:
  int ch, hid, h, w, k, groups = 1;
  int a_sz = 0;
  std::vector<float> afw, afb, ahw, ahb;
  std::vector<float> w1, b1, gw_, gb_, w2, b2, wg, bg;
  float *d_gamma = nullptr, *d_beta = nullptr;
};

}  // namespace

Cell* make_phycell_cell(int precision, int backend, int ch, int hid, int h, int w, int k, const float* conv1_w,
                        const float* conv1_b, const float* gn_w, const float* gn_b, const float* conv2_w,
                        const float* conv2_b, const float* gate_w, const float* gate_b) {
  return new PhyCellCell(precision, backend, ch, hid, h, w, k, conv1_w, conv1_b, gn_w, gn_b, conv2_w, conv2_b, gate_w,
                         gate_b);
}

Cell* make_stlstm_ac_cell(int precision, int backend, int cin, int ch, int h, int w, int k, const float* const* weights,
                          const float* const* biases) {
  return new StLstmAcCell(precision, backend, cin, ch, h, w, k, weights, biases);
}

Cell* make_convlstm_cell(int precision, int backend, int cin, int ch, int h, int w, int k, int gate_order,
                         const float* weight, const float* bias) {
  return new ConvLstmCell(precision, backend, cin, ch, h, w, k, gate_order, weight, bias);
}

Cell* make_stlstm_cell(int precision, int backend, int cin, int ch, int h, int w, int k, const float* w_x,
                       const float* w_h, const float* w_m, const float* w_o, const float* w_last) {
  return new StLstmCell(precision, backend, cin, ch, h, w, k, w_x, w_h, w_m, w_o, w_last);
}

Cell* make_causal_lstm_cell(int precision, int backend, int cin, int cm, int ch, int h, int w, int k,
                            const float* const* weights) {
  return new CausalLstmCell(precision, backend, cin, cm, ch, h, w, k, weights);
}

Cell* make_ghu_cell(int precision, int backend, int ch, int h, int w, int k, const float* w_x, const float* w_z) {
  return new GhuCell(precision, backend, ch, h, w, k, w_x, w_z);
}

}  // namespace vpk
