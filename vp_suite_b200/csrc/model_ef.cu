// convlstm-shi: Encoder-Forecaster ConvLSTM rollout (reference: models/precipitation_nowcasting/ef_blocks.py:52-187,
// ef_conv_lstm.py:7-108, model_blocks/conv_lstm_hzzone.py:38-70).
//
// Schedule: TIME-MAJOR.  The reference runs each layer over all timesteps before the next layer (layer-major,
// ef_blocks.py:67-82,100-114) and materialises [b,t,C,H,W] between layers; since layer l at step t only needs layer
// l-1 at step t and its own state from step t-1, running all layers for step t and then t+1 is the same computation
//  and keeps exactly one h/c state per layer resident.  The forecaster's top RNN is fed
// `inputs=None` by the reference (zeros, conv_lstm_hzzone.py:54-56): its x-side K-steps are dropped.
#include <cstdlib>
#include <cstring>

#include "builders.h"
#include "conv_stem.h"
#include "elementwise.h"
#include "model.h"

namespace vpk {

namespace {

class EfConvLstm : public Model {
 public:
  explicit EfConvLstm(const vpk_model_desc& d) : Model(d) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h > 0 && d.img_w > 0, "bad img_shape");
    // state sizes (ef_blocks.py:145-158; utils/models.py:131-161 and :164-193)
    int hh = d.img_h, ww = d.img_w;
    for (int n = 0; n < 3; ++n) {
      VPK_REQUIRE(d.enc_conv_s[n] == 1 || d.enc_conv_s[n] == 2, "enc_conv_s must be 1 or 2");
      hh = (hh + 2 * d.enc_conv_p[n] - d.enc_conv_k[n]) / d.enc_conv_s[n] + 1;
      ww = (ww + 2 * d.enc_conv_p[n] - d.enc_conv_k[n]) / d.enc_conv_s[n] + 1;
      eh[n] = hh;
      ew[n] = ww;
    }
    dh[0] = hh;
    dw[0] = ww;
    for (int n = 0; n < 3; ++n) {
      // the reference's own (non-standard) formula, utils/models.py:190-191
      hh = (hh - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + (d.dec_conv_k[n] - 1) + d.dec_conv_p[n];
      ww = (ww - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + (d.dec_conv_k[n] - 1) + d.dec_conv_p[n];
      // what ConvTranspose2d really produces (no output_padding in _make_layers, ef_blocks.py:29-31)
      const int th = (dh[n] - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + d.dec_conv_k[n];
      const int tw = (dw[n] - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + d.dec_conv_k[n];
      VPK_REQUIRE(th == hh && tw == ww, "decoder conv hyper-parameters give inconsistent sizes");
      dh[n + 1] = hh;
      dw[n + 1] = ww;
    }
    // AttributeError in the reference (ef_blocks.py:160-167)
    VPK_REQUIRE(dh[3] == d.img_h && dw[3] == d.img_w, "model layer hyper-parameters yield wrong output size");
    for (int n = 0; n < 3; ++n) {
      // forecaster rnn n starts from encoder state 2-n (ef_blocks.py:109-113): sizes must agree
      VPK_REQUIRE(dh[n] == eh[2 - n] && dw[n] == ew[2 - n], "encoder / forecaster state sizes differ");
      VPK_REQUIRE(d.dec_c[2 * n] == d.enc_c[2 * (2 - n) + 1], "encoder / forecaster state channels differ");
      VPK_REQUIRE(d.enc_rnn_k[n] % 2 == 1 && d.dec_rnn_k[n] % 2 == 1, "rnn kernel size must be odd");
    }
    int in_c = d.img_c;
    for (int n = 0; n < 3; ++n) {
      const std::string st = "encoder.stage" + std::to_string(n + 1) + ".conv.";
      const std::string rn = "encoder.rnn" + std::to_string(n + 1) + ".";
      const int mid = d.enc_c[2 * n], outc = d.enc_c[2 * n + 1];
      declare(st + "weight", {mid, in_c, d.enc_conv_k[n], d.enc_conv_k[n]});
      declare(st + "bias", {mid});
      for (const char* pk : {"Wci", "Wcf", "Wco"}) declare(rn + pk, {1, outc, eh[n], ew[n]});
      declare(rn + "_conv.weight", {4 * outc, mid + outc, d.enc_rnn_k[n], d.enc_rnn_k[n]});
      declare(rn + "_conv.bias", {4 * outc});
      in_c = outc;
    }
    for (int n = 0; n < 3; ++n) {
      const int idx = 3 - n;
      const std::string rn = "forecaster.rnn" + std::to_string(idx) + ".";
      const std::string st = "forecaster.stage" + std::to_string(idx) + ".deconv.";
      const int mid = d.dec_c[2 * n], outc = d.dec_c[2 * n + 1];
      for (const char* pk : {"Wci", "Wcf", "Wco"}) declare(rn + pk, {1, mid, dh[n], dw[n]});
      declare(rn + "_conv.weight", {4 * mid, in_c + mid, d.dec_rnn_k[n], d.dec_rnn_k[n]});
      declare(rn + "_conv.bias", {4 * mid});
      declare(st + "weight", {mid, outc, d.dec_conv_k[n], d.dec_conv_k[n]});
      declare(st + "bias", {outc});
      dec_in_c[n] = in_c;
      in_c = outc;
    }
    VPK_REQUIRE(d.final_conv_c == d.dec_c[5], "identity final block must keep the channel count");
    declare("forecaster.stage1.final.weight", {d.img_c, d.final_conv_c, 1, 1});
    declare("forecaster.stage1.final.bias", {d.img_c});
  }

 protected:
  int default_microbatch() const override {
    // keep one microbatch's activations around 8 GB: rnn1 state is the largest tensor
    const double per_seq = static_cast<double>(eh[0]) * ew[0] * 2200.0;   // bytes, all layers, rough
    int mb = static_cast<int>(8e9 / per_seq);
    return std::max(1, std::min(mb, 256));
  }

  std::vector<float> peephole_packed(const std::string& key, int C, int H, int W) const {
    // reference layout [1, C, H, W] -> channel-quad layout [C/4, H, W, 4] (or [H, W, C] when C % 4 != 0), the layout
    // of the cell state it multiplies (epilogue.cuh: state_addr)
    const float* p = hp(key);
    std::vector<float> out(static_cast<size_t>(C) * H * W);
    const bool c4 = C % 4 == 0;
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const size_t src = (static_cast<size_t>(c) * H + y) * W + x;
          const size_t dst = c4 ? ((static_cast<size_t>(c / 4) * H + y) * W + x) * 4 + (c & 3)
                                : (static_cast<size_t>(y) * W + x) * C + c;
          out[dst] = p[src];
        }
    return out;
  }

  // the three peepholes of one cell as bf16, packed [C/8][H][W][3][8] (raw bits carried in a float vector for the
  // upload cache): what the tcgen05 epilogue reads, 48 contiguous bytes per position and 8-channel chunk
  std::vector<float> peephole_bf16_packed(const std::string& rn, int C, int H, int W) const {
    const char* names[3] = {"Wci", "Wcf", "Wco"};
    std::vector<uint16_t> bits(static_cast<size_t>(C) * H * W * 3);
    for (int k = 0; k < 3; ++k) {
      const float* p = hp(rn + names[k]);
      for (int c = 0; c < C; ++c)
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            // halved: lstm_finish folds the 0.5 of sigmoid(z) = 0.5 tanh(0.5 z) + 0.5 into the bias and the peepholes
            // (exact: scaling by a power of two commutes with the bf16 rounding)
            const __nv_bfloat16 v = __float2bfloat16_rn(0.5f * p[(static_cast<size_t>(c) * H + y) * W + x]);
            uint16_t u;
            std::memcpy(&u, &v, 2);
            bits[(((static_cast<size_t>(c >> 3) * H + y) * W + x) * 3 + k) * 8 + (c & 7)] = u;
          }
    }
    std::vector<float> out(bits.size() / 2);
    std::memcpy(out.data(), bits.data(), bits.size() * 2);
    return out;
  }

  // ================================================================================================================
  // Persistent, state-resident form for the latency-bound small-batch case (BASELINE north_star: "a persistent-kernel
  // variant keeps cell state resident across the context and prediction timesteps").  LAYER-MAJOR like the reference
  // (ef_blocks.py:67-82, 100-114): every stage conv / deconv runs ONCE over all B * T frames of its layer, and every
  // ConvLSTM layer is ONE launch of the halo kernel's sequence mode (conv_halo.cu MODE 4): T timesteps inside the kernel,
  // the cell state of a CTA's tiles in shared memory from the first to the last step, h_t written to slot t of the
  // layer's output sequence and read back through TMA after a grid-wide barrier.  ~14 launches per rollout instead of
  // ~120; each (tile, step) performs exactly the per-step program's arithmetic in the same order, so the frames are
  // bit-identical to it.  Sequence tensors are sample-major [b][t] (sample index b * T + t): the input frames and the
  // output frames then already have the boundary's [b, t, ...] order.
  // Returns false (nothing emitted) when a layer's cell state does not fit the shared-memory reserve.
  // ================================================================================================================
  bool build_seq(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) {
    const vpk_model_desc& d = desc;
    const ActInfo act{dtype, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const int cs = 8;                                   // frames are stored with 8 zero-padded channels (TMA-addressable)
    const size_t frame_px = static_cast<size_t>(h) * w;
    char* frames_in = static_cast<char*>(arena.alloc(frame_px * cs * esz * B * t_in));
    float* out_stage = static_cast<float*>(arena.alloc(frame_px * c * sizeof(float) * B * pred));
    void *xin[3], *hseq_e[3], *hseq_f[3], *yseq[3], *hzero[3];
    float* cbuf[3];
    for (int n = 0; n < 3; ++n) {
      const size_t px = static_cast<size_t>(eh[n]) * ew[n];
      xin[n] = arena.alloc(px * d.enc_c[2 * n] * esz * B * t_in);
      hseq_e[n] = arena.alloc(px * d.enc_c[2 * n + 1] * esz * B * t_in);
      hseq_f[n] = arena.alloc(px * d.enc_c[2 * n + 1] * esz * B * pred);      // forecaster layer continuing encoder state n
      cbuf[n] = static_cast<float*>(arena.alloc(px * d.enc_c[2 * n + 1] * sizeof(float) * B));
    }
    for (int n = 0; n < 3; ++n)
      yseq[n] = arena.alloc(static_cast<size_t>(dh[n + 1]) * dw[n + 1] * d.dec_c[2 * n + 1] * esz * B * pred);
    // the zero initial hidden states of the three encoder layers and the six grid-barrier counters: ONE block, one memset node
    size_t zoff[4] = {0, 0, 0, 0};
    for (int n = 0; n < 3; ++n)
      zoff[n + 1] = zoff[n] + (static_cast<size_t>(eh[n]) * ew[n] * d.enc_c[2 * n + 1] * esz * B + 1023) / 1024 * 1024;
    const size_t zero_bytes = zoff[3] + 6 * 128;
    char* zero_block = static_cast<char*>(arena.alloc(zero_bytes));
    if (zero_block != nullptr)
      for (int n = 0; n < 3; ++n) hzero[n] = zero_block + zoff[n];
    unsigned* barriers = zero_block ? reinterpret_cast<unsigned*>(zero_block + zoff[3]) : nullptr;
    if (measure) return true;

    const float* peep[2][3][3] = {};
    const void* peep16[2][3] = {};
    for (int side = 0; side < 2; ++side)
      for (int n = 0; n < 3; ++n) {
        const std::string rn = std::string(side == 0 ? "encoder.rnn" : "forecaster.rnn") + std::to_string(n + 1) + ".";
        const int C = d.enc_c[2 * n + 1];
        if (!has(rn + "Wci") && !has(rn + "Wcf") && !has(rn + "Wco")) continue;
        const char* names[3] = {"Wci", "Wcf", "Wco"};
        for (int k = 0; k < 3; ++k)
          peep[side][n][k] = dev_f32(rn + names[k], peephole_packed(rn + names[k], C, eh[n], ew[n]), stream);
        if (C % 8 == 0) peep16[side][n] = dev_f32(rn + "peepholes.bf16", peephole_bf16_packed(rn, C, eh[n], ew[n]), stream);
      }

    std::vector<Op> pre, body;
    // One ConvLSTM layer over T steps: the single-step spec gives the packed weights, step tables and tiling; its sequence
    // plan re-reads the sources / output as [b][t] sample-major sequences.
    auto seq_layer = [&](LstmArgs la, int T, const void* x_seq, const void* h0, int h0_samples, int h0_sb, int h0_off,
                         void* out_seq, bool c_zero, int slot) -> bool {
      la.x = x_seq;
      la.h_in = out_seq;
      la.h_out = out_seq;
      ConvSpec sp = lstm_spec(la, act);
      std::vector<BuiltConv> built = build_conv(sp, dtype, backend, store, packed_cache, stream, num_sms, false);
      if (built.size() != 1 || !built[0].use_halo) return false;
      BuiltConv& bc = built[0];
      const int n_x = x_seq ? 1 : 0;
      HaloSeqSpec sq{};
      sq.T = T;
      ConvLaunch L = bc.L;
      if (n_x) {
        sq.sb[0] = T; sq.st[0] = 1; sq.off[0] = 0; sq.samples[0] = B * T;
      }
      sq.recur_src = n_x;                                              // the h source follows x in lstm_spec's input list
      sq.sb[n_x] = T; sq.st[n_x] = 1; sq.off[n_x] = -1; sq.samples[n_x] = B * T;     // step t reads slot t - 1
      VPK_REQUIRE(L.nsrc == n_x + 1 && L.nsrc < kMaxSrc, "sequence layer: unexpected source list");
      L.src[L.nsrc] = make_view(h0, la.H, la.W, la.C);                 // initial hidden state (read at t = 0 only)
      sq.h0_src = L.nsrc;
      sq.sb[L.nsrc] = h0_sb; sq.st[L.nsrc] = 0; sq.off[L.nsrc] = h0_off; sq.samples[L.nsrc] = h0_samples;
      L.nsrc += 1;
      sq.out_sb = T; sq.out_st = 1; sq.out_off = 0;
      sq.c_zero = c_zero ? 1 : 0;
      sq.barrier = barriers + slot * 32;
      auto plan = std::make_shared<HaloPlan>();
      if (!halo_make_seq_plan(L, bc.halo.blocks, bc.halo.taps, bc.halo.nblocks, bc.halo.ntaps, bc.halo.P, sq, plan.get(), num_sms))
        return false;
      Op op;
      op.name = la.name + "seq";
      op.flops = bc.L.flops * T;
      op.gate = true;
      op.fn = [plan](cudaStream_t s, const RunCtx&) { launch_conv_halo(*plan, s); };
      body.push_back(std::move(op));
      return true;
    };
    Program tmp;                  // add_conv appends to an op list of its own program argument or to `dst`
    auto conv_ops = [&](const ConvSpec& spec) { add_conv(tmp, spec, false, stream, -1, &body); };
    {
      const int ns = num_sms;
      Op cv;
      cv.name = "frames_to_nhwc";
      const long long chw = static_cast<long long>(c) * h * w;
      cv.fn = [=](cudaStream_t s, const RunCtx& ctx) {
        launch_frames_to_nhwc8(ctx.x, chw, frames_in, nullptr, DT_BF16, B * t_in, 1, c, h, w, ns, s);
      };
      pre.push_back(std::move(cv));
      Op z;
      z.name = "zero_h0_barriers";
      z.is_kernel = false;
      z.fn = [=](cudaStream_t s, const RunCtx&) { VPK_CUDA(cudaMemsetAsync(zero_block, 0, zero_bytes, s)); };
      body.push_back(std::move(z));
    }
    // ---------------- encoder, layer-major ----------------
    const void* in = frames_in;
    int in_h = h, in_w = w, in_c = cs;
    for (int n = 0; n < 3; ++n) {
      const std::string st = "encoder.stage" + std::to_string(n + 1) + ".conv.";
      const std::string rn = "encoder.rnn" + std::to_string(n + 1) + ".";
      const int mid = d.enc_c[2 * n], outc = d.enc_c[2 * n + 1];
      int oh, ow;
      const bool stem = n == 0 && conv_stem_supported(d.enc_conv_k[0], d.enc_conv_s[0], d.enc_conv_p[0], c, mid, in_h, in_w) &&
                        getenv("VPK_NO_STEM") == nullptr;
      if (stem) {
        std::vector<float> hb(hp(st + "bias"), hp(st + "bias") + mid);
        StemArgs sa{in, 0, B * t_in, in_h, in_w, c, d.enc_conv_s[0],
                    dev_f32(st + "stem.w", conv_stem_pack(hp(st + "weight"), mid, c, DT_BF16), stream),
                    dev_f32(st + "stem.b", hb, stream), mid, d.ef_act, xin[n], 0};
        const int ns = num_sms;
        Op op;
        op.name = st + "stem";
        op.fn = [=](cudaStream_t s, const RunCtx&) { launch_conv_stem(sa, ns, s); };
        body.push_back(std::move(op));
      } else {
        ConvArgs ca{st, B * t_in, in_h, in_w, in_c, mid, d.enc_conv_k[n], d.enc_conv_s[n], d.enc_conv_p[n], in,
                    hp(st + "weight"), hp(st + "bias"), d.ef_act, xin[n]};
        if (n == 0) ca.cin_w = c;
        conv_ops(conv_spec(ca, act, &oh, &ow));
      }
      LstmArgs la{rn, B, eh[n], ew[n], mid, outc, d.enc_rnn_k[n], nullptr, nullptr, nullptr, cbuf[n], hp(rn + "_conv.weight"),
                  hp(rn + "_conv.bias"), false, peep[0][n][0], peep[0][n][1], peep[0][n][2]};
      la.c4 = true;
      la.pp16 = peep16[0][n];
      if (!seq_layer(la, t_in, xin[n], hzero[n], B, 1, 0, hseq_e[n], true, n)) return false;
      in = hseq_e[n];
      in_h = eh[n];
      in_w = ew[n];
      in_c = outc;
    }
    // ---------------- forecaster, layer-major ----------------
    const void* fin = nullptr;                                          // rnn3 gets inputs=None
    for (int n = 0; n < 3; ++n) {
      const int idx = 3 - n, e = 2 - n;
      const std::string rn = "forecaster.rnn" + std::to_string(idx) + ".";
      const std::string st = "forecaster.stage" + std::to_string(idx) + ".deconv.";
      const int mid = d.dec_c[2 * n], outc = d.dec_c[2 * n + 1];
      LstmArgs la{rn + (fin ? "" : "h_only."), B, dh[n], dw[n], dec_in_c[n], mid, d.dec_rnn_k[n], nullptr, nullptr, nullptr, cbuf[e],
                  hp(rn + "_conv.weight"), hp(rn + "_conv.bias"), false, peep[1][idx - 1][0], peep[1][idx - 1][1], peep[1][idx - 1][2]};
      la.c4 = true;
      la.pp16 = peep16[1][idx - 1];
      // initial state = the encoder layer's final (h, c): slot t_in - 1 of its output sequence; c through cbuf[e]
      if (!seq_layer(la, pred, fin, hseq_e[e], B * t_in, t_in, t_in - 1, hseq_f[e], false, 3 + n)) return false;
      int oh, ow;
      DeconvArgs da{st, B * pred, dh[n], dw[n], mid, outc, d.dec_conv_k[n], d.dec_conv_s[n], d.dec_conv_p[n], 0,
                    hseq_f[e], hp(st + "weight"), hp(st + "bias"), d.ef_act, yseq[n]};
      if (n == 2) {
        da.name = st + "final_fused.";
        da.out = out_stage;
        da.nchw = true;
        da.oB_nchw = static_cast<long long>(c) * h * w;              // sample (b, t) = frame t of sequence b: [b, t, c, h, w]
        da.proj_n = c;
        const HostParam& fw = params.at("forecaster.stage1.final.weight");
        const HostParam& fb = params.at("forecaster.stage1.final.bias");
        da.proj_w = dev_f32("forecaster.stage1.final.weight", fw.data, stream);
        da.proj_b = dev_f32("forecaster.stage1.final.bias", fb.data, stream);
      }
      const long long tiles = static_cast<long long>(B) * pred * ((dh[n] + 15) / 16) * ((dw[n] + 7) / 8);
      // latency mode: one sub-pixel launch (2.25x the contraction) still beats four per-parity launches up to ~6 tiles per
      // SM (cfg 1, stage 2: 640 tiles, 1.0015 -> 0.9990 ms per rollout)
      bool subpix = deconv_subpix_ok(da) && tiles <= 6ll * num_sms;
      if (const char* sp_env = getenv("VPK_SUBPIX")) subpix = deconv_subpix_ok(da) && atoi(sp_env) != 0;
      if (subpix) conv_ops(deconv_subpix_spec(da, act, &oh, &ow));
      else conv_ops(deconv_spec(da, act, &oh, &ow));
      fin = yseq[n];
    }
    for (Op& o : pre) prog.pre.push_back(std::move(o));
    for (Op& o : body) prog.body.push_back(std::move(o));
    const size_t bytes = frame_px * c * sizeof(float) * B * pred;
    Op post;
    post.name = "copy_out";
    post.is_kernel = false;
    post.fn = [=](cudaStream_t s, const RunCtx& ctx) { VPK_CUDA(cudaMemcpyAsync(ctx.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s)); };
    prog.post.push_back(std::move(post));
    return true;
  }

  // sequence (persistent) form is for the small-batch latency mode: the CUDA-graph flag selects it (VPK_EF_SEQ=0: off)
  bool seq_candidate() const {
    const vpk_model_desc& d = desc;
    const char* env = getenv("VPK_EF_SEQ");
    const bool fused_final_ok = d.dec_conv_s[2] == 1 && d.img_c <= 4 && d.dec_c[4] % 8 == 0 && d.dec_c[5] % 8 == 0 &&
                                d.dec_c[5] <= 64 && d.final_conv_c == d.dec_c[5];
    bool ch8 = true;
    for (int i = 0; i < 6; ++i) ch8 = ch8 && d.enc_c[i] % 8 == 0 && d.dec_c[i] % 8 == 0;
    return desc.use_cuda_graph != 0 && dtype == DT_BF16 && backend == 0 && d.img_c <= 8 && fused_final_ok && ch8 &&
           (env == nullptr || atoi(env) != 0) && getenv("VPK_TC_HALO") == nullptr;
  }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    if (seq_candidate()) {
      if (measure) {            // workspace: the larger of the two layouts
        Arena a2;
        build_seq(prog, a2, B, t_in, pred, true, stream);
        build_steps(prog, arena, B, t_in, pred, true, stream);
        arena.off = std::max(arena.off, a2.off);
        return;
      }
      const size_t off0 = arena.off;
      if (build_seq(prog, arena, B, t_in, pred, false, stream)) {
        seq_built = true;
        return;
      }
      arena.off = off0;         // a layer does not fit: per-step launches
      prog.pre.clear();
      prog.body.clear();
      prog.post.clear();
    }
    seq_built = false;
    build_steps(prog, arena, B, t_in, pred, measure, stream);
  }

  void build_steps(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) {
    const vpk_model_desc& d = desc;
    const ActInfo act{dtype, esize()};
    const int esz = esize();
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const size_t frame_px = static_cast<size_t>(B) * h * w;

    // tcgen05 path: frames are stored with 8 channels per pixel (zero padded) so that the stem conv is TMA-addressable
    const bool pad8 = dtype == DT_BF16 && backend == 0 && c <= 8 && getenv("VPK_NO_PAD8") == nullptr;
    const int cs = pad8 ? 8 : c;
    char* frames_in = static_cast<char*>(arena.alloc(frame_px * cs * esz * t_in));
    float* out_stage = static_cast<float*>(arena.alloc(frame_px * c * sizeof(float) * pred));

    void* xin[3];
    void* hbuf[3][2];
    float* cbuf[3];
    for (int n = 0; n < 3; ++n) {
      const size_t px = static_cast<size_t>(B) * eh[n] * ew[n];
      xin[n] = arena.alloc(px * d.enc_c[2 * n] * esz);
      hbuf[n][0] = arena.alloc(px * d.enc_c[2 * n + 1] * esz);
      hbuf[n][1] = arena.alloc(px * d.enc_c[2 * n + 1] * esz);
      cbuf[n] = static_cast<float*>(arena.alloc(px * d.enc_c[2 * n + 1] * sizeof(float)));
    }
    void* ybuf[3];
    for (int n = 0; n < 3; ++n)
      ybuf[n] = arena.alloc(static_cast<size_t>(B) * dh[n + 1] * dw[n + 1] * d.dec_c[2 * n + 1] * esz);

    // peepholes (device fp32 [H,W,C]); all three absent => plain gates
    const float* peep[2][3][3] = {};
    const void* peep16[2][3] = {};
    if (!measure) {
      for (int side = 0; side < 2; ++side)
        for (int n = 0; n < 3; ++n) {
          const std::string rn = std::string(side == 0 ? "encoder.rnn" : "forecaster.rnn") + std::to_string(n + 1) + ".";
          const int C = d.enc_c[2 * n + 1];
          if (!has(rn + "Wci") && !has(rn + "Wcf") && !has(rn + "Wco")) continue;
          if (dev_env("VPK_EXP_NO_PEEPHOLES") != nullptr) continue;   // perf experiments only: changes the results
          const char* names[3] = {"Wci", "Wcf", "Wco"};
          for (int k = 0; k < 3; ++k)
            peep[side][n][k] = dev_f32(rn + names[k], peephole_packed(rn + names[k], C, eh[n], ew[n]), stream);
          if (dtype == DT_BF16 && C % 8 == 0)
            peep16[side][n] = dev_f32(rn + "peepholes.bf16", peephole_bf16_packed(rn, C, eh[n], ew[n]), stream);
        }
    }

    if (!measure) {
      const int ns = num_sms, dt = dtype;
      if (!streams_input()) {       // CUDA-graph replay: the conversion reads the call's input pointer, so it stays outside
        Op pre;
        pre.name = "frames_to_nhwc";
        pre.fn = [=](cudaStream_t s, const RunCtx& ctx) {
          if (pad8)
            launch_frames_to_nhwc8(ctx.x, static_cast<long long>(t_in) * c * h * w, frames_in, nullptr, DT_BF16, B, t_in, c, h, w, ns, s);
          else
            launch_frames_to_nhwc(ctx.x, frames_in, dt, B, t_in, c, h, w, ns, s);
        };
        prog.pre.push_back(std::move(pre));
      }
      for (int n = 0; n < 3; ++n) {
        const size_t px = static_cast<size_t>(B) * eh[n] * ew[n];
        add_memset(prog, hbuf[n][0], px * d.enc_c[2 * n + 1] * esz, "zero_h");
        add_memset(prog, cbuf[n], px * d.enc_c[2 * n + 1] * sizeof(float), "zero_c");
      }
    }

    int par[3] = {0, 0, 0};
    // tcgen05 path only (the CUDA-core kernels keep the two launches): stride-1 last deconv, <= 4 image channels
    const bool fuse_final = dtype == DT_BF16 && backend == 0 && d.dec_conv_s[2] == 1 && c <= 4 &&
                            d.dec_c[4] % 8 == 0 && d.dec_c[5] % 8 == 0 && d.dec_c[5] <= 64 && d.final_conv_c == d.dec_c[5] &&
                            getenv("VPK_EF_NO_FUSE") == nullptr;
    // ------------------------------------------ encoder (ef_blocks.py:67-82) ---------------------------------
    for (int t = 0; t < t_in; ++t) {
      const void* in = frames_in + static_cast<size_t>(t) * frame_px * cs * esz;
      if (!measure && streams_input()) {
        // input frame t is converted right before the step that reads it; under the host entry this op waits for the
        // frame's own host-to-device copy only
        const int ns = num_sms, dt = dtype;
        char* dst = frames_in + static_cast<size_t>(t) * frame_px * cs * esz;
        const long long bstride = static_cast<long long>(t_in) * c * h * w, foff = static_cast<long long>(t) * c * h * w;
        Op cv;
        cv.name = "frames_to_nhwc";
        cv.needs_input = t;
        cv.fn = [=](cudaStream_t s, const RunCtx& ctx) {
          if (pad8) launch_frames_to_nhwc8(ctx.x + foff, bstride, dst, nullptr, DT_BF16, B, 1, c, h, w, ns, s);
          else launch_frames_to_nhwc_strided(ctx.x + foff, bstride, dst, dt, B, 1, c, h, w, ns, s);
        };
        prog.body.push_back(std::move(cv));
      }
      int in_h = h, in_w = w, in_c = cs;
      for (int n = 0; n < 3; ++n) {
        const std::string st = "encoder.stage" + std::to_string(n + 1) + ".conv.";
        const std::string rn = "encoder.rnn" + std::to_string(n + 1) + ".";
        const int mid = d.enc_c[2 * n], outc = d.enc_c[2 * n + 1];
        int oh, ow;
        ConvArgs ca{st, B, in_h, in_w, in_c, mid, d.enc_conv_k[n], d.enc_conv_s[n], d.enc_conv_p[n], in,
                    hp(st + "weight"), hp(st + "bias"), d.ef_act, xin[n]};
        if (n == 0) ca.cin_w = c;
        // image-channel stem (K = 9 * c): HBM-bound, direct CUDA-core kernel instead of a tensor-core tile
        const bool stem = n == 0 && pad8 && dtype == DT_BF16 && backend == 0 && getenv("VPK_NO_STEM") == nullptr &&
                          (d.ef_act == ACT_LEAKY || d.ef_act == ACT_NONE || d.ef_act == ACT_RELU) &&
                          conv_stem_supported(d.enc_conv_k[0], d.enc_conv_s[0], d.enc_conv_p[0], c, mid, in_h, in_w);
        if (stem) {
          oh = (in_h + 2 - 3) / d.enc_conv_s[0] + 1;
          ow = (in_w + 2 - 3) / d.enc_conv_s[0] + 1;
          if (!measure) {
            std::vector<float> hb(hp(st + "bias"), hp(st + "bias") + mid);
            StemArgs sa{in, 0, B, in_h, in_w, c, d.enc_conv_s[0],
                        dev_f32(st + "stem.w", conv_stem_pack(hp(st + "weight"), mid, c, DT_BF16), stream),
                        dev_f32(st + "stem.b", hb, stream), mid, d.ef_act, xin[n], 0};
            const int ns = num_sms;
            Op op;
            op.name = st + "stem";
            op.flops = 2.0 * static_cast<double>(B) * oh * ow * mid * 9 * c;
            op.fn = [=](cudaStream_t s, const RunCtx&) { launch_conv_stem(sa, ns, s); };
            prog.body.push_back(std::move(op));
          }
        } else
        add_conv(prog, conv_spec(ca, act, &oh, &ow), measure, stream);
        VPK_REQUIRE(oh == eh[n] && ow == ew[n], "encoder stage size mismatch");
        LstmArgs la{rn, B, eh[n], ew[n], mid, outc, d.enc_rnn_k[n], xin[n], hbuf[n][par[n]], hbuf[n][par[n] ^ 1],
                    cbuf[n], hp(rn + "_conv.weight"), hp(rn + "_conv.bias"), false,
                    peep[0][n][0], peep[0][n][1], peep[0][n][2]};
        la.c4 = true;
        la.pp16 = peep16[0][n];
        add_conv(prog, lstm_spec(la, act), measure, stream);
        par[n] ^= 1;
        in = hbuf[n][par[n]];
        in_h = eh[n];
        in_w = ew[n];
        in_c = outc;
      }
    }
    // ------------------------------------------ forecaster (ef_blocks.py:100-114) ----------------------------
    for (int t = 0; t < pred; ++t) {
      const void* in = nullptr;   // rnn3 gets inputs=None
      for (int n = 0; n < 3; ++n) {
        const int idx = 3 - n, e = 2 - n;   // forecaster.rnn{idx} continues encoder state e
        const std::string rn = "forecaster.rnn" + std::to_string(idx) + ".";
        const std::string st = "forecaster.stage" + std::to_string(idx) + ".deconv.";
        const int mid = d.dec_c[2 * n], outc = d.dec_c[2 * n + 1];
        // the packed weights of rnn3 hold only the h-side columns: name them apart from a (hypothetical) full pack
        LstmArgs la{rn + (in ? "" : "h_only."), B, dh[n], dw[n], dec_in_c[n], mid, d.dec_rnn_k[n], in,
                    hbuf[e][par[e]], hbuf[e][par[e] ^ 1], cbuf[e], hp(rn + "_conv.weight"), hp(rn + "_conv.bias"),
                    false, peep[1][idx - 1][0], peep[1][idx - 1][1], peep[1][idx - 1][2]};
        la.c4 = true;
        la.pp16 = peep16[1][idx - 1];
        add_conv(prog, lstm_spec(la, act), measure, stream);
        par[e] ^= 1;
        int oh, ow;
        DeconvArgs da{st, B, dh[n], dw[n], mid, outc, d.dec_conv_k[n], d.dec_conv_s[n], d.dec_conv_p[n], 0,
                      hbuf[e][par[e]], hp(st + "weight"), hp(st + "bias"), d.ef_act, ybuf[n]};
        if (n == 2 && fuse_final) {
          // last deconv + LeakyReLU + final 1x1 conv (ef_conv_lstm.py:99-104) in one launch: the 16-channel map stays
          // in registers and the fp32 NCHW frame t of the staging tensor is written directly
          da.name = st + "final_fused.";
          da.out = out_stage + static_cast<size_t>(t) * c * h * w;
          da.nchw = true;
          da.oB_nchw = static_cast<long long>(pred) * c * h * w;
          da.proj_n = c;
          if (!measure) {
            const HostParam& fw = params.at("forecaster.stage1.final.weight");
            const HostParam& fb = params.at("forecaster.stage1.final.bias");
            da.proj_w = dev_f32("forecaster.stage1.final.weight", fw.data, stream);
            da.proj_b = dev_f32("forecaster.stage1.final.bias", fb.data, stream);
          }
        }
        // small batches: the four per-parity launches of a stride-2 deconv are pure launch latency -- one sub-pixel
        // launch instead (denser contraction, so only while there are fewer tiles than ~2 per SM)
        const long long tiles = static_cast<long long>(B) * ((dh[n] + 15) / 16) * ((dw[n] + 7) / 8);
        const char* sp_env = getenv("VPK_SUBPIX");
        const char* halo_env = getenv("VPK_TC_HALO");     // tests force the per-tap kernels with VPK_TC_HALO=0
        const bool subpix = dtype == DT_BF16 && backend == 0 && deconv_subpix_ok(da) && (halo_env == nullptr || atoi(halo_env) != 0) &&
                            (sp_env ? atoi(sp_env) != 0 : tiles <= 2ll * num_sms);
        if (subpix) add_conv(prog, deconv_subpix_spec(da, act, &oh, &ow), measure, stream);
        else add_conv(prog, deconv_spec(da, act, &oh, &ow), measure, stream);
        VPK_REQUIRE(oh == dh[n + 1] && ow == dw[n + 1], "forecaster stage size mismatch");
        in = ybuf[n];
      }
      auto mark_frame = [&]() {      // the op just added completes predicted frame t (host entry: starts its D2H)
        if (measure || prog.body.empty()) return;
        Op& o = prog.body.back();
        o.frame = t;
        o.frame_src = out_stage + static_cast<size_t>(t) * c * h * w;
        o.frame_pitch = static_cast<long long>(pred) * c * h * w;
        o.frame_elems = static_cast<long long>(c) * h * w;
      };
      if (fuse_final) {
        mark_frame();
        continue;
      }
      // identity + final 1x1 conv (ef_conv_lstm.py:99-104), written as fp32 NCHW frame t of the staging tensor
      int oh, ow;
      ConvArgs fa{"forecaster.stage1.final.", B, h, w, d.final_conv_c, c, 1, 1, 0, ybuf[2],
                  hp("forecaster.stage1.final.weight"), hp("forecaster.stage1.final.bias"), ACT_NONE,
                  out_stage + static_cast<size_t>(t) * c * h * w};
      fa.f32_strided = true;
      fa.oB = static_cast<long long>(pred) * c * h * w;
      fa.oC = static_cast<long long>(h) * w;
      fa.oY = w;
      fa.oX = 1;
      add_conv(prog, conv_spec(fa, act, &oh, &ow), measure, stream);
      mark_frame();
    }
    if (!measure) {
      const size_t bytes = frame_px * c * sizeof(float) * pred;
      Op post;
      post.name = "copy_out";
      post.is_kernel = false;
      post.fn = [=](cudaStream_t s, const RunCtx& ctx) {
        if (ctx.on_frame != nullptr) return;     // host entry with frame streaming: every frame has been copied already
        VPK_CUDA(cudaMemcpyAsync(ctx.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s));
      };
      prog.post.push_back(std::move(post));
    }
  }

  bool streams_input() const override { return !desc.use_cuda_graph && getenv("VPK_NO_INPUT_STREAM") == nullptr; }

 private:
  int eh[3], ew[3], dh[4], dw[4], dec_in_c[3];
  bool seq_built = false;
};

}  // namespace

Model* make_ef_convlstm(const vpk_model_desc& d) { return new EfConvLstm(d); }

}  // namespace vpk
