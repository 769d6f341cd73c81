// trajgru: Encoder-Forecaster with TrajGRU recurrent blocks (reference: models/precipitation_nowcasting/ef_traj_gru.py:8-119,
// ef_blocks.py:52-187, model_blocks/traj_gru.py:70-214).
//
// Same skeleton and the same time-major schedule as convlstm-shi (model_ef.cu): stage conv -> recurrent block per encoder
// layer and step; recurrent block -> stage deconv per forecaster layer and step, the top forecaster block without an input.
// One TrajGRU step (traj_gru.py:187-211, zoneout 0):
//   i2h   = Conv k x k (x)                                   raw fp32 [3C]     (skipped without an input)
//   f     = act(Conv5x5(x) + Conv5x5(h))                     ONE launch, two sources with their own weights (:137-144)
//   flows = Conv5x5(f)                                       fp32 [2L]
//   warped= cat_l bilinear_warp(h, -flow_l)                  trajgru_warp_kernel, from the fp32 state (:192-195)
//   h2h   = Conv1x1(warped)                                  raw fp32 [3C], K = L * C
//   h'    = u * h + (1 - u) * act(i2h_2 + r * h2h_2)         trajgru_gates_kernel
// Every conv is a generalised-conv launch (tcgen05 in 16-bit mode); the warp is a gather from the fp32 state, so operand
// rounding enters the sampled VALUES only through the bf16 copy handed to the 1x1 conv, never the sampling positions'
// source.
#include <cstdlib>

#include "builders.h"
#include "elementwise.h"
#include "model.h"

namespace vpk {

namespace {

class EfTrajGru : public Model {
 public:
  explicit EfTrajGru(const vpk_model_desc& d) : Model(d) {
    VPK_REQUIRE(d.img_c > 0 && d.img_h > 0 && d.img_w > 0, "bad img_shape");
    int hh = d.img_h, ww = d.img_w;
    for (int n = 0; n < 3; ++n) {     // state sizes as in model_ef.cu (ef_blocks.py:145-167)
      VPK_REQUIRE(d.enc_conv_s[n] == 1 || d.enc_conv_s[n] == 2, "enc_conv_s must be 1 or 2");
      hh = (hh + 2 * d.enc_conv_p[n] - d.enc_conv_k[n]) / d.enc_conv_s[n] + 1;
      ww = (ww + 2 * d.enc_conv_p[n] - d.enc_conv_k[n]) / d.enc_conv_s[n] + 1;
      eh[n] = hh;
      ew[n] = ww;
    }
    dh[0] = hh;
    dw[0] = ww;
    for (int n = 0; n < 3; ++n) {
      hh = (hh - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + (d.dec_conv_k[n] - 1) + d.dec_conv_p[n];
      ww = (ww - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + (d.dec_conv_k[n] - 1) + d.dec_conv_p[n];
      const int th = (dh[n] - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + d.dec_conv_k[n];
      const int tw = (dw[n] - 1) * d.dec_conv_s[n] - 2 * d.dec_conv_p[n] + d.dec_conv_k[n];
      VPK_REQUIRE(th == hh && tw == ww, "decoder conv hyper-parameters give inconsistent sizes");
      dh[n + 1] = hh;
      dw[n + 1] = ww;
    }
    VPK_REQUIRE(dh[3] == d.img_h && dw[3] == d.img_w, "model layer hyper-parameters yield wrong output size");
    auto rnn = [&](const std::string& rn, int in_c, int C, int L, int k) {
      VPK_REQUIRE(L >= 1 && L <= 32 && C % 4 == 0 && k % 2 == 1, "bad TrajGRU hyper-parameters");
      declare(rn + "i2h.weight", {3 * C, in_c, k, k});
      declare(rn + "i2h.bias", {3 * C});
      declare(rn + "i2f_conv1.weight", {32, in_c, 5, 5});
      declare(rn + "i2f_conv1.bias", {32});
      declare(rn + "h2f_conv1.weight", {32, C, 5, 5});
      declare(rn + "h2f_conv1.bias", {32});
      declare(rn + "flows_conv.weight", {2 * L, 32, 5, 5});
      declare(rn + "flows_conv.bias", {2 * L});
      declare(rn + "ret.weight", {3 * C, C * L, 1, 1});
      declare(rn + "ret.bias", {3 * C});
    };
    int in_c = d.img_c;
    for (int n = 0; n < 3; ++n) {
      const std::string st = "encoder.stage" + std::to_string(n + 1) + ".conv.";
      const int mid = d.enc_c[2 * n], outc = d.enc_c[2 * n + 1];
      declare(st + "weight", {mid, in_c, d.enc_conv_k[n], d.enc_conv_k[n]});
      declare(st + "bias", {mid});
      rnn("encoder.rnn" + std::to_string(n + 1) + ".", mid, outc, d.enc_rnn_L[n], d.enc_rnn_k[n]);
      in_c = outc;
    }
    for (int n = 0; n < 3; ++n) {
      const int idx = 3 - n;
      const std::string st = "forecaster.stage" + std::to_string(idx) + ".deconv.";
      const int mid = d.dec_c[2 * n], outc = d.dec_c[2 * n + 1];
      VPK_REQUIRE(dh[n] == eh[2 - n] && dw[n] == ew[2 - n] && mid == d.enc_c[2 * (2 - n) + 1], "encoder / forecaster states differ");
      rnn("forecaster.rnn" + std::to_string(idx) + ".", in_c, mid, d.dec_rnn_L[n], d.dec_rnn_k[n]);
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
  int default_microbatch() const override { return 64; }

  struct CellBufs {
    float *i2h, *flows, *h2h;
    void *f1, *warped;
    int fpix;
  };

  // one TrajGRU step: (x or nullptr, h_prev fp32 + activation copy) -> (h_next fp32 + activation copy)
  void add_cell(Program& prog, const std::string& rn, int B, int H, int W, int Cin, int C, int L, int ki, const void* x,
                const float* h32, const void* h_act, float* h32_out, void* h_act_out, const CellBufs& cb, const ActInfo& act,
                bool measure, cudaStream_t stream) {
    int oh, ow;
    if (x != nullptr) {
      ConvArgs a{rn + "i2h.", B, H, W, Cin, 3 * C, ki, 1, ki / 2, x, hp(rn + "i2h.weight"), hp(rn + "i2h.bias"), ACT_NONE, cb.i2h};
      a.out_f32_dense = true;
      ConvSpec sp = conv_spec(a, act, &oh, &ow);
      sp.is_gate_gemm = true;
      add_conv(prog, sp, measure, stream);
    }
    {   // f = act(i2f_conv1(x) + h2f_conv1(h)): two sources, two weight tensors, summed biases
      ConvSpec sp;
      sp.name = rn + (x ? "flow1.xh." : "flow1.h.");
      sp.B = B;
      sp.G = 1;
      sp.C = 32;
      std::vector<ConvInput> ins;
      if (x != nullptr) {
        WeightRef wi;
        wi.w = hp(rn + "i2f_conv1.weight");
        wi.O = 32;
        wi.I = Cin;
        wi.KH = wi.KW = 5;
        sp.wrefs.push_back(wi);
        BiasRef bi;
        bi.b = hp(rn + "i2f_conv1.bias");
        sp.biases.push_back(bi);
        ins.push_back(ConvInput{make_view(x, H, W, Cin), 0, 0});
      }
      WeightRef wh;
      wh.w = hp(rn + "h2f_conv1.weight");
      wh.O = 32;
      wh.I = C;
      wh.KH = wh.KW = 5;
      sp.wrefs.push_back(wh);
      BiasRef bh;
      bh.b = hp(rn + "h2f_conv1.bias");
      sp.biases.push_back(bh);
      ins.push_back(ConvInput{make_view(h_act, H, W, C), static_cast<int>(sp.wrefs.size()) - 1, 0});
      lower_conv(sp, 5, 1, 2, ins, H, W, act.esize, &oh, &ow);
      EpiParams& e = sp.phases[0].epi;
      e.kind = EPI_BIAS_ACT;
      e.act = desc.ef_act;
      dense_out(e, cb.f1, H, W, 32);
      add_conv(prog, sp, measure, stream);
    }
    {
      ConvArgs a{rn + "flows_conv.", B, H, W, 32, 2 * L, 5, 1, 2, cb.f1, hp(rn + "flows_conv.weight"), hp(rn + "flows_conv.bias"),
                 ACT_NONE, cb.flows};
      a.out_f32_dense = true;
      a.out_pix = cb.fpix;
      add_conv(prog, conv_spec(a, act, &oh, &ow), measure, stream);
    }
    const int ns = num_sms, dt = act.dtype, fpix = cb.fpix;
    if (!measure) {
      Op op;
      op.name = rn + "warp";
      float* flows = cb.flows;
      void* warped = cb.warped;
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_trajgru_warp(h32, flows, fpix, warped, dt, B, H, W, C, L, ns, s); };
      prog.body.push_back(std::move(op));
    }
    {
      ConvArgs a{rn + "ret.", B, H, W, L * C, 3 * C, 1, 1, 0, cb.warped, hp(rn + "ret.weight"), hp(rn + "ret.bias"), ACT_NONE, cb.h2h};
      a.out_f32_dense = true;
      ConvSpec sp = conv_spec(a, act, &oh, &ow);
      sp.is_gate_gemm = true;
      add_conv(prog, sp, measure, stream);
    }
    if (!measure) {
      Op op;
      op.name = rn + "gates";
      const float* i2h = x ? cb.i2h : nullptr;
      const float* h2h = cb.h2h;
      const long long P = static_cast<long long>(B) * H * W;
      const int actk = desc.ef_act;
      op.fn = [=](cudaStream_t s, const RunCtx&) { launch_trajgru_gates(i2h, h2h, h32, h32_out, h_act_out, dt, P, C, actk, ns, s); };
      prog.body.push_back(std::move(op));
    }
  }

  void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) override {
    const vpk_model_desc& d = desc;
    const ActInfo act{dtype, esize()};
    const int esz = esize();
    const bool f32 = dtype == DT_F32;
    const int c = d.img_c, h = d.img_h, w = d.img_w;
    const size_t frame_px = static_cast<size_t>(B) * h * w;
    const bool pad8 = !f32 && backend == 0 && c <= 8;
    const int cs = pad8 ? 8 : c;
    char* frames_in = static_cast<char*>(arena.alloc(frame_px * cs * esz * t_in));
    float* out_stage = static_cast<float*>(arena.alloc(frame_px * c * sizeof(float) * pred));
    void* xin[3];
    float* h32[3][2];
    void* hact[3][2];
    CellBufs cb[3];
    for (int n = 0; n < 3; ++n) {
      const size_t px = static_cast<size_t>(B) * eh[n] * ew[n];
      const int C = d.enc_c[2 * n + 1];
      const int Lmax = std::max(d.enc_rnn_L[n], d.dec_rnn_L[2 - n]);
      xin[n] = arena.alloc(px * d.enc_c[2 * n] * esz);
      for (int q = 0; q < 2; ++q) {
        h32[n][q] = static_cast<float*>(arena.alloc(px * C * sizeof(float)));
        hact[n][q] = f32 ? static_cast<void*>(h32[n][q]) : arena.alloc(px * C * esz);
      }
      cb[n].fpix = (2 * Lmax + 3) / 4 * 4;
      cb[n].i2h = static_cast<float*>(arena.alloc(px * 3 * C * sizeof(float)));
      cb[n].h2h = static_cast<float*>(arena.alloc(px * 3 * C * sizeof(float)));
      cb[n].flows = static_cast<float*>(arena.alloc(px * cb[n].fpix * sizeof(float)));
      cb[n].f1 = arena.alloc(px * 32 * esz);
      cb[n].warped = arena.alloc(px * static_cast<size_t>(Lmax) * C * esz);
    }
    void* ybuf[3];
    for (int n = 0; n < 3; ++n) ybuf[n] = arena.alloc(static_cast<size_t>(B) * dh[n + 1] * dw[n + 1] * d.dec_c[2 * n + 1] * esz);

    if (!measure) {
      const int ns = num_sms, dt = dtype;
      Op pre;
      pre.name = "frames_to_nhwc";
      pre.fn = [=](cudaStream_t s, const RunCtx& ctx) {
        if (pad8) launch_frames_to_nhwc8(ctx.x, static_cast<long long>(t_in) * c * h * w, frames_in, nullptr, DT_BF16, B, t_in, c, h, w, ns, s);
        else launch_frames_to_nhwc(ctx.x, frames_in, dt, B, t_in, c, h, w, ns, s);
      };
      prog.pre.push_back(std::move(pre));
      for (int n = 0; n < 3; ++n) {
        const size_t px = static_cast<size_t>(B) * eh[n] * ew[n];
        add_memset(prog, h32[n][0], px * d.enc_c[2 * n + 1] * sizeof(float), "zero_h");
        if (!f32) add_memset(prog, hact[n][0], px * d.enc_c[2 * n + 1] * esz, "zero_h_act");
      }
    }
    int par[3] = {0, 0, 0};
    // ---------------- encoder ----------------
    for (int t = 0; t < t_in; ++t) {
      const void* in = frames_in + static_cast<size_t>(t) * frame_px * cs * esz;
      int in_h = h, in_w = w, in_c = cs;
      for (int n = 0; n < 3; ++n) {
        const std::string st = "encoder.stage" + std::to_string(n + 1) + ".conv.";
        const std::string rn = "encoder.rnn" + std::to_string(n + 1) + ".";
        const int mid = d.enc_c[2 * n], outc = d.enc_c[2 * n + 1];
        int oh, ow;
        ConvArgs ca{st, B, in_h, in_w, in_c, mid, d.enc_conv_k[n], d.enc_conv_s[n], d.enc_conv_p[n], in, hp(st + "weight"),
                    hp(st + "bias"), d.ef_act, xin[n]};
        if (n == 0) ca.cin_w = c;
        add_conv(prog, conv_spec(ca, act, &oh, &ow), measure, stream);
        VPK_REQUIRE(oh == eh[n] && ow == ew[n], "encoder stage size mismatch");
        add_cell(prog, rn, B, eh[n], ew[n], mid, outc, d.enc_rnn_L[n], d.enc_rnn_k[n], xin[n], h32[n][par[n]], hact[n][par[n]],
                 h32[n][par[n] ^ 1], hact[n][par[n] ^ 1], cb[n], act, measure, stream);
        par[n] ^= 1;
        in = hact[n][par[n]];
        in_h = eh[n];
        in_w = ew[n];
        in_c = outc;
      }
    }
    // ---------------- forecaster ----------------
    for (int t = 0; t < pred; ++t) {
      const void* in = nullptr;
      for (int n = 0; n < 3; ++n) {
        const int idx = 3 - n, e = 2 - n;
        const std::string rn = "forecaster.rnn" + std::to_string(idx) + ".";
        const std::string st = "forecaster.stage" + std::to_string(idx) + ".deconv.";
        const int mid = d.dec_c[2 * n], outc = d.dec_c[2 * n + 1];
        add_cell(prog, rn, B, dh[n], dw[n], dec_in_c[n], mid, d.dec_rnn_L[n], d.dec_rnn_k[n], in, h32[e][par[e]], hact[e][par[e]],
                 h32[e][par[e] ^ 1], hact[e][par[e] ^ 1], cb[e], act, measure, stream);
        par[e] ^= 1;
        int oh, ow;
        DeconvArgs da{st, B, dh[n], dw[n], mid, outc, d.dec_conv_k[n], d.dec_conv_s[n], d.dec_conv_p[n], 0, hact[e][par[e]],
                      hp(st + "weight"), hp(st + "bias"), d.ef_act, ybuf[n]};
        add_conv(prog, deconv_spec(da, act, &oh, &ow), measure, stream);
        VPK_REQUIRE(oh == dh[n + 1] && ow == dw[n + 1], "forecaster stage size mismatch");
        in = ybuf[n];
      }
      int oh, ow;
      ConvArgs fa{"forecaster.stage1.final.", B, h, w, d.final_conv_c, c, 1, 1, 0, ybuf[2], hp("forecaster.stage1.final.weight"),
                  hp("forecaster.stage1.final.bias"), ACT_NONE, out_stage + static_cast<size_t>(t) * c * h * w};
      fa.f32_strided = true;
      fa.oB = static_cast<long long>(pred) * c * h * w;
      fa.oC = static_cast<long long>(h) * w;
      fa.oY = w;
      fa.oX = 1;
      add_conv(prog, conv_spec(fa, act, &oh, &ow), measure, stream);
      if (!measure && !prog.body.empty()) {
        Op& o = prog.body.back();
        o.frame = t;
        o.frame_src = out_stage + static_cast<size_t>(t) * c * h * w;
        o.frame_pitch = static_cast<long long>(pred) * c * h * w;
        o.frame_elems = static_cast<long long>(c) * h * w;
      }
    }
    if (!measure) {
      const size_t bytes = frame_px * c * sizeof(float) * pred;
      Op post;
      post.name = "copy_out";
      post.is_kernel = false;
      post.fn = [=](cudaStream_t s, const RunCtx& ctx) {
        if (ctx.on_frame != nullptr) return;
        VPK_CUDA(cudaMemcpyAsync(ctx.out, out_stage, bytes, cudaMemcpyDeviceToDevice, s));
      };
      prog.post.push_back(std::move(post));
    }
  }

 private:
  int eh[3], ew[3], dh[4], dw[4], dec_in_c[3];
};

}  // namespace

Model* make_ef_trajgru(const vpk_model_desc& d) { return new EfTrajGru(d); }

}  // namespace vpk
