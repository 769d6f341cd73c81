// extern "C" surface of libvpk.so (include/vpk.h).  No C++ exception crosses this file.
#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/vpk.h"
#include "cells.h"
#include "elementwise.h"
#include "model.h"

namespace {

thread_local std::string g_last_error;

template <typename F> int guarded(F&& f) {
  try {
    f();
    return VPK_OK;
  } catch (const vpk::Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VPK_ERR_INVALID;
  } catch (...) {
    g_last_error = "unknown error";
    return VPK_ERR_INVALID;
  }
}

}  // namespace

struct vpk_model {
  vpk::Model* impl;
};
struct vpk_cell {
  vpk::Cell* impl;
};

extern "C" {

int vpk_metric_partial_sums(const float* pred, const float* target, int32_t batch, int32_t frames, int64_t chw,
                            double* scratch, double* out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(pred && target && scratch && out, "vpk_metric_partial_sums: null pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      VPK_THROW(VPK_ERR_CUDA, "no CUDA device: libvpk has no CPU fallback");
    }
    vpk::launch_metric_partial_sums(pred, target, batch, frames, chw, scratch, out, static_cast<cudaStream_t>(stream));
  });
}

int64_t vpk_metric_ssim_scratch_elems(int32_t batch, int32_t frames, int32_t c, int32_t h, int32_t w) {
  return vpk::metric_ssim_scratch_elems(batch, frames, c, h, w);
}

int vpk_metric_ssim_sums(const float* pred, const float* target, int32_t batch, int32_t frames, int32_t c, int32_t h,
                         int32_t w, double* scratch, double* out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(pred && target && scratch && out, "vpk_metric_ssim_sums: null pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      VPK_THROW(VPK_ERR_CUDA, "no CUDA device: libvpk has no CPU fallback");
    }
    vpk::launch_metric_ssim_sums(pred, target, batch, frames, c, h, w, scratch, out, static_cast<cudaStream_t>(stream));
  });
}

const char* vpk_last_error(void) { return g_last_error.c_str(); }
const char* vpk_version(void) { return "libvpk 0.1 (sm_100a)"; }

int vpk_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int vpk_model_create(const vpk_model_desc* desc, vpk_model** out) {
  return guarded([&] {
    VPK_REQUIRE(desc != nullptr && out != nullptr, "null argument");
    vpk::Model* m = nullptr;
    switch (desc->kind) {
      case VPK_MODEL_CONVLSTM_SHI: m = vpk::make_ef_convlstm(*desc); break;
      case VPK_MODEL_PREDRNN_PP: m = vpk::make_predrnn(*desc); break;
      case VPK_MODEL_PHY: m = vpk::make_phydnet(*desc, false); break;
      case VPK_MODEL_CONVLSTM_BRANCH: m = vpk::make_phydnet(*desc, true); break;
      case VPK_MODEL_ST_PHY: m = vpk::make_stphy(*desc); break;
      case VPK_MODEL_TRAJGRU: m = vpk::make_ef_trajgru(*desc); break;
      case VPK_MODEL_PREDRNN_PP_CAUSAL: m = vpk::make_predrnnpp_causal(*desc); break;
      default: VPK_THROW(VPK_ERR_INVALID, "unknown model kind");
    }
    *out = new vpk_model{m};
  });
}

int vpk_model_set_param(vpk_model* m, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
  return guarded([&] {
    VPK_REQUIRE(m && key && shape, "null argument");
    m->impl->set_param(key, data, shape, ndim);
  });
}

int vpk_model_num_params(vpk_model* m, int32_t* n) {
  return guarded([&] {
    VPK_REQUIRE(m && n, "null argument");
    *n = static_cast<int32_t>(m->impl->keys.size());
  });
}

int vpk_model_param_info(vpk_model* m, int32_t i, const char** key, int64_t* shape4, int32_t* ndim) {
  return guarded([&] {
    VPK_REQUIRE(m && key && shape4 && ndim, "null argument");
    VPK_REQUIRE(i >= 0 && i < static_cast<int32_t>(m->impl->keys.size()), "parameter index out of range");
    const std::string& k = m->impl->keys[i];
    const vpk::HostParam& p = m->impl->params.at(k);
    *key = k.c_str();
    *ndim = static_cast<int32_t>(p.shape.size());
    for (size_t j = 0; j < p.shape.size() && j < 4; ++j) shape4[j] = p.shape[j];
  });
}

int vpk_model_finalize(vpk_model* m, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->finalize(static_cast<cudaStream_t>(stream));
  });
}

int vpk_model_workspace_bytes(vpk_model* m, int32_t batch, int32_t t_in, int32_t pred_frames, size_t* bytes) {
  return guarded([&] {
    VPK_REQUIRE(m && bytes, "null argument");
    *bytes = m->impl->workspace_bytes(batch, t_in, pred_frames);
  });
}

int vpk_model_forward(vpk_model* m, const float* x, int32_t batch, int32_t t_in, int32_t pred_frames, float* out,
                      float* aux, void* workspace, size_t workspace_bytes, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->forward(x, batch, t_in, pred_frames, out, aux, workspace, workspace_bytes,
                     static_cast<cudaStream_t>(stream));
  });
}

int vpk_model_forward_host(vpk_model* m, const float* x_host, int32_t batch, int32_t t_in, int32_t pred_frames,
                           float* out_host, float* aux_host) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->forward_host(x_host, batch, t_in, pred_frames, out_host, aux_host);
  });
}

int vpk_model_forward_actions(vpk_model* m, const float* x, const float* actions, int32_t action_steps, int32_t batch,
                              int32_t t_in, int32_t pred_frames, float* out, float* aux, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->forward(x, batch, t_in, pred_frames, out, aux, workspace, workspace_bytes, static_cast<cudaStream_t>(stream),
                     actions, action_steps);
  });
}

int vpk_model_forward_host_actions(vpk_model* m, const float* x_host, const float* actions_host, int32_t action_steps,
                                   int32_t batch, int32_t t_in, int32_t pred_frames, float* out_host, float* aux_host) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->forward_host(x_host, batch, t_in, pred_frames, out_host, aux_host, actions_host, action_steps);
  });
}

int vpk_model_microbatch(vpk_model* m, int32_t batch, int32_t* sequences) {
  return guarded([&] {
    VPK_REQUIRE(m && sequences && batch > 0, "bad argument");
    *sequences = m->impl->microbatch(batch);
  });
}

int vpk_model_last_launch_count(vpk_model* m, int64_t* launches) {
  return guarded([&] {
    VPK_REQUIRE(m && launches, "null argument");
    *launches = m->impl->last_launches;
  });
}

int vpk_model_set_timing(vpk_model* m, int32_t enable) {
  return guarded([&] {
    VPK_REQUIRE(m, "null argument");
    m->impl->timing = enable;
  });
}

int vpk_model_last_gemm_ms(vpk_model* m, float* ms, int64_t* gemm_launches, double* gemm_flops) {
  return guarded([&] {
    VPK_REQUIRE(m && ms && gemm_launches && gemm_flops, "null argument");
    m->impl->gemm_stats(ms, gemm_launches, gemm_flops);
  });
}

int vpk_model_profile(vpk_model* m, char* buf, size_t n) {
  return guarded([&] {
    VPK_REQUIRE(m && buf && n > 0, "null argument");
    const std::string t = m->impl->profile_text();
    const size_t k = std::min(n - 1, t.size());
    std::memcpy(buf, t.data(), k);
    buf[k] = 0;
  });
}

void vpk_model_destroy(vpk_model* m) {
  if (m == nullptr) return;
  try {
    delete m->impl;
  } catch (...) {
  }
  delete m;
}

// ---- cells -------------------------------------------------------------------------------------------------------
int vpk_convlstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w,
                             int32_t k, int32_t gate_order, const float* weight, const float* bias, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(weight && out, "null argument");
    *out = new vpk_cell{vpk::make_convlstm_cell(precision, backend, cin, ch, h, w, k, gate_order, weight, bias)};
  });
}

int vpk_convlstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                           const float* wci, const float* wcf, const float* wco, float* h_out, float* c_out,
                           void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && h && c && h_out && c_out, "null argument");
    const float* in[8] = {x, h, c, wci, wcf, wco, nullptr, nullptr};
    float* outp[8] = {h_out, c_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_convlstm_cell_backward(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                               const float* dh_out, const float* dc_out, float* dx, float* dh, float* dc, float* dw, float* db,
                               void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && c && dx && dh && dc && dw, "null argument");
    const float* in[8] = {x, h, c, dh_out, dc_out, nullptr, nullptr, nullptr};
    float* outp[8] = {dx, dh, dc, dw, db, nullptr, nullptr, nullptr};
    cell->impl->backward(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_convlstm_cell_backward_peep(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                                    const float* wci, const float* wcf, const float* wco, const float* dh_out,
                                    const float* dc_out, float* dx, float* dh, float* dc, float* dw, float* db, float* dwci,
                                    float* dwcf, float* dwco, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && h && c && dh && dc && dw, "null argument");
    const float* in[8] = {x, h, c, dh_out, dc_out, wci, wcf, wco};
    float* outp[8] = {dx, dh, dc, dw, db, dwci, dwcf, dwco};
    cell->impl->backward(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_stlstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w,
                           int32_t k, const float* w_x, const float* w_h, const float* w_m, const float* w_o,
                           const float* w_last, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(w_x && w_h && w_m && w_o && w_last && out, "null argument");
    *out = new vpk_cell{vpk::make_stlstm_cell(precision, backend, cin, ch, h, w, k, w_x, w_h, w_m, w_o, w_last)};
  });
}

int vpk_stlstm_cell_set_layer_norm(vpk_cell* cell, const float* gx, const float* bx, const float* gh, const float* bh,
                                   const float* gm, const float* bm, const float* go, const float* bo) {
  return guarded([&] {
    VPK_REQUIRE(cell && gx && bx && gh && bh && gm && bm && go && bo, "null argument");
    const float* p[8] = {gx, bx, gh, bh, gm, bm, go, bo};
    cell->impl->set_layer_norm(p);
  });
}

int vpk_stlstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                         const float* m, float* h_out, float* c_out, float* m_out, float* dc_out, float* dm_out,
                         void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && c && m && h_out && c_out && m_out, "null argument");
    const float* in[8] = {x, h, c, m, nullptr, nullptr, nullptr, nullptr};
    float* outp[8] = {h_out, c_out, m_out, dc_out, dm_out, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_causal_lstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t cm, int32_t ch, int32_t h,
                                int32_t w, int32_t k, const float* const* weights, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(weights && out, "null argument");
    *out = new vpk_cell{vpk::make_causal_lstm_cell(precision, backend, cin, cm, ch, h, w, k, weights)};
  });
}

int vpk_causal_lstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                              const float* m, float* h_out, float* c_out, float* m_out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && c && m && h_out && c_out && m_out, "null argument");
    const float* in[8] = {x, h, c, m, nullptr, nullptr, nullptr, nullptr};
    float* outp[8] = {h_out, c_out, m_out, nullptr, nullptr, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_ghu_cell_create(int32_t precision, int32_t backend, int32_t ch, int32_t h, int32_t w, int32_t k, const float* w_x,
                        const float* w_z, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(w_x && w_z && out, "null argument");
    *out = new vpk_cell{vpk::make_ghu_cell(precision, backend, ch, h, w, k, w_x, w_z)};
  });
}

int vpk_ghu_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* z, float* z_out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && z && z_out, "null argument");
    const float* in[8] = {x, z, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* outp[8] = {z_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_phycell_cell_create(int32_t precision, int32_t backend, int32_t ch, int32_t hid, int32_t h, int32_t w,
                            int32_t k, const float* conv1_w, const float* conv1_b, const float* gn_w,
                            const float* gn_b, const float* conv2_w, const float* conv2_b, const float* gate_w,
                            const float* gate_b, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(conv1_w && conv1_b && gn_w && gn_b && conv2_w && conv2_b && gate_w && gate_b && out, "null argument");
    *out = new vpk_cell{vpk::make_phycell_cell(precision, backend, ch, hid, h, w, k, conv1_w, conv1_b, gn_w, gn_b,
                                               conv2_w, conv2_b, gate_w, gate_b)};
  });
}

int vpk_phycell_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, float* h_out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && h_out, "null argument");
    const float* in[8] = {x, h, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* outp[8] = {h_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_stlstm_ac_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w, int32_t k,
                              const float* const* weights, const float* const* biases, vpk_cell** out) {
  return guarded([&] {
    VPK_REQUIRE(weights && biases && out, "null argument");
    for (int i = 0; i < 6; ++i) VPK_REQUIRE(weights[i] && biases[i], "null weight / bias");
    *out = new vpk_cell{vpk::make_stlstm_ac_cell(precision, backend, cin, ch, h, w, k, weights, biases)};
  });
}

int vpk_stlstm_ac_cell_set_layer_norm(vpk_cell* cell, const float* const* params) {
  return guarded([&] {
    VPK_REQUIRE(cell && params, "null argument");
    for (int i = 0; i < 10; ++i) VPK_REQUIRE(params[i], "null LayerNorm parameter");
    cell->impl->set_layer_norm(params);
  });
}

int vpk_stlstm_ac_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c, const float* m,
                            const float* a, float* h_out, float* c_out, float* m_out, float* dc_out, float* dm_out,
                            void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && c && m && a && h_out && c_out && m_out, "null argument");
    const float* in[8] = {x, h, c, m, a, nullptr, nullptr, nullptr};
    float* outp[8] = {h_out, c_out, m_out, dc_out, dm_out, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

int vpk_phycell_cell_set_action_convs(vpk_cell* cell, int32_t action_size, const float* frame_w, const float* frame_b,
                                      const float* hidden_w, const float* hidden_b) {
  return guarded([&] {
    VPK_REQUIRE(cell && frame_w && frame_b && hidden_w && hidden_b, "null argument");
    cell->impl->set_action_convs(action_size, frame_w, frame_b, hidden_w, hidden_b);
  });
}

int vpk_phycell_cell_step_action(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* action,
                                 float* h_out, void* stream) {
  return guarded([&] {
    VPK_REQUIRE(cell && x && h && action && h_out, "null argument");
    const float* in[8] = {x, h, action, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* outp[8] = {h_out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cell->impl->step(batch, in, outp, static_cast<cudaStream_t>(stream));
  });
}

void vpk_cell_destroy(vpk_cell* cell) {
  if (cell == nullptr) return;
  try {
    delete cell->impl;
  } catch (...) {
  }
  delete cell;
}

}  // extern "C"
