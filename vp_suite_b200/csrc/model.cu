#include "model.h"

#include <cstdlib>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace vpk {

Model::Model(const vpk_model_desc& d) : desc(d) {
  dtype = (d.precision == VPK_PREC_BF16) ? DT_BF16 : DT_F32;
  backend = d.backend;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) num_sms = sms;
  } else {
    cudaGetLastError();
  }
}

struct Model::HostPipe {
  void* d_x[2] = {nullptr, nullptr};
  void* d_out[2] = {nullptr, nullptr};
  void* ws = nullptr;
  float* d_aux = nullptr;
  float* d_actions = nullptr;
  size_t act_bytes = 0;
  size_t x_bytes = 0, out_bytes = 0, ws_bytes = 0;
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2], ev_comp[2], ev_out[2];
  std::vector<cudaEvent_t> ev_frame;       // one per predicted frame (frame streaming)
  std::vector<cudaEvent_t> ev_in_frame;    // [2 buffers][input frames] (input frame streaming)
  bool init = false;
  ~HostPipe() {
    for (int i = 0; i < 2; ++i) {
      if (d_x[i]) cudaFree(d_x[i]);
      if (d_out[i]) cudaFree(d_out[i]);
    }
    if (ws) cudaFree(ws);
    if (d_aux) cudaFree(d_aux);
    if (d_actions) cudaFree(d_actions);
    for (cudaEvent_t e : ev_frame) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_in_frame) cudaEventDestroy(e);
    if (init) {
      for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(ev_in[i]);
        cudaEventDestroy(ev_comp[i]);
        cudaEventDestroy(ev_out[i]);
      }
      cudaStreamDestroy(s_in);
      cudaStreamDestroy(s_comp);
      cudaStreamDestroy(s_out);
    }
  }
};

Model::~Model() {
  programs.clear();
  for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
}

void Model::declare(const std::string& key, std::vector<int64_t> shape) {
  HostParam p;
  size_t n = 1;
  for (int64_t s : shape) n *= static_cast<size_t>(s);
  p.shape = std::move(shape);
  p.data.assign(n, 0.f);
  params.emplace(key, std::move(p));
  keys.push_back(key);
}

bool Model::has(const std::string& key) const {
  auto it = params.find(key);
  return it != params.end() && it->second.provided;
}

const float* Model::hp(const std::string& key) const {
  auto it = params.find(key);
  VPK_REQUIRE(it != params.end(), "unknown parameter " + key);
  return it->second.data.data();
}

void Model::set_param(const std::string& key, const float* data, const int64_t* shape, int ndim) {
  auto it = params.find(key);
  if (it == params.end()) VPK_THROW(1, "unexpected state_dict key '" + key + "'");
  HostParam& p = it->second;
  bool ok = static_cast<size_t>(ndim) == p.shape.size();
  for (int i = 0; ok && i < ndim; ++i) ok = (shape[i] == p.shape[i]);
  if (!ok) VPK_THROW(1, "shape mismatch for state_dict key '" + key + "'");
  VPK_REQUIRE(data != nullptr, "null data for " + key);
  std::memcpy(p.data.data(), data, p.data.size() * sizeof(float));
  p.provided = true;
  if (finalized) {   // weights changed after packing: drop every derived device object
    programs.clear();
    packed_cache.clear();
    f32_cache.clear();
    store.release();
  }
}

void Model::finalize(cudaStream_t) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    VPK_THROW(2, "no CUDA device: libvpk has no CPU fallback");
  }
  finalized = true;
}

float* Model::dev_f32(const std::string& name, const std::vector<float>& host, cudaStream_t stream) {
  auto it = f32_cache.find(name);
  if (it != f32_cache.end()) return it->second;
  float* d = static_cast<float*>(store.upload(host.data(), host.size() * sizeof(float), stream));
  f32_cache[name] = d;
  return d;
}

void Model::add_conv(Program& prog, const ConvSpec& spec, bool measure, cudaStream_t stream, int dt_override,
                     std::vector<Op>* dst) {
  const int dt = dt_override >= 0 ? dt_override : dtype;
  std::vector<BuiltConv> built = build_conv(spec, dt, backend, store, packed_cache, stream, num_sms, measure);
  if (measure) return;
  for (BuiltConv& bc : built) {
    Op op;
    op.name = bc.name;
    op.flops = bc.L.flops;
    op.gate = bc.L.is_gate_gemm != 0;
    if (bc.use_halo) {
      auto plan = std::make_shared<HaloPlan>(bc.halo);
      op.fn = [plan](cudaStream_t s, const RunCtx&) { launch_conv_halo(*plan, s); };
    } else if (bc.use_tc) {
      auto plan = std::make_shared<TcPlan>(bc.tc);
      op.fn = [plan](cudaStream_t s, const RunCtx&) { launch_conv_tc(*plan, s); };
    } else if (bc.use_direct) {
      ConvLaunch L = bc.L;
      const int ns = num_sms;
      op.fn = [L, dt, ns](cudaStream_t s, const RunCtx&) { launch_conv_direct(L, dt, ns, s); };
    } else {
      ConvLaunch L = bc.L;
      op.fn = [L, dt](cudaStream_t s, const RunCtx&) { launch_conv_simt(L, dt, s); };
    }
    (dst != nullptr ? *dst : prog.body).push_back(std::move(op));
  }
}

void Model::add_memset(Program& prog, void* p, size_t bytes, const char* name) {
  Op op;
  op.name = name;
  op.is_kernel = false;
  op.fn = [p, bytes](cudaStream_t s, const RunCtx&) { VPK_CUDA(cudaMemsetAsync(p, 0, bytes, s)); };
  prog.body.push_back(std::move(op));
}

int Model::microbatch(int batch) const {
  // balanced: the fewest passes the cap allows, then equal shares (512 with a cap of 222 -> 171 + 171 + 170, not
  // 222 + 222 + 68: a short tail pass runs the persistent kernels at a fraction of their steady-state efficiency)
  int mb = desc.max_microbatch > 0 ? desc.max_microbatch : default_microbatch();
  mb = std::max(1, std::min(mb, batch));
  const int passes = (batch + mb - 1) / mb;
  return (batch + passes - 1) / passes;
}

size_t Model::workspace_bytes(int batch, int t_in, int pred) {
  validate(t_in, pred);
  VPK_REQUIRE(batch > 0, "batch must be positive");
  Program tmp;
  Arena arena;
  build(tmp, arena, microbatch(batch), t_in, pred, /*measure=*/true, nullptr);
  return arena.off + 4096;
}

Program* Model::get_program(int B, int t_in, int pred, void* ws, size_t ws_bytes, cudaStream_t stream) {
  for (auto& p : programs)
    if (p->B == B && p->t_in == t_in && p->pred == pred && p->ws_base == ws && p->ws_bytes <= ws_bytes) return p.get();
  if (programs.size() >= 8) programs.erase(programs.begin());
  auto prog = std::make_unique<Program>();
  prog->B = B;
  prog->t_in = t_in;
  prog->pred = pred;
  prog->ws_base = ws;
  Arena arena;
  arena.base = static_cast<char*>(ws);
  arena.cap = ws_bytes;
  build(*prog, arena, B, t_in, pred, /*measure=*/false, stream);
  prog->ws_bytes = arena.off;
  // weights uploaded during build() live in pageable staging: make sure they have landed before staging is reused
  VPK_CUDA(cudaStreamSynchronize(stream));
  store.staging.clear();
  if (desc.use_cuda_graph && timing == 0) {
    cudaStream_t cs;
    VPK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    RunCtx ctx{};
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      try {
        for (Op& op : prog->body) op.fn(cs, ctx);
      } catch (...) {
        cudaStreamEndCapture(cs, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        throw;
      }
      VPK_CUDA(cudaStreamEndCapture(cs, &graph));
      VPK_CUDA(cudaGraphInstantiate(&prog->graph, graph, 0));
      cudaGraphDestroy(graph);
    }
    cudaStreamDestroy(cs);
    VPK_CUDA(e);
  }
  programs.push_back(std::move(prog));
  return programs.back().get();
}

void Model::run_ops(std::vector<Op>& ops, cudaStream_t stream, const RunCtx& ctx) {
  for (Op& op : ops) {
    const bool t = (timing == 1 && op.gate) || (timing == 2 && op.is_kernel);
    if (t) {
      while (ev_pool.size() < ev_used + 2) {
        cudaEvent_t e;
        VPK_CUDA(cudaEventCreate(&e));
        ev_pool.push_back(e);
      }
      VPK_CUDA(cudaEventRecord(ev_pool[ev_used], stream));
    }
    if (op.needs_input >= 0 && ctx.on_input != nullptr) (*ctx.on_input)(op.needs_input, stream);
    op.fn(stream, ctx);
    if (t) {
      VPK_CUDA(cudaEventRecord(ev_pool[ev_used + 1], stream));
      ev_used += 2;
      if (op.gate) {
        timed_flops += op.flops;
        timed_launches += 1;
      }
      ev_names.emplace_back(op.name, op.flops);
      gate_flags.push_back(op.gate);
    }
    if (op.is_kernel) ++last_launches;
    if (op.frame >= 0 && ctx.on_frame != nullptr) (*ctx.on_frame)(op, stream);
  }
}

std::string Model::profile_text() {
  std::map<std::string, std::pair<double, std::pair<int, double>>> agg;   // name -> (ms, (count, flops))
  for (size_t i = 0; i + 1 < ev_used && i / 2 < ev_names.size(); i += 2) {
    VPK_CUDA(cudaEventSynchronize(ev_pool[i + 1]));
    float t = 0.f;
    VPK_CUDA(cudaEventElapsedTime(&t, ev_pool[i], ev_pool[i + 1]));
    auto& a = agg[ev_names[i / 2].first];
    a.first += t;
    a.second.first += 1;
    a.second.second += ev_names[i / 2].second;
  }
  std::vector<std::pair<double, std::string>> rows;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof line, "%s %d %.4f %.3f\n", kv.first.c_str(), kv.second.second.first, kv.second.first,
             kv.second.second.second * 1e-9);
    rows.emplace_back(-kv.second.first, line);
  }
  std::sort(rows.begin(), rows.end());
  std::string out;
  for (auto& r : rows) out += r.second;
  return out;
}

void Model::gemm_stats(float* ms, int64_t* launches, double* flops) {
  float total = 0.f;
  for (size_t i = 0; i + 1 < ev_used && i / 2 < ev_names.size(); i += 2) {
    if (timing == 2) {   // every kernel is bracketed: count only the gate GEMMs here
      const std::string& nm = ev_names[i / 2].first;
      (void)nm;
    }
    VPK_CUDA(cudaEventSynchronize(ev_pool[i + 1]));
    float t = 0.f;
    VPK_CUDA(cudaEventElapsedTime(&t, ev_pool[i], ev_pool[i + 1]));
    if (timing != 2 || gate_flags[i / 2]) total += t;
  }
  *ms = total;
  *launches = timed_launches;
  *flops = timed_flops;
}

void Model::check_actions(const float* actions, int action_steps, int t_in, int pred) const {
  const int need = action_steps_needed(t_in, pred);
  // the reference raises ValueError (predrnn_v2.py:149-151, models/phydnet.py:103-105)
  if (need > 0 && (actions == nullptr || action_steps < need))
    VPK_THROW(1, "Given actions are None or of the wrong size! (an action-conditional model needs actions[:, 0:" +
                     std::to_string(need) + "])");
}

void Model::forward(const float* x, int batch, int t_in, int pred, float* out, float* aux, void* ws, size_t ws_bytes,
                    cudaStream_t stream, const float* actions, int action_steps) {
  VPK_REQUIRE(finalized, "forward before finalize");
  VPK_REQUIRE(x != nullptr && out != nullptr && ws != nullptr, "null buffer");
  VPK_REQUIRE(batch > 0 && pred > 0 && t_in > 0, "bad batch / frame counts");
  validate(t_in, pred);
  check_actions(actions, action_steps, t_in, pred);
  const size_t act_stride = static_cast<size_t>(action_steps) * std::max(0, desc.action_size);
  last_launches = 0;
  ev_used = 0;
  ev_names.clear();
  gate_flags.clear();
  timed_flops = 0;
  timed_launches = 0;
  const int mb = microbatch(batch);
  const size_t in_stride = static_cast<size_t>(in_frames(t_in, pred)) * desc.img_c * desc.img_h * desc.img_w;
  const size_t out_stride = static_cast<size_t>(pred) * desc.img_c * desc.img_h * desc.img_w;
  begin_call(batch, t_in, pred, aux, stream);
  for (int mb0 = 0; mb0 < batch; mb0 += mb) {
    const int nb = std::min(mb, batch - mb0);
    Program* prog = get_program(nb, t_in, pred, ws, ws_bytes, stream);
    RunCtx ctx{x + mb0 * in_stride, out + mb0 * out_stride, aux, mb0, nb, batch};
    if (actions != nullptr) {
      ctx.actions = actions + mb0 * act_stride;
      ctx.action_steps = action_steps;
    }
    run_ops(prog->pre, stream, ctx);
    if (prog->graph != nullptr && timing == 0) {
      VPK_CUDA(cudaGraphLaunch(prog->graph, stream));
      for (const Op& op : prog->body)
        if (op.is_kernel) ++last_launches;
    } else {
      run_ops(prog->body, stream, ctx);
    }
    run_ops(prog->post, stream, ctx);
  }
  end_call(batch, aux, stream);
}

void Model::forward_host(const float* x, int batch, int t_in, int pred, float* out, float* aux, const float* actions,
                         int action_steps) {
  VPK_REQUIRE(finalized, "forward before finalize");
  VPK_REQUIRE(x != nullptr && out != nullptr, "null buffer");
  validate(t_in, pred);
  check_actions(actions, action_steps, t_in, pred);
  const int mb = microbatch(batch);
  const size_t in_stride = static_cast<size_t>(in_frames(t_in, pred)) * desc.img_c * desc.img_h * desc.img_w;
  const size_t out_stride = static_cast<size_t>(pred) * desc.img_c * desc.img_h * desc.img_w;
  if (!pipe) pipe = std::make_unique<HostPipe>();
  HostPipe& hpipe = *pipe;
  if (!hpipe.init) {
    VPK_CUDA(cudaStreamCreateWithFlags(&hpipe.s_in, cudaStreamNonBlocking));
    VPK_CUDA(cudaStreamCreateWithFlags(&hpipe.s_comp, cudaStreamNonBlocking));
    VPK_CUDA(cudaStreamCreateWithFlags(&hpipe.s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      VPK_CUDA(cudaEventCreateWithFlags(&hpipe.ev_in[i], cudaEventDisableTiming));
      VPK_CUDA(cudaEventCreateWithFlags(&hpipe.ev_comp[i], cudaEventDisableTiming));
      VPK_CUDA(cudaEventCreateWithFlags(&hpipe.ev_out[i], cudaEventDisableTiming));
    }
    VPK_CUDA(cudaMalloc(&hpipe.d_aux, 256));
    hpipe.init = true;
  }
  const size_t xb = mb * in_stride * sizeof(float), ob = mb * out_stride * sizeof(float);
  const size_t wb = workspace_bytes(batch, t_in, pred);
  if (xb > hpipe.x_bytes || ob > hpipe.out_bytes || wb > hpipe.ws_bytes) {
    VPK_CUDA(cudaDeviceSynchronize());
    programs.clear();
    for (int i = 0; i < 2; ++i) {
      if (hpipe.d_x[i]) cudaFree(hpipe.d_x[i]);
      if (hpipe.d_out[i]) cudaFree(hpipe.d_out[i]);
      VPK_CUDA(cudaMalloc(&hpipe.d_x[i], xb));
      VPK_CUDA(cudaMalloc(&hpipe.d_out[i], ob));
    }
    if (hpipe.ws) cudaFree(hpipe.ws);
    VPK_CUDA(cudaMalloc(&hpipe.ws, wb));
    hpipe.x_bytes = xb;
    hpipe.out_bytes = ob;
    hpipe.ws_bytes = wb;
  }
  last_launches = 0;
  ev_used = 0;
  ev_names.clear();
  gate_flags.clear();
  timed_flops = 0;
  timed_launches = 0;
  begin_call(batch, t_in, pred, hpipe.d_aux, hpipe.s_comp);
  // actions of the whole call (a few KB): one copy ahead of the first microbatch, on the compute stream
  const size_t act_stride = static_cast<size_t>(action_steps) * std::max(0, desc.action_size);
  if (actions != nullptr && act_stride > 0) {
    const size_t ab = static_cast<size_t>(batch) * act_stride * sizeof(float);
    if (ab > hpipe.act_bytes) {
      if (hpipe.d_actions) cudaFree(hpipe.d_actions);
      VPK_CUDA(cudaMalloc(&hpipe.d_actions, ab));
      hpipe.act_bytes = ab;
    }
    VPK_CUDA(cudaMemcpyAsync(hpipe.d_actions, actions, ab, cudaMemcpyHostToDevice, hpipe.s_comp));
  }
  int it = 0;
  for (int mb0 = 0; mb0 < batch; mb0 += mb, ++it) {
    const int nb = std::min(mb, batch - mb0);
    const int buf = it & 1;
    // H2D of this microbatch; d_x[buf] was last read by the compute of iteration it-2
    if (it >= 2) VPK_CUDA(cudaStreamWaitEvent(hpipe.s_in, hpipe.ev_comp[buf], 0));
    const bool frame_in = streams_input() && getenv("VPK_NO_FRAME_STREAM") == nullptr;
    const size_t chw = static_cast<size_t>(desc.img_c) * desc.img_h * desc.img_w;
    if (frame_in) {
      // one 2-D copy per input frame (all sequences of the microbatch), each with its own event: the rollout's step t
      // waits for frame t only, so compute starts after 1 / t_in of the microbatch's input has arrived
      const int nf = used_in_frames(t_in, pred);
      while (hpipe.ev_in_frame.size() < static_cast<size_t>(2 * nf)) {
        cudaEvent_t e;
        VPK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hpipe.ev_in_frame.push_back(e);
      }
      for (int f = 0; f < nf; ++f) {
        VPK_CUDA(cudaMemcpy2DAsync(static_cast<float*>(hpipe.d_x[buf]) + f * chw, in_stride * sizeof(float),
                                   x + mb0 * in_stride + f * chw, in_stride * sizeof(float), chw * sizeof(float), nb,
                                   cudaMemcpyHostToDevice, hpipe.s_in));
        VPK_CUDA(cudaEventRecord(hpipe.ev_in_frame[buf * nf + f], hpipe.s_in));
      }
    } else {
      const size_t used = static_cast<size_t>(used_in_frames(t_in, pred)) * desc.img_c * desc.img_h * desc.img_w;
      if (used == in_stride)
        VPK_CUDA(cudaMemcpyAsync(hpipe.d_x[buf], x + mb0 * in_stride, nb * in_stride * sizeof(float),
                                 cudaMemcpyHostToDevice, hpipe.s_in));
      else      // only the frames the rollout reads, same device layout (sequence pitch unchanged)
        VPK_CUDA(cudaMemcpy2DAsync(hpipe.d_x[buf], in_stride * sizeof(float), x + mb0 * in_stride, in_stride * sizeof(float),
                                   used * sizeof(float), nb, cudaMemcpyHostToDevice, hpipe.s_in));
    }
    VPK_CUDA(cudaEventRecord(hpipe.ev_in[buf], hpipe.s_in));
    // compute; d_out[buf] was last read by the D2H of iteration it-2
    if (!frame_in) VPK_CUDA(cudaStreamWaitEvent(hpipe.s_comp, hpipe.ev_in[buf], 0));
    if (it >= 2) VPK_CUDA(cudaStreamWaitEvent(hpipe.s_comp, hpipe.ev_out[buf], 0));
    host_build = true;
    Program* prog = get_program(nb, t_in, pred, hpipe.ws, hpipe.ws_bytes, hpipe.s_comp);
    host_build = false;
    // Frame streaming: as soon as the op that completes predicted frame p is enqueued, the frame is copied into this
    // microbatch's output buffer and from there to the host on the copy stream, so that only the LAST frame's transfer
    // (1 / pred of the output) is left when the rollout ends, not the whole microbatch's.
    int frames_streamed = 0;
    float* dout = static_cast<float*>(hpipe.d_out[buf]);
    float* hout = out + mb0 * out_stride;
    const std::function<void(const Op&, cudaStream_t)> on_frame = [&](const Op& op, cudaStream_t s) {
      const size_t width = static_cast<size_t>(op.frame_elems) * sizeof(float);
      const size_t pitch = out_stride * sizeof(float);
      float* d_frame = dout + static_cast<size_t>(op.frame) * op.frame_elems;
      VPK_CUDA(cudaMemcpy2DAsync(d_frame, pitch, op.frame_src, static_cast<size_t>(op.frame_pitch) * sizeof(float), width, nb,
                                 cudaMemcpyDeviceToDevice, s));
      while (hpipe.ev_frame.size() <= static_cast<size_t>(op.frame)) {
        cudaEvent_t e;
        VPK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hpipe.ev_frame.push_back(e);
      }
      cudaEvent_t e = hpipe.ev_frame[op.frame];
      VPK_CUDA(cudaEventRecord(e, s));
      VPK_CUDA(cudaStreamWaitEvent(hpipe.s_out, e, 0));
      VPK_CUDA(cudaMemcpy2DAsync(hout + static_cast<size_t>(op.frame) * op.frame_elems, pitch, d_frame, pitch, width, nb,
                                 cudaMemcpyDeviceToHost, hpipe.s_out));
      ++frames_streamed;
    };
    RunCtx ctx{static_cast<const float*>(hpipe.d_x[buf]), static_cast<float*>(hpipe.d_out[buf]), hpipe.d_aux, mb0, nb,
               batch};
    if (actions != nullptr && act_stride > 0) {
      ctx.actions = hpipe.d_actions + mb0 * act_stride;
      ctx.action_steps = action_steps;
    }
    const int nf_in = used_in_frames(t_in, pred);
    const std::function<void(int, cudaStream_t)> on_input = [&](int f, cudaStream_t s) {
      VPK_CUDA(cudaStreamWaitEvent(s, hpipe.ev_in_frame[buf * nf_in + f], 0));
    };
    if (frame_in) ctx.on_input = &on_input;
    // (a CUDA-graph replay enqueues the whole body at once: no per-op hook, the whole-microbatch copy below is used)
    if (getenv("VPK_NO_FRAME_STREAM") == nullptr && !(prog->graph != nullptr && timing == 0)) ctx.on_frame = &on_frame;
    run_ops(prog->pre, hpipe.s_comp, ctx);
    if (prog->graph != nullptr && timing == 0) {
      VPK_CUDA(cudaGraphLaunch(prog->graph, hpipe.s_comp));
      for (const Op& op : prog->body)
        if (op.is_kernel) ++last_launches;
    } else {
      run_ops(prog->body, hpipe.s_comp, ctx);
    }
    run_ops(prog->post, hpipe.s_comp, ctx);
    VPK_CUDA(cudaEventRecord(hpipe.ev_comp[buf], hpipe.s_comp));
    VPK_REQUIRE(frames_streamed == 0 || frames_streamed == pred, "frame streaming: a rollout must mark every predicted frame");
    // D2H (whole microbatch unless every frame has already been streamed)
    VPK_CUDA(cudaStreamWaitEvent(hpipe.s_out, hpipe.ev_comp[buf], 0));
    if (frames_streamed != pred)
      VPK_CUDA(cudaMemcpyAsync(out + mb0 * out_stride, hpipe.d_out[buf], nb * out_stride * sizeof(float),
                               cudaMemcpyDeviceToHost, hpipe.s_out));
    VPK_CUDA(cudaEventRecord(hpipe.ev_out[buf], hpipe.s_out));
  }
  end_call(batch, hpipe.d_aux, hpipe.s_comp);
  if (aux != nullptr)
    VPK_CUDA(cudaMemcpyAsync(aux, hpipe.d_aux, sizeof(float), cudaMemcpyDeviceToHost, hpipe.s_comp));
  VPK_CUDA(cudaStreamSynchronize(hpipe.s_comp));
  VPK_CUDA(cudaStreamSynchronize(hpipe.s_out));
}

}  // namespace vpk
