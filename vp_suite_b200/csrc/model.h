// Model runtime of libvpk: parameters, launch programs (one per microbatch shape), workspace planning, CUDA-graph
// replay, microbatch loop and the host-buffer pipeline.  Concrete rollouts (EF-ConvLSTM, PredRNN-V2, PhyDNet) only
// implement build().
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vpk.h"
#include "common.h"
#include "lowering.h"

namespace vpk {

struct HostParam {
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool provided = false;
};

struct Op;
struct RunCtx {          // pointers of the current microbatch (already offset)
  const float* x;
  float* out;
  float* aux;
  int mb0;               // first sequence of the microbatch
  int nb;                // sequences in it
  int batch;             // sequences of the whole call
  // host-buffer entry only: called right after an op that completes a predicted frame (Op::frame >= 0) was enqueued,
  // so that the frame's device-to-host copy can start while the rollout continues
  const std::function<void(const Op&, cudaStream_t)>* on_frame = nullptr;
  // host-buffer entry only: called right BEFORE an op that is the first to read input frame Op::needs_input, so that
  // the compute stream waits for that frame's host-to-device copy only (the frames arrive one by one)
  const std::function<void(int frame, cudaStream_t)>* on_input = nullptr;
  // action-conditional models: DEVICE fp32 [nb, action_steps, action_size] of this microbatch (nullptr otherwise)
  const float* actions = nullptr;
  int action_steps = 0;
};

struct Op {
  std::string name;
  std::function<void(cudaStream_t, const RunCtx&)> fn;
  double flops = 0;
  bool gate = false;     // counted in the gate-GEMM roofline figure
  bool is_kernel = true; // false for memsets / copies
  // >= 0: this op writes the last bytes of predicted frame `frame` of every sequence of the microbatch, at
  // frame_src + b * frame_pitch (fp32, frame_elems values each)
  int frame = -1;
  int needs_input = -1;  // >= 0: the first op that reads input frame `needs_input` of the microbatch
  const float* frame_src = nullptr;
  long long frame_pitch = 0, frame_elems = 0;
};

struct Program {
  int B = 0, t_in = 0, pred = 0;
  void* ws_base = nullptr;
  size_t ws_bytes = 0;
  std::vector<Op> pre, body, post;
  cudaGraphExec_t graph = nullptr;
  ~Program() {
    if (graph) cudaGraphExecDestroy(graph);
  }
};

class Model {
 public:
  explicit Model(const vpk_model_desc& d);
  virtual ~Model();

  void set_param(const std::string& key, const float* data, const int64_t* shape, int ndim);
  void finalize(cudaStream_t stream);
  size_t workspace_bytes(int batch, int t_in, int pred);
  void forward(const float* x, int batch, int t_in, int pred, float* out, float* aux, void* ws, size_t ws_bytes,
               cudaStream_t stream, const float* actions = nullptr, int action_steps = 0);
  void forward_host(const float* x, int batch, int t_in, int pred, float* out, float* aux, const float* actions = nullptr,
                    int action_steps = 0);

  int microbatch(int batch) const;

  vpk_model_desc desc;
  std::vector<std::string> keys;
  std::map<std::string, HostParam> params;
  int64_t last_launches = 0;
  int timing = 0;          // 0 off, 1 events around the gate GEMMs, 2 events around every kernel (per-layer profile)
  std::string profile_text();
  void gemm_stats(float* ms, int64_t* launches, double* flops);

 protected:
  // ---- implemented by the concrete rollouts ----
  virtual void build(Program& prog, Arena& arena, int B, int t_in, int pred, bool measure, cudaStream_t stream) = 0;
  virtual void validate(int t_in, int pred) const {}
  // rollout steps that read an action vector (0: the model is not action-conditional)
  virtual int action_steps_needed(int t_in, int pred) const { return 0; }
  void check_actions(const float* actions, int action_steps, int t_in, int pred) const;
  virtual int default_microbatch() const { return 64; }
  virtual int in_frames(int t_in, int pred) const { return t_in; }
  // frames of each input sequence the rollout really reads (the host entry copies only these to the device)
  virtual int used_in_frames(int t_in, int pred) const { return in_frames(t_in, pred); }
  // true: the program converts input frame t in its own op (marked Op::needs_input = t) instead of converting all frames
  // up front, so the host entry may deliver the frames one by one
  virtual bool streams_input() const { return false; }
  // true while forward_host() builds a program: a rollout may lay its ops out differently for the host pipeline
  // (PhyDNet encodes the context frames in growing groups so that compute starts after the first frame's copy)
  bool host_build = false;
  virtual void begin_call(int batch, int t_in, int pred, float* aux, cudaStream_t stream) {}
  virtual void end_call(int batch, float* aux, cudaStream_t stream) {}

  // ---- helpers for build() ----
  void declare(const std::string& key, std::vector<int64_t> shape);
  const float* hp(const std::string& key) const;           // host data of a parameter
  bool has(const std::string& key) const;
  float* dev_f32(const std::string& name, const std::vector<float>& host, cudaStream_t stream);  // cached upload
  // dt: operand type of this launch (-1: the model's precision); lets one program mix bf16 cells with fp32 convs
  // dst: op list the launches are appended to (nullptr: prog.body); ops that read RunCtx pointers of the CALL (input
  // frames, actions) and everything feeding only on them may go to prog.pre, which is never part of a captured graph
  void add_conv(Program& prog, const ConvSpec& spec, bool measure, cudaStream_t stream, int dt = -1,
                std::vector<Op>* dst = nullptr);
  void add_memset(Program& prog, void* p, size_t bytes, const char* name);
  int esize() const { return static_cast<int>(dtype_size(dtype)); }
  SrcView dense_view(void* p, int H, int W, int C) const {
    return SrcView{p, H, W, C, static_cast<long long>(H) * W * C, static_cast<long long>(W) * C, C};
  }

  int dtype = DT_F32;
  int backend = 0;
  int num_sms = 148;
  bool finalized = false;
  DeviceStore store;
  std::map<std::string, std::vector<PackedWeights>> packed_cache;
  std::map<std::string, float*> f32_cache;

 private:
  Program* get_program(int B, int t_in, int pred, void* ws, size_t ws_bytes, cudaStream_t stream);
  void run_ops(std::vector<Op>& ops, cudaStream_t stream, const RunCtx& ctx);
  std::vector<std::unique_ptr<Program>> programs;
  // gate-GEMM timing
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<bool> gate_flags;
  std::vector<std::pair<std::string, double>> ev_names;   // (layer name, flops) per recorded event pair
  double timed_flops = 0;
  int64_t timed_launches = 0;
  // host pipeline resources
  struct HostPipe;
  std::unique_ptr<HostPipe> pipe;
};

Model* make_ef_convlstm(const vpk_model_desc& d);
Model* make_predrnn(const vpk_model_desc& d);
Model* make_phydnet(const vpk_model_desc& d, bool branch_only);
Model* make_stphy(const vpk_model_desc& d);
Model* make_ef_trajgru(const vpk_model_desc& d);
Model* make_predrnnpp_causal(const vpk_model_desc& d);

}  // namespace vpk
