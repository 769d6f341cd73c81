/*
 * vpk.h -- C ABI of libvpk.so: the B200-native (sm_100a) recurrent video-prediction hot path.
 *
 * The reference (AIS-Bonn/vp-suite) is pure Python/PyTorch and has no FFI for this path; the boundary it offers
 * is the VPModel / VPModelBlock class contract (vp_suite/base/base_model.py:11-146,
 * vp_suite/base/base_model_block.py:4-13).  Each entry point below cites the reference interface it replaces.
 * The Python drop-ins in vp_suite_b200/ bind these symbols through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; every function returns an int status (0 = VPK_OK); no C++ exception
 *     crosses the ABI; vpk_last_error() gives the message of the last failure on the calling thread.
 *   - The caller owns every tensor and the workspace.  The library owns only the handle (packed weights, launch
 *     programs, tensor maps, CUDA graphs).  All work is enqueued on the caller's stream (a cudaStream_t passed
 *     as void*); nothing synchronises the device except vpk_*_forward_host, which is the host-buffer entry.
 *   - Device = the current CUDA device of the calling thread.  A handle is not thread-safe; distinct handles are.
 *   - Tensors at the boundary are fp32, contiguous, in the reference's layouts (frames [b, t, c, h, w]).
 *   - Inference only (the reference wraps this path in torch.no_grad(): vp_suite/vpsuite.py:533).
 *   - There is no CPU fallback: without a CUDA device every compute entry fails with VPK_ERR_CUDA.
 */
#ifndef VPK_H_
#define VPK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPK_OK 0
#define VPK_ERR_INVALID 1   /* bad argument / shape mismatch (the reference raises ValueError/AttributeError) */
#define VPK_ERR_CUDA 2      /* CUDA runtime / driver failure, or no device */
#define VPK_ERR_STATE 3     /* call order violated (e.g. forward before finalize) */
#define VPK_ERR_WORKSPACE 4 /* workspace too small */

/* Operand precision of the conv contractions (accumulation is always fp32; cell state c/m is always fp32). */
#define VPK_PREC_FP32 0     /* fp32 operands on CUDA cores: matches the reference to <= 1e-4 */
#define VPK_PREC_BF16 1     /* bf16 operands, tcgen05 tensor cores (TMEM accumulators, TMA-fed) */

/* Kernel family for the bf16 contractions (testing aid; VPK_BACKEND_AUTO is what ships). */
#define VPK_BACKEND_AUTO 0  /* tcgen05 wherever the layout allows, CUDA-core kernel for the few tiny-K layers */
#define VPK_BACKEND_SIMT 1  /* force the CUDA-core implicit-GEMM kernel (same operands) */

typedef struct vpk_model vpk_model;

/* Model kinds: keys of vp_suite.models.MODEL_CLASSES (vp_suite/models/__init__.py:14-26) on the hot path, plus the
 * BASELINE config-2 composition of reference blocks. */
#define VPK_MODEL_CONVLSTM_SHI 0    /* EF_ConvLSTM      models/precipitation_nowcasting/ef_conv_lstm.py:7-108   */
#define VPK_MODEL_PREDRNN_PP 1      /* PredRNN_V2       models/predrnn_v2.py:11-230 (non action-conditional)    */
#define VPK_MODEL_PHY 2             /* PhyDNet          models/phydnet.py:12-137   (non action-conditional)     */
#define VPK_MODEL_CONVLSTM_BRANCH 3 /* DCGANEncoder->EncoderSplit->SingleStepConvLSTM->DecoderSplit->DCGANDecoder */
#define VPK_MODEL_TRAJGRU 5         /* EF_TrajGRU       models/precipitation_nowcasting/ef_traj_gru.py:8-119                 */
#define VPK_MODEL_PREDRNN_PP_CAUSAL 6 /* PredRNN++ as published (Causal LSTM stack + gradient highway unit; Wang et al., ICML 2018):
                                       named by the north star, ABSENT from the reference checkout -> parity unpinned
                                       (checker: oracle/causal.py); uses patch_size, num_layers >= 2, num_hidden, filter_size */
#define VPK_MODEL_ST_PHY 4          /* STPhy            models/st_phy.py:16-181    (non action-conditional); uses num_layers,
                                       num_hidden[0] = st_cell_channels, phycell_channels, phycell_kernel_size          */

/* Hyper-parameters.  Field names follow the reference's class attributes. Unused fields are ignored per kind. */
typedef struct vpk_model_desc {
  int32_t kind;              /* VPK_MODEL_*                                                                    */
  int32_t precision;         /* VPK_PREC_*                                                                     */
  int32_t backend;           /* VPK_BACKEND_*                                                                  */
  int32_t img_c, img_h, img_w; /* VPModel.img_shape (base_model.py:63-64)                                      */
  /* convlstm-shi (ef_conv_lstm.py:31-65); num_layers is fixed to 3 by Forecaster.forward (ef_blocks.py:109-110) */
  int32_t enc_c[6], dec_c[6];
  int32_t enc_conv_k[3], enc_conv_s[3], enc_conv_p[3];
  int32_t dec_conv_k[3], dec_conv_s[3], dec_conv_p[3];
  int32_t enc_rnn_k[3], dec_rnn_k[3];          /* rnn stride is 1 and padding k/2 (only those keep the state size) */
  int32_t final_conv_c;                        /* final_conv_1_c (identity block's channel count)                */
  int32_t ef_act;                              /* stage activation from the layer names (ef_blocks.py:33-36,42-45):
                                                  1 = LeakyReLU(0.2) ("leaky"), 3 = ReLU ("relu"), 0 = none       */
  /* predrnn-pp (predrnn_v2.py:34-43) */
  int32_t patch_size, num_layers, num_hidden[8], filter_size;
  float decoupling_loss_scale;
  int32_t layer_norm;        /* 1: nn.LayerNorm([k*C, H/p, W/p]) after conv_x/h/m/o (model_blocks/predrnn.py:24-40)  */
  /* phy / convlstm-branch (models/phydnet.py:28-33) */
  int32_t phycell_n_layers, phycell_channels, phycell_kernel_size;
  int32_t convlstm_n_layers, convlstm_hidden_dims[8], convlstm_kernel_size;
  /* execution */
  int32_t max_microbatch;    /* sequences processed per pass over the layers (0 = library default)             */
  int32_t use_cuda_graph;    /* 1: capture the per-microbatch launch program into a CUDA graph and replay it   */
  /* action-conditional variants (VPModel.action_conditional / action_size, base_model.py:33-34): predrnn-pp
   * (predrnn_v2.py:65-90: stride-2 input / action convs, ActionConditionalSpatioTemporalLSTMCell, output deconvs) and phy
   * (model_blocks/phydnet.py:44-55, 153-155: frame / hidden action convs in PhyCell, action channels into the ConvLSTM);
   * st-phy (st_phy.py:48-56, 142-150: Linear + (5,1) / (1,5) convs to the cells' action tensor, PhyCell action convs) */
  int32_t action_conditional;
  int32_t action_size;
  int32_t residual_on_action_conv;   /* predrnn-pp (predrnn_v2.py:46, 213-218) */
  /* trajgru (ef_traj_gru.py:47-61): flows per recurrent block; the i2h kernel sizes travel in enc_rnn_k / dec_rnn_k */
  int32_t enc_rnn_L[3], dec_rnn_L[3];
  /* st-phy, action-conditional (st_phy.py:32, 48-56): channels of the Linear-inflated action map */
  int32_t inflated_action_dim;
} vpk_model_desc;

/* Replaces: MODEL_CLASSES[key](device, **model_kwargs)  (vp_suite/vpsuite.py:170; base_model.py:38-69). */
int vpk_model_create(const vpk_model_desc* desc, vpk_model** out);

/* Replaces: nn.Module.load_state_dict.  `key` is the reference's state_dict key, `data` a HOST
 * fp32 contiguous array of `shape[0..ndim)`.  Unknown keys and wrong shapes fail with VPK_ERR_INVALID.  Keys not
 * supplied keep their default (zeros; this is how missing Wci/Wcf/Wco of CUDA-built checkpoints are handled,
 * conv_lstm_hzzone.py:30-32). */
int vpk_model_set_param(vpk_model* m, const char* key, const float* data, const int64_t* shape, int32_t ndim);

/* Number of state_dict entries the model expects, and the i-th key / shape (for layout checks). */
int vpk_model_num_params(vpk_model* m, int32_t* n);
int vpk_model_param_info(vpk_model* m, int32_t i, const char** key, int64_t* shape4, int32_t* ndim);

/* Packs all weights into the kernels' layouts and uploads them (enqueued on `stream`). */
int vpk_model_finalize(vpk_model* m, void* stream);

/* Workspace the caller must supply to vpk_model_forward for `batch` sequences of `t_in` input frames and
 * `pred_frames` predicted frames. */
int vpk_model_workspace_bytes(vpk_model* m, int32_t batch, int32_t t_in, int32_t pred_frames, size_t* bytes);

/* Replaces: VPModel.forward(x, pred_frames)  (ef_blocks.py:184-187, predrnn_v2.py:131-230, models/phydnet.py:94-137).
 *   x    DEVICE fp32 [batch, t_in, c, h, w]   (predrnn-pp: t_in = context + pred_frames, predrnn_v2.py:134-137)
 *   out  DEVICE fp32 [batch, pred_frames, c, h, w]
 *   aux  DEVICE fp32 [1] or NULL: model loss the reference returns in eval (predrnn-pp: decouple loss, :229-230)
 */
int vpk_model_forward(vpk_model* m, const float* x, int32_t batch, int32_t t_in, int32_t pred_frames, float* out,
                      float* aux, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces: VPModel.forward(x, pred_frames, actions=a) of an action-conditional model (predrnn_v2.py:147-152, 181-191;
 * models/phydnet.py:100-105).  actions: DEVICE fp32 [batch, action_steps, action_size]; step t of the rollout reads
 * actions[:, t]; action_steps must cover the rollout (context + pred_frames - 1).  A model created without
 * action_conditional ignores them (pass NULL / use vpk_model_forward); an action-conditional model given NULL actions or
 * too few steps fails with VPK_ERR_INVALID ("Given actions are None or of the wrong size!"). */
int vpk_model_forward_actions(vpk_model* m, const float* x, const float* actions, int32_t action_steps, int32_t batch,
                              int32_t t_in, int32_t pred_frames, float* out, float* aux, void* workspace,
                              size_t workspace_bytes, void* stream);

/* Same call with HOST buffers (x, out, aux on the host; pinned memory recommended): the library stages
 * microbatches through the device, overlapping copies with compute, and returns when `out` is complete.
 * The library allocates its own device workspace for this entry. */
int vpk_model_forward_host(vpk_model* m, const float* x_host, int32_t batch, int32_t t_in, int32_t pred_frames,
                           float* out_host, float* aux_host);
/* Host-buffer entry of an action-conditional model: actions_host fp32 [batch, action_steps, action_size] on the host. */
int vpk_model_forward_host_actions(vpk_model* m, const float* x_host, const float* actions_host, int32_t action_steps,
                                   int32_t batch, int32_t t_in, int32_t pred_frames, float* out_host, float* aux_host);

/* Sequences the library processes per pass over the layers (its microbatch) for a call with `batch` sequences. */
int vpk_model_microbatch(vpk_model* m, int32_t batch, int32_t* sequences);

/* Number of kernel launches (the library's own kernels) enqueued by the last forward call. */
int vpk_model_last_launch_count(vpk_model* m, int64_t* launches);

/* Duration in ms of the gate-GEMM kernels in the last forward (CUDA events on the launch stream) when timing was
 * enabled with vpk_model_set_timing(m, 1); used by bench.py for the roofline line. */
int vpk_model_set_timing(vpk_model* m, int32_t enable);
int vpk_model_last_gemm_ms(vpk_model* m, float* ms, int64_t* gemm_launches, double* gemm_flops);

/* Per-layer device times of the last forward when timing was enabled with vpk_model_set_timing(m, 2): writes lines
 * "<layer name> <launches> <ms> <gflop>" (sorted by time) into buf (NUL-terminated, truncated to n). */
int vpk_model_profile(vpk_model* m, char* buf, size_t n);

void vpk_model_destroy(vpk_model* m);

/* ---- single-step cells: the VPModelBlock boundary ------------------------------------------------------------ */

/* Replaces one iteration of ConvLSTM.forward's time loop (model_blocks/conv_lstm_hzzone.py:52-69; gate order
 * i,f,g,o; peepholes) when peephole pointers are given, and ConvLSTMCell.forward (model_blocks/conv_lstm_ndrplz.py:
 * 28-43; gate order i,f,o,g; no peepholes) when gate_order == 1.
 *   x [b, cin, h, w] or NULL (zero input, conv_lstm_hzzone.py:54-56); h, c [b, ch, h, w]; all DEVICE fp32 NCHW.
 *   weight HOST fp32 [4*ch, cin+ch, k, k], bias HOST fp32 [4*ch] or NULL; wci/wcf/wco DEVICE fp32 [ch, h, w] or NULL.
 *   h_out, c_out DEVICE fp32 [b, ch, h, w] (may not alias h). */
typedef struct vpk_cell vpk_cell;
int vpk_convlstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w,
                             int32_t k, int32_t gate_order, const float* weight, const float* bias, vpk_cell** out);
int vpk_convlstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                           const float* wci, const float* wcf, const float* wco, float* h_out, float* c_out,
                           void* stream);

/* Backward of one ConvLSTMCell step (gate_order == 1: conv_lstm_ndrplz.py:28-43), for training through the drop-in block
 * (the reference's train_iter back-propagates through every cell step, base_model.py:148-179).
 * Takes the step's inputs x, h, c and the upstream gradients of its outputs dh_out / dc_out (DEVICE fp32 NCHW; either may be
 * NULL = zero), recomputes the gates, and writes dx [b, cin, h, w], dh, dc [b, ch, h, w], dw [4ch, cin + ch, k, k] and
 * db [4ch] (NULL allowed) -- the gradients torch.autograd gives for the reference cell.  The cell keeps the weights it was
 * created with; dw / db are per call (the caller accumulates).  Deterministic. */
int vpk_convlstm_cell_backward(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                               const float* dh_out, const float* dc_out, float* dx, float* dh, float* dc, float* dw, float* db,
                               void* stream);

/* The same for one timestep of the Shi et al. ConvLSTM with peepholes (conv_lstm_hzzone.py:57-69; a cell created with
 * gate_order 0): torch.autograd's gradients of (h', c') = step(x, h, c; W, b, Wci, Wcf, Wco).  wci / wcf / wco DEVICE fp32
 * [1, ch, h, w] (NULL = zero); x may be NULL (the forecaster's all-zero input; dx is then not written).  Also writes the
 * peephole gradients dwci / dwcf / dwco [1, ch, h, w] (sums over the batch; NULL = not wanted).  BPTT over a sequence is the
 * caller's loop (one call per timestep, newest first), as with the reference module under autograd.  Deterministic. */
int vpk_convlstm_cell_backward_peep(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                                    const float* wci, const float* wcf, const float* wco, const float* dh_out,
                                    const float* dc_out, float* dx, float* dh, float* dc, float* dw, float* db, float* dwci,
                                    float* dwcf, float* dwco, void* stream);

/* Replaces SpatioTemporalLSTMCell.forward with layer_norm=False (model_blocks/predrnn.py:57-83).
 *   weights HOST fp32: w_x [7ch, cin, k, k], w_h [4ch, ch, k, k], w_m [3ch, ch, k, k], w_o [ch, 2ch, k, k],
 *   w_last [ch, 2ch, 1, 1].  x [b, cin, h, w]; h, c, m [b, ch, h, w]; outputs h', c', m', delta_c, delta_m. */
int vpk_stlstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w,
                           int32_t k, const float* w_x, const float* w_h, const float* w_m, const float* w_o,
                           const float* w_last, vpk_cell** out);
/* layer_norm=True (model_blocks/predrnn.py:24-40): the affine parameters of the nn.LayerNorm([k*ch, h, w]) that follows
 * conv_x (k = 7), conv_h (4), conv_m (3), conv_o (1); host fp32 in the reference layout [k*ch, h, w].  Call once after
 * vpk_stlstm_cell_create; the cell then normalises the four conv outputs per sample (eps 1e-5). */
int vpk_stlstm_cell_set_layer_norm(vpk_cell* cell, const float* gx, const float* bx, const float* gh, const float* bh,
                                   const float* gm, const float* bm, const float* go, const float* bo);
int vpk_stlstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                         const float* m, float* h_out, float* c_out, float* m_out, float* dc_out, float* dm_out,
                         void* stream);

/* PredRNN++ cells (Wang et al., ICML 2018).  The north star names them; the reference checkout has no CausalLSTMCell / GHU
 * module, so these entries replace no reference file: they are what a `model_blocks/predrnn_pp.py` written
 * to the paper would bind.  PARITY UNPINNED (checker: oracle/causal.py).
 *   Causal LSTM: weights = HOST fp32 arrays conv_x [7ch, cin, k, k] (i, f, g, i', f', g', o), conv_h [4ch, ch, k, k] (i, f, g, o),
 *   conv_c [3ch, ch, k, k] (i, f, g), conv_m [3ch, cm, k, k] (i', f', m_m; cm = channels of the memory the cell reads,
 *   the width of the cell that wrote it), conv_c2m [4ch, ch, k, k] (i', g', f', o),
 *   conv_om [ch, ch, k, k], conv_last [ch, 2ch, 1, 1]; no biases, forget bias 1.  x [b, cin, h, w]; h, c [b, ch, h, w]; m [b, cm, h, w]; m' [b, ch, h, w]. */
int vpk_causal_lstm_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t cm, int32_t ch, int32_t h,
                                int32_t w, int32_t k, const float* const* weights, vpk_cell** out);
int vpk_causal_lstm_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c,
                              const float* m, float* h_out, float* c_out, float* m_out, void* stream);
/*   Gradient highway unit: w_x, w_z HOST fp32 [2ch, ch, k, k] (rows p, u); z' = sig(u) z + (1 - sig(u)) tanh(p). */
int vpk_ghu_cell_create(int32_t precision, int32_t backend, int32_t ch, int32_t h, int32_t w, int32_t k, const float* w_x,
                        const float* w_z, vpk_cell** out);
int vpk_ghu_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* z, float* z_out, void* stream);

/* Replaces ActionConditionalSpatioTemporalLSTMCell.forward (model_blocks/predrnn.py:142-169).  weights / biases: HOST fp32
 * arrays of the six convs in the order conv_x [7ch, cin, k, k], conv_h [4ch, ch, k, k], conv_a [4ch, ch, k, k],
 * conv_m [3ch, ch, k, k], conv_o [ch, 2ch, k, k], conv_last [ch, 2ch, 1, 1] (every conv of this cell has a bias, :104-140).
 * layer_norm=True: vpk_stlstm_ac_cell_set_layer_norm with (weight, bias) of the LayerNorms after conv_x, conv_h, conv_a,
 * conv_m, conv_o (10 host pointers, reference layout [k*ch, h, w]).  Step: x [b, cin, h, w]; h, c, m, a [b, ch, h, w]
 * (a = the action tensor the model convolved to the latent size); outputs h', c', m', delta_c, delta_m. */
int vpk_stlstm_ac_cell_create(int32_t precision, int32_t backend, int32_t cin, int32_t ch, int32_t h, int32_t w, int32_t k,
                              const float* const* weights, const float* const* biases, vpk_cell** out);
int vpk_stlstm_ac_cell_set_layer_norm(vpk_cell* cell, const float* const* params);
int vpk_stlstm_ac_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* c, const float* m,
                            const float* a, float* h_out, float* c_out, float* m_out, float* dc_out, float* dm_out,
                            void* stream);

/* Replaces PhyCell_Cell.forward with action_conditional=False (model_blocks/phydnet.py:49-62).
 *   weights HOST fp32: conv1 [hid, ch, k, k] + bias, GroupNorm(groups, hid) weight/bias, conv2 [ch, hid, 1, 1] + bias,
 *   convgate [ch, 2ch, 3, 3] + bias.  x, h [b, ch, hh, ww]; output h'. */
int vpk_phycell_cell_create(int32_t precision, int32_t backend, int32_t ch, int32_t hid, int32_t h, int32_t w,
                            int32_t k, const float* conv1_w, const float* conv1_b, const float* gn_w,
                            const float* gn_b, const float* conv2_w, const float* conv2_b, const float* gate_w,
                            const float* gate_b, vpk_cell** out);
int vpk_phycell_cell_step(vpk_cell* cell, int32_t batch, const float* x, const float* h, float* h_out, void* stream);
/* action_conditional=True (model_blocks/phydnet.py:44-55): frame_action_conv / hidden_action_conv, HOST fp32
 * [ch, ch + action_size, 1, 1] + bias [ch] each (action_size <= 8); the step then takes the action vectors, DEVICE fp32
 * [b, action_size], inflates them to the frame size and runs frame / hidden through those 1x1 convs first. */
int vpk_phycell_cell_set_action_convs(vpk_cell* cell, int32_t action_size, const float* frame_w, const float* frame_b,
                                      const float* hidden_w, const float* hidden_b);
int vpk_phycell_cell_step_action(vpk_cell* cell, int32_t batch, const float* x, const float* h, const float* action,
                                 float* h_out, void* stream);

void vpk_cell_destroy(vpk_cell* cell);

/* ---- evaluation metrics (replaces the per-measure torch reductions + .item() syncs of
 *      vp_suite/measure/metric_provider.py:34-73 for MSE / PSNR) ----------------------
 * pred, target: device fp32 [batch, frames, chw] (contiguous [b, p, c, h, w] tensors).  Writes the fp64 device vector
 * out[2 * frames + 1] = [ sum_b sum_chw (p - y)^2 per frame | sum_b 10 * log10(mean_chw (p - y)^2) per frame | batch ]
 * (vp_suite/measure/image_wise.py:19-31, 53-75 before the means): the partial sums one NCCL all-reduce then adds over
 * ranks.  scratch: device fp64 [batch * frames].  Deterministic (fixed-order reductions).  Enqueued on `stream`. */
int vpk_metric_partial_sums(const float* pred, const float* target, int32_t batch, int32_t frames, int64_t chw,
                            double* scratch, double* out, void* stream);

/* SSIM partial sums (vp_suite/measure/image_wise.py:100-121: `1 - piqa.ssim.SSIM()(pred, target)` after
 * base_measure.py:59-75's reshape_clamp).  piqa 1.1.7 is a third-party dependency absent from the reference checkout and
 * from this image: its published algorithm is restated (11-tap Gaussian window, sigma 1.5, valid region, K1 = 0.01,
 * K2 = 0.03, value range 1, mean over channels and positions per image) -- VALUE PARITY UNPINNED; checked against the
 * reference's own SSIM tests' axioms (tests/test_measure.py:26-50) and the oracle's torch restatement.
 * pred, target: device fp32 [batch, frames, c, h, w] in the model's [-1, 1]-mapped convention of reshape_clamp
 * ((v + 1) / 2, clamped to [0, 1]).  out[frames] (device fp64) = sum over the batch of SSIM(pred[b, t], target[b, t]);
 * the reference's measure is 1 - mean of these.  scratch: device fp64 [vpk_metric_ssim_scratch_elems(...)]
 * (-1: image smaller than the 11 x 11 window or too wide).  Deterministic.  Enqueued on `stream`. */
int64_t vpk_metric_ssim_scratch_elems(int32_t batch, int32_t frames, int32_t c, int32_t h, int32_t w);
int vpk_metric_ssim_sums(const float* pred, const float* target, int32_t batch, int32_t frames, int32_t c, int32_t h,
                         int32_t w, double* scratch, double* out, void* stream);

/* ---- misc ------------------------------------------------------------------------------------------------------ */
const char* vpk_last_error(void);
const char* vpk_version(void);
/* 1 when a CUDA device of compute capability 10.x is present, else 0 (never fails). */
int vpk_device_ok(void);

#ifdef __cplusplus
}
#endif
#endif /* VPK_H_ */
