// Single-step cells behind the VPModelBlock boundary (fp32 NCHW device tensors in and out).
#pragma once
#include "common.h"

namespace vpk {

class Cell {
 public:
  virtual ~Cell() {}
  // in / out: up to 8 device pointers each, meaning defined per cell kind (see api.cu)
  virtual void step(int batch, const float* const* in, float* const* out, cudaStream_t stream) = 0;
  // ST-LSTM only: LayerNorm affine parameters (gamma, beta) x (conv_x, conv_h, conv_m, conv_o), host, [k*C, H, W]
  virtual void set_layer_norm(const float* const* params) { VPK_THROW(1, "this cell kind has no LayerNorm variant"); }
  // ConvLSTM cell (ndrplz form) only: gradients of one step.  in: x, h, c, dh_out, dc_out (either gradient may be null);
  // out: dx, dh, dc [b, ., h, w] and dw [4ch, cin + ch, k, k], db [4ch] (all device fp32)
  virtual void backward(int batch, const float* const* in, float* const* out, cudaStream_t stream) {
    VPK_THROW(1, "this cell kind has no backward pass");
  }
  // PhyCell only: the two 1x1 action convs of the action-conditional cell (host; [ch, ch + a, 1, 1] + bias each)
  virtual void set_action_convs(int action_size, const float* fw, const float* fb, const float* hw, const float* hb) {
    VPK_THROW(1, "this cell kind has no action-conditional variant");
  }
};

Cell* make_convlstm_cell(int precision, int backend, int cin, int ch, int h, int w, int k, int gate_order,
                         const float* weight, const float* bias);
Cell* make_stlstm_cell(int precision, int backend, int cin, int ch, int h, int w, int k, const float* w_x,
                       const float* w_h, const float* w_m, const float* w_o, const float* w_last);
// ActionConditionalSpatioTemporalLSTMCell: weights / biases of conv_x, conv_h, conv_a, conv_m, conv_o, conv_last (host)
Cell* make_stlstm_ac_cell(int precision, int backend, int cin, int ch, int h, int w, int k, const float* const* weights,
                          const float* const* biases);
Cell* make_phycell_cell(int precision, int backend, int ch, int hid, int h, int w, int k, const float* conv1_w,
                        const float* conv1_b, const float* gn_w, const float* gn_b, const float* conv2_w,
                        const float* conv2_b, const float* gate_w, const float* gate_b);

// PredRNN++ (causal.h): weights = conv_x, conv_h, conv_c, conv_m, conv_c2m, conv_om, conv_last (host)
Cell* make_causal_lstm_cell(int precision, int backend, int cin, int cm, int ch, int h, int w, int k,
                            const float* const* weights);   // cm: channels of the spatial memory the cell READS
Cell* make_ghu_cell(int precision, int backend, int ch, int h, int w, int k, const float* w_x, const float* w_z);

}  // namespace vpk
