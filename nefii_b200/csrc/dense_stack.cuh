// Native sequencer of the trainable dense stacks (see dense_stack.cu).
#pragma once
#include "common.cuh"

namespace nefii {

constexpr int kDenseMaxLayers = 17;     // hidden layers + the output layer

struct DenseStack {
  int rows = 0;
  int n_hidden = 0;                       // Linear + activation layers before the (1..4)-wide output layer
  int act = 0;                            // Act of the hidden layers
  int n_seg = 0;                          // input = concat of segments; n_freqs >= 0: positional encoding of a 3-vector, -1: raw copy
  const float* seg_src[4] = {};
  int seg_width[4] = {};
  int seg_freqs[4] = {};
  const float* weights[kDenseMaxLayers] = {};   // effective fp32 [out, in] row-major, hidden layers then the output layer
  const float* biases[kDenseMaxLayers] = {};
  int dim_in[kDenseMaxLayers] = {};
  int dim_out[kDenseMaxLayers] = {};
  int need_grad = 0;                      // forward: keep every activation plane in the workspace for dense_stack_bwd
  void* workspace = nullptr;
  long long workspace_bytes = 0;
  float* y = nullptr;                     // forward out [rows, n_out]
  const float* gy = nullptr;              // backward in [rows, n_out]
  float* grad_w[kDenseMaxLayers] = {};    // backward out, same shapes as weights / biases (overwritten)
  float* grad_b[kDenseMaxLayers] = {};
};

long long dense_stack_workspace_bytes(const DenseStack& d);
int dense_stack_fwd(cudaStream_t stream, const DenseStack& d);
int dense_stack_bwd(cudaStream_t stream, const DenseStack& d);

}  // namespace nefii
