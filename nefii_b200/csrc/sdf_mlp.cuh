// SDF / feature MLP handle (see sdf_mlp.cu).
#pragma once
#include "common.cuh"

namespace nefii {

struct SdfConfig {
  int d_in = 3;
  int n_freqs = 6;      // positional-encoding frequencies (multires)
  int width = 512;      // hidden width == feature size
  int n_hidden = 8;     // Linear+Softplus layers before the 1-wide output layer
  int skip_layer = 4;   // layer whose input is cat([h, PE]) / sqrt(2); <= 0: none
  int d_out = 1;
  // 0: feature vector = input of the last layer (`use_last_as_f`, conf.conf); > 0: the last Linear has 1 + d_feat outputs,
  // row 0 the SDF and rows 1.. the feature vector (use_last_as_f = False, conf_neus.conf)
  int d_feat = 0;
};

// 1: the positional encoding is computed inside layer 0's GEMM (PE prologue); 0 (default): by one encode kernel per evaluation
int sdf_set_pe_prologue(int on);

class SdfNet {
 public:
  SdfNet();
  ~SdfNet();
  int init(const SdfConfig& cfg);
  // PlaneFormat of the inference chain (mlp_gemm.cuh); drops the packed weights: call set_weights afterwards
  int set_format(int fmt);
  int format() const;
  // weights[l]: device fp32 [out_l, in_l] (weight-norm already folded), biases[l]: [out_l]; l = 0..n_hidden
  int set_weights(cudaStream_t stream, const float* const* weights, const float* const* biases);
  size_t workspace_bytes(int rows_cap, bool with_grad) const;
  // x [rows,3] -> sdf [rows], feat [rows,width] (optional), grad [rows,3] (optional).  k_flush: K blocks per TMEM partial of
  // the layer GEMMs (accuracy tier, see gemm_set_k_flush); 0 = the library default.
  int eval(cudaStream_t stream, int rows_cap, const int* count, const float* x, void* workspace, size_t ws_bytes,
           float* sdf, float* feat, float* grad, int k_flush = 0) const;
  const SdfConfig& config() const;

 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace nefii
