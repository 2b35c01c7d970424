// Native sequencer of the trainable dense stacks (reference RenderingNetwork.forward, implicit_differentiable_renderer.py:196-241;
// EnvmapMaterialNetwork.diffuse_albedo_layers, sg_envmap_material.py:357-366): ONE host call runs the whole forward (input
// assembly, weight packing, the hidden-layer GEMMs, the fused output layer) and ONE the whole backward (output layer, then per layer
// the two transposes, the split-K weight-gradient GEMM and its reduction, the data-gradient GEMM).  nefii_b200/mlp.py used to issue
// these ~20 + ~60 launches one ctypes call at a time: at the 1/8-batch size of an 8-GPU step the stacks were bound by the host's
// ~20 us per call, not by the GPU.  Same kernels, same order, same arithmetic as the Python sequence it replaces.
//
// All temporaries live in one caller-owned workspace (dense_stack_workspace_bytes); the forward leaves the activation planes
// there for the backward (need_grad) -- the caller keeps the workspace alive between the two.
#include <algorithm>
#include <cuda_bf16.h>
#include "common.cuh"
#include "mlp_gemm.cuh"
#include "dense_stack.cuh"

namespace nefii {

int assemble_input(cudaStream_t, int, int, const float* const*, const int*, const int*, __nv_bfloat16*, __nv_bfloat16*, int, int);
int transpose_planes(cudaStream_t, const __nv_bfloat16*, const __nv_bfloat16*, int, int, int, __nv_bfloat16*, __nv_bfloat16*, int, int, int, float*);
int last_layer_bwd(cudaStream_t, int, int, int, int, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, int,
                   __nv_bfloat16*, __nv_bfloat16*, int, float*, float*);
int reduce_splits(cudaStream_t, const float*, int, long long, int, int, int, float*);
int gemm_device_sms();

namespace {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct Layout {
  int d_in = 0, k0 = 0, n_out = 0, width_last = 0, n_pad = 0, max_out = 0, max_in = 0;
  size_t off_in0 = 0, off_act[kDenseMaxLayers] = {}, off_wfwd[kDenseMaxLayers] = {}, off_g[2] = {}, off_gt = 0, off_ht = 0, off_partial = 0, off_wt = 0, total = 0;
  int act_ld[kDenseMaxLayers] = {};      // plane row stride of the output of hidden layer l
};

int make_layout(const DenseStack& d, int n_sms, Layout& L) {
  NEFII_CHECK_ARG(d.n_hidden >= 1 && d.n_hidden < kDenseMaxLayers, "dense_stack: n_hidden out of range");
  NEFII_CHECK_ARG(d.n_seg >= 1 && d.n_seg <= 4, "dense_stack: 1..4 input segments");
  L.d_in = 0;
  for (int s = 0; s < d.n_seg; ++s) L.d_in += d.seg_freqs[s] >= 0 ? 3 + 6 * d.seg_freqs[s] : d.seg_width[s];
  NEFII_CHECK_ARG(L.d_in == d.dim_in[0], "dense_stack: the segments give %d input columns, layer 0 takes %d", L.d_in, d.dim_in[0]);
  for (int l = 0; l < d.n_hidden; ++l)
    NEFII_CHECK_ARG(d.dim_in[l + 1] == d.dim_out[l], "dense_stack: layer %d output %d != next input %d", l, d.dim_out[l], d.dim_in[l + 1]);
  L.n_out = d.dim_out[d.n_hidden];
  NEFII_CHECK_ARG(L.n_out >= 1 && L.n_out <= 4, "dense_stack: the output layer has 1..4 outputs");
  L.width_last = d.dim_in[d.n_hidden];
  L.k0 = round_up(L.d_in, 64);
  const size_t rows = (size_t)std::max(d.rows, 1);
  L.n_pad = round_up(std::max(d.rows, 1), 64);
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t o = p; p = align256(p + bytes); return o; };
  L.off_in0 = take(2 * rows * L.k0 * 2);
  size_t partial = 0;
  for (int l = 0; l < d.n_hidden; ++l) {
    L.act_ld[l] = round_up(d.dim_out[l], 64);
    L.max_out = std::max(L.max_out, d.dim_out[l]);
    L.max_in = std::max(L.max_in, d.dim_in[l]);
    const int gt_tiles = round_up(d.dim_out[l], 128) / 128;
    const int splits = std::max(1, std::min(L.n_pad / 128, (n_sms + gt_tiles - 1) / gt_tiles));
    partial = std::max(partial, (size_t)splits * d.dim_out[l] * d.dim_in[l] * 4);
  }
  const int max_ld = round_up(L.max_out, 64);
  if (d.need_grad) {
    for (int l = 0; l < d.n_hidden; ++l) L.off_act[l] = take(2 * rows * L.act_ld[l] * 2);
  } else {        // no backward follows: two buffers, used in turn
    const size_t o0 = take(2 * rows * max_ld * 2), o1 = take(2 * rows * max_ld * 2);
    for (int l = 0; l < d.n_hidden; ++l) L.off_act[l] = (l & 1) ? o1 : o0;
  }
  // every layer's packed weights have their own buffer: all of them are packed before the first GEMM, so the layer GEMMs
  // follow one another in the stream (programmatic dependent launch overlaps layer l+1's prologue with layer l's tail)
  for (int l = 0; l < d.n_hidden; ++l) L.off_wfwd[l] = take((size_t)2 * round_up(d.dim_out[l], 256) * round_up(d.dim_in[l], 64) * 2);
  if (d.need_grad) {             // backward temporaries
    const size_t g_bytes = 2 * rows * round_up(std::max(L.max_out, L.max_in), 64) * 2;
    L.off_g[0] = take(g_bytes);
    L.off_g[1] = take(g_bytes);
    L.off_gt = take((size_t)2 * round_up(L.max_out, 128) * L.n_pad * 2);
    L.off_ht = take((size_t)2 * round_up(L.max_in, 256) * L.n_pad * 2);
    L.off_partial = take(partial);
    L.off_wt = take((size_t)2 * round_up(L.max_in, 256) * round_up(L.max_out, 64) * 2);
  }
  L.total = p + 256;
  return NEFII_OK;
}

Planes planes_at(char* base, size_t off, size_t rows, int ld) {
  Planes pl;
  pl.hi = (__nv_bfloat16*)(base + off);
  pl.lo = (__nv_bfloat16*)(base + off + rows * ld * 2);
  pl.ld = ld;
  return pl;
}

}  // namespace

long long dense_stack_workspace_bytes(const DenseStack& d) {
  Layout L;
  if (make_layout(d, gemm_device_sms(), L)) return -1;
  return (long long)L.total;
}

int dense_stack_fwd(cudaStream_t stream, const DenseStack& d) {
  int rc;
  if ((rc = gemm_prepare_device())) return rc;
  Layout L;
  if ((rc = make_layout(d, gemm_device_sms(), L))) return rc;
  NEFII_CHECK_ARG(d.rows > 0, "dense_stack_fwd: no rows");
  NEFII_CHECK_ARG(d.workspace && d.workspace_bytes >= (long long)L.total, "dense_stack_fwd: workspace too small");
  NEFII_CHECK_ARG(d.y != nullptr, "dense_stack_fwd: null output");
  char* base = (char*)(((uintptr_t)d.workspace + 255) & ~(uintptr_t)255);
  const size_t rows = (size_t)d.rows;
  Planes in0 = planes_at(base, L.off_in0, rows, L.k0);
  if ((rc = assemble_input(stream, d.rows, d.n_seg, d.seg_src, d.seg_width, d.seg_freqs, in0.hi, in0.lo, L.k0, L.k0))) return rc;
  const int H = d.n_hidden;
  for (int l = 0; l < H; ++l) {
    const int out = d.dim_out[l], in = d.dim_in[l];
    const int k_pad = round_up(in, 64), n_pad = round_up(out, 256);
    __nv_bfloat16* w_hi = (__nv_bfloat16*)(base + L.off_wfwd[l]);
    if ((rc = split_to_planes(stream, d.weights[l], out, in, in, 0, 1.f, w_hi, w_hi + (size_t)n_pad * k_pad, n_pad, k_pad))) return rc;
  }
  Planes a = in0;
  for (int l = 0; l < H; ++l) {
    const int out = d.dim_out[l], in = d.dim_in[l];
    const int k_pad = round_up(in, 64), n_pad = round_up(out, 256);
    __nv_bfloat16* w_hi = (__nv_bfloat16*)(base + L.off_wfwd[l]);
    __nv_bfloat16* w_lo = w_hi + (size_t)n_pad * k_pad;
    GemmProblem g{};
    g.a_hi = a.hi; g.a_lo = a.lo; g.a_ld = a.ld; g.rows_cap = d.rows;
    g.b_hi = w_hi; g.b_lo = w_lo; g.b_ld = k_pad; g.n_pad = n_pad; g.k_pad = k_pad;
    g.count = nullptr;
    g.epi.mode = 0; g.epi.act = d.act; g.epi.n_valid = out; g.epi.bias = d.biases[l];
    Planes dst = planes_at(base, L.off_act[l], rows, L.act_ld[l]);      // without a backward: two buffers used in turn
    if (l < H - 1) {
      g.epi.dst = dst; g.epi.dst_ncols = out; g.epi.dst_zero_to = L.act_ld[l];
    } else {
      if (d.need_grad) { g.epi.dst = dst; g.epi.dst_ncols = out; g.epi.dst_zero_to = L.act_ld[l]; }
      g.epi.w_last = d.weights[H]; g.epi.b_last = d.biases[H]; g.epi.n_last = L.n_out; g.epi.w_last_ld = L.width_last;
      g.epi.dst_last = d.y;
    }
    if ((rc = gemm_split_bf16(stream, g))) return rc;
    a = dst;
  }
  return NEFII_OK;
}

int dense_stack_bwd(cudaStream_t stream, const DenseStack& d) {
  int rc;
  if ((rc = gemm_prepare_device())) return rc;
  const int n_sms = gemm_device_sms();
  Layout L;
  if ((rc = make_layout(d, n_sms, L))) return rc;
  NEFII_CHECK_ARG(d.rows > 0 && d.need_grad, "dense_stack_bwd: needs the workspace of a forward with need_grad");
  NEFII_CHECK_ARG(d.workspace && d.workspace_bytes >= (long long)L.total, "dense_stack_bwd: workspace too small");
  const int H = d.n_hidden, n = d.rows;
  NEFII_CHECK_ARG(d.gy && d.grad_w[H] && d.grad_b[H], "dense_stack_bwd: null gradient pointers of the output layer");
  for (int l = 0; l < H; ++l) NEFII_CHECK_ARG(d.grad_w[l] && d.grad_b[l], "dense_stack_bwd: null gradient pointer of layer %d", l);
  char* base = (char*)(((uintptr_t)d.workspace + 255) & ~(uintptr_t)255);
  const size_t rows = (size_t)n;
  Planes in0 = planes_at(base, L.off_in0, rows, L.k0);
  auto act_of = [&](int l) { return l == 0 ? in0 : planes_at(base, L.off_act[l - 1], rows, L.act_ld[l - 1]); };   // input of layer l
  // output layer: d y / d h_L, its weight / bias gradients (accumulated: zeroed here)
  NEFII_CUDA(cudaMemsetAsync(d.grad_w[H], 0, (size_t)L.n_out * L.width_last * 4, stream));
  NEFII_CUDA(cudaMemsetAsync(d.grad_b[H], 0, (size_t)L.n_out * 4, stream));
  const Planes hL = act_of(H);
  Planes G = planes_at(base, L.off_g[0], rows, hL.ld);
  int which = 0;
  if ((rc = last_layer_bwd(stream, d.act, n, L.width_last, L.n_out, d.gy, d.weights[H], hL.hi, hL.lo, hL.ld, G.hi, G.lo, G.ld,
                           d.grad_w[H], d.grad_b[H])))
    return rc;
  const int n_pad = L.n_pad;
  for (int l = H - 1; l >= 0; --l) {
    const int out = d.dim_out[l], in = d.dim_in[l];
    NEFII_CUDA(cudaMemsetAsync(d.grad_b[l], 0, (size_t)out * 4, stream));
    const int gt_rows = round_up(out, 128), ht_rows = round_up(in, 256);
    Planes GT = planes_at(base, L.off_gt, (size_t)gt_rows, n_pad);
    if ((rc = transpose_planes(stream, G.hi, G.lo, G.ld, n, out, GT.hi, GT.lo, n_pad, n_pad, gt_rows, d.grad_b[l]))) return rc;
    const Planes h_prev = act_of(l);
    Planes HT = planes_at(base, L.off_ht, (size_t)ht_rows, n_pad);
    if ((rc = transpose_planes(stream, h_prev.hi, h_prev.lo, h_prev.ld, n, in, HT.hi, HT.lo, n_pad, n_pad, ht_rows, nullptr))) return rc;
    const int splits = std::max(1, std::min(n_pad / 128, (n_sms + (gt_rows / 128) - 1) / (gt_rows / 128)));
    float* partial = (float*)(base + L.off_partial);
    int used = 1;
    {
      GemmProblem g{};
      g.a_hi = GT.hi; g.a_lo = GT.lo; g.a_ld = GT.ld; g.rows_cap = out;
      g.b_hi = HT.hi; g.b_lo = HT.lo; g.b_ld = HT.ld; g.n_pad = ht_rows; g.k_pad = n_pad;
      g.count = nullptr;
      g.k_splits = splits; g.f32_split_stride = (long long)out * in; g.k_splits_used = &used;
      g.epi.mode = 0; g.epi.act = ACT_NONE; g.epi.n_valid = in;
      g.epi.dst_f32 = partial; g.epi.f32_ld = in; g.epi.f32_begin = 0; g.epi.f32_end = in;
      if ((rc = gemm_split_bf16(stream, g))) return rc;
    }
    if ((rc = reduce_splits(stream, partial, used, (long long)out * in, out, in, in, d.grad_w[l]))) return rc;
    if (l > 0) {
      const int wt_rows = round_up(in, 256), wt_ld = round_up(out, 64);
      __nv_bfloat16* wt_hi = (__nv_bfloat16*)(base + L.off_wt);
      __nv_bfloat16* wt_lo = wt_hi + (size_t)wt_rows * wt_ld;
      if ((rc = split_to_planes(stream, d.weights[l], out, in, in, 1, 1.f, wt_hi, wt_lo, wt_rows, wt_ld))) return rc;
      Planes G_prev = planes_at(base, L.off_g[which ^ 1], rows, round_up(in, 64));
      GemmProblem g{};
      g.a_hi = G.hi; g.a_lo = G.lo; g.a_ld = G.ld; g.rows_cap = n;
      g.b_hi = wt_hi; g.b_lo = wt_lo; g.b_ld = wt_ld; g.n_pad = wt_rows; g.k_pad = wt_ld;
      g.count = nullptr;
      g.epi.mode = 1; g.epi.act = d.act; g.epi.n_valid = in;
      g.epi.dst = G_prev; g.epi.dst_ncols = in; g.epi.dst_zero_to = round_up(in, 64);
      g.epi.sav_hi = h_prev.hi; g.epi.sav_lo = h_prev.lo; g.epi.sav_ld = h_prev.ld; g.epi.sav_ncols = in; g.epi.sav_scale = 1.f;
      if ((rc = gemm_split_bf16(stream, g))) return rc;
      G = G_prev;
      which ^= 1;
    }
  }
  return NEFII_OK;
}

}  // namespace nefii
