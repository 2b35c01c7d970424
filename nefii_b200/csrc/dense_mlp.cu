// Helper kernels around the tcgen05 layer GEMM for the trainable dense stacks (radiance net:
// RenderingNetwork, implicit_differentiable_renderer.py:196-241; material net: EnvmapMaterialNetwork,
// sg_envmap_material.py:357-425): input assembly (positional encodings + concat -> bf16 planes), plane
// transposes that feed the weight-gradient GEMMs (K = points), the backward of the fused tiny output layer,
// and the reduction of split-K partials.
#include <cuda_bf16.h>
#include <algorithm>
#include "mlp_gemm.cuh"

namespace nefii {

namespace {

__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

struct Segments {
  const float* src[4];
  int width[4];     // row width of the source
  int n_freqs[4];   // >= 0: positional encoding of a 3-vector with that many frequencies; -1: raw copy
  int begin[5];     // output column range of each segment
  int n_seg;
};

__global__ void assemble_input_kernel(Segments S, int rows, int k_pad, Planes dst) {
  const long long total = (long long)rows * k_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / k_pad), col = (int)(i % k_pad);
    float v = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < S.n_seg && col >= S.begin[s] && col < S.begin[s + 1]) {
        const int j = col - S.begin[s];
        const float* p = S.src[s] + (size_t)row * S.width[s];
        if (S.n_freqs[s] < 0) {
          v = p[j];
        } else if (j < 3) {
          v = p[j];
        } else {
          const int k = (j - 3) / 6, r = (j - 3) % 6;
          const float a = p[r % 3] * exp2f((float)k);
          v = r < 3 ? sinf(a) : cosf(a);
        }
      }
    }
    __nv_bfloat16 h, l;
    split2(v, h, l);
    dst.hi[(size_t)row * dst.ld + col] = h;
    dst.lo[(size_t)row * dst.ld + col] = l;
  }
}

// dst[c][r] = src[r][c] for r < rows, c < cols; dst is [cols_pad, ld_dst] and zero-filled elsewhere.
// Optional column sums of (hi + lo) into col_sum[c] (bias gradient).
__global__ void __launch_bounds__(256)
transpose_planes_kernel(const __nv_bfloat16* __restrict__ s_hi, const __nv_bfloat16* __restrict__ s_lo, int ld_src, int rows,
                        int cols, __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, int ld_dst,
                        int rows_pad, int cols_pad, float* __restrict__ col_sum) {
  __shared__ __nv_bfloat16 th[32][33], tl[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
    if (r < rows && c < cols) { h = s_hi[(size_t)r * ld_src + c]; l = s_lo[(size_t)r * ld_src + c]; }
    th[j][tx] = h; tl[j][tx] = l;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < cols_pad && r < rows_pad) { d_hi[(size_t)c * ld_dst + r] = th[tx][j]; d_lo[(size_t)c * ld_dst + r] = tl[tx][j]; }
  }
  if (col_sum != nullptr && ty == 0) {
    float acc = 0.f;
    for (int j = 0; j < 32; ++j) acc += __bfloat162float(th[j][tx]) + __bfloat162float(tl[j][tx]);
    if (c0 + tx < cols) atomicAdd(col_sum + c0 + tx, acc);
  }
}

// Backward of y = h W_last^T + b_last (n_out <= 4) with h = act(z):
//   G[row, n] = (sum_q gy[row, q] W_last[q, n]) * act'(h[row, n])     -> planes
//   gW_last[q, n] += sum_rows gy[row, q] h[row, n] ; gb_last[q] += sum_rows gy[row, q]
template <int ACT>
__global__ void __launch_bounds__(256)
last_layer_bwd_kernel(int rows, int width, int n_out, const float* __restrict__ gy, const float* __restrict__ w_last,
                      Planes h, Planes g_out, float* __restrict__ gw_last, float* __restrict__ gb_last, int rows_per_block) {
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(rows, r_begin + rows_per_block);
  for (int n = threadIdx.x; n < width; n += blockDim.x) {
    float wq[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < n_out; ++q) wq[q] = w_last[(size_t)q * width + n];
    for (int r = r_begin; r < r_end; ++r) {
      const size_t off = (size_t)r * h.ld + n;
      const float hv = __bfloat162float(h.hi[off]) + __bfloat162float(h.lo[off]);
      float g = 0.f;
      for (int q = 0; q < n_out; ++q) {
        const float gq = gy[(size_t)r * n_out + q];
        g += gq * wq[q];
        acc[q] += gq * hv;
      }
      float d;
      if (ACT == ACT_RELU) d = hv > 0.f ? 1.f : 0.f;
      else if (ACT == ACT_ELU) d = hv > 0.f ? 1.f : hv + 1.f;
      else if (ACT == ACT_SOFTPLUS100) d = 1.f - expf(-100.f * hv);
      else d = 1.f;
      __nv_bfloat16 a, b;
      split2(g * d, a, b);
      g_out.hi[(size_t)r * g_out.ld + n] = a;
      g_out.lo[(size_t)r * g_out.ld + n] = b;
    }
    for (int q = 0; q < n_out; ++q) atomicAdd(gw_last + (size_t)q * width + n, acc[q]);
  }
  if (threadIdx.x < n_out) {
    float acc = 0.f;
    for (int r = r_begin; r < r_end; ++r) acc += gy[(size_t)r * n_out + threadIdx.x];
    atomicAdd(gb_last + threadIdx.x, acc);
  }
}

__global__ void reduce_splits_kernel(const float* __restrict__ partial, int n_splits, long long stride, long long n,
                                     int ld_src, int cols, float* __restrict__ out) {
  // partial s: [rows, ld_src]; out: [rows, cols] contiguous
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i % cols;
    float acc = 0.f;
    for (int s = 0; s < n_splits; ++s) acc += partial[(long long)s * stride + r * ld_src + c];
    out[i] = acc;
  }
}

}  // namespace

int assemble_input(cudaStream_t stream, int rows, int n_seg, const float* const* src, const int* width, const int* n_freqs,
                   __nv_bfloat16* hi, __nv_bfloat16* lo, int ld, int k_pad) {
  NEFII_CHECK_ARG(n_seg >= 1 && n_seg <= 4 && hi && lo && k_pad <= ld, "assemble_input: bad arguments");
  if (rows <= 0) return NEFII_OK;
  Segments S{};
  S.n_seg = n_seg;
  int col = 0;
  for (int s = 0; s < n_seg; ++s) {
    NEFII_CHECK_ARG(src[s] != nullptr, "assemble_input: null segment");
    S.src[s] = src[s]; S.width[s] = width[s]; S.n_freqs[s] = n_freqs[s];
    S.begin[s] = col;
    col += n_freqs[s] >= 0 ? 3 + 6 * n_freqs[s] : width[s];
  }
  S.begin[n_seg] = col;
  NEFII_CHECK_ARG(col <= k_pad, "assemble_input: segments (%d columns) exceed k_pad (%d)", col, k_pad);
  Planes dst; dst.hi = hi; dst.lo = lo; dst.ld = ld;
  const long long total = (long long)rows * k_pad;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  assemble_input_kernel<<<blocks, 256, 0, stream>>>(S, rows, k_pad, dst);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int transpose_planes(cudaStream_t stream, const __nv_bfloat16* s_hi, const __nv_bfloat16* s_lo, int ld_src, int rows, int cols,
                     __nv_bfloat16* d_hi, __nv_bfloat16* d_lo, int ld_dst, int rows_pad, int cols_pad, float* col_sum) {
  NEFII_CHECK_ARG(s_hi && s_lo && d_hi && d_lo && rows_pad <= ld_dst && rows <= rows_pad && cols <= cols_pad,
                  "transpose_planes: bad arguments");
  if (rows_pad <= 0 || cols_pad <= 0) return NEFII_OK;
  dim3 grid(ceil_div(rows_pad, 32), ceil_div(cols_pad, 32));
  transpose_planes_kernel<<<grid, 256, 0, stream>>>(s_hi, s_lo, ld_src, rows, cols, d_hi, d_lo, ld_dst, rows_pad, cols_pad, col_sum);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int last_layer_bwd(cudaStream_t stream, int act, int rows, int width, int n_out, const float* gy, const float* w_last,
                   const __nv_bfloat16* h_hi, const __nv_bfloat16* h_lo, int h_ld, __nv_bfloat16* g_hi, __nv_bfloat16* g_lo,
                   int g_ld, float* gw_last, float* gb_last) {
  NEFII_CHECK_ARG(n_out >= 1 && n_out <= 4 && gy && w_last && h_hi && h_lo && g_hi && g_lo && gw_last && gb_last,
                  "last_layer_bwd: bad arguments");
  if (rows <= 0) return NEFII_OK;
  Planes h; h.hi = const_cast<__nv_bfloat16*>(h_hi); h.lo = const_cast<__nv_bfloat16*>(h_lo); h.ld = h_ld;
  Planes g; g.hi = g_hi; g.lo = g_lo; g.ld = g_ld;
  const int rows_per_block = std::max(16, ceil_div(rows, kNumSMs * 4));
  const int blocks = ceil_div(rows, rows_per_block);
  switch (act) {
    case ACT_RELU: last_layer_bwd_kernel<ACT_RELU><<<blocks, 256, 0, stream>>>(rows, width, n_out, gy, w_last, h, g, gw_last, gb_last, rows_per_block); break;
    case ACT_ELU: last_layer_bwd_kernel<ACT_ELU><<<blocks, 256, 0, stream>>>(rows, width, n_out, gy, w_last, h, g, gw_last, gb_last, rows_per_block); break;
    case ACT_SOFTPLUS100: last_layer_bwd_kernel<ACT_SOFTPLUS100><<<blocks, 256, 0, stream>>>(rows, width, n_out, gy, w_last, h, g, gw_last, gb_last, rows_per_block); break;
    default: last_layer_bwd_kernel<ACT_NONE><<<blocks, 256, 0, stream>>>(rows, width, n_out, gy, w_last, h, g, gw_last, gb_last, rows_per_block); break;
  }
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int reduce_splits(cudaStream_t stream, const float* partial, int n_splits, long long stride, int rows, int ld_src, int cols,
                  float* out) {
  NEFII_CHECK_ARG(partial && out && n_splits >= 1, "reduce_splits: bad arguments");
  const long long n = (long long)rows * cols;
  if (n <= 0) return NEFII_OK;
  int blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  reduce_splits_kernel<<<blocks, 256, 0, stream>>>(partial, n_splits, stride, n, ld_src, cols, out);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
