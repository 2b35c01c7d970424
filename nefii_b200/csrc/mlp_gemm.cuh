// Split-bf16 ("bf16x3") tcgen05 layer GEMM -- declarations shared by the MLP sequencers.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace nefii {

enum Act : int { ACT_NONE = 0, ACT_SOFTPLUS100 = 1, ACT_RELU = 2, ACT_ELU = 3 };
// Plane format of one launch (operands, output planes, seed planes and saved activations alike; see mlp_gemm_kernel.cuh)
enum PlaneFormat : int { PLANES_BF16 = 0, PLANES_FP16 = 1 };

// A value x is carried between layers as two 16-bit planes (hi = r16(x), lo = r16(x - hi); r16 = bf16 or fp16 rounding).
struct Planes {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int ld = 0;  // elements per row (multiple of 64)
};

struct GemmEpilogue {
  int mode = 0;             // 0: forward layer (bias + activation), 1: backward layer (x act'(saved))
  int act = ACT_NONE;
  int fmt = PLANES_BF16;    // PlaneFormat of every plane this launch reads or writes
  int n_valid = 0;          // output columns [0, n_valid) are real
  const float* bias = nullptr;
  float out_scale = 1.f;    // applied to what is written to dst planes (e.g. 1/sqrt(2) before a skip concat)
  // destination planes: column n -> dst.{hi,lo}[m * dst.ld + dst_col0 + n] for n < dst_ncols
  Planes dst;
  int dst_col0 = 0;
  int dst_ncols = 0;
  int dst_zero_to = 0;      // plane columns [dst_ncols, dst_zero_to) are written as zeros
  int dst_pad_ok = 0;       // 1: a ragged last 128-column span stays on the bulk-store epilogue (the store clips at dst_ncols,
                            // nothing beyond is written: those plane columns may hold other data, e.g. a skip layer's PE half)
  // optional fp32 copy of columns [f32_begin, f32_end): dst_f32[m * f32_ld + n - f32_begin]
  float* dst_f32 = nullptr;
  int f32_ld = 0, f32_begin = 0, f32_end = 0;
  // forward only: fused tiny output layer y[m, q] = sum_n act(z[m, n]) * w_last[q * w_last_ld + n] + b_last[q]
  const float* w_last = nullptr;
  const float* b_last = nullptr;
  int n_last = 0, w_last_ld = 0;
  float* dst_last = nullptr;   // [rows, n_last]
  // forward only: seed of the input-gradient chain, split(w_last[n] * act'(z[m, n])) -> seed planes
  Planes seed;
  // backward only: forward activations that this layer's output gradient is multiplied with
  // (act' is recovered from the saved post-activation value * sav_scale), columns [0, sav_ncols)
  const __nv_bfloat16* sav_hi = nullptr;
  const __nv_bfloat16* sav_lo = nullptr;
  int sav_ld = 0, sav_ncols = 0;
  float sav_scale = 1.f;
};

// A operand generated inside the kernel ("PE prologue"): row m of A = the positional encoding of x[m] (reference embedder.py:22-36:
// [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...], 3 + 6 n_freqs <= 64 columns, zero padded to one 64-wide K block).  Two idle warps
// of the kernel's role warpgroup compute it straight into the swizzled shared-memory tile the tensor core reads; no plane ever
// holds it.  Optionally the same encoding times side_scale is also written to plane columns [side_col0, side_col0 + d_pe) of
// `side` (the PE half of a later skip layer's input).
struct PeSource {
  const float* x = nullptr;   // [rows, 3] fp32; nullptr: A comes from planes (a_hi / a_lo) by TMA
  int n_freqs = 0;
  Planes side;                // optional
  int side_col0 = 0;
  float side_scale = 1.f;
};

struct GemmProblem {
  // A: activations [rows_cap, k_pad] as planes; B: weights [n_pad, k_pad] as planes (K-major both)
  const __nv_bfloat16* a_hi; const __nv_bfloat16* a_lo; int a_ld; int rows_cap;
  const __nv_bfloat16* b_hi; const __nv_bfloat16* b_lo; int b_ld; int n_pad;  // n_pad multiple of 256
  int k_pad;                // multiple of 64, <= a_ld, <= b_ld
  const int* count;         // device: number of valid rows (nullptr -> rows_cap)
  int k_splits = 1;         // >1: split-K over gridDim.y, partial s is written to dst_f32 + s * f32_split_stride
  long long f32_split_stride = 0;
  int* k_splits_used = nullptr;   // host out: number of partials actually produced
  int k_flush = 0;          // K blocks per TMEM partial for this launch (accuracy tier); 0: the library default
  PeSource pe;              // pe.x != nullptr: a_hi / a_lo are ignored, k_pad must be 64
  GemmEpilogue epi;
};

int gemm_split_bf16(cudaStream_t stream, const GemmProblem& p);

// Per-launch device timing of the layer GEMM (bench.py roofline): while enabled every launch is bracketed by CUDA
// events on its stream and its row count is copied back; fetch() -> {total ms, total algorithmic flops, launches}.
// upper bound for the TMA-multicast cluster size the launcher picks (tuning / A-B measurements)
int gemm_set_cluster(int cl);
int gemm_set_debug(int mask);
// programmatic dependent launch of the layer GEMMs (default on; NEFII_GEMM_PDL=0 at load)
int gemm_set_pdl(int on);
// accuracy / overlap knob: number of 64-wide K blocks accumulated inside TMEM (truncating adder) before the partial sum is
// added to the fp32 register accumulators (round to nearest).  1 = most accurate, 2 = default, >= K/64 = everything in TMEM.
int gemm_set_k_flush(int k);
int gemm_set_k_flush_head(int k);
// First-order compensation of the truncating accumulator: a TMEM partial sum of `k_blocks` K blocks is multiplied by
// (1 + rho) when it is added to the register accumulators.  rho = 0 (default) leaves the sum untouched.
int gemm_set_trunc_comp(int k_blocks, float rho, int fmt = PLANES_BF16);
int gemm_profile_enable(int on);
int gemm_profile_fetch(double* out3);
bool gemm_profile_active();
// per-device one-time set-up (dynamic shared memory opt-in of every kernel variant, SM count); cheap when already done
int gemm_prepare_device();
// SM count of the current device (after gemm_prepare_device)
int gemm_device_sms();
// upper bound on the persistent grid of the following launches (0 = every SM); NOT part of the epoch below
int gemm_set_grid_cap(int sms);
int gemm_grid_cap();
// bumped by every setter above: anything that caches launches (trace graphs) keys on it
long long gemm_config_epoch();

// fp32 [rows, cols] (row stride ld_src) -> zero-padded bf16 planes [rows_pad, cols_pad];
// transpose=1 writes src^T.
int split_to_planes(cudaStream_t stream, const float* src, int rows, int cols, int ld_src, int transpose,
                    float scale, __nv_bfloat16* hi, __nv_bfloat16* lo, int rows_pad, int cols_pad, int fmt = PLANES_BF16);

}  // namespace nefii
