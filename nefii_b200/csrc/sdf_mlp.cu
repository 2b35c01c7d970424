// SDF / feature MLP (reference ImplicitNetwork, code/model/implicit_differentiable_renderer.py:18-123)
// sequenced over the tcgen05 layer GEMM: positional encoding (one kernel, or inside layer 0's GEMM: mlp_gemm_kernel.cuh "PE
// prologue") -> n_hidden x (Linear + Softplus(100)),
// skip concat at `skip_layer`, fused 1-wide output layer, and the closed-form input gradient
// (d sdf / d x, what ImplicitNetwork.gradient obtains through autograd) as a reverse chain of GEMMs.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "mlp_gemm.cuh"
#include "sdf_mlp.cuh"

namespace nefii {

namespace {

constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kSqrt2 = 1.41421356237309505f;

// two neighbouring plane elements (hi words, lo words) of format fmt (PlaneFormat, mlp_gemm.cuh)
__device__ __forceinline__ void split_pair(float a, float b, int fmt, uint32_t& wh, uint32_t& wl) {
  if (fmt == PLANES_FP16) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
    wh = *reinterpret_cast<const uint32_t*>(&h2);
    wl = *reinterpret_cast<const uint32_t*>(&l2);
  } else {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    wh = *reinterpret_cast<const uint32_t*>(&h2);
    wl = *reinterpret_cast<const uint32_t*>(&l2);
  }
}

// ONE pass over the points writes both copies of the positional encoding an evaluation needs: PE(x) as the 64-column input
// planes of layer 0 (zero padded), and PE(x) / sqrt(2) into plane columns [side_col0, side_col0 + d_pe) of the skip layer's
// input (next to the h columns that layer skip - 1 writes later; its bulk store leaves these columns alone).
// One lane per point: 3 x n_freqs sincosf calls give the whole encoding of a row (layout of embedder.py:22-36:
// [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...], 3 components each); the warp's 32 rows are staged in shared memory and
// written back with consecutive lanes on consecutive columns (16-byte stores where the destination allows it).
constexpr int kEncWarps = 4;
constexpr int kEncMaxWidth = 64;
constexpr int kEncStride = kEncMaxWidth + 1;   // odd stride: conflict-free row-major writes by lane = row

__global__ void __launch_bounds__(kEncWarps * 32)
encode_kernel(const float* __restrict__ x, const int* __restrict__ count, int rows_cap, int n_freqs, Planes dst, Planes side,
              int side_col0, float side_scale, int fmt) {
  __shared__ float s_v[kEncWarps][32 * kEncStride];
  int limit = rows_cap;
  if (count) limit = min(limit, *count);
  const int d_pe = 3 + 6 * n_freqs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sv = s_v[warp];
  const int n_groups = (limit + 31) / 32;
  for (int grp = blockIdx.x * kEncWarps + warp; grp < n_groups; grp += gridDim.x * kEncWarps) {
    const int row0 = grp * 32;
    const int row = row0 + lane;
    if (row < limit) {
      const float p[3] = {x[(size_t)row * 3 + 0], x[(size_t)row * 3 + 1], x[(size_t)row * 3 + 2]};
      float* o = sv + lane * kEncStride;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = p[c];
      for (int k = 0; k < n_freqs; ++k) {
        const float f = exp2f((float)k);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sn, cs;
          sincosf(p[c] * f, &sn, &cs);
          o[3 + 6 * k + c] = sn;
          o[6 + 6 * k + c] = cs;
        }
      }
      for (int j = d_pe; j < kEncMaxWidth; ++j) o[j] = 0.f;
    }
    __syncwarp();
    const int n_rows = min(32, limit - row0);
    // layer 0's input planes: whole 64-column rows, 16 bytes per store
    for (int e = lane; e < n_rows * (kEncMaxWidth / 8); e += 32) {
      const int r = e / (kEncMaxWidth / 8), c0 = (e % (kEncMaxWidth / 8)) * 8;
      const float* v = sv + r * kEncStride + c0;
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_pair(v[2 * q], v[2 * q + 1], fmt, wh[q], wl[q]);
      const size_t off = (size_t)(row0 + r) * dst.ld + c0;
      *reinterpret_cast<uint4*>(dst.hi + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      *reinterpret_cast<uint4*>(dst.lo + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
    // the skip layer's PE columns (d_pe of them, at an odd column offset): consecutive lanes on consecutive columns
    if (side.hi != nullptr) {
      for (int e = lane; e < n_rows * d_pe; e += 32) {
        const int r = e / d_pe, j = e % d_pe;
        uint32_t wh, wl;
        split_pair(sv[r * kEncStride + j] * side_scale, 0.f, fmt, wh, wl);
        const size_t off = (size_t)(row0 + r) * side.ld + side_col0 + j;
        reinterpret_cast<unsigned short*>(side.hi)[off] = (unsigned short)(wh & 0xFFFFu);
        reinterpret_cast<unsigned short*>(side.lo)[off] = (unsigned short)(wl & 0xFFFFu);
      }
    }
    __syncwarp();
  }
}

// d sdf/dx from the gradient w.r.t. the encoding: g = g0 + g_skip / sqrt(2)
__global__ void pe_backward_kernel(const float* __restrict__ x, const int* __restrict__ count, int rows_cap, int n_freqs,
                                   const float* __restrict__ g0, const float* __restrict__ g_skip, int ld,
                                   float* __restrict__ grad) {
  int limit = rows_cap;
  if (count) limit = min(limit, *count);
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < limit; row += gridDim.x * blockDim.x) {
    const float* a = g0 + (size_t)row * ld;
    const float* b = g_skip ? g_skip + (size_t)row * ld : nullptr;
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xc = x[(size_t)row * 3 + c];
      float acc = a[c] + (b ? b[c] * kInvSqrt2 : 0.f);
      for (int k = 0; k < n_freqs; ++k) {
        const float f = exp2f((float)k);
        const float gs = a[3 + 6 * k + c] + (b ? b[3 + 6 * k + c] * kInvSqrt2 : 0.f);
        const float gc = a[6 + 6 * k + c] + (b ? b[6 + 6 * k + c] * kInvSqrt2 : 0.f);
        acc += f * (cosf(xc * f) * gs - sinf(xc * f) * gc);
      }
      out[c] = acc;
    }
    grad[(size_t)row * 3 + 0] = out[0];
    grad[(size_t)row * 3 + 1] = out[1];
    grad[(size_t)row * 3 + 2] = out[2];
  }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// fp16 split by default: the depth / shading parity with the reference rests on SDF values that agree with an fp32 evaluation
// to a few 1e-7; NEFII_SDF_FORMAT=bf16 selects the wide-range format at load.
// Where the positional encoding is computed: 0 = one encode_kernel launch per evaluation (default), 1 = inside layer 0's GEMM
// ("PE prologue", mlp_gemm_kernel.cuh).  The prologue removes the launch and the in0 round trip but is SLOWER on B200: the 2 304
// sincosf of a 128-row tile fall to the kernel's two idle warps, which are latency-bound on them (measured: 2^20 points 10.99 ms
// against 9.39 ms, a latency-bound 4096-point evaluation 0.161 against 0.151 ms).  NEFII_SDF_PE_PROLOGUE=1.
int env_pe_prologue() {
  const char* e = getenv("NEFII_SDF_PE_PROLOGUE");
  return (e && e[0] == '1') ? 1 : 0;
}
int g_pe_prologue = env_pe_prologue();

int default_format() {
  const char* e = getenv("NEFII_SDF_FORMAT");
  return (e && (!strcmp(e, "bf16") || !strcmp(e, "0"))) ? PLANES_BF16 : PLANES_FP16;
}

}  // namespace

int sdf_set_pe_prologue(int on) { g_pe_prologue = on ? 1 : 0; return NEFII_OK; }

struct SdfNet::Impl {
  SdfConfig cfg;
  int d_pe = 0;
  int n_lin = 0;                      // hidden layers + output layer
  std::vector<int> in_dim, out_dim;   // logical sizes per linear layer
  std::vector<int> k_pad, n_pad;      // forward padding (K multiple of 64, N multiple of 256)
  // device storage
  std::vector<Planes> w_fwd;          // [n_pad, k_pad]
  std::vector<Planes> w_bwd;          // W^T: [round_up(in,256), round_up(out,64)]
  std::vector<float*> bias;
  float* w_last = nullptr;            // [d_out, width] fp32
  float* b_last = nullptr;
  Planes w_feat;                      // d_feat > 0: rows 1.. of the last Linear as planes [round_up(d_feat,256), width]
  float* b_feat = nullptr;
  void* blob = nullptr;
  bool has_weights = false;
  int fmt = default_format();
};

SdfNet::SdfNet() : impl_(new Impl) {}
SdfNet::~SdfNet() {
  if (impl_->blob) cudaFree(impl_->blob);
  delete impl_;
}

const SdfConfig& SdfNet::config() const { return impl_->cfg; }
int SdfNet::format() const { return impl_->fmt; }
int SdfNet::set_format(int fmt) {
  NEFII_CHECK_ARG(fmt == PLANES_BF16 || fmt == PLANES_FP16, "sdf: unknown plane format %d", fmt);
  if (fmt != impl_->fmt) impl_->has_weights = false;
  impl_->fmt = fmt;
  return NEFII_OK;
}

int SdfNet::init(const SdfConfig& cfg) {
  NEFII_CHECK_ARG(cfg.d_in == 3, "sdf: d_in must be 3");
  NEFII_CHECK_ARG(cfg.n_freqs >= 0 && cfg.n_freqs <= 10, "sdf: n_freqs out of range");
  NEFII_CHECK_ARG(cfg.width >= 64 && cfg.width % 64 == 0 && cfg.width <= 1024, "sdf: width must be a multiple of 64 in [64,1024]");
  NEFII_CHECK_ARG(cfg.n_hidden >= 2 && cfg.n_hidden <= 16, "sdf: n_hidden out of range");
  NEFII_CHECK_ARG(cfg.skip_layer < cfg.n_hidden, "sdf: skip_layer out of range");
  NEFII_CHECK_ARG(cfg.d_out == 1, "sdf: d_out must be 1");
  NEFII_CHECK_ARG(cfg.d_feat >= 0 && cfg.d_feat <= 1024, "sdf: d_feat out of range");
  Impl& s = *impl_;
  s.cfg = cfg;
  s.d_pe = 3 + 6 * cfg.n_freqs;
  NEFII_CHECK_ARG(s.d_pe <= 64, "sdf: positional encoding wider than 64 is not supported");
  s.n_lin = cfg.n_hidden + 1;
  s.in_dim.assign(s.n_lin, cfg.width);
  s.out_dim.assign(s.n_lin, cfg.width);
  s.in_dim[0] = s.d_pe;
  s.out_dim[s.n_lin - 1] = cfg.d_out;
  if (cfg.skip_layer > 0) s.out_dim[cfg.skip_layer - 1] = cfg.width - s.d_pe;
  s.k_pad.resize(s.n_lin);
  s.n_pad.resize(s.n_lin);
  size_t bytes = 0;
  auto take = [&](size_t n) { size_t o = bytes; bytes += (n + 255) / 256 * 256; return o; };
  std::vector<size_t> off_fh(s.n_lin), off_fl(s.n_lin), off_bh(s.n_lin), off_bl(s.n_lin), off_bias(s.n_lin);
  for (int l = 0; l < s.n_lin - 1; ++l) {
    s.k_pad[l] = round_up(s.in_dim[l], 64);
    s.n_pad[l] = round_up(s.out_dim[l], 256);
    const size_t nf = (size_t)s.n_pad[l] * s.k_pad[l] * 2;
    off_fh[l] = take(nf); off_fl[l] = take(nf);
    const size_t nb = (size_t)round_up(s.in_dim[l], 256) * round_up(s.out_dim[l], 64) * 2;
    off_bh[l] = take(nb); off_bl[l] = take(nb);
    off_bias[l] = take((size_t)s.out_dim[l] * 4);
  }
  const size_t off_wl = take((size_t)cfg.d_out * cfg.width * 4);
  const size_t off_bl_last = take((size_t)cfg.d_out * 4);
  const size_t n_wf = (size_t)round_up(cfg.d_feat > 0 ? cfg.d_feat : 1, 256) * cfg.width * 2;
  const size_t off_wfh = take(cfg.d_feat > 0 ? n_wf : 0), off_wfl = take(cfg.d_feat > 0 ? n_wf : 0);
  const size_t off_bf = take((size_t)(cfg.d_feat > 0 ? cfg.d_feat : 0) * 4);
  if (s.blob) cudaFree(s.blob);
  NEFII_CUDA(cudaMalloc(&s.blob, bytes));
  NEFII_CUDA(cudaMemset(s.blob, 0, bytes));
  char* base = (char*)s.blob;
  s.w_fwd.resize(s.n_lin - 1);
  s.w_bwd.resize(s.n_lin - 1);
  s.bias.resize(s.n_lin - 1);
  for (int l = 0; l < s.n_lin - 1; ++l) {
    s.w_fwd[l].hi = (__nv_bfloat16*)(base + off_fh[l]);
    s.w_fwd[l].lo = (__nv_bfloat16*)(base + off_fl[l]);
    s.w_fwd[l].ld = s.k_pad[l];
    s.w_bwd[l].hi = (__nv_bfloat16*)(base + off_bh[l]);
    s.w_bwd[l].lo = (__nv_bfloat16*)(base + off_bl[l]);
    s.w_bwd[l].ld = round_up(s.out_dim[l], 64);
    s.bias[l] = (float*)(base + off_bias[l]);
  }
  s.w_last = (float*)(base + off_wl);
  s.b_last = (float*)(base + off_bl_last);
  s.w_feat.hi = (__nv_bfloat16*)(base + off_wfh); s.w_feat.lo = (__nv_bfloat16*)(base + off_wfl); s.w_feat.ld = cfg.width;
  s.b_feat = (float*)(base + off_bf);
  s.has_weights = false;
  return NEFII_OK;
}

int SdfNet::set_weights(cudaStream_t stream, const float* const* weights, const float* const* biases) {
  Impl& s = *impl_;
  NEFII_CHECK_ARG(s.blob != nullptr, "sdf: set_weights before init");
  for (int l = 0; l < s.n_lin; ++l) NEFII_CHECK_ARG(weights[l] && biases[l], "sdf: null weight pointer for layer %d", l);
  int rc;
  for (int l = 0; l < s.n_lin - 1; ++l) {
    if ((rc = split_to_planes(stream, weights[l], s.out_dim[l], s.in_dim[l], s.in_dim[l], 0, 1.f, s.w_fwd[l].hi,
                              s.w_fwd[l].lo, s.n_pad[l], s.k_pad[l], s.fmt)))
      return rc;
    if ((rc = split_to_planes(stream, weights[l], s.out_dim[l], s.in_dim[l], s.in_dim[l], 1, 1.f, s.w_bwd[l].hi,
                              s.w_bwd[l].lo, round_up(s.in_dim[l], 256), round_up(s.out_dim[l], 64), s.fmt)))
      return rc;
    NEFII_CUDA(cudaMemcpyAsync(s.bias[l], biases[l], (size_t)s.out_dim[l] * 4, cudaMemcpyDeviceToDevice, stream));
  }
  // the last Linear: row 0 (the SDF) is applied in fp32 inside the last hidden layer's epilogue; with d_feat > 0 its
  // remaining rows (the feature vector of the use_last_as_f = False layout) become one more plain layer
  NEFII_CUDA(cudaMemcpyAsync(s.w_last, weights[s.n_lin - 1], (size_t)s.cfg.d_out * s.cfg.width * 4, cudaMemcpyDeviceToDevice, stream));
  NEFII_CUDA(cudaMemcpyAsync(s.b_last, biases[s.n_lin - 1], (size_t)s.cfg.d_out * 4, cudaMemcpyDeviceToDevice, stream));
  if (s.cfg.d_feat > 0) {
    if ((rc = split_to_planes(stream, weights[s.n_lin - 1] + (size_t)s.cfg.d_out * s.cfg.width, s.cfg.d_feat, s.cfg.width, s.cfg.width, 0,
                              1.f, s.w_feat.hi, s.w_feat.lo, round_up(s.cfg.d_feat, 256), s.cfg.width, s.fmt)))
      return rc;
    NEFII_CUDA(cudaMemcpyAsync(s.b_feat, biases[s.n_lin - 1] + s.cfg.d_out, (size_t)s.cfg.d_feat * 4, cudaMemcpyDeviceToDevice, stream));
  }
  s.has_weights = true;
  return NEFII_OK;
}

// Workspace layout (all plane buffers hold hi then lo, [rows_cap, width] 16-bit each):
//   in0 (ld 64: the encoding, layer 0's input) | inference: two ping-pong activation buffers + the skip layer's input buffer (its
//   PE half is written together with in0, its h half by layer skip - 1); with the gradient: one buffer per hidden layer + seed +
//   spare | fp32 g0 [rows, 64] | g_skip [rows, 64]
size_t SdfNet::workspace_bytes(int rows_cap, bool with_grad) const {
  const Impl& s = *impl_;
  const size_t rows = (size_t)round_up(rows_cap > 0 ? rows_cap : 1, 128);
  const size_t plane_w = rows * s.cfg.width * 2 * 2;
  const size_t plane_0 = rows * 64 * 2 * 2;
  const int n_act = with_grad ? (s.cfg.n_hidden - 1) + 2 : 3;
  size_t b = plane_0 + (size_t)n_act * plane_w;
  if (with_grad) b += 2 * rows * 64 * 4;
  return b + 1024;
}

int SdfNet::eval(cudaStream_t stream, int rows_cap, const int* count, const float* x, void* workspace, size_t ws_bytes,
                 float* sdf, float* feat, float* grad, int k_flush) const {
  const Impl& s = *impl_;
  NEFII_CHECK_ARG(s.has_weights, "sdf: eval before set_weights");
  if (rows_cap <= 0) return NEFII_OK;
  NEFII_CHECK_ARG(x && sdf && workspace, "sdf: null pointer");
  const bool with_grad = grad != nullptr;
  NEFII_CHECK_ARG(ws_bytes >= workspace_bytes(rows_cap, with_grad), "sdf: workspace too small (%zu < %zu)", ws_bytes,
                  workspace_bytes(rows_cap, with_grad));
  const int W = s.cfg.width, H = s.cfg.n_hidden, skip = s.cfg.skip_layer;
  const size_t rows = (size_t)round_up(rows_cap, 128);
  char* p = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  auto planes = [&](int ld) {
    Planes pl;
    pl.hi = (__nv_bfloat16*)p; p += rows * ld * 2;
    pl.lo = (__nv_bfloat16*)p; p += rows * ld * 2;
    pl.ld = ld;
    return pl;
  };
  Planes in0 = planes(64);
  const int n_act = with_grad ? (H - 1) + 2 : 3;
  std::vector<Planes> act(n_act);
  for (int i = 0; i < n_act; ++i) act[i] = planes(W);
  float* g0 = nullptr;
  float* g_skip = nullptr;
  if (with_grad) {
    g0 = (float*)p; p += rows * 64 * 4;
    g_skip = (float*)p; p += rows * 64 * 4;
  }
  // input planes of hidden layer l (l >= 1); the skip layer's input has its own buffer in the ping-pong scheme
  auto in_of = [&](int l) -> Planes& { return with_grad ? act[l - 1] : (l == skip ? act[2] : act[(l - 1) & 1]); };
  int rc;
  const bool pe_prologue = g_pe_prologue != 0;
  if (!pe_prologue) {
    Planes side;
    if (skip > 0) side = in_of(skip);
    encode_kernel<<<kNumSMs * 6, kEncWarps * 32, 0, stream>>>(x, count, rows_cap, s.cfg.n_freqs, in0, side, W - s.d_pe, kInvSqrt2, s.fmt);
    NEFII_LAUNCH_CHECK();
  }

  // ---------------------------------------------------------------- forward
  Planes seed;  // gradient seed G_{H-1}
  if (with_grad) seed = act[H - 1];
  // d_feat > 0: planes of the last hidden activation, input of the feature rows of the last Linear.  Buffer: the one
  // the last hidden layer does not read (ping-pong) / the spare buffer of the gradient chain (with_grad).
  Planes feat_in = with_grad ? act[H] : act[(H - 1) & 1];
  for (int l = 0; l < H; ++l) {
    GemmProblem g{};
    g.k_flush = k_flush;
    g.epi.fmt = s.fmt;
    g.rows_cap = rows_cap;
    if (l == 0 && pe_prologue) {
      // PE prologue: the encoding is computed inside the kernel; its 1/sqrt(2)-scaled copy goes next to the h columns of the
      // skip layer's input (layer skip - 1 writes those later and leaves the PE columns alone)
      g.pe.x = x; g.pe.n_freqs = s.cfg.n_freqs;
      if (skip > 0) { g.pe.side = in_of(skip); g.pe.side_col0 = W - s.d_pe; g.pe.side_scale = kInvSqrt2; }
    } else {
      const Planes& a = (l == 0) ? in0 : in_of(l);
      g.a_hi = a.hi; g.a_lo = a.lo; g.a_ld = a.ld;
    }
    g.b_hi = s.w_fwd[l].hi; g.b_lo = s.w_fwd[l].lo; g.b_ld = s.w_fwd[l].ld; g.n_pad = s.n_pad[l];
    g.k_pad = s.k_pad[l];
    g.count = count;
    g.epi.mode = 0;
    g.epi.act = ACT_SOFTPLUS100;
    g.epi.n_valid = s.out_dim[l];
    g.epi.bias = s.bias[l];
    if (l < H - 1) {
      g.epi.dst = in_of(l + 1);
      g.epi.dst_ncols = s.out_dim[l];
      g.epi.out_scale = (l + 1 == skip) ? kInvSqrt2 : 1.f;
      g.epi.dst_pad_ok = (l + 1 == skip) ? 1 : 0;   // the PE columns next to h are already there
    } else {
      g.epi.w_last = s.w_last; g.epi.b_last = s.b_last; g.epi.n_last = s.cfg.d_out; g.epi.w_last_ld = W;
      g.epi.dst_last = sdf;
      if (feat && s.cfg.d_feat == 0) { g.epi.dst_f32 = feat; g.epi.f32_ld = W; g.epi.f32_begin = 0; g.epi.f32_end = W; }
      if (feat && s.cfg.d_feat > 0) { g.epi.dst = feat_in; g.epi.dst_ncols = W; }   // planes of the last hidden activation
      if (with_grad) g.epi.seed = seed;
    }
    if ((rc = gemm_split_bf16(stream, g))) return rc;
    if (l == H - 1 && feat && s.cfg.d_feat > 0) {
      GemmProblem f{};     // feature = W_last[1:] h + b_last[1:]  (no activation), fp32 out
      f.k_flush = k_flush;
      f.epi.fmt = s.fmt;
      f.a_hi = feat_in.hi; f.a_lo = feat_in.lo; f.a_ld = feat_in.ld; f.rows_cap = rows_cap;
      f.b_hi = s.w_feat.hi; f.b_lo = s.w_feat.lo; f.b_ld = s.w_feat.ld; f.n_pad = round_up(s.cfg.d_feat, 256);
      f.k_pad = W;
      f.count = count;
      f.epi.mode = 0; f.epi.act = ACT_NONE; f.epi.n_valid = s.cfg.d_feat; f.epi.bias = s.b_feat;
      f.epi.dst_f32 = feat; f.epi.f32_ld = s.cfg.d_feat; f.epi.f32_begin = 0; f.epi.f32_end = s.cfg.d_feat;
      if ((rc = gemm_split_bf16(stream, f))) return rc;
    }
  }
  if (!with_grad) return NEFII_OK;

  // ---------------------------------------------------------------- reverse chain for d sdf / d x
  // G_l = d sdf / d z_l as planes; G_{H-1} = seed.  For l = H-1 .. 1:  G_{l-1} = (G_l W_l) * act'(z_{l-1})
  Planes cur = seed;
  Planes spare = act[H];
  for (int l = H - 1; l >= 1; --l) {
    GemmProblem g{};
    g.k_flush = k_flush;
    g.epi.fmt = s.fmt;
    g.a_hi = cur.hi; g.a_lo = cur.lo; g.a_ld = cur.ld; g.rows_cap = rows_cap;
    g.b_hi = s.w_bwd[l].hi; g.b_lo = s.w_bwd[l].lo; g.b_ld = s.w_bwd[l].ld; g.n_pad = round_up(s.in_dim[l], 256);
    g.k_pad = round_up(s.out_dim[l], 64);
    g.count = count;
    g.epi.mode = 1;
    g.epi.act = ACT_SOFTPLUS100;
    g.epi.n_valid = s.in_dim[l];
    const Planes& saved = in_of(l);   // forward input of layer l == activation output of layer l-1 (scaled at the skip)
    g.epi.sav_hi = saved.hi; g.epi.sav_lo = saved.lo; g.epi.sav_ld = saved.ld;
    g.epi.dst = spare;
    if (l == skip) {
      g.epi.sav_ncols = W - s.d_pe; g.epi.sav_scale = kSqrt2;
      g.epi.out_scale = kInvSqrt2;
      g.epi.dst_ncols = W - s.d_pe;
      g.epi.dst_zero_to = round_up(W - s.d_pe, 64);   // K padding of the next GEMM must be finite zeros
      g.epi.dst_f32 = g_skip; g.epi.f32_ld = 64; g.epi.f32_begin = W - s.d_pe; g.epi.f32_end = W;
    } else {
      g.epi.sav_ncols = s.in_dim[l]; g.epi.sav_scale = 1.f;
      g.epi.dst_ncols = s.in_dim[l];
    }
    if ((rc = gemm_split_bf16(stream, g))) return rc;
    Planes t = cur; cur = spare; spare = t;
  }
  {
    GemmProblem g{};   // layer 0: gradient w.r.t. the encoding, fp32 out
    g.k_flush = k_flush;
    g.epi.fmt = s.fmt;
    g.a_hi = cur.hi; g.a_lo = cur.lo; g.a_ld = cur.ld; g.rows_cap = rows_cap;
    g.b_hi = s.w_bwd[0].hi; g.b_lo = s.w_bwd[0].lo; g.b_ld = s.w_bwd[0].ld; g.n_pad = round_up(s.in_dim[0], 256);
    g.k_pad = round_up(s.out_dim[0], 64);
    g.count = count;
    g.epi.mode = 1;
    g.epi.act = ACT_NONE;
    g.epi.n_valid = s.d_pe;
    g.epi.dst_f32 = g0; g.epi.f32_ld = 64; g.epi.f32_begin = 0; g.epi.f32_end = s.d_pe;
    if ((rc = gemm_split_bf16(stream, g))) return rc;
  }
  pe_backward_kernel<<<ceil_div(rows_cap, 128), 128, 0, stream>>>(x, count, rows_cap, s.cfg.n_freqs, g0,
                                                                  skip > 0 ? g_skip : nullptr, 64, grad);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
