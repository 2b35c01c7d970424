// Split-bf16 tcgen05 layer GEMM for the MLPs on the hot path (SDF net, radiance net, material net).
//
//   D[m, n] = sum_k  Ahi[m,k]*Bhi[n,k] + Ahi[m,k]*Blo[n,k] + Alo[m,k]*Bhi[n,k]        (fp32 accumulate in TMEM)
//
// The reference runs these layers as FP32 SGEMM (nn.Linear).  tcgen05 has no fp32 MMA, so every
// fp32 value x travels as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)); three kind::f16 MMAs
// per K step recover ~16 mantissa bits per operand (measured against the fp32 oracle in
// tests/test_gemm.py and reported in DESIGN.md).
//
// Persistent CTAs (one per SM) loop over 128-row tiles of points and, per tile, over the output columns in chunks of 256:
//   warps 0..7  : epilogue (tcgen05.ld of every partial sum into fp32 registers -> bias / activation / derivative ->
//                 bf16 planes by bulk tensor store / fp32 / fused output layer), overlapped with the MMAs of the next chunk
//   warp 8      : TMA producer (A/B hi+lo tiles, 128B swizzle, mbarrier ring)
//   warp 9      : TMEM allocator + single-thread tcgen05.mma issuer (2 x 256-column partial-sum buffers)
// Two instantiations: single CTA (cta_group::1, M = 128) and CTA pairs (cta_group::2, M = 256 over two SMs).
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "mlp_gemm_kernel.cuh"

namespace nefii {

namespace {

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Encoded tensor maps are pure functions of their arguments: a per-thread cache takes cuTensorMapEncodeTiled (six calls per
// launch otherwise) off the host path of the launch-bound phases.
struct MapKey {
  const void* base; int rows, ld, ext, box_rows, bk;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && ld == o.ld && ext == o.ext && box_rows == o.box_rows && bk == o.bk;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    h ^= ((size_t)k.rows << 32) ^ ((size_t)k.ld << 8) ^ ((size_t)k.ext << 20) ^ ((size_t)k.box_rows << 44) ^ (size_t)k.bk;
    return h ^ (h >> 29);
  }
};
typedef std::unordered_map<MapKey, CUtensorMap, MapKeyHash> MapCache;
MapCache& map_cache() {
  static thread_local MapCache cache;
  if (cache.size() > 8192) cache.clear();
  return cache;
}

int make_map_uncached(CUtensorMap* map, const void* base, int rows, int ld, int k_extent, int box_rows, int bk) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(NEFII_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(NEFII_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d ld=%d", (int)r, rows, ld);
  return NEFII_OK;
}

int make_map(CUtensorMap* map, const void* base, int rows, int ld, int k_extent, int box_rows, int bk) {
  MapCache& c = map_cache();
  const MapKey key{base, rows, ld, k_extent, box_rows, bk};
  auto it = c.find(key);
  if (it != c.end()) { *map = it->second; return NEFII_OK; }
  const int rc = make_map_uncached(map, base, rows, ld, k_extent, box_rows, bk);
  if (!rc) c.emplace(key, *map);
  return rc;
}

int make_store_map_uncached(CUtensorMap* map, const void* base, int rows, int ld, int cols);
// [rows, cols] bf16 plane view (row stride ld) for the epilogue's 32 x 32 bulk stores out of 64 B-swizzled staging tiles
int make_store_map(CUtensorMap* map, const void* base, int rows, int ld, int cols) {
  MapCache& c = map_cache();
  const MapKey key{base, rows, ld, cols, 32, -1};
  auto it = c.find(key);
  if (it != c.end()) { *map = it->second; return NEFII_OK; }
  const int rc = make_store_map_uncached(map, base, rows, ld, cols);
  if (!rc) c.emplace(key, *map);
  return rc;
}
int make_store_map_uncached(CUtensorMap* map, const void* base, int rows, int ld, int cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(NEFII_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(NEFII_ERR_CUDA, "cuTensorMapEncodeTiled (store) failed (%d) rows=%d ld=%d cols=%d", (int)r, rows, ld, cols);
  return NEFII_OK;
}

__global__ void split_to_planes_kernel(const float* __restrict__ src, int rows, int cols, int ld_src, int transpose,
                                       float scale, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                       int rows_pad, int cols_pad, int fmt) {
  const long long total = (long long)rows_pad * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
    float x = 0.f;
    if (!transpose) {
      if (r < rows && c < cols) x = src[(size_t)r * ld_src + c];
    } else {
      if (r < cols && c < rows) x = src[(size_t)c * ld_src + r];
    }
    x *= scale;
    __nv_bfloat16 h, l;
    split2(x, fmt, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace

namespace {
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return (v >= lo && v <= hi) ? v : dflt;
}
// K blocks (of 64) accumulated inside TMEM before the sum moves to registers; see gemm_set_k_flush (NEFII_GEMM_KFLUSH
// overrides the default at load, for A/B runs of whole programs).  Default 8 = a whole 512-wide layer per partial, the fastest
// schedule (one hand-off per 256-column chunk).  Its truncation shortfall (22 ulp of 2^-24, systematic) is removed by the
// first-order compensation below: measured SDF-MLP error vs f64 (tools/diag_gpu.py trunccomp, mean |err|) 1.9e-6 at 8 blocks
// with compensation against 1.22e-5 without, 2.0e-6 at ONE block without and 1.6e-6 with -- what is left (1.9e-6 rms about
// the mean at every setting) is the 16-bit operand representation of the bf16 hi/lo split, not the accumulation.
constexpr int kDefaultKFlush = 8;
int g_k_flush = env_int("NEFII_GEMM_KFLUSH", kDefaultKFlush, 1, 64);
int g_store_tma = getenv("NEFII_GEMM_NO_TMA_STORE") ? 0 : 1;   // development switch for A/B timing
int g_k_flush_head = env_int("NEFII_GEMM_KFLUSH", kDefaultKFlush, 1, 64);   // ... for the first two partials of a column chunk (gemm_set_k_flush_head)
int g_pdl = env_int("NEFII_GEMM_PDL", 1, 0, 1);   // programmatic dependent launch of the layer GEMMs (gemm_set_pdl)
int g_debug = 0;          // development only: bit mask that disables pipeline pieces for timing experiments
// 1: single-CTA kernel, 2: cta_group::2 pairs (nefii_gemm_set_cluster; NEFII_GEMM_CLUSTER overrides the default at load).
// Pairs are the default: the single-CTA kernel is bound by shared-memory bandwidth (TMA writes + tensor-core operand reads +
// staging = 2.2 MB per 128 x 256 chunk against 128 B/clk), pairs read a third less.
int env_cluster_pref() {
  const char* e = getenv("NEFII_GEMM_CLUSTER");
  return (e && e[0] == '1') ? 1 : 2;
}
int g_cluster_pref = env_cluster_pref();
struct ProfRec {
  cudaEvent_t a, b;
  double flops_per_row;
  int rows_cap;
  int* count_host;   // pinned; -1 when the launch had no device row count
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::atomic<long long> g_epoch{0};
// Relative shortfall of a TMEM partial sum of L K blocks (64 columns each).  tcgen05 aligns every addend to the accumulator and
// truncates it towards zero, so a partial sum comes out short by a factor that depends on its length only.  Measured on B200
// with `tools/diag_gpu.py trunccomp` (slope of the accumulation error against the exact sum of the three bf16 products, 16384 x
// 512 x 512, Gaussian and all-positive activations agree to 3 %): 1.02 / 3.9 / 9.9 / 22.1 ulp of 2^-24 at 1 / 2 / 4 / 8 blocks.
// fp16-split planes (22-bit products instead of 16-bit ones): 1.50 / 4.82 / 11.4 / 24.4 ulp (DIAG_FMT=1, same tool).
// Other lengths: log-log interpolation.  gemm_set_trunc_comp overrides an entry; NEFII_GEMM_TRUNC_COMP=0 switches it off.
float default_rho(int L, int fmt) {
  static const float kL[4] = {1.f, 2.f, 4.f, 8.f};
  static const float kRhoFmt[2][4] = {{6.10e-8f, 2.34e-7f, 5.92e-7f, 1.32e-6f}, {8.93e-8f, 2.87e-7f, 6.82e-7f, 1.456e-6f}};
  const float* kRho = kRhoFmt[fmt];
  if (L <= 1) return kRho[0];
  int i = 0;
  while (i < 2 && (float)L > kL[i + 1]) ++i;
  const float t = (logf((float)L) - logf(kL[i])) / (logf(kL[i + 1]) - logf(kL[i]));
  return expf(logf(kRho[i]) + t * (logf(kRho[i + 1]) - logf(kRho[i])));
}
struct RhoTable {
  float v[65];
  explicit RhoTable(int fmt) {
    const char* e = getenv("NEFII_GEMM_TRUNC_COMP");
    const bool on = !(e && e[0] == '0');
    v[0] = 0.f;
    for (int L = 1; L <= 64; ++L) v[L] = on ? default_rho(L, fmt) : 0.f;
  }
};
RhoTable g_trunc[2] = {RhoTable(PLANES_BF16), RhoTable(PLANES_FP16)};   // per PlaneFormat
std::mutex g_dev_mu;
bool g_dev_ready[64] = {};
int g_dev_sms[64] = {};
}  // namespace

long long gemm_config_epoch() { return g_epoch.load(std::memory_order_relaxed); }
bool gemm_profile_active() { return g_prof_on; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: every device this process launches on
// gets its own opt-in (and its own SM count for the persistent grids).
int gemm_prepare_device() {
  int dev = 0;
  NEFII_CUDA(cudaGetDevice(&dev));
  NEFII_CHECK_ARG(dev >= 0 && dev < 64, "gemm: device index out of range");
  std::lock_guard<std::mutex> lock(g_dev_mu);
  if (g_dev_ready[dev]) return NEFII_OK;
  for (int key = 0; key < kGemmKernelKeys; ++key) {
    GemmKernelFn f1 = select_gemm_kernel<1>(key);
    GemmKernelFn f2 = reinterpret_cast<GemmKernelFn>(gemm_pair_kernel(key));
    if (f1) NEFII_CUDA(cudaFuncSetAttribute(f1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    if (f2) NEFII_CUDA(cudaFuncSetAttribute(f2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  }
  int sms = 0;
  NEFII_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g_dev_sms[dev] = sms > 0 ? sms : kNumSMs;
  g_dev_ready[dev] = true;
  return NEFII_OK;
}
namespace {
int device_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_dev_ready[dev]) return kNumSMs;
  return g_dev_sms[dev];
}
}  // namespace

int gemm_device_sms() { return device_sms(); }

// Upper bound on the persistent grid of the launches that follow (0 = every SM).  A caller that runs a bulk evaluation next to a
// latency-bound chain on another stream (IDRNetwork.prefetch_trace) leaves a few SMs to that chain: persistent CTAs hold their SM
// for the whole kernel, so without the bound a 15 us kernel of the other stream waits for a ~1 ms bulk kernel to END.  Not part of
// the configuration epoch (it is toggled around single calls); the trace-graph cache keys on it separately.
int g_grid_cap = 0;
int gemm_set_grid_cap(int sms) {
  NEFII_CHECK_ARG(sms >= 0 && sms <= 4096, "gemm_set_grid_cap: SM count out of range");
  g_grid_cap = sms;
  return NEFII_OK;
}
int gemm_grid_cap() { return g_grid_cap; }

int gemm_set_cluster(int cl) {
  NEFII_CHECK_ARG(cl == 1 || cl == 2, "gemm_set_cluster: 1 (single CTA) or 2 (cta_group::2 pair)");
  g_cluster_pref = cl;
  ++g_epoch;
  return NEFII_OK;
}

int gemm_set_debug(int mask) { g_debug = mask; ++g_epoch; return NEFII_OK; }
int gemm_set_pdl(int on) { g_pdl = on ? 1 : 0; ++g_epoch; return NEFII_OK; }
int gemm_set_k_flush(int k) {
  NEFII_CHECK_ARG(k >= 1 && k <= 64, "gemm_set_k_flush: out of range");
  g_k_flush = k;
  g_k_flush_head = k;
  ++g_epoch;
  return NEFII_OK;
}
int gemm_set_k_flush_head(int k) {
  NEFII_CHECK_ARG(k >= 1 && k <= 64, "gemm_set_k_flush_head: out of range");
  g_k_flush_head = k;
  ++g_epoch;
  return NEFII_OK;
}

int gemm_set_trunc_comp(int k_blocks, float rho, int fmt) {
  NEFII_CHECK_ARG(k_blocks >= 1 && k_blocks <= 64, "gemm_set_trunc_comp: partial length out of range");
  NEFII_CHECK_ARG(rho > -1e-3f && rho < 1e-3f, "gemm_set_trunc_comp: |rho| must be below 1e-3");
  NEFII_CHECK_ARG(fmt == PLANES_BF16 || fmt == PLANES_FP16, "gemm_set_trunc_comp: unknown plane format %d", fmt);
  g_trunc[fmt].v[k_blocks] = rho;
  ++g_epoch;
  return NEFII_OK;
}

int gemm_profile_enable(int on) {
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    cudaFreeHost(r.count_host);
  }
  g_prof.clear();
  g_prof_on = on != 0;
  return NEFII_OK;
}

int gemm_profile_fetch(double* out3) {
  NEFII_CHECK_ARG(out3 != nullptr, "gemm_profile_fetch: null output");
  double ms = 0, flops = 0;
  for (auto& r : g_prof) {
    NEFII_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    NEFII_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    int rows = r.rows_cap;
    if (*r.count_host >= 0 && *r.count_host < rows) rows = *r.count_host;
    if (rows > 0) {   // launches whose row count was 0 exit immediately: neither time nor work is attributed
      ms += t;
      flops += r.flops_per_row * rows;
    }
  }
  out3[0] = ms; out3[1] = flops; out3[2] = (double)g_prof.size();
  return NEFII_OK;
}

int gemm_split_bf16(cudaStream_t stream, const GemmProblem& p) {
  const bool pe_mode = p.pe.x != nullptr;
  NEFII_CHECK_ARG((pe_mode || (p.a_hi && p.a_lo)) && p.b_hi && p.b_lo, "gemm_split_bf16: null operand");
  NEFII_CHECK_ARG(p.k_pad > 0 && p.k_pad % BK == 0 && (pe_mode || p.k_pad <= p.a_ld) && p.k_pad <= p.b_ld,
                  "gemm_split_bf16: k_pad=%d must be a multiple of %d and <= ld (a_ld=%d b_ld=%d)", p.k_pad, BK, p.a_ld, p.b_ld);
  NEFII_CHECK_ARG(!pe_mode || (p.k_pad == BK && p.pe.n_freqs >= 0 && 3 + 6 * p.pe.n_freqs <= BK && p.k_splits <= 1),
                  "gemm_split_bf16: the PE prologue produces one %d-wide K block (n_freqs=%d)", BK, p.pe.n_freqs);
  NEFII_CHECK_ARG(!pe_mode || p.pe.side.hi == nullptr || p.pe.side.lo != nullptr, "gemm_split_bf16: PE side planes need hi and lo");
  NEFII_CHECK_ARG(p.n_pad > 0 && p.n_pad % BN == 0, "gemm_split_bf16: n_pad=%d must be a multiple of %d", p.n_pad, BN);
  NEFII_CHECK_ARG((pe_mode || p.a_ld % 8 == 0) && p.b_ld % 8 == 0, "gemm_split_bf16: leading dimensions must be multiples of 8");
  NEFII_CHECK_ARG(p.epi.n_valid > 0 && p.epi.n_valid <= p.n_pad, "gemm_split_bf16: n_valid out of range");
  NEFII_CHECK_ARG(p.epi.n_last <= kMaxLast, "gemm_split_bf16: fused output layer supports at most %d outputs", kMaxLast);
  NEFII_CHECK_ARG(p.k_flush >= 0 && p.k_flush <= 64, "gemm_split_bf16: k_flush out of range");
  NEFII_CHECK_ARG(p.epi.fmt == PLANES_BF16 || p.epi.fmt == PLANES_FP16, "gemm_split_bf16: unknown plane format %d", p.epi.fmt);
  if (p.rows_cap <= 0) return NEFII_OK;
  int rc;
  if ((rc = gemm_prepare_device())) return rc;
  const int n_sms = device_sms();
  const int m_tiles = ceil_div(p.rows_cap, BM);
  const int cl = (g_cluster_pref == 2 && m_tiles >= 2) ? 2 : 1;   // CTA pairs need two row tiles to work on
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  const int bk = cl == 2 ? Ring<2>::kBK : Ring<1>::kBK;
  if ((rc = make_map(&mb_hi, p.b_hi, p.n_pad, p.b_ld, p.k_pad, BN / cl, bk))) return rc;
  if ((rc = make_map(&mb_lo, p.b_lo, p.n_pad, p.b_ld, p.k_pad, BN / cl, bk))) return rc;
  if (pe_mode) {   // the kernel never touches the A maps
    ma_hi = mb_hi; ma_lo = mb_lo;
  } else {
    if ((rc = make_map(&ma_hi, p.a_hi, p.rows_cap, p.a_ld, p.k_pad, BM, bk))) return rc;
    if ((rc = make_map(&ma_lo, p.a_lo, p.rows_cap, p.a_ld, p.k_pad, BM, bk))) return rc;
  }
  const int n_chunks = ceil_div(p.epi.dst_zero_to > p.epi.n_valid ? p.epi.dst_zero_to : p.epi.n_valid, BN);
  // plain hidden layers (the kernel's fast_layer): the output planes as tensors for the epilogue's bulk stores
  CUtensorMap md_hi = ma_hi, md_lo = ma_lo;
  int store_tma = 0;
  const bool planes_out = p.epi.dst.hi != nullptr && p.epi.dst_f32 == nullptr && (p.epi.dst_col0 & 7) == 0 && (p.epi.dst.ld & 7) == 0 &&
                          p.epi.dst_ncols >= 32;
  const bool fwd_fast = p.epi.mode == 0 && p.epi.w_last == nullptr && p.epi.bias != nullptr && n_chunks * BN <= kBiasSmemFloats;
  const bool bwd_fast = p.epi.mode == 1 && p.epi.sav_hi != nullptr && (p.epi.sav_ld & 7) == 0;
  if (g_store_tma && planes_out && (fwd_fast || bwd_fast)) {
    if ((rc = make_store_map(&md_hi, p.epi.dst.hi + p.epi.dst_col0, p.rows_cap, p.epi.dst.ld, p.epi.dst_ncols))) return rc;
    if ((rc = make_store_map(&md_lo, p.epi.dst.lo + p.epi.dst_col0, p.rows_cap, p.epi.dst.ld, p.epi.dst_ncols))) return rc;
    store_tma = 1;
  }
  NEFII_CHECK_ARG(n_chunks * BN <= p.n_pad, "gemm_split_bf16: dst_zero_to beyond n_pad");
  NEFII_CHECK_ARG(p.epi.mode == 0 || p.epi.sav_hi == nullptr || p.epi.sav_ld >= n_chunks * BN || p.epi.sav_ncols % 32 == 0,
                  "gemm_split_bf16: saved-activation rows must cover whole 32-column groups");
  int grid = ceil_div(m_tiles, cl) * cl;
  // few row tiles: room for the kernel's device-side split of the column chunks over clusters (see the kernel)
  if (n_chunks > 1 && !(p.epi.mode == 0 && p.epi.w_last != nullptr) && p.k_splits <= 1) grid *= n_chunks;
  const int grid_max = (g_grid_cap > 0 && g_grid_cap < n_sms) ? (g_grid_cap < cl ? cl : g_grid_cap) : n_sms;
  if (grid > grid_max) grid = grid_max / cl * cl;   // persistent CTAs: one per SM, looping over row tiles
  GemmKernelFn fn = nullptr;
  const bool fuse = p.epi.mode == 0 && p.epi.w_last != nullptr;
  NEFII_CHECK_ARG(!fuse || (p.epi.dst_last != nullptr && p.epi.n_last >= 1), "gemm_split_bf16: fused output layer needs dst_last");
  NEFII_CHECK_ARG(p.epi.seed.hi == nullptr || fuse, "gemm_split_bf16: seed planes need the fused output layer");
  NEFII_CHECK_ARG(!pe_mode || (p.epi.mode == 0 && !fuse && (p.epi.act == ACT_SOFTPLUS100 || p.epi.act == ACT_NONE)),
                  "gemm_split_bf16: the PE prologue is instantiated for plain forward layers (softplus / no activation)");
  const int key = pe_mode ? (p.epi.act == ACT_SOFTPLUS100 ? 12 : 13) : (fuse ? 8 : 0) + p.epi.mode * 4 + p.epi.act;
  fn = cl == 2 ? reinterpret_cast<GemmKernelFn>(gemm_pair_kernel(key)) : select_gemm_kernel<1>(key);
  if (fn == nullptr) return set_error(NEFII_ERR_ARG, "gemm_split_bf16: bad mode/act (%d/%d)", p.epi.mode, p.epi.act);
  const int k_blocks = p.k_pad / BK;
  int splits = p.k_splits > 1 ? p.k_splits : 1;
  if (splits > k_blocks) splits = k_blocks;
  const int kb_per = ceil_div(k_blocks, splits);
  splits = ceil_div(k_blocks, kb_per);
  NEFII_CHECK_ARG(splits == 1 || (p.epi.dst_f32 != nullptr && p.epi.dst.hi == nullptr && !fuse && p.epi.mode == 0 && p.epi.act == ACT_NONE),
                  "gemm_split_bf16: split-K needs a plain fp32 output");
  if (p.k_splits_used) *p.k_splits_used = splits;
  ProfRec* rec = nullptr;
  if (g_prof_on) {
    ProfRec r;
    NEFII_CUDA(cudaEventCreate(&r.a));
    NEFII_CUDA(cudaEventCreate(&r.b));
    NEFII_CUDA(cudaMallocHost(&r.count_host, sizeof(int)));
    *r.count_host = -1;
    r.rows_cap = p.rows_cap;
    r.flops_per_row = 2.0 * (double)p.epi.n_valid * (double)p.k_pad;
    if (p.count) NEFII_CUDA(cudaMemcpyAsync(r.count_host, p.count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    g_prof.push_back(r);
    rec = &g_prof.back();
    NEFII_CUDA(cudaEventRecord(rec->a, stream));
  }
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (g_pdl) {
      // programmatic dependent launch: this kernel's prologue (barriers, TMEM, cluster sync) may run while the previous kernel
      // of the stream is still draining; the kernel waits (griddepcontrol.wait) before it reads anything that kernel wrote
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 2;
    }
    const int kf_tail = p.k_flush > 0 ? p.k_flush : g_k_flush;
    const int kf_head = p.k_flush > 0 ? p.k_flush : g_k_flush_head;
    // Partial schedule of a column chunk (PartSched): partials of kf K blocks, the last one takes what is left.  Split-K
    // launches (kb_per blocks per split; the last split may be shorter) and head != tail schedules are compensated with the
    // full-length factor only where the lengths are known to be uniform.
    float scale_full = 1.0f, scale_last = 1.0f;
    if (kf_head == kf_tail) {
      const int len_full = kf_tail < kb_per ? kf_tail : kb_per;
      const int n_parts = ceil_div(kb_per, len_full);
      const int len_last = kb_per - (n_parts - 1) * len_full;
      const RhoTable& rho = g_trunc[p.epi.fmt];
      scale_full = 1.0f + rho.v[len_full < 64 ? len_full : 64];
      scale_last = (splits == 1) ? 1.0f + rho.v[len_last < 64 ? len_last : 64] : scale_full;
    }
    NEFII_CUDA(cudaLaunchKernelEx(&cfg, fn, ma_hi, ma_lo, mb_hi, mb_lo, md_hi, md_lo, store_tma, p.count, p.rows_cap, k_blocks, n_chunks, kb_per,
                                  (long long)p.f32_split_stride, g_debug, kf_tail | (kf_head << 8), scale_full, scale_last, p.epi, p.pe));
  }
  if (rec) NEFII_CUDA(cudaEventRecord(rec->b, stream));
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int split_to_planes(cudaStream_t stream, const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                    __nv_bfloat16* hi, __nv_bfloat16* lo, int rows_pad, int cols_pad, int fmt) {
  NEFII_CHECK_ARG(src && hi && lo, "split_to_planes: null pointer");
  NEFII_CHECK_ARG(fmt == PLANES_BF16 || fmt == PLANES_FP16, "split_to_planes: unknown plane format %d", fmt);
  const int out_rows = transpose ? cols : rows, out_cols = transpose ? rows : cols;
  NEFII_CHECK_ARG(rows_pad >= out_rows && cols_pad >= out_cols, "split_to_planes: padded shape too small");
  const long long total = (long long)rows_pad * cols_pad;
  if (total == 0) return NEFII_OK;
  int blocks = ceil_div(total, 256);
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  split_to_planes_kernel<<<blocks, 256, 0, stream>>>(src, rows, cols, ld_src, transpose, scale, hi, lo, rows_pad, cols_pad, fmt);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
