// Device side of the split-bf16 tcgen05 layer GEMM (see mlp_gemm.cu for the description): constants, PTX wrappers, epilogue
// math and the kernel template.  Included by the translation units that instantiate the kernel (mlp_gemm.cu: single-CTA
// variants and the host launcher; mlp_gemm_pair.cu: CTA-pair variants) -- two units only so that they compile in parallel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include "mlp_gemm.cuh"

namespace nefii {

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
#ifndef NEFII_GEMM_BK1
#define NEFII_GEMM_BK1 64
#endif
#ifndef NEFII_GEMM_BK2
#define NEFII_GEMM_BK2 64
#endif
// Operand ring.  CL = 1: one CTA multiplies its 128 rows by the whole 256-row weight tile (96 KB per stage, 2 stages).
// CL = 2: a CTA pair (cta_group::2, one 256 x 256 x 16 MMA across two SMs): each CTA stages its own 128 rows of A and
// HALF of the weight tile, the tensor cores read the other half from the peer's shared memory -- 64 KB per stage, 3 stages,
// and a third less shared-memory traffic per FLOP (the single-CTA kernel is bound by exactly that, profiles/r1_gemm_*.md).
// A 64-wide K block (the unit of the host interface and of the partial-sum schedule) travels as 64 / kBK ring stages:
// smaller stages mean more of the 192 KB ring is in flight while one stage is being multiplied -- the tensor pipe idles
// whenever the operand feed (L2 -> shared memory, ~1 us under load) falls behind, and one 96 KB stage of lookahead does.
constexpr int kRingBytes = 196608;
template <int CL> struct Ring {
  static constexpr int kBK = CL == 2 ? NEFII_GEMM_BK2 : NEFII_GEMM_BK1;   // K columns per stage (64: SWIZZLE_128B rows, 32: SWIZZLE_64B)
  static constexpr int kSub = BK / kBK;
  static constexpr int kBRows = BN / CL;
  static constexpr int kATileBytes = BM * kBK * 2;
  static constexpr int kBTileBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
  static constexpr int kStages = kRingBytes / kStageBytes;
  static_assert(kBK == 64 || kBK == 32, "stage width");
  static_assert(kStages * kStageBytes == kRingBytes && kStages >= 2, "ring size");
};
constexpr int kMaxStages = 6;
static_assert(Ring<1>::kStages <= kMaxStages && Ring<2>::kStages <= kMaxStages, "barrier slots");
// mbarriers: [0..2] operand stage full, [3..5] stage empty, [6..7] TMEM partial full, [8..9] TMEM partial empty
constexpr int kBarFull = 0, kBarEmpty = kMaxStages, kBarTFull = 2 * kMaxStages, kBarTEmpty = 2 * kMaxStages + 2;
constexpr int kEpiWarps = 8;
constexpr int kPeWarps = 2;               // warps 10, 11: idle unless the A operand is a positional encoding computed in the kernel
constexpr int kThreads = 128 + kEpiWarps * 32;  // warpgroup 0: TMA + MMA (+2 idle warps); warpgroups 1,2: epilogue
constexpr int kTmemCols = 512;
constexpr int kMaxLast = 4;
constexpr int kStageOutBytes = 4096;      // per epilogue warp: one 32 x 32 tile of both bf16 planes (64 B rows) for coalesced stores
constexpr int kBiasSmemFloats = 512;      // the bias of layers up to 512 outputs is staged in shared memory once per CTA
// After the operand ring and the barriers: the plane-store staging + bias (plain layers) share their bytes with the
// fused output layer's row partials (FUSE kernels never take the staged path).  The ring must start 1024-byte aligned
// (SWIZZLE_128B atoms); dynamic shared memory starts at the CTA window's base, the kernel traps if that ever changes.
constexpr size_t kTailBytes = (size_t)kEpiWarps * kStageOutBytes + kBiasSmemFloats * sizeof(float);
static_assert(kTailBytes >= 2 * BM * kMaxLast * sizeof(float) + (1 + kMaxLast) * kBiasSmemFloats * sizeof(float),
              "fused-layer partials, bias and output weights alias the staging area");
constexpr int kBarBytes = 512;            // barriers + TMEM slot; keeps the staging tiles 512-byte aligned (SWIZZLE_64B pattern)
constexpr size_t kSmemBytes = (size_t)kRingBytes + kBarBytes + kTailBytes;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// CTA-pair load: lands in this CTA's shared memory, completes bytes on the LEADER CTA's barrier (bit 24 of a
// shared::cluster address selects the odd CTA of a pair; clearing it names the same barrier in the even one).
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(x), "r"(y)
      : "memory");
}
// plane stores of whole 32-row tiles: bulk tensor store from the staged (64 B-swizzled) tile, one instruction per plane
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// commit of a CTA pair's MMAs: one arrive on the barrier at this offset in each CTA of the mask
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// K-major, 128-byte-swizzled shared-memory matrix descriptor (8-row atoms of 1024 B).
// ROW_BYTES = 128: SWIZZLE_128B rows of 64 bf16; 64: SWIZZLE_64B rows of 32 bf16.  8-row atoms either way.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;    // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;   // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// kind::f16 instruction descriptor: 16-bit operands -> fp32, both operands K-major, M=128, N=256.  Operand format bits
// (a: 7..9, b: 10..12): 1 = bf16 (kIdescBf16 added for PLANES_BF16), 0 = fp16 (PLANES_FP16).
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// CTA pair: M = 256 (128 rows in each CTA's TMEM), N = 256
constexpr uint32_t kIdescPair = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr uint32_t kIdescBf16 = (1u << 7) | (1u << 10);

// ---- branch-free activation helpers (fast intrinsics: the error they add, <= 1e-9 absolute on a
// softplus output, is far below the bf16x3 product error; the SG kernels never use them) ----------
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int ACT> __device__ __forceinline__ float act_fwd(float v) {
  if (ACT == ACT_SOFTPLUS100) {
    // softplus(beta=100, threshold=20): log1p(exp(100 v)) / 100, and v itself once 100 v > 20.
    // 8 instructions: exp and log through the MUFU ex2 / lg2 units (absolute error of the result < 1e-9).
    const float x = v * 144.26950408889634f;                  // 100 v log2(e)
    const float e = fast_ex2(fminf(x, 28.853900817779268f));  // exp(min(100 v, 20))
    const float h = fast_lg2(1.f + e) * 0.0069314718055994530f;   // ln(1 + e) / 100
    return (x > 28.853900817779268f) ? v : h;
  } else if (ACT == ACT_RELU) {
    return fmaxf(v, 0.f);
  } else if (ACT == ACT_ELU) {
    const float m = fminf(v, 0.f);
    const float em = (m > -1e-2f) ? m * (1.f + m * (0.5f + 0.16666667f * m)) : __expf(m) - 1.f;
    return (v > 0.f) ? v : em;
  }
  return v;
}
// derivative of the activation recovered from its saved output h
template <int ACT> __device__ __forceinline__ float act_bwd_from_output(float h) {
  if (ACT == ACT_SOFTPLUS100) {
    const float u = 100.f * h;   // sigmoid(100 z) = 1 - exp(-100 h)
    const float small = u * (1.f - 0.5f * u + 0.16666667f * u * u);
    return (u < 0.01f) ? small : 1.f - __expf(-u);
  } else if (ACT == ACT_RELU) {
    return h > 0.f ? 1.f : 0.f;
  } else if (ACT == ACT_ELU) {
    return h > 0.f ? 1.f : h + 1.f;
  }
  return 1.f;
}

// Plane formats (`fmt`, uniform over a launch).  PLANES_BF16: hi = bf16(x), lo = bf16(x - hi): 16 significant bits, fp32's
// exponent range (gradients of any magnitude).  PLANES_FP16: hi = fp16(x), lo = fp16(x - hi): 22 significant bits while
// |x| >= 2^-3 and an absolute error <= 2^-25 below that (lo becomes subnormal), |x| < 65504 -- the format of the SDF
// network's inference chain, whose activations and input gradients are O(1) (csrc/sdf_mlp.cu).  The planes are 16-bit storage
// either way (the pointers are typed __nv_bfloat16 for both).
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
// hi + lo of two neighbouring plane elements
__device__ __forceinline__ float2 join_pair(uint32_t wh, uint32_t wl, int fmt) {
  float2 hf, lf;
  if (fmt == PLANES_FP16) {
    hf = __half22float2(*reinterpret_cast<const __half2*>(&wh));
    lf = __half22float2(*reinterpret_cast<const __half2*>(&wl));
  } else {
    hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wh));
    lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wl));
  }
  return make_float2(hf.x + lf.x, hf.y + lf.y);
}
__device__ __forceinline__ void split2(float x, int fmt, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  uint32_t wh, wl;
  split_pair(x, 0.f, fmt, wh, wl);
  const unsigned short h = (unsigned short)(wh & 0xFFFFu), l = (unsigned short)(wl & 0xFFFFu);
  hi = *reinterpret_cast<const __nv_bfloat16*>(&h);
  lo = *reinterpret_cast<const __nv_bfloat16*>(&l);
}

// pack 8 floats into 8 hi (uint4) + 8 lo (uint4) plane elements
__device__ __forceinline__ void split8(const float* v, int fmt, uint4& hi, uint4& lo) {
  uint32_t wh[4], wl[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_pair(v[2 * j], v[2 * j + 1], fmt, wh[j], wl[j]);
  hi = make_uint4(wh[0], wh[1], wh[2], wh[3]);
  lo = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

// Store 32 consecutive values of one row as planes starting at plane column `col` (= col_base + n0);
// elements with n0 + j >= n_limit are not written.  16-byte stores when a group of 8 is complete and aligned.
__device__ __forceinline__ void store_planes32(const Planes& dst, int fmt, long long row, int col, int n0, int n_limit,
                                               const float* vals) {
  __nv_bfloat16* ph = dst.hi + row * dst.ld + col;
  __nv_bfloat16* pl = dst.lo + row * dst.ld + col;
  const bool aligned = (col & 7) == 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (aligned && n0 + 8 * g + 8 <= n_limit) {
      uint4 h, l;
      split8(vals + 8 * g, fmt, h, l);
      *reinterpret_cast<uint4*>(ph + 8 * g) = h;
      *reinterpret_cast<uint4*>(pl + 8 * g) = l;
    } else if (n0 + 8 * g < n_limit) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + 8 * g + j < n_limit) {
          __nv_bfloat16 h, l;
          split2(vals[8 * g + j], fmt, h, l);
          ph[8 * g + j] = h;
          pl[8 * g + j] = l;
        }
      }
    }
  }
}

// fp32 side output of columns [begin, end): dst[row * ld + n - begin]
__device__ __forceinline__ void store_f32_32(float* dst, int ld, long long row, int n0, int begin, int end,
                                             const float* vals) {
  float* p = dst + row * ld + (n0 - begin);
  const bool aligned = (((size_t)p) & 15) == 0;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int n = n0 + 4 * g;
    if (aligned && n >= begin && n + 4 <= end) {
      *reinterpret_cast<float4*>(p + 4 * g) = make_float4(vals[4 * g], vals[4 * g + 1], vals[4 * g + 2], vals[4 * g + 3]);
    } else if (n + 4 > begin && n < end) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j >= begin && n + j < end) p[4 * g + j] = vals[4 * g + j];
    }
  }
}

constexpr int kColsPerWarp = BN / 2;   // each epilogue warp owns one TMEM lane quarter x 128 columns

// Fast path (the hidden layers of every MLP): this warp's whole 128-column span holds real outputs that go to aligned
// planes only.  One straight-line block for the four 32-column groups (no per-element predicates), so the math of one
// group is scheduled under the shared-memory transpose and the global stores of the previous one.
//   * bias: broadcast 16-byte loads from the copy staged in shared memory;
//   * each group's hi and lo tiles are transposed through shared memory so that every store instruction writes 8 rows x
//     64 contiguous bytes (whole sectors) instead of 32 rows x 16 bytes: lane = row on the way in, (row, 16-byte piece)
//     = (8 i + lane / 4, lane % 4) on the way out; pieces are XOR-swizzled with the row, both directions conflict free;
//   * the global row pointers and row predicates are computed once per tile (p_hi / p_lo / ok_mask).
struct FastStore {
  uint4* st_in;                 // staging, this lane's row: [plane][32 rows][4 pieces]
  const uint4* st_out;          // staging, (row lane / 4, piece lane % 4) after swizzle
  int sw_in;                    // (lane >> 1) & 3
  __nv_bfloat16* p_hi;          // plane element (row_warp0 + lane / 4, dst_col0 + 8 (lane % 4))
  __nv_bfloat16* p_lo;
  long long stride8;            // 8 rows, in elements
  unsigned ok_mask;             // bit i: row_warp0 + 8 i + lane / 4 is a valid row
  const CUtensorMap* map_hi;    // bulk-store path (whole 32-row tiles): the planes as 2-D tensors, origin at dst_col0
  const CUtensorMap* map_lo;
  uint32_t st_u32;              // staging base (shared-window address)
  int y0;                       // row_warp0
};

// KEEP: the span's last real column (keep_from) is not on a 16-byte boundary (a separate instance: the common one stays lean)
// FMT: the plane format as a compile-time constant -- the caller branches once per span, so that each instance is straight-line
// code (a run-time format test inside the unrolled conversion groups doubles the instruction footprint of the hot loop).
template <int ACT, bool TMA, int FMT, bool KEEP = false>
__device__ __forceinline__ void finish_span_fast(float* acc, const float4* s_bias4, int n_span0, float scale, const FastStore& fs, int dbg,
                                                 int keep_from = -1, const uint4* keep_hi = nullptr, const uint4* keep_lo = nullptr) {
  constexpr int fmt = FMT;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < kColsPerWarp / 32; ++c) {
    const float* v = acc + c * 32;
    uint4 hq[4], lq[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 ba = s_bias4[(n_span0 + c * 32) / 4 + 2 * g], bb = s_bias4[(n_span0 + c * 32) / 4 + 2 * g + 1];
      float o[8];
      o[0] = act_fwd<ACT>(v[8 * g + 0] + ba.x) * scale; o[1] = act_fwd<ACT>(v[8 * g + 1] + ba.y) * scale;
      o[2] = act_fwd<ACT>(v[8 * g + 2] + ba.z) * scale; o[3] = act_fwd<ACT>(v[8 * g + 3] + ba.w) * scale;
      o[4] = act_fwd<ACT>(v[8 * g + 4] + bb.x) * scale; o[5] = act_fwd<ACT>(v[8 * g + 5] + bb.y) * scale;
      o[6] = act_fwd<ACT>(v[8 * g + 6] + bb.z) * scale; o[7] = act_fwd<ACT>(v[8 * g + 7] + bb.w) * scale;
      split8(o, fmt, hq[g], lq[g]);
      // The bulk store writes whole 16-byte pieces: when the layer's last real column (keep_from) falls inside this piece, the
      // plane elements from keep_from on are not this layer's -- they keep what is there (the PE half of a skip layer's input,
      // written by layer 0's kernel): patched in from global memory before the tile is staged.
      if (KEEP && TMA) {
        const int g0 = n_span0 + c * 32 + 8 * g;
        if (g0 < keep_from && g0 + 8 > keep_from) {
          const uint4 kh = __ldg(keep_hi + (g0 >> 3)), kl = __ldg(keep_lo + (g0 >> 3));
          unsigned short* ph = reinterpret_cast<unsigned short*>(&hq[g]);
          unsigned short* pl = reinterpret_cast<unsigned short*>(&lq[g]);
          const unsigned short* qh = reinterpret_cast<const unsigned short*>(&kh);
          const unsigned short* ql = reinterpret_cast<const unsigned short*>(&kl);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (g0 + j >= keep_from) { ph[j] = qh[j]; pl[j] = ql[j]; }
        }
      }
    }
    if (dbg & 16) {   // timing ablation: math only
      if (hq[0].x == 0x7fc07fc1u && lq[3].w == 0x12345678u) fs.p_hi[c] = __float2bfloat16(1.f);
      continue;
    }
    if (lane == 0) bulk_wait_read_all();   // an earlier bulk store may still be reading the staging tile
    __syncwarp();     // the previous group's tiles have been read
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      fs.st_in[g ^ fs.sw_in] = hq[g];
      fs.st_in[128 + (g ^ fs.sw_in)] = lq[g];
    }
    if (TMA) {
      // the staged layout IS the tensor map's SWIZZLE_64B layout (16-byte piece ^= (row >> 1) & 3): one bulk store per
      // plane writes the 32 x 64 B tile; the warp goes straight on to the next group's math
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && !(dbg & 32)) {
        tma_store_2d(fs.map_hi, fs.st_u32, n_span0 + c * 32, fs.y0);
        tma_store_2d(fs.map_lo, fs.st_u32 + 2048, n_span0 + c * 32, fs.y0);
        bulk_commit();
      }
      continue;
    }
    __syncwarp();
    uint4 oh[4], ol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      oh[i] = fs.st_out[i * 32];
      ol[i] = fs.st_out[128 + i * 32];
    }
    if (dbg & 32) {   // timing ablation: no global stores
      if (oh[0].x == 0x7fc07fc1u && ol[3].w == 0x12345678u) fs.p_hi[c] = __float2bfloat16(1.f);
      continue;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (fs.ok_mask & (1u << i)) {
        *reinterpret_cast<uint4*>(fs.p_hi + i * fs.stride8 + (n_span0 + c * 32)) = oh[i];
        *reinterpret_cast<uint4*>(fs.p_lo + i * fs.stride8 + (n_span0 + c * 32)) = ol[i];
      }
    }
  }
}

// Fast path of the data-gradient layers (MODE 1) on whole 32-row tiles whose column span is entirely real: the products are
// scaled by the activation derivative recovered from the saved forward output (read as this lane's row, 16 bytes at a time) and
// leave as bf16 planes through the same staged bulk tensor stores as the forward path; no per-element predicates.
template <int ACT, int FMT>
__device__ __forceinline__ void finish_span_bwd_fast(float* acc, const __nv_bfloat16* sav_hi_row, const __nv_bfloat16* sav_lo_row,
                                                     float sav_scale, int n_span0, float scale, const FastStore& fs) {
  constexpr int fmt = FMT;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < kColsPerWarp / 32; ++c) {
    const float* v = acc + c * 32;
    const uint4* sh = reinterpret_cast<const uint4*>(sav_hi_row + n_span0 + c * 32);
    const uint4* sl = reinterpret_cast<const uint4*>(sav_lo_row + n_span0 + c * 32);
    uint4 hq[4], lq[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 h4 = __ldg(sh + g), l4 = __ldg(sl + g);
      const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 sv = join_pair(hw[j], lw[j], fmt);
        o[2 * j] = v[8 * g + 2 * j] * act_bwd_from_output<ACT>(sv.x * sav_scale) * scale;
        o[2 * j + 1] = v[8 * g + 2 * j + 1] * act_bwd_from_output<ACT>(sv.y * sav_scale) * scale;
      }
      split8(o, fmt, hq[g], lq[g]);
    }
    if (lane == 0) bulk_wait_read_all();   // the previous bulk stores have read the staging tile
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      fs.st_in[g ^ fs.sw_in] = hq[g];
      fs.st_in[128 + (g ^ fs.sw_in)] = lq[g];
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(fs.map_hi, fs.st_u32, n_span0 + c * 32, fs.y0);
      tma_store_2d(fs.map_lo, fs.st_u32 + 2048, n_span0 + c * 32, fs.y0);
      bulk_commit();
    }
  }
}

// Fast path of the fused output layer when nothing but the n_last dot products leaves the tile (the SDF value of the
// tracer's evaluations): activation and FMA only, bias and output weights read as broadcast 16-byte loads from the copies
// staged in shared memory (s_w: [kMaxLast][kBiasSmemFloats]).
template <int ACT>
__device__ __forceinline__ void finish_span_fused_fast(const float* acc, const float4* s_bias4, const float4* s_w4, int n_span0,
                                                       int n_last, float* part) {
#pragma unroll
  for (int g = 0; g < kColsPerWarp / 4; ++g) {
    const float4 b = s_bias4[n_span0 / 4 + g];
    const float h0 = act_fwd<ACT>(acc[4 * g + 0] + b.x), h1 = act_fwd<ACT>(acc[4 * g + 1] + b.y);
    const float h2 = act_fwd<ACT>(acc[4 * g + 2] + b.z), h3 = act_fwd<ACT>(acc[4 * g + 3] + b.w);
#pragma unroll
    for (int q = 0; q < kMaxLast; ++q) {
      if (q < n_last) {
        // group of four summed as a tree, then one add into the running sum: the chain of roundings per row is 32 long
        // instead of 128 (the output layer's own rounding error was 2.4e-7 rms on an SDF value, as much as the layers before it)
        const float4 w = s_w4[q * (kBiasSmemFloats / 4) + n_span0 / 4 + g];
        part[q] += fmaf(h1, w.y, h0 * w.x) + fmaf(h3, w.w, h2 * w.z);
      }
    }
  }
}

// Final per-tile math on the fp32 sums of one 32-column group held in registers (v[0..31]).
template <int MODE, int ACT, bool FUSE>
__device__ __forceinline__ void finish_group(const GemmEpilogue& epi, float* v, int n0, long long row, bool row_ok,
                                             float* part) {
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      const bool real = n < epi.n_valid;
      const float b = (epi.bias != nullptr && real) ? __ldg(epi.bias + n) : 0.f;
      const float h = act_fwd<ACT>(v[j] + b);
      v[j] = real ? h : 0.f;
    }
    if (FUSE) {
#pragma unroll
      for (int q = 0; q < kMaxLast; ++q) {
        if (q < epi.n_last) {
          // the same association as finish_span_fused_fast (groups of four as a tree, one add per group): an evaluation
          // returns the same SDF bits whether or not features / gradient seeds leave the tile
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = n0 + 4 * g + j;
              w4[j] = (n < epi.n_valid) ? __ldg(epi.w_last + (size_t)q * epi.w_last_ld + n) : 0.f;
            }
            part[q] += fmaf(v[4 * g + 1], w4[1], v[4 * g] * w4[0]) + fmaf(v[4 * g + 3], w4[3], v[4 * g + 2] * w4[2]);
          }
        }
      }
      if (epi.seed.hi != nullptr && row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float sv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = n0 + 8 * g + j;
            const float w = (n < epi.n_valid) ? __ldg(epi.w_last + n) : 0.f;
            sv[j] = w * act_bwd_from_output<ACT>(v[8 * g + j]);
          }
          // the seed buffer is a full-width plane buffer: 16-byte stores are always aligned and in range
          uint4 h, l;
          split8(sv, epi.fmt, h, l);
          *reinterpret_cast<uint4*>(epi.seed.hi + row * epi.seed.ld + n0 + 8 * g) = h;
          *reinterpret_cast<uint4*>(epi.seed.lo + row * epi.seed.ld + n0 + 8 * g) = l;
        }
      }
    }
  } else {
    if (epi.sav_hi != nullptr && row_ok && n0 < epi.sav_ncols) {
      const uint4* sh = reinterpret_cast<const uint4*>(epi.sav_hi + row * epi.sav_ld + n0);
      const uint4* sl = reinterpret_cast<const uint4*>(epi.sav_lo + row * epi.sav_ld + n0);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 h4 = __ldg(sh + g), l4 = __ldg(sl + g);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 sv = join_pair(hw[j], lw[j], epi.fmt);
          const int n = n0 + 8 * g + 2 * j;
          const float d0 = act_bwd_from_output<ACT>(sv.x * epi.sav_scale);
          const float d1 = act_bwd_from_output<ACT>(sv.y * epi.sav_scale);
          v[8 * g + 2 * j] *= (n < epi.sav_ncols) ? d0 : 1.f;
          v[8 * g + 2 * j + 1] *= (n + 1 < epi.sav_ncols) ? d1 : 1.f;
        }
      }
    }
  }
  if (!row_ok) return;
  if (epi.dst_f32 != nullptr && n0 < epi.f32_end && n0 + 32 > epi.f32_begin)
    store_f32_32(epi.dst_f32, epi.f32_ld, row, n0, epi.f32_begin, epi.f32_end, v);
  const int dst_end = epi.dst_zero_to > epi.dst_ncols ? epi.dst_zero_to : epi.dst_ncols;
  if (epi.dst.hi != nullptr && n0 < dst_end) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (n0 + j < epi.dst_ncols) ? v[j] * epi.out_scale : 0.f;
    store_planes32(epi.dst, epi.fmt, row, epi.dst_col0 + n0, n0, dst_end, v);
  }
}

// Accumulation scheme (accuracy): tcgen05 adds into its fp32 accumulator with truncation, which biases long
// sums (measured: -4 ulp at K=512, -39 ulp at K=2048, tools/diag_gpu.py trunc).  Every 64-wide K block is
// therefore multiplied into a *fresh* TMEM buffer (the two small cross terms first, hi*hi last) and the
// partial products are summed across K blocks by the epilogue warps in registers with round-to-nearest.
// CL = 2: the CTAs of a pair (cluster of 2, consecutive row tiles) run ONE tcgen05.mma.cta_group::2 stream issued by the
// even CTA; both load operands (completing on the leader's barrier), both run their own epilogue on their own TMEM.
// Partial schedule of one column chunk: the first two partials take `head` K blocks each, the rest `tail`.  Only two
// partials fit in TMEM, so the MMAs of everything after the second partial wait for the epilogue's final math of the
// previous chunk: longer head partials move work under that math, shorter tail partials bound the truncation error.
struct PartSched {
  int head, tail, k_blocks;
  __device__ __forceinline__ int count() const {
    return k_blocks <= 2 * head ? (k_blocks + head - 1) / head : 2 + (k_blocks - 2 * head + tail - 1) / tail;
  }
  __device__ __forceinline__ int begin(int p) const {
    const int b = p <= 2 ? p * head : 2 * head + (p - 2) * tail;
    return b < k_blocks ? b : k_blocks;
  }
};

// PE = true: the "PE prologue" instantiation (A operand computed in the kernel, see PeSource); a separate instantiation so that
// its code -- sincosf in warps limited to 40 registers, with a stack frame -- costs the other variants nothing.
template <int MODE, int ACT, bool FUSE, int CL, bool PE = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_split_bf16_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const __grid_constant__ CUtensorMap map_d_hi, const __grid_constant__ CUtensorMap map_d_lo, int store_tma,
                       const int* __restrict__ count_ptr, int rows_cap, int k_blocks_total, int n_chunks, int kb_per_split,
                       long long f32_split_stride, int dbg, int k_flush, float part_scale_full, float part_scale_last, const __grid_constant__ GemmEpilogue epi_in,
                       const __grid_constant__ PeSource pe) {
  // Persistent over row tiles: CTA x handles tiles x, x + gridDim.x, ... so that the final epilogue math of one tile
  // overlaps the MMAs of the next (the pipelines and barrier phases simply keep running across tiles).
  int m_tile0 = blockIdx.x;
  int tile_stride = gridDim.x;
  int nc_begin = 0, nc_end = n_chunks;
  // split-K (weight gradients): CTA (x, y) reduces K blocks [y * kb_per_split, ...) into its own fp32 partial
  const int kb_begin = blockIdx.y * kb_per_split;
  const int k_blocks = min(k_blocks_total, kb_begin + kb_per_split) - kb_begin;
  GemmEpilogue epi = epi_in;
  if (epi.dst_f32 != nullptr) epi.dst_f32 += (long long)blockIdx.y * f32_split_stride;
  // uniform over the cluster: leave only if the cluster's FIRST tile is already past the valid rows
  static_assert(CL == 1 || CL == 2, "single CTA or a cta_group::2 pair");
  using R = Ring<CL>;
  constexpr int kStages = R::kStages;
  constexpr int kStageBytes = R::kStageBytes;
  constexpr int kATileBytes = R::kATileBytes;
  constexpr int kBTileBytes = R::kBTileBytes;
  constexpr int kBK = R::kBK;
  const int cta_rank = (CL > 1) ? (int)cluster_ctarank() : 0;

  // ---- prologue: nothing here reads what the previous kernel in the stream wrote, so under programmatic dependent launch
  // (cudaLaunchAttributeProgrammaticStreamSerialization, set by the launcher) it runs while that kernel is still draining ----
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw;   // no integer round-trip: keeps every access below a shared-window (LDS/STS) access
  if (smem_u32(smem_raw) & 1023u) __trap();
  unsigned char* tiles = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kRingBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  float* s_last = reinterpret_cast<float*>(smem + (size_t)kRingBytes + kBarBytes);   // [2][BM][kMaxLast]

  // Warp roles: warps 0..7 epilogue, warp 8 TMA producer, warp 9 MMA issuer (+TMEM alloc), warps 10, 11 idle.  The role
  // warps carry the highest warp ids of their schedulers: the issue arbiter prefers the highest id, and the single
  // MMA / TMA threads must never queue behind the epilogue's long ALU streams.
  const int warp = threadIdx.x >> 5;
  const int role = warp - kEpiWarps;   // 0 TMA, 1 MMA, <0 epilogue
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      // the (leader's) producer arrives once, TMA completes the bytes; PE prologue: plus the two encoding warps of every CTA
      mbar_init(smem_u32(&bars[kBarFull + s]), PE ? 1 + kPeWarps * CL : 1);
      mbar_init(smem_u32(&bars[kBarEmpty + s]), 1);   // one commit (multicast to both CTAs of a pair)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars[kBarTFull + b]), 1);
      mbar_init(smem_u32(&bars[kBarTEmpty + b]), kEpiWarps * CL);   // the leader hears from both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (role == 1) {
    if (CL == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {   // the same warp of both CTAs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- from here on the kernel reads what its predecessor produced (row count, activation planes) ----
  asm volatile("griddepcontrol.wait;" ::: "memory");                // no-op without the launch attribute
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel may start ITS prologue on free SMs

  int m_limit = rows_cap;
  if (count_ptr != nullptr) {
    int c = *count_ptr;
    if (c < m_limit) m_limit = c;
  }
  bool live = true;
  // Few live row tiles (the march / bisection rounds of a trace; decided from the DEVICE-side row count): the column chunks of
  // a tile go to different clusters instead of running one after the other in the same one, which halves the latency of a
  // layer when less than half of the grid has work.  Cluster c takes (tile group c % G, chunk c / G), G = live tile groups.
  // Not for the fused output layer (its dot product needs the whole row) and not together with split-K.
  if (!FUSE && n_chunks > 1 && gridDim.y == 1) {
    const int live_tiles = (m_limit + BM - 1) / BM;
    const int groups = (live_tiles + CL - 1) / CL;
    const int clusters = (int)gridDim.x / CL;
    if (groups > 0 && groups * n_chunks <= clusters) {
      const int c = (int)blockIdx.x / CL;
      if (c >= groups * n_chunks) live = false;         // uniform over the cluster
      m_tile0 = (c % groups) * CL + cta_rank;
      nc_begin = c / groups;
      nc_end = nc_begin + 1;
      tile_stride = 1 << 28;                            // one tile per CTA
    }
  }
  // every loop below runs while the PAIR's first tile is live, so both CTAs of a pair take the same trips
  if ((long long)(m_tile0 - cta_rank) * BM >= m_limit || k_blocks <= 0) live = false;
#define NEFII_TILE_LIVE(t) ((long long)((t) - cta_rank) * BM < m_limit)

  if (live) {
  // Register re-partitioning (168 regs/thread at launch): the role warpgroup keeps 40, each epilogue
  // warpgroup grows to 232 so the 128 fp32 partial sums per thread stay in registers.
  if (role >= 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
  if (role == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m_tile = m_tile0; NEFII_TILE_LIVE(m_tile); m_tile += tile_stride)
      for (int nc = nc_begin; nc < nc_end; ++nc) {
        for (int ks = 0; ks < k_blocks * R::kSub; ++ks) {
          mbar_wait(smem_u32(&bars[kBarEmpty + stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bars[kBarFull + stage]);
          if (dbg & 2) { if (cta_rank == 0) mbar_arrive(full); if (++stage == kStages) { stage = 0; phase ^= 1; } continue; }
          unsigned char* st = tiles + (size_t)stage * kStageBytes;
          const int kx = kb_begin * BK + ks * kBK;
          const bool load_a = !PE;                                      // PE prologue: warps 10, 11 write the A tiles
          const int tx = load_a ? kStageBytes : 2 * kBTileBytes;
          if (CL == 1) {
            mbar_expect_tx(full, tx);
            if (load_a) {
              tma_load_2d(smem_u32(st), &map_a_hi, full, kx, m_tile * BM);
              tma_load_2d(smem_u32(st + kATileBytes), &map_a_lo, full, kx, m_tile * BM);
            }
            tma_load_2d(smem_u32(st + 2 * kATileBytes), &map_b_hi, full, kx, nc * BN);
            tma_load_2d(smem_u32(st + 2 * kATileBytes + kBTileBytes), &map_b_lo, full, kx, nc * BN);
          } else {
            if (cta_rank == 0) mbar_expect_tx(full, 2 * tx);            // both CTAs' bytes land on the leader's barrier
            const int brow = nc * BN + cta_rank * R::kBRows;            // this CTA's half of the weight tile
            if (load_a) {
              tma_load_2d_pair(smem_u32(st), &map_a_hi, full, kx, m_tile * BM);
              tma_load_2d_pair(smem_u32(st + kATileBytes), &map_a_lo, full, kx, m_tile * BM);
            }
            tma_load_2d_pair(smem_u32(st + 2 * kATileBytes), &map_b_hi, full, kx, brow);
            tma_load_2d_pair(smem_u32(st + 2 * kATileBytes + kBTileBytes), &map_b_lo, full, kx, brow);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (role == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // K blocks are accumulated in TMEM in groups of k_flush ("partials"); each partial goes to a fresh buffer and is
      // added to the register accumulators by the epilogue warps (bounds the tensor core's truncation bias), and the two
      // buffers let the MMAs of up to two partials run ahead of the epilogue's final math.
      const PartSched sched{(k_flush >> 8) ? (k_flush >> 8) : (k_flush & 255), k_flush & 255, k_blocks};
      const int parts_per_chunk = sched.count();
      const uint32_t idesc = (CL == 1 ? kIdesc : kIdescPair) | (epi.fmt == PLANES_FP16 ? 0u : kIdescBf16);
      uint32_t pcount = 0;
      for (int m_tile = m_tile0; NEFII_TILE_LIVE(m_tile); m_tile += tile_stride)
      for (int nc = nc_begin; nc < nc_end; ++nc) {
        for (int pi = 0; pi < parts_per_chunk; ++pi, ++pcount) {
          const int buf = pcount & 1;
          mbar_wait(smem_u32(&bars[kBarTEmpty + buf]), ((pcount >> 1) & 1) ^ 1);   // the partial of two groups ago was read
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
          const int kb_first = sched.begin(pi), kb_end = sched.begin(pi + 1);
          for (int ks = kb_first * R::kSub; ks < kb_end * R::kSub; ++ks) {
            mbar_wait(smem_u32(&bars[kBarFull + stage]), phase);
            tc_fence_after();
            const uint32_t st = smem_u32(tiles + (size_t)stage * kStageBytes);
            const uint64_t a_hi = make_smem_desc<2 * kBK>(st);
            const uint64_t a_lo = make_smem_desc<2 * kBK>(st + kATileBytes);
            const uint64_t b_hi = make_smem_desc<2 * kBK>(st + 2 * kATileBytes);
            const uint64_t b_lo = make_smem_desc<2 * kBK>(st + 2 * kATileBytes + kBTileBytes);
            const uint32_t fresh = (ks == kb_first * R::kSub) ? 0u : 1u;
            if (!(dbg & 4)) {
#pragma unroll
              for (int k = 0; k < kBK / UMMA_K; ++k) {
                const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);   // 32 B per K step inside the swizzle row
                if (CL == 1) {
                  tc_mma_bf16(tmem_d, a_hi + koff, b_lo + koff, idesc, (k != 0) ? 1u : fresh);
                  tc_mma_bf16(tmem_d, a_lo + koff, b_hi + koff, idesc, 1);
                } else {
                  tc_mma_bf16_pair(tmem_d, a_hi + koff, b_lo + koff, idesc, (k != 0) ? 1u : fresh);
                  tc_mma_bf16_pair(tmem_d, a_lo + koff, b_hi + koff, idesc, 1);
                }
              }
#pragma unroll
              for (int k = 0; k < kBK / UMMA_K; ++k) {
                const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                if (CL == 1) tc_mma_bf16(tmem_d, a_hi + koff, b_hi + koff, idesc, 1);
                else tc_mma_bf16_pair(tmem_d, a_hi + koff, b_hi + koff, idesc, 1);
              }
            }
            // frees the operand stage (in both CTAs of a pair) once these MMAs retire
            if (CL == 1) tc_commit(smem_u32(&bars[kBarEmpty + stage]));
            else tc_commit_pair(smem_u32(&bars[kBarEmpty + stage]), 3);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (CL == 1) tc_commit(smem_u32(&bars[kBarTFull + buf]));     // this partial product is ready
          else tc_commit_pair(smem_u32(&bars[kBarTFull + buf]), 3);       // ... in both CTAs' TMEM
        }
      }
    }
  } else if (PE) {
    // ---------------------------------------------------------------- PE prologue (warps 10, 11)
    // Row r of the tile = encoding of x[m_tile * 128 + r], written as hi / lo planes into the stage's A tiles in the layout TMA
    // would have produced (K-major rows of 128 bytes, SWIZZLE_128B: 16-byte piece j of row r sits at piece j ^ (r & 7)), then
    // made visible to the tensor core (async proxy) and signalled on the (leader's) full barrier like a completed load.
    static_assert(kBK == 64, "the PE prologue writes whole 64-column K blocks");
    const int pw = role - 2;                      // 0, 1
    int stage = 0;
    uint32_t phase = 0;
    for (int m_tile = m_tile0; NEFII_TILE_LIVE(m_tile); m_tile += tile_stride)
    for (int nc = nc_begin; nc < nc_end; ++nc) {
      mbar_wait(smem_u32(&bars[kBarEmpty + stage]), phase ^ 1);
      unsigned char* st = tiles + (size_t)stage * kStageBytes;
#pragma unroll 1
      for (int half_rows = 0; half_rows < 2; ++half_rows) {
        const int r = pw * 64 + half_rows * 32 + lane;
        const long long row = (long long)m_tile * BM + r;
        const bool ok = row < m_limit;
        float p[3] = {0.f, 0.f, 0.f};
        if (ok) { p[0] = pe.x[row * 3 + 0]; p[1] = pe.x[row * 3 + 1]; p[2] = pe.x[row * 3 + 2]; }
        unsigned short* a_hi = reinterpret_cast<unsigned short*>(st + (size_t)r * 128);
        unsigned short* a_lo = reinterpret_cast<unsigned short*>(st + kATileBytes + (size_t)r * 128);
        const int sw = r & 7;
        const bool side = ok && nc == 0 && pe.side.hi != nullptr;
        unsigned short* s_hi = reinterpret_cast<unsigned short*>(pe.side.hi) + row * pe.side.ld + pe.side_col0;
        unsigned short* s_lo = reinterpret_cast<unsigned short*>(pe.side.lo) + row * pe.side.ld + pe.side_col0;
        auto put = [&](int col, float v) {
          uint32_t wh, wl;
          split_pair(v, 0.f, epi.fmt, wh, wl);
          const int pos = (((col >> 3) ^ sw) << 3) | (col & 7);
          a_hi[pos] = (unsigned short)wh;
          a_lo[pos] = (unsigned short)wl;
          if (side) {
            split_pair(v * pe.side_scale, 0.f, epi.fmt, wh, wl);
            s_hi[col] = (unsigned short)wh;
            s_lo[col] = (unsigned short)wl;
          }
        };
#pragma unroll
        for (int c = 0; c < 3; ++c) put(c, p[c]);
        for (int k = 0; k < pe.n_freqs; ++k) {
          const float f = exp2f((float)k);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(p[c] * f, &sn, &cs);
            put(3 + 6 * k + c, sn);
            put(6 + 6 * k + c, cs);
          }
        }
        // zero padding up to the 64-wide K block (finite values: the weight planes' padding is zero, but 0 * NaN is not)
        const int d_pe = 3 + 6 * pe.n_freqs;
        for (int col = d_pe; col < BK; ++col) {
          const int pos = (((col >> 3) ^ sw) << 3) | (col & 7);
          a_hi[pos] = 0;
          a_lo[pos] = 0;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1) mbar_arrive(smem_u32(&bars[kBarFull + stage]));
        else mbar_arrive_cluster(smem_u32(&bars[kBarFull + stage]), 0);
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    const int e = warp;                 // 0..7
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int half = e >> 2;            // which 128-column half of the 256-column chunk
    const int row_in_tile = quarter * 32 + lane;
    unsigned char* tail = smem + (size_t)kRingBytes + kBarBytes;
    uint4* stage_out = reinterpret_cast<uint4*>(tail + (size_t)e * kStageOutBytes);
    float* s_bias = reinterpret_cast<float*>(tail + (size_t)kEpiWarps * kStageOutBytes);
    // plain hidden layer writing aligned planes: bias staged in shared memory, spans take finish_span_fast
    const bool fast_layer = MODE == 0 && !FUSE && epi.bias != nullptr && epi.dst.hi != nullptr && epi.dst_f32 == nullptr &&
                            (epi.dst_col0 & 7) == 0 && (epi.dst.ld & 7) == 0 && n_chunks * BN <= kBiasSmemFloats;
    if (fast_layer) {
      for (int i = threadIdx.x; i < n_chunks * BN; i += kEpiWarps * 32) s_bias[i] = i < epi.n_valid ? __ldg(epi.bias + i) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
    }
    // data-gradient layer writing aligned planes with a saved forward output for every column it produces
    const bool fast_bwd = MODE == 1 && store_tma && epi.dst.hi != nullptr && epi.dst_f32 == nullptr && epi.sav_hi != nullptr &&
                          (epi.dst_col0 & 7) == 0 && (epi.dst.ld & 7) == 0 && (epi.sav_ld & 7) == 0;
    // fused output layer with only its dot products as output: bias and output weights staged behind the row partials
    float* s_fbias = reinterpret_cast<float*>(tail + 2 * BM * kMaxLast * sizeof(float));
    float* s_fw = s_fbias + kBiasSmemFloats;
    const bool fast_fused = MODE == 0 && FUSE && epi.bias != nullptr && epi.dst.hi == nullptr && epi.dst_f32 == nullptr &&
                            epi.seed.hi == nullptr && n_chunks * BN <= kBiasSmemFloats;
    if (fast_fused) {
      for (int i = threadIdx.x; i < n_chunks * BN; i += kEpiWarps * 32) {
        s_fbias[i] = i < epi.n_valid ? __ldg(epi.bias + i) : 0.f;
        for (int q = 0; q < kMaxLast; ++q)
          s_fw[q * kBiasSmemFloats + i] = (q < epi.n_last && i < epi.n_valid) ? __ldg(epi.w_last + (size_t)q * epi.w_last_ld + i) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
    }
    FastStore fs;
    fs.sw_in = (lane >> 1) & 3;
    fs.st_in = stage_out + lane * 4;
    fs.st_out = stage_out + (lane >> 2) * 4 + ((lane & 3) ^ ((lane >> 3) & 3));
    fs.stride8 = 8ll * epi.dst.ld;
    fs.map_hi = &map_d_hi;
    fs.map_lo = &map_d_lo;
    fs.st_u32 = smem_u32(stage_out);
    const int n_loop = epi.dst_zero_to > epi.n_valid ? epi.dst_zero_to : epi.n_valid;
    const int n_real = epi.n_valid < epi.dst_ncols ? epi.n_valid : epi.dst_ncols;
    const int n_fast = epi.dst_pad_ok ? (n_real + kColsPerWarp - 1) / kColsPerWarp * kColsPerWarp : n_real;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * kColsPerWarp);
    float acc[kColsPerWarp];
    const PartSched sched{(k_flush >> 8) ? (k_flush >> 8) : (k_flush & 255), k_flush & 255, k_blocks};
    const int parts_per_chunk = sched.count();
    uint32_t pcount = 0;
    for (int m_tile = m_tile0; NEFII_TILE_LIVE(m_tile); m_tile += tile_stride) {
    const long long row = (long long)m_tile * BM + row_in_tile;
    const bool row_ok = row < m_limit;
    float part[kMaxLast] = {0.f, 0.f, 0.f, 0.f};
    if (fast_layer || fast_bwd) {
      const long long r0 = row - lane + (lane >> 2);
      fs.p_hi = epi.dst.hi + r0 * epi.dst.ld + epi.dst_col0 + 8 * (lane & 3);
      fs.p_lo = epi.dst.lo + r0 * epi.dst.ld + epi.dst_col0 + 8 * (lane & 3);
      fs.ok_mask = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) fs.ok_mask |= (r0 + 8 * i < m_limit) ? (1u << i) : 0u;
      fs.y0 = (int)(row - lane);
    }
    for (int nc = nc_begin; nc < nc_end; ++nc) {
      for (int pi = 0; pi < parts_per_chunk; ++pi, ++pcount) {
        const int buf = pcount & 1;
        mbar_wait(smem_u32(&bars[kBarTFull + buf]), (pcount >> 1) & 1);
        tc_fence_after();
        // four TMEM loads in flight per wait: the flush is latency-bound otherwise
        if (!(dbg & 8))
#pragma unroll
        for (int hc = 0; hc < kColsPerWarp / 64; ++hc) {
          uint32_t r[64];
#pragma unroll
          for (int q = 0; q < 4; ++q) tc_ld16_nowait(t_lane + (uint32_t)(buf * BN + hc * 64 + q * 16), r + q * 16);
          tc_wait_ld();
          const float part_scale = (pi == parts_per_chunk - 1) ? part_scale_last : part_scale_full;
          // part_scale = 1 + rho: first-order compensation of the tensor core's round-toward-zero accumulation (every addend
          // is truncated towards zero when it is aligned to the accumulator: a partial sum comes out short by a factor that
          // depends on its length only; gemm_set_trunc_comp).  1.0f reproduces the plain sum bit for bit.
          if (pi == 0) {
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[hc * 64 + j] = __uint_as_float(r[j]) * part_scale;
          } else {
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[hc * 64 + j] = fmaf(__uint_as_float(r[j]), part_scale, acc[hc * 64 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 1) mbar_arrive(smem_u32(&bars[kBarTEmpty + buf]));
          else mbar_arrive_cluster(smem_u32(&bars[kBarTEmpty + buf]), 0);   // the leader issues the pair's MMAs
        }
      }
      const int n_span0 = nc * BN + half * kColsPerWarp;
      if (dbg & 1) continue;
      // dst_pad_ok: the plane columns up to the next multiple of 128 are never written -- the bulk store clips at the tensor
      // map's extent (dst_ncols); a ragged span that cannot take the bulk store falls through to the predicated path below
      // (the PE half of a skip layer's input may already sit in those columns)
      const bool whole_rows = store_tma && row - lane + 32 <= m_limit;
      if (fast_layer && n_span0 + kColsPerWarp <= n_fast && (whole_rows || n_span0 + kColsPerWarp <= n_real)) {
        // whole 32-row tiles leave by bulk tensor store; the ragged last tile keeps per-row predicates
        const float4* sb4 = reinterpret_cast<const float4*>(s_bias);
        if (whole_rows) {
          // a ragged span (dst_pad_ok) whose last real column is not on a 16-byte boundary: see finish_span_fast
          if (n_span0 + kColsPerWarp > n_real && (n_real & 7) != 0) {
            const uint4* kh = reinterpret_cast<const uint4*>(epi.dst.hi + row * epi.dst.ld + epi.dst_col0);
            const uint4* kl = reinterpret_cast<const uint4*>(epi.dst.lo + row * epi.dst.ld + epi.dst_col0);
            if (epi.fmt == PLANES_FP16) finish_span_fast<ACT, true, PLANES_FP16, true>(acc, sb4, n_span0, epi.out_scale, fs, dbg, n_real, kh, kl);
            else finish_span_fast<ACT, true, PLANES_BF16, true>(acc, sb4, n_span0, epi.out_scale, fs, dbg, n_real, kh, kl);
          } else if (epi.fmt == PLANES_FP16) {
            finish_span_fast<ACT, true, PLANES_FP16>(acc, sb4, n_span0, epi.out_scale, fs, dbg);
          } else {
            finish_span_fast<ACT, true, PLANES_BF16>(acc, sb4, n_span0, epi.out_scale, fs, dbg);
          }
        } else if (epi.fmt == PLANES_FP16) {
          finish_span_fast<ACT, false, PLANES_FP16>(acc, sb4, n_span0, epi.out_scale, fs, dbg);
        } else {
          finish_span_fast<ACT, false, PLANES_BF16>(acc, sb4, n_span0, epi.out_scale, fs, dbg);
        }
        continue;
      }
      if (fast_bwd && row - lane + 32 <= m_limit && n_span0 + kColsPerWarp <= epi.n_valid && n_span0 + kColsPerWarp <= epi.dst_ncols &&
          n_span0 + kColsPerWarp <= epi.sav_ncols) {
        if (epi.fmt == PLANES_FP16)
          finish_span_bwd_fast<ACT, PLANES_FP16>(acc, epi.sav_hi + row * epi.sav_ld, epi.sav_lo + row * epi.sav_ld, epi.sav_scale, n_span0,
                                                 epi.out_scale, fs);
        else
          finish_span_bwd_fast<ACT, PLANES_BF16>(acc, epi.sav_hi + row * epi.sav_ld, epi.sav_lo + row * epi.sav_ld, epi.sav_scale, n_span0,
                                                 epi.out_scale, fs);
        continue;
      }
      if (fast_fused && n_span0 + kColsPerWarp <= epi.n_valid) {
        finish_span_fused_fast<ACT>(acc, reinterpret_cast<const float4*>(s_fbias), reinterpret_cast<const float4*>(s_fw), n_span0,
                                    epi.n_last, part);
        continue;
      }
#pragma unroll
      for (int c = 0; c < kColsPerWarp / 32; ++c) {
        const int n0 = n_span0 + c * 32;
        if (n0 < n_loop) finish_group<MODE, ACT, FUSE>(epi, acc + c * 32, n0, row, row_ok, part);
      }
    }

    if (MODE == 0 && FUSE) {
      // combine the two column halves of each row (deterministic order) and add the bias
#pragma unroll
      for (int q = 0; q < kMaxLast; ++q) s_last[(half * BM + row_in_tile) * kMaxLast + q] = part[q];
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      if (half == 0 && row_ok) {
        for (int q = 0; q < epi.n_last; ++q) {
          float y = s_last[(0 * BM + row_in_tile) * kMaxLast + q] + s_last[(1 * BM + row_in_tile) * kMaxLast + q];
          if (epi.b_last) y += __ldg(epi.b_last + q);
          epi.dst_last[(size_t)row * epi.n_last + q] = y;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");   // s_last is reused by the next tile
    }
    }   // tile loop
    if (lane == 0) bulk_wait_all();   // this warp's bulk stores have left shared memory and are on their way
  }

  }   // live

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into / signal this CTA
  if (role == 1) {
    tc_fence_after();
    if (CL == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
#undef NEFII_TILE_LIVE
}

using GemmKernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, int, const int*, int, int, int,
                              int, long long, int, int, float, float, GemmEpilogue, PeSource);

// kernel of one (mode, act, fused-output-layer) combination for cluster size CL; key = fuse * 8 + mode * 4 + act, 12 / 13 = the
// PE prologue instantiations
constexpr int kGemmKernelKeys = 14;
template <int CL>
GemmKernelFn select_gemm_kernel(int key) {
  switch (key) {
    case 0: return gemm_split_bf16_kernel<0, ACT_NONE, false, CL>;
    case 1: return gemm_split_bf16_kernel<0, ACT_SOFTPLUS100, false, CL>;
    case 2: return gemm_split_bf16_kernel<0, ACT_RELU, false, CL>;
    case 3: return gemm_split_bf16_kernel<0, ACT_ELU, false, CL>;
    case 4: return gemm_split_bf16_kernel<1, ACT_NONE, false, CL>;
    case 5: return gemm_split_bf16_kernel<1, ACT_SOFTPLUS100, false, CL>;
    case 6: return gemm_split_bf16_kernel<1, ACT_RELU, false, CL>;
    case 7: return gemm_split_bf16_kernel<1, ACT_ELU, false, CL>;
    case 8: return gemm_split_bf16_kernel<0, ACT_NONE, true, CL>;
    case 9: return gemm_split_bf16_kernel<0, ACT_SOFTPLUS100, true, CL>;
    case 10: return gemm_split_bf16_kernel<0, ACT_RELU, true, CL>;
    case 11: return gemm_split_bf16_kernel<0, ACT_ELU, true, CL>;
    case 12: return gemm_split_bf16_kernel<0, ACT_SOFTPLUS100, false, CL, true>;   // PE prologue (layer 0 of the SDF network)
    case 13: return gemm_split_bf16_kernel<0, ACT_NONE, false, CL, true>;          // ... without activation (tests)
    default: return nullptr;
  }
}

}  // namespace

// defined in mlp_gemm_pair.cu
void* gemm_pair_kernel(int key);

}  // namespace nefii
