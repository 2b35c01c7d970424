// IDR sphere tracing -- reference RayTracing (code/model/ray_tracing.py:29-337) and
// get_sphere_intersection (code/utils/rend_util.py:200-221) re-built as per-ray state machines.
//
// The reference marches all rays in lock step with boolean-mask gathers and ~2 host syncs per
// iteration.  Here every ray carries its own state in HBM (SoA), each march round is one kernel that
// consumes the SDF values it asked for in the previous round, advances the state machine and emits
// a *compacted* list of the points it needs next (warp ballot + block scan + one atomic per CTA);
// the SDF MLP (tcgen05 GEMM chain) then runs on exactly that list -- dead lanes never reach the
// tensor cores.  One host sync per trace (to size the sampler passes) replaces ~60 in the reference.
//
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true: every value that feeds a comparison is
// a rounded multiply followed by a rounded add, as torch evaluates it.
#include <cuda_bf16.h>
#include <algorithm>
#include "tracer.cuh"

namespace nefii {

namespace {

constexpr int kBlock = 256;
enum : unsigned char { F_HIT = 1, F_UNF_S = 2, F_UNF_E = 4, F_PEND_S = 8, F_PEND_E = 16, F_NET = 32, F_SAMP = 64, F_MIN = 128 };
enum { ROUND_INIT = 0, ROUND_HEAD = 1, ROUND_LS = 2, ROUND_FINAL = 3 };
constexpr int kMaxCounters = 128;

struct RayState {
  // inputs
  const float* cam_loc; const float* dirs; const unsigned char* object_mask; int n_pix; int n_rays;
  // per-ray state
  float *acc_s, *acc_e, *min_dis, *max_dis, *cur_s, *cur_e, *nxt_s, *nxt_e;
  int *slot_s, *slot_e;
  unsigned char* flags;
  // request list
  float* req_pts; float* req_sdf; int* counters;
  // sampler / root-find / min-sdf lists
  int *samp_list, *min_list, *root_list;
  float *z_lo, *z_hi, *s_lo, *s_hi;
  int* any_work;       // [n_rootfind_steps + 2]
  int* blocks_done;    // [n_rootfind_steps + 2]
  // outputs
  float* points; unsigned char* hit; float* dists;
};

__device__ __forceinline__ float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

__device__ __forceinline__ void ray_od(const RayState& S, int r, float* o, float* d) {
  const int b = r / S.n_pix;
#pragma unroll
  for (int c = 0; c < 3; ++c) { o[c] = S.cam_loc[b * 3 + c]; d[c] = S.dirs[(size_t)r * 3 + c]; }
}

// Block-level compaction: every thread contributes n (0..2) requests; returns the global base slot of the thread.
__device__ __forceinline__ int reserve_slots(int n, int* counter) {
  __shared__ int warp_tot[kBlock / 32];
  __shared__ int block_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = n;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) { int t = warp_tot[w]; warp_tot[w] = tot; tot += t; }
    block_base = tot ? atomicAdd(counter, tot) : 0;
  }
  __syncthreads();
  return block_base + warp_tot[warp] + incl - n;
}

__device__ __forceinline__ void emit_point(const RayState& S, int slot, const float* o, float t, const float* d) {
  S.req_pts[(size_t)slot * 3 + 0] = o[0] + t * d[0];
  S.req_pts[(size_t)slot * 3 + 1] = o[1] + t * d[1];
  S.req_pts[(size_t)slot * 3 + 2] = o[2] + t * d[2];
}

__global__ void __launch_bounds__(kBlock)
march_round_kernel(RayState S, int kind, int first_head, float radius, float thr, float back, int counter_out) {
  const int r = blockIdx.x * kBlock + threadIdx.x;
  const bool live = r < S.n_rays;
  float o[3] = {0, 0, 0}, d[3] = {0, 0, 0};
  unsigned char f = 0;
  float acc_s = 0, acc_e = 0, cur_s = 0, cur_e = 0, nxt_s = 0, nxt_e = 0;
  bool req_s = false, req_e = false;
  if (live) {
    ray_od(S, r, o, d);
    if (kind == ROUND_INIT) {
      // rend_util.get_sphere_intersection
      const float b = dot3(d, o);
      const float onorm = sqrtf((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]);
      const float under = b * b - (onorm * onorm - radius * radius);
      const bool hits = under > 0.f;
      float t0 = 0.f, t1 = 0.f;
      if (hits) {
        const float root = sqrtf(under);
        t0 = root * -1.0f - b;
        t1 = root * 1.0f - b;
        t0 = fmaxf(t0, 0.01f);
        t1 = fmaxf(t1, 0.01f);
      }
      acc_s = t0; acc_e = t1;
      S.min_dis[r] = acc_s; S.max_dis[r] = acc_e;
      f = hits ? (F_HIT | F_UNF_S | F_UNF_E) : 0;
      req_s = req_e = hits;
    } else {
      f = S.flags[r];
      acc_s = S.acc_s[r]; acc_e = S.acc_e[r]; cur_s = S.cur_s[r]; cur_e = S.cur_e[r];
      nxt_s = S.nxt_s[r]; nxt_e = S.nxt_e[r];
      if (f & F_PEND_S) nxt_s = S.req_sdf[S.slot_s[r]];
      if (f & F_PEND_E) nxt_e = S.req_sdf[S.slot_e[r]];
      f &= ~(F_PEND_S | F_PEND_E);
      bool unf_s = f & F_UNF_S, unf_e = f & F_UNF_E;
      if (kind == ROUND_HEAD || kind == ROUND_FINAL) {
        if (!first_head) {   // end of the previous iteration (ray_tracing.py:190-191)
          const bool crossed = acc_s < acc_e;
          unf_s = unf_s && crossed;
          unf_e = unf_e && crossed;
        }
        cur_s = unf_s ? nxt_s : 0.f;
        if (cur_s <= thr) cur_s = 0.f;
        cur_e = unf_e ? nxt_e : 0.f;
        if (cur_e <= thr) cur_e = 0.f;
        unf_s = unf_s && (cur_s > thr);
        unf_e = unf_e && (cur_e > thr);
        if (kind == ROUND_HEAD && (unf_s || unf_e)) {
          acc_s = acc_s + cur_s;
          acc_e = acc_e - cur_e;
          nxt_s = 0.f; nxt_e = 0.f;
          req_s = unf_s; req_e = unf_e;
        }
      } else {   // ROUND_LS: step back where the march crossed the surface (ray_tracing.py:170-188)
        const bool bad_s = nxt_s < 0.f, bad_e = nxt_e < 0.f;
        if (bad_s) { acc_s = acc_s - back * cur_s; req_s = true; }
        if (bad_e) { acc_e = acc_e + back * cur_e; req_e = true; }
      }
      f = (f & ~(F_UNF_S | F_UNF_E)) | (unf_s ? F_UNF_S : 0) | (unf_e ? F_UNF_E : 0);
    }
  }
  const int n_req = (req_s ? 1 : 0) + (req_e ? 1 : 0);
  int slot = reserve_slots(n_req, S.counters + counter_out);
  if (!live) return;
  if (req_s) { emit_point(S, slot, o, acc_s, d); S.slot_s[r] = slot; f |= F_PEND_S; ++slot; }
  if (req_e) { emit_point(S, slot, o, acc_e, d); S.slot_e[r] = slot; f |= F_PEND_E; }
  S.flags[r] = f;
  S.acc_s[r] = acc_s; S.acc_e[r] = acc_e; S.cur_s[r] = cur_s; S.cur_e[r] = cur_e; S.nxt_s[r] = nxt_s; S.nxt_e[r] = nxt_e;
}

// After sphere tracing: network mask, sampler list, (training) projection of sphere-missing rays and min-SDF list.
// counters[c_samp], counters[c_min] receive the list lengths.
__global__ void __launch_bounds__(kBlock)
post_march_kernel(RayState S, int training, int want_min, int c_samp, int c_min) {
  const int r = blockIdx.x * kBlock + threadIdx.x;
  if (r >= S.n_rays) return;
  float o[3], d[3];
  ray_od(S, r, o, d);
  unsigned char f = S.flags[r];
  float acc_s = S.acc_s[r];
  const float acc_e = S.acc_e[r];
  const bool hits = f & F_HIT;
  const bool net = acc_s < acc_e;
  const bool samp = f & F_UNF_S;
  const bool obj = S.object_mask ? (S.object_mask[r] != 0) : true;
  f &= ~(F_NET | F_SAMP | F_MIN);
  if (net) f |= F_NET;
  if (samp) {
    f |= F_SAMP;
    S.samp_list[atomicAdd(S.counters + c_samp, 1)] = r;
  }
  if (training) {
    const bool in_mask = !net && obj && !samp;
    const bool out_mask = !obj && !samp;
    if ((in_mask || out_mask) && !hits) {
      acc_s = -dot3(d, o);                       // closest approach to the origin (ray_tracing.py:82-87)
    }
    if ((in_mask || out_mask) && hits) {
      if (net && out_mask) S.min_dis[r] = acc_s;  // ray_tracing.py:92
      if (want_min) {
        f |= F_MIN;
        S.min_list[atomicAdd(S.counters + c_min, 1)] = r;
      }
    }
  }
  S.flags[r] = f;
  S.acc_s[r] = acc_s;
  S.dists[r] = acc_s;
  S.hit[r] = net ? 1 : 0;
  S.points[(size_t)r * 3 + 0] = o[0] + acc_s * d[0];
  S.points[(size_t)r * 3 + 1] = o[1] + acc_s * d[1];
  S.points[(size_t)r * 3 + 2] = o[2] + acc_s * d[2];
}

// n_steps sample points per listed ray: t_j = lo + frac_j * (hi - lo)  (uniform: frac = linspace, lo/hi = march
// interval; min-SDF: frac = shared uniforms, lo/hi = min_dis/max_dis)
__global__ void __launch_bounds__(kBlock)
sample_emit_kernel(RayState S, const int* __restrict__ list, int list_begin, int n_list, int n_steps,
                   const float* __restrict__ frac, int use_minmax) {
  const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= (long long)n_list * n_steps) return;
  const int k = (int)(i / n_steps), j = (int)(i % n_steps);
  const int r = list[list_begin + k];
  float o[3], d[3];
  ray_od(S, r, o, d);
  const float lo = use_minmax ? S.min_dis[r] : S.acc_s[r];
  const float hi = use_minmax ? S.max_dis[r] : S.acc_e[r];
  const float t = use_minmax ? (frac[j] * (hi - lo) + lo) : (lo + frac[j] * (hi - lo));
  emit_point(S, (int)i, o, t, d);
}

// One warp per sampled ray: first sign change / minimal SDF, candidate interval for the bisection.
__global__ void __launch_bounds__(kBlock)
sample_reduce_kernel(RayState S, const int* __restrict__ list, int list_begin, int n_list, int n_steps,
                     const float* __restrict__ frac, int training, int c_root) {
  const int k = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n_list) return;
  const int r = list[list_begin + k];
  const float* v = S.req_sdf + (size_t)k * n_steps;
  int first_neg = n_steps, first_zero = n_steps, best_j = n_steps;
  float best = 0.f;
  for (int j = lane; j < n_steps; j += 32) {
    const float x = v[j];
    if (x < 0.f && first_neg == n_steps) first_neg = j;
    if (x == 0.f && first_zero == n_steps) first_zero = j;
    if (best_j == n_steps || x < best) { best = x; best_j = j; }   // j ascending: strict < keeps the first minimum
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    first_neg = min(first_neg, __shfl_xor_sync(0xffffffffu, first_neg, off));
    first_zero = min(first_zero, __shfl_xor_sync(0xffffffffu, first_zero, off));
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
    if (oj < n_steps && (best_j == n_steps || ob < best || (ob == best && oj < best_j))) { best = ob; best_j = oj; }
  }
  if (lane != 0) return;
  const int first = first_neg < n_steps ? first_neg : (first_zero < n_steps ? first_zero : n_steps - 1);
  float o[3], d[3];
  ray_od(S, r, o, d);
  const float lo = S.acc_s[r], hi = S.acc_e[r];
  const bool inside_net = v[first] < 0.f;
  const bool inside_gt = S.object_mask ? (S.object_mask[r] != 0) : true;
  const int sel = (inside_gt && inside_net) ? first : best_j;
  const float t_sel = lo + frac[sel] * (hi - lo);
  unsigned char f = S.flags[r];
  f = inside_net ? (f | F_NET) : (f & ~F_NET);
  S.flags[r] = f;
  S.hit[r] = inside_net ? 1 : 0;
  S.dists[r] = t_sel;
  S.points[(size_t)r * 3 + 0] = o[0] + t_sel * d[0];
  S.points[(size_t)r * 3 + 1] = o[1] + t_sel * d[1];
  S.points[(size_t)r * 3 + 2] = o[2] + t_sel * d[2];
  const bool refine = training ? (inside_net && inside_gt) : inside_net;
  if (refine) {
    const int lo_i = (first + n_steps - 1) % n_steps;   // python's [-1]: sample 0 pairs with the last sample
    const int slot = atomicAdd(S.counters + c_root, 1);
    S.root_list[slot] = r;
    const float zl = lo + frac[lo_i] * (hi - lo), zh = lo + frac[first] * (hi - lo);
    const float sl = v[lo_i], sh = v[first];
    S.z_lo[slot] = zl; S.z_hi[slot] = zh; S.s_lo[slot] = sl; S.s_hi[slot] = sh;
    if (sl > 0.f && sh < 0.f && zh > zl) S.any_work[0] = 1;
  }
}

// Bisection round `step` (RayTracing.rootfind, ray_tracing.py:259-280).  The reference bisects *every* passed ray
// while ANY ray still has work, so the loop is batch-coupled: any_work[i] says whether iteration i runs.
// Kernel `step` applies iteration step-1 (if it ran), ORs the surviving work bits into any_work[step], emits the
// next mid-points, and its last CTA publishes the request count of iteration `step` (n_root or 0).
__global__ void __launch_bounds__(kBlock)
bisect_round_kernel(RayState S, int step, int c_root, int counter_out, int last_step, unsigned char* __restrict__ work) {
  const int n_root = S.counters[c_root];
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const bool ran_prev = step > 0 && S.any_work[step - 1] != 0;
  if (k < n_root) {
    float zl = S.z_lo[k], zh = S.z_hi[k];
    if (ran_prev) {
      const float zm = (zl + zh) * 0.5f;
      const float sm = S.req_sdf[k];
      if (sm > 0.f) { zl = zm; S.z_lo[k] = zl; S.s_lo[k] = sm; }
      if (sm <= 0.f) { zh = zm; S.z_hi[k] = zh; S.s_hi[k] = sm; }
      const unsigned char w = work[k] && ((zh - zl) > 1e-6f);
      work[k] = w;
      if (w) S.any_work[step] = 1;
    }
    const int r = S.root_list[k];
    float o[3], d[3];
    ray_od(S, r, o, d);
    const float zm = (zl + zh) * 0.5f;
    if (!last_step) {
      emit_point(S, k, o, zm, d);
    } else {
      S.dists[r] = zm;
      S.points[(size_t)r * 3 + 0] = o[0] + zm * d[0];
      S.points[(size_t)r * 3 + 1] = o[1] + zm * d[1];
      S.points[(size_t)r * 3 + 2] = o[2] + zm * d[2];
    }
  }
  if (last_step) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(S.blocks_done + step, 1);
    if (done == (int)gridDim.x - 1) {
      __threadfence();
      const int alive = *(volatile int*)(S.any_work + step);
      S.counters[counter_out] = alive ? n_root : 0;
    }
  }
}

// work_mask = (sdf_low > 0) & (sdf_high < 0) & (z_high > z_low)   (ray_tracing.py:261)
__global__ void __launch_bounds__(kBlock) bisect_init_kernel(RayState S, int c_root, unsigned char* __restrict__ work) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k < S.counters[c_root]) work[k] = (S.s_lo[k] > 0.f && S.s_hi[k] < 0.f && S.z_hi[k] > S.z_lo[k]) ? 1 : 0;
}

// minimal_sdf_points (ray_tracing.py:309-337): argmin over the n_steps shared random depths
__global__ void __launch_bounds__(kBlock)
minsdf_reduce_kernel(RayState S, const int* __restrict__ list, int list_begin, int n_list, int n_steps,
                     const float* __restrict__ frac) {
  const int k = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= n_list) return;
  const int r = list[list_begin + k];
  const float* v = S.req_sdf + (size_t)k * n_steps;
  int best_j = n_steps;
  float best = 0.f;
  for (int j = lane; j < n_steps; j += 32) {
    const float x = v[j];
    if (best_j == n_steps || x < best) { best = x; best_j = j; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, off);
    if (oj < n_steps && (best_j == n_steps || ob < best || (ob == best && oj < best_j))) { best = ob; best_j = oj; }
  }
  if (lane != 0) return;
  float o[3], d[3];
  ray_od(S, r, o, d);
  const float lo = S.min_dis[r], hi = S.max_dis[r];
  const float t = frac[best_j] * (hi - lo) + lo;
  S.dists[r] = t;
  S.points[(size_t)r * 3 + 0] = o[0] + t * d[0];
  S.points[(size_t)r * 3 + 1] = o[1] + t * d[1];
  S.points[(size_t)r * 3 + 2] = o[2] + t * d[2];
}

// Analytic test scene (oracle/tracer.py: analytic_sdf): union of spheres and boxes, fixed operation order.
__global__ void __launch_bounds__(kBlock)
analytic_sdf_kernel(const float* __restrict__ prims, int n_prims, int n, const int* __restrict__ count,
                    const float* __restrict__ x, float* __restrict__ out) {
  int limit = n;
  if (count) limit = min(limit, *count);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < limit; i += gridDim.x * kBlock) {
    const float p0 = x[(size_t)i * 3 + 0], p1 = x[(size_t)i * 3 + 1], p2 = x[(size_t)i * 3 + 2];
    float best = 0.f;
    for (int k = 0; k < n_prims; ++k) {
      const float* P = prims + k * 8;
      const float q0 = p0 - P[1], q1 = p1 - P[2], q2 = p2 - P[3];
      float val;
      if (P[0] == 0.f) {
        val = sqrtf((q0 * q0 + q1 * q1) + q2 * q2) - P[4];
      } else {
        const float a0 = fabsf(q0) - P[4], a1 = fabsf(q1) - P[5], a2 = fabsf(q2) - P[6];
        const float m0 = fmaxf(a0, 0.f), m1 = fmaxf(a1, 0.f), m2 = fmaxf(a2, 0.f);
        const float outside = sqrtf((m0 * m0 + m1 * m1) + m2 * m2);
        const float inside = fminf(fmaxf(a0, fmaxf(a1, a2)), 0.f);
        val = outside + inside;
      }
      best = (k == 0) ? val : fminf(best, val);
    }
    out[i] = best;
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Layout {
  size_t off_f[8], off_slot[2], off_flags, off_req_pts, off_req_sdf, off_counters, off_lists[3], off_z[4], off_any, off_done,
      off_work, off_mlp, total;
  int cap_pts;
};

Layout make_layout(const SdfSource& src, int n_rays, int n_steps) {
  Layout L;
  const size_t R = (size_t)std::max(n_rays, 1);
  long long cap = std::max<long long>(2 * (long long)R, std::min<long long>((long long)R * n_steps, 1 << 20));
  cap = std::max<long long>(cap, n_steps);
  L.cap_pts = (int)cap;
  size_t p = 0;
  auto take = [&](size_t bytes) { size_t o = p; p = align_up(p + bytes, 256); return o; };
  for (int i = 0; i < 8; ++i) L.off_f[i] = take(R * 4);
  for (int i = 0; i < 2; ++i) L.off_slot[i] = take(R * 4);
  L.off_flags = take(R);
  L.off_req_pts = take((size_t)cap * 12);
  L.off_req_sdf = take((size_t)cap * 4);
  L.off_counters = take(kMaxCounters * 4);
  for (int i = 0; i < 3; ++i) L.off_lists[i] = take(R * 4);
  for (int i = 0; i < 4; ++i) L.off_z[i] = take(R * 4);
  L.off_any = take(kMaxCounters * 4);
  L.off_done = take(kMaxCounters * 4);
  L.off_work = take(R);
  L.off_mlp = p;
  if (src.net) p += align_up(src.net->workspace_bytes((int)cap, false), 256);
  L.total = p + 1024;
  return L;
}

}  // namespace

size_t trace_workspace_bytes(const SdfSource& src, int n_rays, int n_steps) { return make_layout(src, n_rays, n_steps).total; }

int analytic_sdf_eval(cudaStream_t stream, const float* prims, int n_prims, int n, const int* count, const float* x, float* sdf) {
  if (n <= 0) return NEFII_OK;
  NEFII_CHECK_ARG(prims && n_prims > 0 && x && sdf, "analytic_sdf_eval: null pointer");
  int blocks = std::min(ceil_div(n, kBlock), kNumSMs * 16);
  analytic_sdf_kernel<<<blocks, kBlock, 0, stream>>>(prims, n_prims, n, count, x, sdf);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int ray_trace(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
              const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
              const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
              unsigned char* hit, float* dists, long long* stats) {
  const long long n_rays_ll = (long long)n_batch * n_pix;
  NEFII_CHECK_ARG(n_batch >= 0 && n_pix >= 0 && n_rays_ll < (1ll << 30), "ray_trace: bad ray count");
  const int R = (int)n_rays_ll;
  if (stats) for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (R == 0) return NEFII_OK;
  NEFII_CHECK_ARG(cam_loc && ray_dirs && points && hit && dists && workspace && linspace, "ray_trace: null pointer");
  NEFII_CHECK_ARG(src.net != nullptr || (src.prims != nullptr && src.n_prims > 0), "ray_trace: no SDF source");
  NEFII_CHECK_ARG(cfg.n_steps >= 2 && cfg.n_steps <= 4096, "ray_trace: n_steps out of range");
  NEFII_CHECK_ARG(cfg.n_rootfind_steps >= 0 && cfg.n_rootfind_steps + 2 <= kMaxCounters / 2, "ray_trace: n_rootfind_steps out of range");
  const int n_rounds = 1 + cfg.sphere_tracing_iters * (1 + cfg.line_step_iters) + 1;
  NEFII_CHECK_ARG(cfg.sphere_tracing_iters >= 0 && cfg.line_step_iters >= 0 && n_rounds + cfg.n_rootfind_steps + 8 <= kMaxCounters,
                  "ray_trace: too many march rounds (%d)", n_rounds);
  const bool training = flags & TRACE_TRAINING;
  const bool want_min = training && !(flags & TRACE_SKIP_MIN_SDF);
  NEFII_CHECK_ARG(!want_min || uniforms != nullptr, "ray_trace: training mode needs the %d uniforms", cfg.n_steps);
  const Layout L = make_layout(src, R, cfg.n_steps);
  NEFII_CHECK_ARG(ws_bytes >= L.total, "ray_trace: workspace too small (%zu < %zu)", ws_bytes, L.total);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);

  RayState S;
  S.cam_loc = cam_loc; S.dirs = ray_dirs; S.object_mask = object_mask; S.n_pix = n_pix; S.n_rays = R;
  float** fptr[8] = {&S.acc_s, &S.acc_e, &S.min_dis, &S.max_dis, &S.cur_s, &S.cur_e, &S.nxt_s, &S.nxt_e};
  for (int i = 0; i < 8; ++i) *fptr[i] = (float*)(base + L.off_f[i]);
  S.slot_s = (int*)(base + L.off_slot[0]); S.slot_e = (int*)(base + L.off_slot[1]);
  S.flags = (unsigned char*)(base + L.off_flags);
  S.req_pts = (float*)(base + L.off_req_pts); S.req_sdf = (float*)(base + L.off_req_sdf);
  S.counters = (int*)(base + L.off_counters);
  S.samp_list = (int*)(base + L.off_lists[0]); S.min_list = (int*)(base + L.off_lists[1]); S.root_list = (int*)(base + L.off_lists[2]);
  S.z_lo = (float*)(base + L.off_z[0]); S.z_hi = (float*)(base + L.off_z[1]); S.s_lo = (float*)(base + L.off_z[2]); S.s_hi = (float*)(base + L.off_z[3]);
  S.any_work = (int*)(base + L.off_any); S.blocks_done = (int*)(base + L.off_done);
  unsigned char* work = (unsigned char*)(base + L.off_work);
  S.points = points; S.hit = hit; S.dists = dists;
  void* mlp_ws = base + L.off_mlp;
  const size_t mlp_ws_bytes = src.net ? src.net->workspace_bytes(L.cap_pts, false) : 0;

  // counters: [0, n_rounds) march rounds; then sampler, min-sdf, root list lengths; then bisection rounds
  const int c_samp = n_rounds, c_min = n_rounds + 1, c_root = n_rounds + 2, c_bis0 = n_rounds + 3;
  NEFII_CUDA(cudaMemsetAsync(base + L.off_counters, 0, (L.off_work - L.off_counters) + (size_t)R, stream));

  auto eval = [&](int rows_cap, const int* count) -> int {
    if (src.net) return src.net->eval(stream, rows_cap, count, S.req_pts, mlp_ws, mlp_ws_bytes, S.req_sdf, nullptr, nullptr);
    return analytic_sdf_eval(stream, src.prims, src.n_prims, rows_cap, count, S.req_pts, S.req_sdf);
  };

  const int grid = ceil_div(R, kBlock);
  int rc;
  int round = 0;
  march_round_kernel<<<grid, kBlock, 0, stream>>>(S, ROUND_INIT, 1, cfg.radius, cfg.sdf_threshold, 0.f, round);
  NEFII_LAUNCH_CHECK();
  if ((rc = eval(2 * R, S.counters + round))) return rc;
  ++round;
  for (int it = 0; it < cfg.sphere_tracing_iters; ++it) {
    march_round_kernel<<<grid, kBlock, 0, stream>>>(S, ROUND_HEAD, it == 0, cfg.radius, cfg.sdf_threshold, 0.f, round);
    NEFII_LAUNCH_CHECK();
    if ((rc = eval(2 * R, S.counters + round))) return rc;
    ++round;
    for (int ls = 0; ls < cfg.line_step_iters; ++ls) {
      const float back = (float)((1.0 - (double)cfg.line_search_step) / (double)(1 << ls));
      march_round_kernel<<<grid, kBlock, 0, stream>>>(S, ROUND_LS, 0, cfg.radius, cfg.sdf_threshold, back, round);
      NEFII_LAUNCH_CHECK();
      if ((rc = eval(2 * R, S.counters + round))) return rc;
      ++round;
    }
  }
  march_round_kernel<<<grid, kBlock, 0, stream>>>(S, ROUND_FINAL, cfg.sphere_tracing_iters == 0, cfg.radius, cfg.sdf_threshold, 0.f, round);
  NEFII_LAUNCH_CHECK();
  ++round;
  post_march_kernel<<<grid, kBlock, 0, stream>>>(S, training ? 1 : 0, want_min ? 1 : 0, c_samp, c_min);
  NEFII_LAUNCH_CHECK();

  // the single host sync of the trace: how many rays go to the sampler / to min-SDF sampling
  int h_counts[2] = {0, 0};
  NEFII_CUDA(cudaMemcpyAsync(h_counts, S.counters + c_samp, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
  NEFII_CUDA(cudaStreamSynchronize(stream));
  const int n_samp = h_counts[0], n_min = h_counts[1];
  const int chunk_rays = std::max(1, L.cap_pts / cfg.n_steps);

  if (n_samp > 0) {
    for (int begin = 0; begin < n_samp; begin += chunk_rays) {
      const int n = std::min(chunk_rays, n_samp - begin);
      const long long pts = (long long)n * cfg.n_steps;
      sample_emit_kernel<<<ceil_div(pts, kBlock), kBlock, 0, stream>>>(S, S.samp_list, begin, n, cfg.n_steps, linspace, 0);
      NEFII_LAUNCH_CHECK();
      if ((rc = eval((int)pts, nullptr))) return rc;
      sample_reduce_kernel<<<ceil_div(n, kBlock / 32), kBlock, 0, stream>>>(S, S.samp_list, begin, n, cfg.n_steps, linspace,
                                                                            training ? 1 : 0, c_root);
      NEFII_LAUNCH_CHECK();
    }
    // Rays with a bracketed root are often none at all (smooth scenes): a second, cheap host read of that count saves the
    // n_rootfind_steps x (8 layer launches + 3) empty launches of the bisection (~1 ms per trace) when there is nothing to refine.
    int n_root = 0;
    NEFII_CUDA(cudaMemcpyAsync(&n_root, S.counters + c_root, sizeof(int), cudaMemcpyDeviceToHost, stream));
    NEFII_CUDA(cudaStreamSynchronize(stream));
    bisect_init_kernel<<<ceil_div(n_samp, kBlock), kBlock, 0, stream>>>(S, c_root, work);
    NEFII_LAUNCH_CHECK();
    const int bgrid = ceil_div(n_samp, kBlock);
    for (int step = 0; step <= cfg.n_rootfind_steps && n_root > 0; ++step) {
      const int last = step == cfg.n_rootfind_steps;
      bisect_round_kernel<<<bgrid, kBlock, 0, stream>>>(S, step, c_root, c_bis0 + step, last, work);
      NEFII_LAUNCH_CHECK();
      if (!last && (rc = eval(n_samp, S.counters + c_bis0 + step))) return rc;
    }
  }
  if (want_min && n_min > 0) {
    for (int begin = 0; begin < n_min; begin += chunk_rays) {
      const int n = std::min(chunk_rays, n_min - begin);
      const long long pts = (long long)n * cfg.n_steps;
      sample_emit_kernel<<<ceil_div(pts, kBlock), kBlock, 0, stream>>>(S, S.min_list, begin, n, cfg.n_steps, uniforms, 1);
      NEFII_LAUNCH_CHECK();
      if ((rc = eval((int)pts, nullptr))) return rc;
      minsdf_reduce_kernel<<<ceil_div(n, kBlock / 32), kBlock, 0, stream>>>(S, S.min_list, begin, n, cfg.n_steps, uniforms);
      NEFII_LAUNCH_CHECK();
    }
  }
  if (stats) {
    int h[kMaxCounters];
    NEFII_CUDA(cudaMemcpyAsync(h, S.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    NEFII_CUDA(cudaStreamSynchronize(stream));
    long long evals = 0;
    for (int i = 0; i < n_rounds; ++i) evals += h[i];
    evals += (long long)n_samp * cfg.n_steps + (long long)(want_min ? n_min : 0) * cfg.n_steps;
    for (int i = 0; i < cfg.n_rootfind_steps; ++i) evals += h[c_bis0 + i];
    stats[0] = n_samp; stats[1] = h[c_root]; stats[2] = want_min ? n_min : 0; stats[3] = evals;
    stats[4] = h[0];
  }
  return NEFII_OK;
}

}  // namespace nefii
