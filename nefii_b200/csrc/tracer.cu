// IDR sphere tracing -- reference RayTracing (code/model/ray_tracing.py:29-337) and get_sphere_intersection
// (code/utils/rend_util.py:200-221) re-built as per-ray state machines with a device-driven schedule.
//
// The reference marches all rays in lock step with boolean-mask gathers and ~2 host syncs per iteration.  Here
//   * every ray carries its own state in HBM (SoA) and runs through ITS OWN sequence of iterations and line-search steps
//     (csrc/tracer_math.cuh: march_advance): a ray that needs no line search never waits for the rays that do, so a
//     trace takes max-over-rays(evaluations per ray) rounds instead of iters x (1 + line_step_iters);
//   * each round is one kernel that consumes the SDF values it asked for, advances the state machines and emits a
//     *compacted* list of the points it needs next (warp scan + one atomic per CTA); the SDF MLP (tcgen05 GEMM chain) runs
//     on exactly that list through a device-side row count -- dead lanes never reach the tensor cores;
//   * the sampler, bisection and min-SDF passes are sized on the device too (list lengths, chunk counters and the
//     batch-coupled "any ray still has work" flag live in a control block): the trace never synchronises with the host.
//     The whole trace is a fixed sequence of launches (empty rounds exit at once) or -- captured once per shape -- a CUDA
//     graph whose loops are conditional WHILE nodes driven by that control block (tracer_graph.cu).
//
// Accuracy tiers: evaluations that decide WHERE a ray stops (every march round, every bisection round) run the layer GEMM
// with one K block per TMEM partial (most accurate); the 100-sample scans of the sampler and of min-SDF sampling only pick
// brackets / arg-mins and run with the bulk setting (TraceTiers).
//
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true: every value that feeds a comparison is a rounded multiply
// followed by a rounded add, as torch evaluates it.
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "tracer.cuh"
#include "tracer_math.cuh"

namespace nefii {

namespace {

using namespace trm;

constexpr int kBlock = 256;

// device control block of one trace (ints; zeroed at the start of every trace)
struct Ctrl {
  int cnt_acc;       // requests accumulated by the kernel that is running
  int cnt_eval;      // rows of the next SDF evaluation
  int ticket;        // last-block detection
  int rounds;        // march rounds that had work
  int n_samp, n_min, n_root;   // list lengths
  int chunk;         // chunk of the sampler / min-SDF pass in flight
  int any_work;      // bisection: OR of the surviving work bits
  int bis_step;      // bisection iterations applied
  int bis_alive;     // the bisection loop is still running
  int bis_depth;     // bisection iterations applied per round (1..kMaxBisDepth; > 1 while few rays are refined)
  int bis_rewalk;    // -1, or the number of iterations of the LAST round that the reference executes (< bis_depth)
  int spec_work[4];  // OR of the surviving work bits after each of a round's iterations
  int march_spec;    // the requests in flight carry the line-search points of their step too (few rays: latency-bound rounds)
  int ends_acc, march_ends;   // marching ends that asked for a value: being counted / in the round in flight
  long long evals;   // SDF point evaluations (stats)
};

struct RayState {
  // inputs
  const float* cam_loc; const float* dirs; const unsigned char* object_mask; int n_pix; int n_rays;
  // per-ray state
  float *acc_s, *acc_e, *min_dis, *max_dis, *cur_s, *cur_e, *nxt_s, *nxt_e;
  int *slot_s, *slot_e;
  unsigned char *flags, *iter, *ls;
  // request list
  float* req_pts; float* req_sdf;
  Ctrl* ctrl;
  // sampler / root-find / min-sdf lists
  int *samp_list, *min_list, *root_list;
  float *z_lo, *z_hi, *s_lo, *s_hi;
  float *z_lo_prev, *z_hi_prev;   // bracket before the round in flight (speculative bisection)
  unsigned char* work;
  // outputs
  float* points; unsigned char* hit; float* dists;
};

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

// true in exactly one CTA of the grid: the one that finishes last (all other CTAs' global writes are visible to it)
__device__ __forceinline__ bool last_block(int* ticket) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    is_last = (t == (int)(gridDim.x * gridDim.y) - 1) ? 1 : 0;
    if (is_last) { *ticket = 0; __threadfence(); }
  }
  __syncthreads();
  return is_last != 0;
}

__device__ __forceinline__ void set_cond(unsigned long long handle, bool on) {
  if (handle != 0ull) cudaGraphSetConditional((cudaGraphConditionalHandle)handle, on ? 1u : 0u);
}

__device__ __forceinline__ void emit_point(const RayState& S, int slot, const float* o, float t, const float* d) {
  S.req_pts[(size_t)slot * 3 + 0] = o[0] + t * d[0];
  S.req_pts[(size_t)slot * 3 + 1] = o[1] + t * d[1];
  S.req_pts[(size_t)slot * 3 + 2] = o[2] + t * d[2];
}

// One round of the march.  first = 1: ray set-up (bounding-sphere intersection, both ends ask for their first SDF value);
// first = 0: consume the values of the previous request, run the ray's state machine, emit the next request.
// The last CTA publishes the request count (Ctrl::cnt_eval) and drives the WHILE node of the graph.
// Speculative line search (few requests in flight: a round is latency-bound -- 8 dependent layer GEMMs on a handful of row
// tiles): a ray that starts a new iteration asks for the SDF at its new position AND at the ls_iters positions the line search
// would step back to (ray_tracing.py:170-188: back by (1 - line_search_step) / 2^j of the step, one after the other), computed
// with the very operations march_advance applies.  The next round then runs the whole iteration -- step, up to ls_iters
// back-offs, head of the next iteration -- from values that are already there: same results, one round per iteration instead of
// one per evaluation.  Decided on the device per round from the previous round's request count (requests never increase).
constexpr int kMaxLineSearch = 3;

__global__ void __launch_bounds__(kBlock)
march_round_kernel(RayState S, int first, float radius, float thr, float back0, int ls_iters, int max_iters, int spec_rows,
                   unsigned long long cond) {
  Ctrl* C = S.ctrl;
  if (!first && C->cnt_eval == 0) return;   // nothing was asked for: every ray has left the march (uniform over the grid)
  const int spec_in = first ? 0 : C->march_spec;                                  // mode of the requests being consumed
  const int n_spec = 1 + ls_iters;                                                // values per speculative request
  const int spec_out = (!first && ls_iters >= 1 && ls_iters <= kMaxLineSearch && (long long)C->march_ends * n_spec <= spec_rows) ? 1 : 0;
  const int r = blockIdx.x * kBlock + threadIdx.x;
  const bool live = r < S.n_rays;
  float o[3] = {0, 0, 0}, d[3] = {0, 0, 0};
  March m;
  m.acc_s = m.acc_e = m.cur_s = m.cur_e = m.nxt_s = m.nxt_e = 0.f;
  m.flags = 0; m.iter = 0; m.ls = 0;
  int req = 0;
  bool touched = false;
  if (live) {
    if (first) {
      ray_od(S, r, o, d);
      float t0, t1;
      const bool hits = sphere_intersection(o, d, radius, t0, t1);
      req = march_begin(m, hits, t0, t1);
      S.min_dis[r] = m.acc_s; S.max_dis[r] = m.acc_e;
      touched = true;
    } else {
      m.flags = S.flags[r];
      if (m.flags & (F_PEND_S | F_PEND_E)) {
        ray_od(S, r, o, d);
        m.acc_s = S.acc_s[r]; m.acc_e = S.acc_e[r]; m.cur_s = S.cur_s[r]; m.cur_e = S.cur_e[r];
        m.nxt_s = S.nxt_s[r]; m.nxt_e = S.nxt_e[r];
        m.iter = S.iter[r]; m.ls = S.ls[r];
        const int base_s = S.slot_s[r], base_e = S.slot_e[r];
        // a request made at the start of an iteration (ls == 0, iter > 0) in speculative mode carries the back-off points
        const bool have_ls = spec_in && m.ls == 0 && m.iter > 0;
        if (m.flags & F_PEND_S) m.nxt_s = S.req_sdf[base_s];
        if (m.flags & F_PEND_E) m.nxt_e = S.req_sdf[base_e];
        m.flags &= ~(F_PEND_S | F_PEND_E);
        req = march_advance(m, thr, back0, ls_iters, max_iters);
        while (have_ls && req != 0 && m.ls > 0) {      // a line-search step whose value is already here: level m.ls
          if (req & REQ_S) m.nxt_s = S.req_sdf[base_s + m.ls];
          if (req & REQ_E) m.nxt_e = S.req_sdf[base_e + m.ls];
          req = march_advance(m, thr, back0, ls_iters, max_iters);
        }
        touched = true;
      }
    }
  }
  // points per requesting end: the position itself, plus (speculative mode, start of an iteration) its back-off positions
  const int per_end = (spec_out && req != 0 && m.ls == 0 && m.iter > 0) ? n_spec : 1;
  const int n_ends = ((req & REQ_S) ? 1 : 0) + ((req & REQ_E) ? 1 : 0);
  const int n_req = n_ends * per_end;
  int slot = reserve_slots(n_req, &C->cnt_acc);
  {
    const int block_ends = __syncthreads_count(n_ends >= 1) + __syncthreads_count(n_ends >= 2);
    if (threadIdx.x == 0 && block_ends) atomicAdd(&C->ends_acc, block_ends);
  }
  if (touched) {
    if (req & REQ_S) {
      float t = m.acc_s;
      emit_point(S, slot, o, t, d);
      for (int j = 1; j < per_end; ++j) { t = t - ldexpf(back0, -(j - 1)) * m.cur_s; emit_point(S, slot + j, o, t, d); }
      S.slot_s[r] = slot; m.flags |= F_PEND_S; slot += per_end;
    }
    if (req & REQ_E) {
      float t = m.acc_e;
      emit_point(S, slot, o, t, d);
      for (int j = 1; j < per_end; ++j) { t = t + ldexpf(back0, -(j - 1)) * m.cur_e; emit_point(S, slot + j, o, t, d); }
      S.slot_e[r] = slot; m.flags |= F_PEND_E;
    }
    S.flags[r] = m.flags; S.iter[r] = m.iter; S.ls[r] = m.ls;
    S.acc_s[r] = m.acc_s; S.acc_e[r] = m.acc_e; S.cur_s[r] = m.cur_s; S.cur_e[r] = m.cur_e;
    S.nxt_s[r] = m.nxt_s; S.nxt_e[r] = m.nxt_e;
  }
  if (last_block(&C->ticket) && threadIdx.x == 0) {
    const int n = *(volatile int*)&C->cnt_acc;
    C->cnt_acc = 0;
    C->cnt_eval = n;
    C->evals += n;
    C->march_spec = spec_out;
    C->march_ends = *(volatile int*)&C->ends_acc;
    C->ends_acc = 0;
    if (n > 0) C->rounds += 1;
    set_cond(cond, n > 0);
  }
}

// After sphere tracing: network mask, sampler list, (training) projection of sphere-missing rays and min-SDF list.
// The last CTA arms the sampler loop.
__global__ void __launch_bounds__(kBlock)
post_march_kernel(RayState S, int training, int want_min, unsigned long long cond) {
  Ctrl* C = S.ctrl;
  const int r = blockIdx.x * kBlock + threadIdx.x;
  if (r < S.n_rays) {
    float o[3], d[3];
    ray_od(S, r, o, d);
    unsigned char f = S.flags[r];
    float acc_s = S.acc_s[r];
    const float acc_e = S.acc_e[r];
    const bool hits = f & F_HIT;
    const bool net = acc_s < acc_e;
    const bool samp = f & F_UNF_S;
    const bool obj = S.object_mask ? (S.object_mask[r] != 0) : true;
    f &= ~(F_NET | F_SAMP | F_MIN | F_PEND_S | F_PEND_E);
    if (net) f |= F_NET;
    if (samp) {
      f |= F_SAMP;
      S.samp_list[atomicAdd(&C->n_samp, 1)] = r;
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
          S.min_list[atomicAdd(&C->n_min, 1)] = r;
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
  if (last_block(&C->ticket) && threadIdx.x == 0) {
    C->chunk = 0;
    C->cnt_eval = 0;
    set_cond(cond, *(volatile int*)&C->n_samp > 0);
  }
}

// n_steps sample points per listed ray of the chunk in flight: t_j = lo + frac_j * (hi - lo)  (sampler: frac = linspace,
// lo/hi = march interval; min-SDF: frac = shared uniforms, lo/hi = min_dis/max_dis).  Publishes the chunk's point count.
__global__ void __launch_bounds__(kBlock)
sample_emit_kernel(RayState S, const int* __restrict__ list, const int* __restrict__ n_list_ptr, int chunk_rays, int n_steps,
                   const float* __restrict__ frac, int use_minmax) {
  Ctrl* C = S.ctrl;
  const int begin = C->chunk * chunk_rays;
  const int n_list = max(0, min(chunk_rays, *n_list_ptr - begin));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->cnt_eval = n_list * n_steps;
    C->evals += (long long)n_list * n_steps;
  }
  const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i >= (long long)n_list * n_steps) return;
  const int k = (int)(i / n_steps), j = (int)(i % n_steps);
  const int r = list[begin + k];
  float o[3], d[3];
  ray_od(S, r, o, d);
  const float lo = use_minmax ? S.min_dis[r] : S.acc_s[r];
  const float hi = use_minmax ? S.max_dis[r] : S.acc_e[r];
  const float t = use_minmax ? (frac[j] * (hi - lo) + lo) : (lo + frac[j] * (hi - lo));
  emit_point(S, (int)i, o, t, d);
}

// next chunk / end of a chunked pass; `cond_next` arms the loop that follows the pass (value computed from *next_count)
__device__ __forceinline__ void finish_chunk(Ctrl* C, const int* n_list_ptr, int chunk_rays, unsigned long long cond) {
  if (last_block(&C->ticket) && threadIdx.x == 0) {
    const int c = C->chunk + 1;
    C->chunk = c;
    C->cnt_eval = 0;
    set_cond(cond, (long long)c * chunk_rays < (long long)*n_list_ptr);
  }
}

// One warp per sampled ray: first sign change / minimal SDF, candidate interval for the bisection.
__global__ void __launch_bounds__(kBlock)
sample_reduce_kernel(RayState S, int chunk_rays, int n_steps, const float* __restrict__ frac, int training, unsigned long long cond) {
  Ctrl* C = S.ctrl;
  const int begin = C->chunk * chunk_rays;
  const int n_list = max(0, min(chunk_rays, C->n_samp - begin));
  const int k = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k < n_list) {
    const int r = S.samp_list[begin + k];
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
    if (lane == 0) {
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
        const int slot = atomicAdd(&C->n_root, 1);
        S.root_list[slot] = r;
        S.z_lo[slot] = lo + frac[lo_i] * (hi - lo);
        S.z_hi[slot] = lo + frac[first] * (hi - lo);
        S.s_lo[slot] = v[lo_i];
        S.s_hi[slot] = v[first];
      }
    }
  }
  finish_chunk(C, &C->n_samp, chunk_rays, cond);
}

// Bisection (RayTracing.rootfind, ray_tracing.py:259-280).  The reference bisects *every* passed ray while ANY ray still has
// work, so the loop is batch-coupled: Ctrl::bis_alive says whether the next iteration runs.
//   phase INIT  : work = (s_lo > 0) & (s_hi < 0) & (z_hi > z_lo) (:261); emits the first mid-points when any ray has work
//   phase STEP  : applies the iterations of one round (:265-277), emits the next mid-points
//   finish      : z_pred = (z_lo + z_hi) / 2 -> dists / points
// Several iterations per round ("speculative" mode) when few rays are being refined: a round is latency-bound then (8
// dependent layer GEMMs on a handful of row tiles), so a round evaluates the whole binary tree of mid-points that D
// iterations can visit -- 2^D - 1 candidates per ray, node j of the tree (heap order, root 1, left child = the half that
// is kept when sdf(mid) <= 0) in request rows [(j - 1) n, j n) -- and applies D iterations of the reference at once: the
// same float operations on the same values, 1 / D of the rounds.  D is decided on the device at INIT from the number of
// rays being refined (largest D <= max_depth with (2^D - 1) n <= spec_rows).
// The batch-coupled stop inside a round: a round keeps the bracket it started from (z_*_prev) and records, per iteration,
// whether any ray still had work (Ctrl::spec_work); if the reference's loop ends after e < D iterations of the LAST round,
// the finish kernel re-walks e iterations from the saved bracket.  While the loop goes on, all D iterations were valid.
enum { BIS_INIT = 0, BIS_STEP = 1 };
constexpr int kMaxBisDepth = 4;

__device__ __forceinline__ void bisect_emit(const RayState& S, int k, int n_root, int depth, float zl, float zh) {
  const int r = S.root_list[k];
  float o[3], d[3];
  ray_od(S, r, o, d);
  float lo[1 << kMaxBisDepth], hi[1 << kMaxBisDepth];
  lo[1] = zl; hi[1] = zh;
#pragma unroll
  for (int j = 1; j < (1 << kMaxBisDepth); ++j) {
    if (j < (1 << depth)) {
      const float m = (lo[j] + hi[j]) * 0.5f;
      emit_point(S, (j - 1) * n_root + k, o, m, d);          // only evaluated if the loop goes on
      if (2 * j + 1 < (1 << kMaxBisDepth)) {
        lo[2 * j] = lo[j]; hi[2 * j] = m;                    // sdf(m) <= 0: the root lies in [lo, m]
        lo[2 * j + 1] = m; hi[2 * j + 1] = hi[j];            // sdf(m) >  0: in [m, hi]
      }
    }
  }
}

// `n_iter` iterations of the reference from bracket (zl, zh) with the SDF values of the round's candidate tree
__device__ __forceinline__ void bisect_walk(const RayState& S, int k, int n_root, int n_iter, float& zl, float& zh, float& sl, float& sh,
                                            unsigned char& w, Ctrl* C) {
  int node = 1;
  for (int i = 0; i < n_iter; ++i) {
    const float zm = (zl + zh) * 0.5f;
    const float sm = S.req_sdf[(size_t)(node - 1) * n_root + k];
    if (sm > 0.f) { zl = zm; sl = sm; node = 2 * node + 1; }
    if (sm <= 0.f) { zh = zm; sh = sm; node = 2 * node; }
    w = (w && ((zh - zl) > 1e-6f)) ? 1 : 0;
    if (C != nullptr && w) C->spec_work[i] = 1;
  }
}

__global__ void __launch_bounds__(kBlock)
bisect_kernel(RayState S, int phase, int n_rootfind_steps, int spec_rows, int max_depth, unsigned long long cond) {
  Ctrl* C = S.ctrl;
  if (phase != BIS_INIT && !C->bis_alive) return;          // uniform over the grid: the flag only changes in the last CTA
  const int n_root = C->n_root;
  int depth;
  if (phase == BIS_INIT) {
    depth = 1;
    for (int d = 2; d <= max_depth && d <= kMaxBisDepth; ++d)
      if (n_root > 0 && (long long)((1 << d) - 1) * n_root <= spec_rows) depth = d;
  } else {
    depth = C->bis_depth;
  }
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k < n_root) {
    float zl = S.z_lo[k], zh = S.z_hi[k];
    unsigned char w;
    if (phase == BIS_INIT) {
      w = (S.s_lo[k] > 0.f && S.s_hi[k] < 0.f && zh > zl) ? 1 : 0;
      if (w) C->any_work = 1;
    } else {
      S.z_lo_prev[k] = zl; S.z_hi_prev[k] = zh;
      float sl = S.s_lo[k], sh = S.s_hi[k];
      w = S.work[k];
      bisect_walk(S, k, n_root, depth, zl, zh, sl, sh, w, C);
      S.z_lo[k] = zl; S.z_hi[k] = zh; S.s_lo[k] = sl; S.s_hi[k] = sh;
    }
    S.work[k] = w;
    bisect_emit(S, k, n_root, depth, zl, zh);
  }
  if (last_block(&C->ticket) && threadIdx.x == 0) {
    bool alive;
    if (phase == BIS_INIT) {
      alive = (*(volatile int*)&C->any_work != 0) && 0 < n_rootfind_steps && n_root > 0;
      C->bis_step = 0;
      C->any_work = 0;
      C->bis_depth = depth;
      C->bis_rewalk = -1;
    } else {
      // how many of this round's iterations the reference's `while work.any() and i < n_steps` executes
      const int step0 = C->bis_step;
      int e = 0;
      alive = true;
      for (int i = 0; i < depth && alive; ++i) {
        e = i + 1;
        alive = (*(volatile int*)&C->spec_work[i] != 0) && (step0 + e < n_rootfind_steps);
      }
      C->bis_step = step0 + e;
      C->bis_rewalk = (e < depth) ? e : -1;
    }
    for (int i = 0; i < kMaxBisDepth; ++i) C->spec_work[i] = 0;
    C->bis_alive = alive ? 1 : 0;
    const int rows = ((1 << depth) - 1) * n_root;
    C->cnt_eval = alive ? rows : 0;
    if (alive) C->evals += rows;
    set_cond(cond, alive);
  }
}

__global__ void __launch_bounds__(kBlock) bisect_finish_kernel(RayState S, unsigned long long cond_next) {
  Ctrl* C = S.ctrl;
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const int n_root = C->n_root;
  if (k < n_root) {
    const int r = S.root_list[k];
    float o[3], d[3];
    ray_od(S, r, o, d);
    float zl = S.z_lo[k], zh = S.z_hi[k];
    const int e = C->bis_rewalk;
    if (e >= 0) {
      // the loop ended inside the last round: only its first e iterations count
      zl = S.z_lo_prev[k]; zh = S.z_hi_prev[k];
      float sl = 0.f, sh = 0.f;
      unsigned char w = 1;
      bisect_walk(S, k, n_root, e, zl, zh, sl, sh, w, nullptr);
    }
    const float zm = (zl + zh) * 0.5f;
    S.dists[r] = zm;
    S.points[(size_t)r * 3 + 0] = o[0] + zm * d[0];
    S.points[(size_t)r * 3 + 1] = o[1] + zm * d[1];
    S.points[(size_t)r * 3 + 2] = o[2] + zm * d[2];
  }
  if (last_block(&C->ticket) && threadIdx.x == 0) {
    C->chunk = 0;
    C->cnt_eval = 0;
    set_cond(cond_next, *(volatile int*)&C->n_min > 0);
  }
}

// minimal_sdf_points (ray_tracing.py:309-337): argmin over the n_steps shared random depths
__global__ void __launch_bounds__(kBlock)
minsdf_reduce_kernel(RayState S, int chunk_rays, int n_steps, const float* __restrict__ frac, unsigned long long cond) {
  Ctrl* C = S.ctrl;
  const int begin = C->chunk * chunk_rays;
  const int n_list = max(0, min(chunk_rays, C->n_min - begin));
  const int k = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k < n_list) {
    const int r = S.min_list[begin + k];
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
    if (lane == 0) {
      float o[3], d[3];
      ray_od(S, r, o, d);
      const float lo = S.min_dis[r], hi = S.max_dis[r];
      const float t = frac[best_j] * (hi - lo) + lo;
      S.dists[r] = t;
      S.points[(size_t)r * 3 + 0] = o[0] + t * d[0];
      S.points[(size_t)r * 3 + 1] = o[1] + t * d[1];
      S.points[(size_t)r * 3 + 2] = o[2] + t * d[2];
    }
  }
  finish_chunk(C, &C->n_min, chunk_rays, cond);
}

// Analytic test scene (oracle/tracer.py: analytic_sdf): union of spheres and boxes, fixed operation order.
__global__ void __launch_bounds__(kBlock)
analytic_sdf_kernel(const float* __restrict__ prims, int n_prims, int n, const int* __restrict__ count,
                    const float* __restrict__ x, float* __restrict__ out) {
  int limit = n;
  if (count) limit = min(limit, *count);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < limit; i += gridDim.x * kBlock) {
    const float p[3] = {x[(size_t)i * 3 + 0], x[(size_t)i * 3 + 1], x[(size_t)i * 3 + 2]};
    out[i] = analytic_sdf(prims, n_prims, p);
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Layout {
  size_t off_f[8], off_slot[2], off_bytes[4], off_req_pts, off_req_sdf, off_ctrl, off_lists[3], off_z[6], off_mlp, total;
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
  L.off_ctrl = take(sizeof(Ctrl));
  for (int i = 0; i < 4; ++i) L.off_bytes[i] = take(R);     // flags, iter, ls, work
  for (int i = 0; i < 8; ++i) L.off_f[i] = take(R * 4);
  for (int i = 0; i < 2; ++i) L.off_slot[i] = take(R * 4);
  L.off_req_pts = take((size_t)cap * 12);
  L.off_req_sdf = take((size_t)cap * 4);
  for (int i = 0; i < 3; ++i) L.off_lists[i] = take(R * 4);
  for (int i = 0; i < 6; ++i) L.off_z[i] = take(R * 4);
  L.off_mlp = p;
  if (src.net) p += align_up(src.net->workspace_bytes((int)cap, false), 256);
  L.total = p + 1024;
  return L;
}

// NEFII_TRACE_TIERS="march,bulk" overrides the defaults at load (A/B runs of whole programs)
TraceTiers env_tiers() {
  TraceTiers t;
  const char* e = getenv("NEFII_TRACE_TIERS");
  int a = 0, b = 0;
  if (e && sscanf(e, "%d,%d", &a, &b) == 2 && a >= 0 && a <= 64 && b >= 0 && b <= 64) { t.march_flush = a; t.bulk_flush = b; }
  return t;
}
TraceTiers g_tiers = env_tiers();
// bisection: D iterations per round while (2^D - 1) * (rays being refined) <= this many rows; 0 switches the mode off
// (NEFII_TRACE_QUAD_ROWS at load); D <= NEFII_TRACE_BISECT_DEPTH (1..4, default 4)
int env_quad_rows() {
  const char* e = getenv("NEFII_TRACE_QUAD_ROWS");
  const int v = e ? atoi(e) : 12288;
  return v < 0 ? 0 : v;
}
int g_bisect_quad_rows = env_quad_rows();
int env_bisect_depth() {
  const char* e = getenv("NEFII_TRACE_BISECT_DEPTH");
  const int v = e ? atoi(e) : kMaxBisDepth;
  return v < 1 ? 1 : (v > kMaxBisDepth ? kMaxBisDepth : v);
}
int g_bisect_max_depth = env_bisect_depth();

}  // namespace

int trace_set_tiers(int march_flush, int bulk_flush) {
  NEFII_CHECK_ARG(march_flush >= 0 && march_flush <= 64 && bulk_flush >= 0 && bulk_flush <= 64, "trace_set_tiers: out of range");
  g_tiers.march_flush = march_flush;
  g_tiers.bulk_flush = bulk_flush;
  return NEFII_OK;
}
TraceTiers trace_tiers() { return g_tiers; }
int trace_set_quad_rows(int rows) {
  NEFII_CHECK_ARG(rows >= 0, "trace_set_quad_rows: negative");
  g_bisect_quad_rows = rows;
  return NEFII_OK;
}
int trace_quad_rows() { return g_bisect_quad_rows; }
int trace_set_bisect_depth(int depth) {
  NEFII_CHECK_ARG(depth >= 1 && depth <= kMaxBisDepth, "trace_set_bisect_depth: 1..%d", kMaxBisDepth);
  g_bisect_max_depth = depth;
  return NEFII_OK;
}
int trace_bisect_depth() { return g_bisect_max_depth; }

size_t trace_workspace_bytes(const SdfSource& src, int n_rays, int n_steps) { return make_layout(src, n_rays, n_steps).total; }

int analytic_sdf_eval(cudaStream_t stream, const float* prims, int n_prims, int n, const int* count, const float* x, float* sdf) {
  if (n <= 0) return NEFII_OK;
  NEFII_CHECK_ARG(prims && n_prims > 0 && x && sdf, "analytic_sdf_eval: null pointer");
  int blocks = std::min(ceil_div(n, kBlock), kNumSMs * 16);
  analytic_sdf_kernel<<<blocks, kBlock, 0, stream>>>(prims, n_prims, n, count, x, sdf);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int trace_max_rounds(const TraceConfig& cfg) { return 1 + cfg.sphere_tracing_iters * (1 + cfg.line_step_iters); }

// Enqueues the whole trace on `stream`.  loops == nullptr: fixed schedule (every loop unrolled to its worst case; rounds
// without work exit at once).  loops != nullptr: the caller is capturing a CUDA graph and `loops` turns each loop into a
// conditional WHILE node (tracer_graph.cu).
int ray_trace_enqueue(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
                      const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
                      const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
                      unsigned char* hit, float* dists, TraceLoops* loops) {
  const long long n_rays_ll = (long long)n_batch * n_pix;
  NEFII_CHECK_ARG(n_batch >= 0 && n_pix >= 0 && n_rays_ll < (1ll << 30), "ray_trace: bad ray count");
  const int R = (int)n_rays_ll;
  if (R == 0) return NEFII_OK;
  NEFII_CHECK_ARG(cam_loc && ray_dirs && points && hit && dists && workspace && linspace, "ray_trace: null pointer");
  NEFII_CHECK_ARG(src.net != nullptr || (src.prims != nullptr && src.n_prims > 0), "ray_trace: no SDF source");
  NEFII_CHECK_ARG(cfg.n_steps >= 2 && cfg.n_steps <= 4096, "ray_trace: n_steps out of range");
  NEFII_CHECK_ARG(cfg.n_rootfind_steps >= 0 && cfg.n_rootfind_steps <= 4096, "ray_trace: n_rootfind_steps out of range");
  NEFII_CHECK_ARG(cfg.sphere_tracing_iters >= 0 && cfg.sphere_tracing_iters <= 250 && cfg.line_step_iters >= 0 &&
                      cfg.line_step_iters <= 30, "ray_trace: march iteration counts out of range");
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
  S.flags = (unsigned char*)(base + L.off_bytes[0]); S.iter = (unsigned char*)(base + L.off_bytes[1]);
  S.ls = (unsigned char*)(base + L.off_bytes[2]); S.work = (unsigned char*)(base + L.off_bytes[3]);
  S.req_pts = (float*)(base + L.off_req_pts); S.req_sdf = (float*)(base + L.off_req_sdf);
  S.ctrl = (Ctrl*)(base + L.off_ctrl);
  S.samp_list = (int*)(base + L.off_lists[0]); S.min_list = (int*)(base + L.off_lists[1]); S.root_list = (int*)(base + L.off_lists[2]);
  S.z_lo = (float*)(base + L.off_z[0]); S.z_hi = (float*)(base + L.off_z[1]); S.s_lo = (float*)(base + L.off_z[2]); S.s_hi = (float*)(base + L.off_z[3]);
  S.z_lo_prev = (float*)(base + L.off_z[4]); S.z_hi_prev = (float*)(base + L.off_z[5]);
  S.points = points; S.hit = hit; S.dists = dists;
  void* mlp_ws = base + L.off_mlp;
  const size_t mlp_ws_bytes = src.net ? src.net->workspace_bytes(L.cap_pts, false) : 0;
  int* cnt_eval = &S.ctrl->cnt_eval;

  // control block + the per-ray byte arrays (flags, iter, ls, work)
  NEFII_CUDA(cudaMemsetAsync(base + L.off_ctrl, 0, L.off_f[0] - L.off_ctrl, stream));

  const TraceTiers tiers = g_tiers;
  auto eval = [&](cudaStream_t st, int rows_cap, int flush) -> int {
    if (src.net) return src.net->eval(st, rows_cap, cnt_eval, S.req_pts, mlp_ws, mlp_ws_bytes, S.req_sdf, nullptr, nullptr, flush);
    return analytic_sdf_eval(st, src.prims, src.n_prims, rows_cap, cnt_eval, S.req_pts, S.req_sdf);
  };
  // run `body` as a device-driven loop (graph capture) or unrolled `max_trips` times (fixed schedule)
  auto loop = [&](int max_trips, auto&& body) -> int {
    int rc;
    if (loops) {
      cudaStream_t bs = nullptr;
      unsigned long long h = 0;
      if ((rc = loops->begin(stream, &bs, &h))) return rc;
      rc = body(bs, h);
      const int rc2 = loops->end(stream, bs);
      return rc ? rc : rc2;
    }
    for (int i = 0; i < max_trips; ++i)
      if ((rc = body(stream, 0ull))) return rc;
    return NEFII_OK;
  };
  // a conditional handle must exist before the kernel that arms it is captured
  auto next_cond = [&]() -> unsigned long long { return loops ? loops->next_handle(stream) : 0ull; };

  const int grid = ceil_div(R, kBlock);
  const float back0 = (float)(1.0 - (double)cfg.line_search_step);
  int rc;

  // ---- sphere tracing --------------------------------------------------------------------------------------------
  unsigned long long h_march = next_cond();
  // row budget of the speculative rounds (line search here, bisection below); 0 = off
  const int spec_rows = std::min(g_bisect_quad_rows, L.cap_pts);
  march_round_kernel<<<grid, kBlock, 0, stream>>>(S, 1, cfg.radius, cfg.sdf_threshold, back0, cfg.line_step_iters,
                                                  cfg.sphere_tracing_iters, spec_rows, h_march);
  NEFII_LAUNCH_CHECK();
  if ((rc = loop(trace_max_rounds(cfg), [&](cudaStream_t st, unsigned long long h) -> int {
        int rc2;
        if ((rc2 = eval(st, std::min(std::max(2 * R, spec_rows), L.cap_pts), tiers.march_flush))) return rc2;
        march_round_kernel<<<grid, kBlock, 0, st>>>(S, 0, cfg.radius, cfg.sdf_threshold, back0, cfg.line_step_iters,
                                                    cfg.sphere_tracing_iters, spec_rows, h);
        NEFII_LAUNCH_CHECK();
        return NEFII_OK;
      })))
    return rc;

  // ---- sampler: 100 uniform samples per unconverged ray, chunked ---------------------------------------------------
  const int chunk_rays = std::max(1, L.cap_pts / cfg.n_steps);
  const int max_chunks = ceil_div(R, chunk_rays);
  const long long chunk_pts = (long long)std::min(chunk_rays, R) * cfg.n_steps;
  unsigned long long h_samp = next_cond();
  post_march_kernel<<<grid, kBlock, 0, stream>>>(S, training ? 1 : 0, want_min ? 1 : 0, h_samp);
  NEFII_LAUNCH_CHECK();
  if ((rc = loop(max_chunks, [&](cudaStream_t st, unsigned long long h) -> int {
        int rc2;
        sample_emit_kernel<<<ceil_div(chunk_pts, kBlock), kBlock, 0, st>>>(S, S.samp_list, &S.ctrl->n_samp, chunk_rays, cfg.n_steps, linspace, 0);
        NEFII_LAUNCH_CHECK();
        if ((rc2 = eval(st, (int)chunk_pts, tiers.bulk_flush))) return rc2;
        sample_reduce_kernel<<<ceil_div(std::min(chunk_rays, R), kBlock / 32), kBlock, 0, st>>>(S, chunk_rays, cfg.n_steps, linspace,
                                                                                               training ? 1 : 0, h);
        NEFII_LAUNCH_CHECK();
        return NEFII_OK;
      })))
    return rc;

  // ---- bisection between the bracketing samples ----------------------------------------------------------------------
  unsigned long long h_bis = next_cond();
  // up to g_bisect_max_depth iterations per round while the candidate tree still fits a latency-bound evaluation (decided on
  // the device from n_root)
  const int max_depth = spec_rows > 0 ? g_bisect_max_depth : 1;
  bisect_kernel<<<grid, kBlock, 0, stream>>>(S, BIS_INIT, cfg.n_rootfind_steps, spec_rows, max_depth, h_bis);
  NEFII_LAUNCH_CHECK();
  if ((rc = loop(cfg.n_rootfind_steps, [&](cudaStream_t st, unsigned long long h) -> int {
        int rc2;
        if ((rc2 = eval(st, std::min(std::max(R, spec_rows), L.cap_pts), tiers.march_flush))) return rc2;
        bisect_kernel<<<grid, kBlock, 0, st>>>(S, BIS_STEP, cfg.n_rootfind_steps, spec_rows, max_depth, h);
        NEFII_LAUNCH_CHECK();
        return NEFII_OK;
      })))
    return rc;
  unsigned long long h_min = want_min ? next_cond() : 0ull;
  bisect_finish_kernel<<<grid, kBlock, 0, stream>>>(S, h_min);
  NEFII_LAUNCH_CHECK();

  // ---- min-SDF sampling (training) -------------------------------------------------------------------------------------
  if (want_min) {
    if ((rc = loop(max_chunks, [&](cudaStream_t st, unsigned long long h) -> int {
          int rc2;
          sample_emit_kernel<<<ceil_div(chunk_pts, kBlock), kBlock, 0, st>>>(S, S.min_list, &S.ctrl->n_min, chunk_rays, cfg.n_steps, uniforms, 1);
          NEFII_LAUNCH_CHECK();
          if ((rc2 = eval(st, (int)chunk_pts, tiers.bulk_flush))) return rc2;
          minsdf_reduce_kernel<<<ceil_div(std::min(chunk_rays, R), kBlock / 32), kBlock, 0, st>>>(S, chunk_rays, cfg.n_steps, uniforms, h);
          NEFII_LAUNCH_CHECK();
          return NEFII_OK;
        })))
      return rc;
  }
  return NEFII_OK;
}

// stats[0..4] = sampler rays, root-find rays, min-SDF rays, SDF point evaluations, march rounds with work (synchronises)
int trace_read_stats(cudaStream_t stream, const SdfSource& src, int n_rays, int n_steps, void* workspace, long long* stats) {
  if (!stats) return NEFII_OK;
  for (int i = 0; i < 8; ++i) stats[i] = 0;
  if (n_rays <= 0) return NEFII_OK;
  const Layout L = make_layout(src, n_rays, n_steps);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  Ctrl h;
  NEFII_CUDA(cudaMemcpyAsync(&h, base + L.off_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
  NEFII_CUDA(cudaStreamSynchronize(stream));
  stats[0] = h.n_samp; stats[1] = h.n_root; stats[2] = h.n_min; stats[3] = h.evals; stats[4] = h.rounds;
  return NEFII_OK;
}

}  // namespace nefii
