// IDR sphere tracing (reference RayTracing, code/model/ray_tracing.py) -- declarations.
#pragma once
#include "common.cuh"
#include "sdf_mlp.cuh"

namespace nefii {

struct TraceConfig {
  float radius = 1.0f;             // object_bounding_sphere
  float sdf_threshold = 5.0e-5f;
  float line_search_step = 0.5f;
  int line_step_iters = 3;
  int sphere_tracing_iters = 10;
  int n_steps = 100;
  int n_rootfind_steps = 32;
};

// SDF provider: `net` (the MLP) or, for bit-exact control-flow tests, an analytic primitive table
// (device float [n_prims, 8]: kind(0 sphere / 1 box), centre xyz, radius|half extents, pad).
struct SdfSource {
  const SdfNet* net = nullptr;
  const float* prims = nullptr;
  int n_prims = 0;
};

enum TraceFlags : int {
  TRACE_TRAINING = 1,        // RayTracing.training: restrict root finding to object_mask rays, run min-SDF sampling
  TRACE_SKIP_MIN_SDF = 2,    // skip minimal_sdf_points (its outputs only reach lanes the caller discards)
};

// Accuracy tiers of the SDF evaluations inside a trace: K blocks (of 64) accumulated in TMEM before the partial sum moves
// to the fp32 register accumulators (see gemm_set_k_flush).  0 = the library default.
//   march_flush: march and bisection rounds -- their values decide where a ray stops;
//   bulk_flush : the n_steps-sample scans of the sampler and of min-SDF sampling -- they only select brackets / arg-mins.
// Both default to 0 since the truncation compensation of the layer GEMM (gemm_set_trunc_comp) made the longest partials as
// accurate as the shortest; the knobs remain for A/B measurements (NEFII_TRACE_TIERS="march,bulk").
struct TraceTiers {
  int march_flush = 0;
  int bulk_flush = 0;
};
int trace_set_tiers(int march_flush, int bulk_flush);
TraceTiers trace_tiers();
// bisection: two iterations per round (3 n rows per evaluation) while 3 * n_root <= rows; 0 = always one iteration per round
int trace_set_quad_rows(int rows);
int trace_quad_rows();
int trace_set_bisect_depth(int depth);
int trace_bisect_depth();

size_t trace_workspace_bytes(const SdfSource& src, int n_rays, int n_steps);
int trace_max_rounds(const TraceConfig& cfg);

// Hooks that turn the loops of a trace into conditional WHILE nodes while a CUDA graph is being captured (tracer_graph.cu).
struct TraceLoops {
  virtual ~TraceLoops() {}
  // creates the handle of the NEXT loop (the kernel that arms it is captured before the loop itself)
  virtual unsigned long long next_handle(cudaStream_t stream) = 0;
  // adds the WHILE node after what `stream` has captured so far; the body is captured on *body_stream
  virtual int begin(cudaStream_t stream, cudaStream_t* body_stream, unsigned long long* handle) = 0;
  virtual int end(cudaStream_t stream, cudaStream_t body_stream) = 0;
};

// cam_loc [B,3], ray_dirs [B,P,3], object_mask [B*P] (u8) -> points [B*P,3], hit [B*P] (u8), dists [B*P].
// linspace: device [n_steps] = torch.linspace(0,1,n_steps); uniforms: device [n_steps] (training only).
// Enqueues the trace on `stream`; never synchronises.  loops == nullptr: fixed launch schedule.
int ray_trace_enqueue(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
                      const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
                      const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
                      unsigned char* hit, float* dists, TraceLoops* loops);

// The same trace through the graph cache: the first call with a given argument tuple captures the launch sequence into a
// CUDA graph (loops as conditional WHILE nodes), later calls replay it.  stats (host, optional, 8 ints: sampler rays,
// root-find rays, min-SDF rays, SDF point evaluations, march rounds) synchronises the stream; without it there is no sync.
int ray_trace(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
              const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
              const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
              unsigned char* hit, float* dists, long long* stats);

int trace_read_stats(cudaStream_t stream, const SdfSource& src, int n_rays, int n_steps, void* workspace, long long* stats);

// 0: fixed launch schedule, 1: CUDA graph with conditional WHILE nodes (default; NEFII_TRACE_GRAPH overrides at load)
int trace_set_graph_mode(int mode);
int trace_graph_mode();
long long trace_graph_captures();
// drops every cached graph (call when a workspace or a network the graphs point into is freed)
int trace_graph_clear();

// evaluate the analytic primitive table at n points (used by tests)
int analytic_sdf_eval(cudaStream_t stream, const float* prims, int n_prims, int n, const int* count, const float* x, float* sdf);

}  // namespace nefii
