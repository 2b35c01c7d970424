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

size_t trace_workspace_bytes(const SdfSource& src, int n_rays, int n_steps);

// cam_loc [B,3], ray_dirs [B,P,3], object_mask [B*P] (u8) -> points [B*P,3], hit [B*P] (u8), dists [B*P].
// linspace: device [n_steps] = torch.linspace(0,1,n_steps); uniforms: device [n_steps] (training only).
// stats (host, optional, 8 ints): sampler rays, root-find rays, min-SDF rays, SDF point evaluations, ...
// Synchronises `stream` once (to size the sampler passes).
int ray_trace(cudaStream_t stream, const TraceConfig& cfg, const SdfSource& src, int n_batch, int n_pix,
              const float* cam_loc, const float* ray_dirs, const unsigned char* object_mask, int flags,
              const float* linspace, const float* uniforms, void* workspace, size_t ws_bytes, float* points,
              unsigned char* hit, float* dists, long long* stats);

// evaluate the analytic primitive table at n points (used by tests)
int analytic_sdf_eval(cudaStream_t stream, const float* prims, int n_prims, int n, const int* count, const float* x, float* sdf);

}  // namespace nefii
