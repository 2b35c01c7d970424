/* nefii_b200 -- C ABI of the B200-native NeFII rendering hot path.
 *
 * The reference (FuxiComputerVision/Nefii) has no native code and no FFI: its seam is Python
 * (SURVEY.md section 8b).  Every entry point below replaces a reference *function* on the hot path; the
 * reference file:line it stands in for is cited next to it.  INTEGRATION.md shows the ctypes stub
 * a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked `host`;
 *     all tensors are contiguous row-major float32 unless noted; masks are uint8 (0/1).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); no call synchronises the
 *     host unless stated.
 *   - the caller owns every buffer; the library keeps nothing past the call except inside a
 *     handle created by nefii_create (workspace + TMA descriptors), released by nefii_destroy.
 *   - return value: 0 ok, <0 error (-1 bad argument, -2 CUDA error, -3 bad state); the text is
 *     available from nefii_last_error() (thread-local).  Nothing throws, nothing exits.
 */
#ifndef NEFII_B200_H_
#define NEFII_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* nefii_last_error(void);
/* ABI version of this header (bumped on any signature change) */
int nefii_abi_version(void);
/* number of CUDA kernels this library has launched since load (bench.py reports the delta per timed region) */
int64_t nefii_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * SG shading -- replaces render_with_sg, code/model/sg_render.py:164-295 (forward).
 *   lgt_sgs [M,7] raw parameter (axis, sharpness, rgb amplitude; abs()/normalise applied inside)
 *   specular [K,3], roughness [K,1], albedo/normal/view [N,3] (unit normals / view dirs),
 *   blending [N,K] or NULL.  Outputs [N,3]: sg_rgb, sg_specular_rgb, sg_diffuse_rgb.
 * ------------------------------------------------------------------------------------------- */
int nefii_sg_render_fwd(void* stream, int n_rays, int n_sg, int n_mat,
                        const float* lgt_sgs, const float* specular, const float* roughness,
                        const float* albedo, const float* normal, const float* view,
                        const float* blending,
                        float* out_rgb, float* out_specular, float* out_diffuse);

/* Backward of nefii_sg_render_fwd (what autograd does through sg_render.py:164-295).  out_specular / out_diffuse: the forward
 * outputs (the reference clamps the summed radiance at 0; the gradient passes where they are positive); g_rgb / g_specular /
 * g_diffuse: upstream gradients [N,3] (any may be NULL).  ACCUMULATED outputs (zero them first): g_lgt_acc [M,7] in the unit
 * parametrisation (convert with nefii_sg_param_grad, eps 1e-6), g_roughness [K], g_specular_refl [K,3]; written: g_albedo [N,3]
 * and, when not NULL, g_normal [N,3] (gradient w.r.t. the `normal` input as given, i.e. after the caller's normalisation).
 * blending: the forward's per-point weights [N,K] or NULL; g_blending [N,K] (written) or NULL.
 * Hand-derived adjoint of the closed-form SG integrals (csrc/sg_adjoint_math.cuh).  View directions carry no gradient. */
int nefii_sg_render_bwd(void* stream, int n_rays, int n_sg, int n_mat, const float* lgt_sgs, const float* specular,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* out_specular, const float* out_diffuse, const float* g_rgb, const float* g_specular,
                        const float* g_diffuse, float* g_lgt_acc, float* g_roughness, float* g_specular_refl, float* g_albedo,
                        float* g_normal, const float* blending, float* g_blending);

/* Environment radiance along miss rays -- replaces IDRNetwork.get_background_rgb (light_type 'sg'),
 * code/model/implicit_differentiable_renderer.py:646-663 + sg_fn path_tracing_render.py:404-413. */
int nefii_background_sg_fwd(void* stream, int n_rays, int n_sg,
                            const float* lgt_sgs, const float* dirs, float* out_rgb);

/* ---------------------------------------------------------------------------------------------
 * MLP layer GEMM on tcgen05 -- the building block that replaces nn.Linear (+ Softplus/ReLU/ELU) in
 * ImplicitNetwork.forward/.gradient (implicit_differentiable_renderer.py:85-123),
 * RenderingNetwork.forward (:196-241) and EnvmapMaterialNetwork.forward (sg_envmap_material.py:357-425).
 * fp32 values travel as two 16-bit planes (hi, lo); see csrc/mlp_gemm.cu.  Plane format `fmt` of a launch: 0 = bf16 split
 * (hi = bf16(x), lo = bf16(x - hi): 16 significant bits, fp32's exponent range; the trainable stacks and every backward pass),
 * 1 = fp16 split (22 significant bits for |x| >= 2^-3, absolute error <= 2^-25 below, |x| < 65504; the SDF network's
 * inference chain).  Three kind::f16 MMAs per product either way.
 * The struct is a host-side argument block (all pointers inside are device pointers).
 * ------------------------------------------------------------------------------------------- */
typedef struct nefii_gemm_desc {
  const void* a_hi; const void* a_lo; int32_t a_ld; int32_t rows_cap;   /* activations [rows_cap, a_ld] bf16 planes */
  const void* b_hi; const void* b_lo; int32_t b_ld; int32_t n_pad;      /* weights [n_pad, b_ld] bf16 planes, n_pad % 256 == 0 */
  int32_t k_pad;                  /* reduction length, multiple of 64 */
  const int32_t* count;           /* device int: valid rows (NULL = rows_cap) */
  int32_t mode;                   /* 0 forward (bias + act), 1 backward (x act'(saved forward output)) */
  int32_t act;                    /* 0 none, 1 softplus(beta=100), 2 relu, 3 elu */
  int32_t n_valid;                /* real output columns */
  const float* bias;              /* [n_valid] or NULL */
  float out_scale;
  void* dst_hi; void* dst_lo; int32_t dst_ld; int32_t dst_col0; int32_t dst_ncols; int32_t dst_zero_to; /* output planes (may be NULL); cols [dst_ncols,dst_zero_to) zero-filled */
  float* dst_f32; int32_t f32_ld; int32_t f32_begin; int32_t f32_end;                /* optional fp32 output columns */
  const float* w_last; const float* b_last; int32_t n_last; int32_t w_last_ld; float* dst_last; /* fused tiny output layer */
  void* seed_hi; void* seed_lo; int32_t seed_ld;                                     /* input-gradient seed planes */
  const void* sav_hi; const void* sav_lo; int32_t sav_ld; int32_t sav_ncols; float sav_scale; /* backward: saved activations */
  int32_t k_splits; int64_t f32_split_stride; int32_t k_splits_used;   /* split-K: partial s -> dst_f32 + s*stride (floats); k_splits_used is an output */
  int32_t k_flush;                /* K blocks per TMEM partial for this launch (1 = most accurate); 0 = library default */
  int32_t dst_pad_ok;             /* 1: a ragged last 128-column span stays on the bulk-store epilogue; plane columns >= dst_ncols are never written */
  int32_t fmt;                    /* plane format of a, b, dst, seed and sav: 0 bf16 split, 1 fp16 split */
  /* "PE prologue": pe_x != NULL -> row m of A is the positional encoding of pe_x[m] (embedder.py:22-36: [x, sin(2^0 x), cos(2^0 x),
   * ...], 3 + 6 pe_n_freqs <= 64 columns), computed inside the kernel straight into the shared-memory operand tile; a_hi / a_lo
   * are ignored and k_pad must be 64.  pe_side_*: optional planes that receive the encoding times pe_side_scale at columns
   * [pe_side_col0, pe_side_col0 + 3 + 6 pe_n_freqs) (the PE half of a skip layer's input). */
  const float* pe_x; int32_t pe_n_freqs; void* pe_side_hi; void* pe_side_lo; int32_t pe_side_ld; int32_t pe_side_col0; float pe_side_scale;
} nefii_gemm_desc;

int nefii_gemm_split_bf16(void* stream, nefii_gemm_desc* desc /* host */);

/* Measurement aid (bench.py): `blocks` x 256 threads run `iters` rounds of 8 independent FP32 FMAs each
 * (flops = blocks * 256 * iters * 16); timed by the caller with CUDA events -> the FP32 peak `fp32_fraction` is quoted against */
int nefii_probe_fp32(void* stream, int blocks, int iters, float* sink);

/* Measurement aid (bench.py roofline): while enabled, every nefii layer-GEMM launch is bracketed by CUDA events on its
 * own stream.  fetch() synchronises those events and returns {total ms, total algorithmic flops (2*rows*n*k, each
 * fp32 product counted once), number of launches}. */
int nefii_gemm_profile_enable(int on);
/* tuning knob: 1 = single-CTA layer GEMM, 2 = CTA pairs (tcgen05 cta_group::2: two SMs run one 256-row MMA and share the
 * weight tile through each other's shared memory; default).  The environment variable NEFII_GEMM_CLUSTER sets the default at load. */
int nefii_gemm_set_cluster(int cluster_size);
/* development only: disables pieces of the GEMM pipeline (1 epilogue math, 2 TMA loads, 4 MMAs, 8 TMEM flush, 16 plane staging +
 * stores, 32 plane stores) for timing experiments; results are then garbage */
int nefii_gemm_set_debug(int mask);
/* programmatic dependent launch of the layer GEMMs (a launch's prologue overlaps the tail of the previous kernel in the stream;
 * default on, NEFII_GEMM_PDL=0 switches it off at load) */
int nefii_gemm_set_pdl(int on);
/* upper bound on the persistent grid (SMs) of the layer-GEMM launches that follow, 0 = every SM (default).  For a caller that runs
 * a bulk evaluation on one stream next to a latency-bound chain on another (IDRNetwork.prefetch_trace: the primary trace of the next
 * batch beside the shading of this one; the reference runs them back to back, implicit_differentiable_renderer.py:343-349): persistent
 * CTAs keep their SM for the whole kernel, the bound leaves the other stream SMs to start on.  Part of the trace-graph cache key. */
int nefii_gemm_set_grid_cap(int sms);
/* accuracy / overlap knob of the layer GEMM: 64-wide K blocks accumulated inside TMEM before the partial sum moves to the fp32
 * register accumulators (1 = most accurate; default 4, NEFII_GEMM_KFLUSH sets the default at load) */
int nefii_gemm_set_k_flush(int k_blocks);
/* ... for the first two partial sums of every 256-column chunk only (set_k_flush resets it to the same value) */
int nefii_gemm_set_k_flush_head(int k_blocks);
/* first-order compensation of the tensor core's round-toward-zero accumulation: TMEM partial sums of `k_blocks` K blocks are
 * scaled by (1 + rho) when they are added to the fp32 register accumulators; rho = 0 (default): plain sum */
int nefii_gemm_set_trunc_comp(int k_blocks, float rho);
/* ... for the partial sums of launches in plane format `fmt` (nefii_gemm_set_trunc_comp sets the bf16 table) */
int nefii_gemm_set_trunc_comp_fmt(int fmt, int k_blocks, float rho);
int nefii_gemm_profile_fetch(double* out3 /* host */);

/* fp32 [rows, cols] (row stride ld_src) -> zero-padded bf16 hi/lo planes [rows_pad, cols_pad];
 * transpose != 0 writes the transpose.  Used to pack weights (and test inputs). */
int nefii_split_to_planes(void* stream, const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                          void* dst_hi, void* dst_lo, int rows_pad, int cols_pad);
/* ... into planes of format `fmt` (0 bf16 split = nefii_split_to_planes, 1 fp16 split) */
int nefii_split_to_planes_fmt(void* stream, const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                              void* dst_hi, void* dst_lo, int rows_pad, int cols_pad, int fmt);

/* ---------------------------------------------------------------------------------------------
 * SDF / feature MLP -- replaces ImplicitNetwork.forward and .gradient(x, no_grad=True),
 * code/model/implicit_differentiable_renderer.py:85-123 (+ embedder.py:5-50).
 * The handle owns the packed (bf16 hi/lo, transposed) copy of the weights; the caller owns the
 * workspace and every input / output buffer.
 * ------------------------------------------------------------------------------------------- */
typedef struct nefii_sdf_config {
  int32_t d_in;        /* 3 */
  int32_t n_freqs;     /* multires (6) */
  int32_t width;       /* hidden width == feature_vector_size (512) */
  int32_t n_hidden;    /* len(dims) (8) */
  int32_t skip_layer;  /* skip_in[0] (4); <= 0: none */
  int32_t d_out;       /* 1 */
  int32_t d_feat;      /* 0: feature vector = input of the last layer (use_last_as_f, conf.conf); > 0: the last Linear has
                          1 + d_feat outputs, rows 1.. are the feature vector (use_last_as_f = False, conf_neus.conf) */
} nefii_sdf_config;

int nefii_sdf_create(void** handle, const nefii_sdf_config* cfg /* host */);
int nefii_sdf_destroy(void* handle);
/* Plane format of the network's inference chain (weights, activations, gradient chain): 1 = fp16 split (default; the SDF value
 * then agrees with an fp32 evaluation to ~2e-7, which is what the depth / shading parity with the reference rests on),
 * 0 = bf16 split (~2e-6; for networks whose activations or input gradients could exceed fp16's range, |x| < 65504).
 * NEFII_SDF_FORMAT=bf16|fp16 sets the default at load.  Call before nefii_sdf_set_weights (it drops the packed weights). */
int nefii_sdf_set_format(void* handle, int fmt);
int nefii_sdf_get_format(void* handle);
/* Where an evaluation's positional encoding is computed: 0 (default) one encode kernel writes layer 0's input planes and the PE
 * half of the skip layer's input in a single pass; 1 inside layer 0's GEMM kernel ("PE prologue": two otherwise idle warps
 * write the encoding straight into the shared-memory operand tile -- no launch, no round trip, but measured SLOWER on B200
 * because those two warps are latency-bound on the tile's 2 304 sincosf; NEFII_SDF_PE_PROLOGUE=1 at load). */
int nefii_sdf_set_pe_prologue(int on);
/* weights / biases: host arrays of n_hidden+1 device pointers; weights[l] is the EFFECTIVE fp32 matrix
 * [out_l, in_l] (weight_norm folded: g * v / |v|), row-major contiguous; the last one is [1 + d_feat, width].
 * Call again whenever the parameters change. */
int nefii_sdf_set_weights(void* handle, void* stream, const float* const* weights, const float* const* biases);
int64_t nefii_sdf_workspace_bytes(void* handle, int rows_cap, int with_grad);
/* x [rows_cap,3]; count: device int32 with the number of valid rows or NULL; sdf [rows_cap];
 * feat [rows_cap, d_feat > 0 ? d_feat : width] or NULL; grad [rows_cap,3] or NULL (d sdf / d x).
 * k_flush: accuracy tier of the layer GEMMs (K blocks per TMEM partial, 1 = most accurate); 0 = library default. */
int nefii_sdf_eval(void* handle, void* stream, int rows_cap, const int32_t* count, const float* x,
                   void* workspace, int64_t workspace_bytes, float* sdf, float* feat, float* grad, int k_flush);

/* ---------------------------------------------------------------------------------------------
 * Sphere tracing -- replaces RayTracing.forward, code/model/ray_tracing.py:29-101 (sphere_tracing
 * :104-193, ray_sampler :195-257, rootfind :259-280, minimal_sdf_points :309-337) and
 * rend_util.get_sphere_intersection, code/utils/rend_util.py:200-221.
 * The reference's `sdf` callable argument becomes an SDF source: the MLP handle (sdf_kind 0) or an
 * analytic primitive table used by the bit-exact control-flow tests (sdf_kind 1: device float
 * [n_prims, 8] rows = kind (0 sphere, 1 box), centre xyz, radius | half extents, pad).
 * ------------------------------------------------------------------------------------------- */
typedef struct nefii_trace_config {
  float object_bounding_sphere;   /* 1.0 */
  float sdf_threshold;            /* 5e-5 */
  float line_search_step;         /* 0.5 */
  int32_t line_step_iters;        /* 3 */
  int32_t sphere_tracing_iters;   /* 10 */
  int32_t n_steps;                /* 100 */
  int32_t n_rootfind_steps;       /* 32 */
} nefii_trace_config;

#define NEFII_TRACE_TRAINING 1      /* RayTracing.training */
#define NEFII_TRACE_SKIP_MIN_SDF 2  /* skip minimal_sdf_points (outputs only reach lanes the caller masks out) */

int64_t nefii_trace_workspace_bytes(int sdf_kind, const void* sdf, int n_rays, int n_steps);
/* cam_loc [B,3], ray_dirs [B,P,3], object_mask [B*P] uint8 (NULL = all true);
 * linspace: device [n_steps] = linspace(0,1,n_steps); uniforms: device [n_steps] U(0,1) draws shared by
 * all rays (training only; the reference draws them on the CPU generator, ray_tracing.py:316).
 * Outputs: points [B*P,3], hit [B*P] uint8 (network_object_mask), dists [B*P].
 * stats: host int64[8] or NULL: {sampler rays, root-find rays, min-SDF rays, SDF point evaluations, march rounds, ...}.
 * The trace never waits for the host: every loop of the reference (march iterations, sampler chunks, bisection steps,
 * min-SDF chunks) is sized by a device-side control block.  Only stats != NULL synchronises `stream` (to read that block).
 * In graph mode (default) the launch sequence is captured once per argument tuple into a CUDA graph whose loops are
 * conditional WHILE nodes and replayed by later calls: keep the buffers at the same addresses to hit the cache. */
int nefii_ray_trace(void* stream, const nefii_trace_config* cfg, int sdf_kind, const void* sdf, int n_prims,
                    int n_batch, int n_pix, const float* cam_loc, const float* ray_dirs, const uint8_t* object_mask,
                    int flags, const float* linspace, const float* uniforms, void* workspace, int64_t workspace_bytes,
                    float* points, uint8_t* hit, float* dists, int64_t* stats);
/* accuracy tiers of the SDF evaluations inside a trace (K blocks per TMEM partial, 0 = library default): march_flush for the
 * march and bisection rounds (they decide where a ray stops; default 1), bulk_flush for the n_steps-sample scans of the
 * sampler and of min-SDF sampling (they only select brackets / arg-mins; default 0) */
int nefii_trace_set_tiers(int march_flush, int bulk_flush);
/* 0: fixed launch schedule (every loop unrolled to its worst case, empty rounds exit at once); 1: CUDA graph with
 * conditional WHILE nodes (default; NEFII_TRACE_GRAPH=0 selects the fixed schedule at load) */
int nefii_trace_set_graph_mode(int mode);
int nefii_trace_graph_mode(void);   /* the mode in force */
/* trace graphs captured so far in this process (a capture + instantiate costs tens of milliseconds: callers that time a loop
 * check that the count did not move inside it) */
int64_t nefii_trace_graph_captures(void);
/* Speculative rounds while few rays are in flight (a round is then latency-bound: 8 dependent layer GEMMs on a handful of row
 * tiles).  Sphere tracing: a ray that starts an iteration also asks for the SDF at the positions its line search would step back
 * to (1 + line_step_iters points per marching end), so the next round runs the whole iteration; used while
 * (1 + line_step_iters) * (ends in flight) <= rows.  Bisection:
 * bisection rounds apply D iterations of the reference's loop at once while few rays are refined: the whole binary tree of
 * mid-points that D iterations can visit is evaluated in one round ((2^D - 1) n rows instead of n, bit-identical results,
 * 1 / D of the latency-bound rounds); D = the largest depth <= max depth with (2^D - 1) * n_root <= rows, decided on the
 * device.  rows = 0 switches it off (default 12288; NEFII_TRACE_QUAD_ROWS at load); max depth 1..4 (default 4;
 * NEFII_TRACE_BISECT_DEPTH at load). */
int nefii_trace_set_quad_rows(int rows);
int nefii_trace_set_bisect_depth(int depth);
/* drops the cached trace graphs (they hold raw pointers into workspaces and SDF handles) */
int nefii_trace_graph_clear(void);
/* the analytic test SDF alone: x [n,3] -> sdf [n] */
int nefii_analytic_sdf_eval(void* stream, const float* prims, int n_prims, int n, const float* x, float* sdf);

/* ---------------------------------------------------------------------------------------------
 * Ray set-up -- replaces rend_util.get_camera_params + lift, code/utils/rend_util.py:90-142 (pose-matrix form; the [B,7]
 * quaternion form is converted to a matrix by the caller, :91-96).  uv [B,P,2], pose [B,4,4] camera-to-world, intrinsics
 * [B,4,4] -> dirs [B,P,3] (unit), cam_loc [B,3] (may be NULL).  order: summation order of the 4-term products of
 * torch.bmm(pose, cam_points) to reproduce (0: FMA chain, 1: separate multiply / add).
 * ------------------------------------------------------------------------------------------- */
int nefii_camera_rays(void* stream, int n_batch, int n_pix, const float* uv, const float* pose, const float* intrinsics,
                      int order, float* dirs, float* cam_loc);

/* ---------------------------------------------------------------------------------------------
 * Differentiable intersection -- replaces SampleNetwork.forward, code/model/sample_network.py:10-24 (IDR eq. 3):
 *   x = c + (t0 - (s - s0) / (grad . v0)) v,   |grad . v0| < 1e-8 -> 1e-8.
 * surface_output / surface_sdf_values / surface_dists [n] (the reference's [n,1]), the others [n,3].
 * bwd: g_points [n,3] -> gradients w.r.t. every input (any output pointer may be NULL).
 * ------------------------------------------------------------------------------------------- */
int nefii_sample_network_fwd(void* stream, int n, const float* surface_output, const float* surface_sdf_values,
                             const float* surface_points_grad, const float* surface_dists, const float* surface_cam_loc,
                             const float* surface_ray_dirs, float* out_points);
int nefii_sample_network_bwd(void* stream, int n, const float* surface_output, const float* surface_sdf_values,
                             const float* surface_points_grad, const float* surface_dists, const float* surface_ray_dirs,
                             const float* g_points, float* g_output, float* g_sdf_values, float* g_dists, float* g_cam_loc,
                             float* g_ray_dirs, float* g_points_grad);

/* ---------------------------------------------------------------------------------------------
 * Near-field indirect-illumination integrator -- replaces the sampling and shading halves of
 * pt_render_indirect_mlp, code/model/path_tracing_render.py:1255-1487 (cos_sampling :128-156,
 * brdf_sampling :61-125, mix_sg_sampling :168-271, power_heuristic_list :390-401, shading :1406-1476).
 * The secondary trace between the two halves is nefii_ray_trace; the radiance query is the MLP path.
 * n = surface points; all per-sample arrays are laid out [3 (cos, ggx, mixture), n, ...].
 * ------------------------------------------------------------------------------------------- */
/* u [n,7]: the uniforms in the reference's torch.rand order (cos r1 r2, ggx r1 r2, mix r0 r1 r2).
 * Outputs: wi [3,n,3], pdf [3,n] (clamped at 1e-6), weight [3,n] (power heuristic),
 * pdf_matrix [3,3,n] or NULL (pdf of strategy j at direction i). */
int nefii_mis_sample(void* stream, int n, int n_sg, const float* lgt_sgs, const float* roughness, const float* normal,
                     const float* view, const float* u, float* wi, float* pdf, float* weight, float* pdf_matrix);
/* specular: [1,3] (spec_per_point = 0) or [n,3]; hit [3,n] uint8 (secondary_mask); indirect [3,n,3].
 * Outputs [n,3]: sg_rgb, sg_specular_rgb, sg_diffuse_rgb; light [3,n,3] (environment radiance along wi,
 * kept for the backward pass) or NULL. */
int nefii_mis_shade_fwd(void* stream, int n, int n_sg, const float* lgt_sgs, const float* specular, int spec_per_point,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* wi, const float* pdf, const float* weight, const uint8_t* hit,
                        const float* indirect, float* out_rgb, float* out_specular, float* out_diffuse, float* light);
/* g_rgb / g_specular / g_diffuse: upstream gradients [n,3] (any may be NULL).  Outputs: g_roughness [n],
 * g_albedo [n,3], g_specular_refl [n,3] or NULL, g_indirect [3,n,3]; g_lgt_acc [n_sg,7] is ACCUMULATED
 * (atomicAdd) in the unit parametrisation {d axis, d sharpness, d amplitude} -- convert with nefii_sg_param_grad;
 * g_normal [n,3] or NULL: d / d normal (needed when the geometry trains: the normal is d sdf/dx with a graph). */
int nefii_mis_shade_bwd(void* stream, int n, int n_sg, const float* lgt_sgs, const float* specular, int spec_per_point,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* wi, const float* pdf, const float* weight, const uint8_t* hit,
                        const float* indirect, const float* light, const float* g_rgb, const float* g_specular,
                        const float* g_diffuse, float* g_roughness, float* g_albedo, float* g_specular_refl,
                        float* g_indirect, float* g_lgt_acc, float* g_normal);
/* backward of nefii_background_sg_fwd: accumulates into g_lgt_acc [n_sg,7] (unit parametrisation, eps 1e-8) */
int nefii_background_sg_bwd(void* stream, int n_rays, int n_sg, const float* lgt_sgs, const float* dirs,
                            const float* g_out, float* g_lgt_acc);
/* unit-parametrisation accumulator -> gradient of the raw lgtSGs parameter (abs(), lobe / (|lobe| + eps));
 * accumulate != 0 adds to g_lgt instead of overwriting */
int nefii_sg_param_grad(void* stream, int n_sg, const float* lgt_sgs, const float* acc, float eps, float* g_lgt, int accumulate);

/* Helpers of the trainable dense stacks (RenderingNetwork :196-241, EnvmapMaterialNetwork :357-425):
 * input assembly: up to 4 segments, each a positional encoding (n_freqs >= 0) of a 3-vector or a raw copy
 * (n_freqs = -1) of `width` floats per row, concatenated and zero padded to k_pad -> bf16 planes [rows, ld] */
int nefii_assemble_input(void* stream, int rows, int n_seg, const float* const* src /* host array of device ptrs */,
                         const int32_t* width /* host */, const int32_t* n_freqs /* host */, void* dst_hi, void* dst_lo,
                         int ld, int k_pad);
/* dst[c][r] = src[r][c] as planes, zero padded to [cols_pad, rows_pad] (row stride ld_dst); col_sum (or NULL)
 * accumulates the column sums of src (bias gradient) */
int nefii_transpose_planes(void* stream, const void* src_hi, const void* src_lo, int ld_src, int rows, int cols,
                           void* dst_hi, void* dst_lo, int ld_dst, int rows_pad, int cols_pad, float* col_sum);
/* backward of the fused tiny output layer y = act(z) W_last^T + b_last: G = (gy W_last) * act'(h) as planes,
 * gw_last [n_out,width] and gb_last [n_out] are ACCUMULATED */
int nefii_last_layer_bwd(void* stream, int act, int rows, int width, int n_out, const float* gy, const float* w_last,
                         const void* h_hi, const void* h_lo, int h_ld, void* g_hi, void* g_lo, int g_ld,
                         float* gw_last, float* gb_last);
/* out[r, c] = sum_s partial[s * stride + r * ld_src + c] */
int nefii_reduce_splits(void* stream, const float* partial, int n_splits, int64_t stride, int rows, int ld_src, int cols,
                        float* out);


/* The whole trainable dense stack in ONE host call each way (RenderingNetwork.forward :196-241; EnvmapMaterialNetwork's
 * diffuse_albedo / roughness layers, sg_envmap_material.py:357-366, :403-425): input assembly, weight packing, the hidden-layer
 * GEMMs and the fused 1..4-wide output layer; backward = output layer, then per hidden layer the two transposes, the split-K
 * weight-gradient GEMM + reduction and the data-gradient GEMM.  Same kernels and arithmetic as the helpers above, sequenced
 * natively (at small batches the per-call host cost of sequencing them one by one was the limiter).
 * Host arrays: seg_* [n_seg <= 4]; weights / biases / dim_in / dim_out / grad_w / grad_b [n_hidden + 1] (hidden layers, then
 * the output layer; effective fp32 weights [out, in] row-major, device).  workspace: device, >= nefii_dense_stack_workspace_bytes;
 * with need_grad the forward leaves the activation planes in it and the backward must get the same workspace, untouched.
 * y [rows, n_out]; gy [rows, n_out]; grad_w / grad_b are overwritten. */
typedef struct nefii_dense_stack_desc {
  int32_t rows, n_hidden, act, n_seg;
  const float* seg_src[4];
  int32_t seg_width[4];
  int32_t seg_freqs[4];            /* >= 0: positional encoding of a 3-vector with that many octaves, -1: raw copy */
  const float* const* weights;     /* host array [n_hidden + 1] of device pointers */
  const float* const* biases;
  const int32_t* dim_in;           /* host [n_hidden + 1] */
  const int32_t* dim_out;
  int32_t need_grad;
  void* workspace;
  int64_t workspace_bytes;
  float* y;                        /* forward out */
  const float* gy;                 /* backward in */
  float* const* grad_w;            /* backward out, host array [n_hidden + 1] of device pointers */
  float* const* grad_b;
} nefii_dense_stack_desc;
int64_t nefii_dense_stack_workspace_bytes(const nefii_dense_stack_desc* desc /* host */);   /* < 0: error */
int nefii_dense_stack_fwd(void* stream, const nefii_dense_stack_desc* desc);
int nefii_dense_stack_bwd(void* stream, const nefii_dense_stack_desc* desc);


/* ---- IDRLoss, the step right after the rendering path (reference code/model/loss.py:122-320) -----------------
 * One launch for the five terms live in the step-2 recipe, means formed on the device (an empty mask yields 0 like the
 * reference's `mask.sum() == 0` early-outs, without their host round trips):
 *   terms[0] idr_rgb_loss, [1] sg_rgb_loss   image loss over net & obj                       (loss.py:162-174)
 *   terms[2] background_rgb_loss             env image loss over ~net & ~obj                 (loss.py:176-186)
 *   terms[3] mask_loss                       (1/alpha) BCE(-alpha sdf, obj) over ~(net&obj), / n   (loss.py:228-235)
 *   terms[4] normalsmooth_loss               mean unbiased variance of the normals of fully masked patches of `patch`
 *                                            consecutive pixels (4 r_patch^2; 0 or 1 = off)  (loss.py:255-264)
 *   terms[5..7] the three mask counts (pixels in net&obj, pixels in ~net&~obj, whole patches) -- inputs of the backward.
 * loss_type / env_loss_type: 0 L1, 1 L2 (MSE), 2 smooth L1 (image loss only).  All pointers device; masks uint8 [n];
 * rgb / normal [n,3]; sdf [n]; terms [8]. */
int nefii_idr_loss_fwd(void* stream, int n, int patch, const float* idr_rgb, const float* sg_rgb, const float* rgb_gt,
                       const float* normal, const float* sdf_output, const uint8_t* net_mask, const uint8_t* obj_mask,
                       int loss_type, int env_loss_type, float alpha, float* terms);
/* gradients of sum_k g_terms[k] * terms[k] (k < 5, device) w.r.t. the inputs; every output pointer may be null */
int nefii_idr_loss_bwd(void* stream, int n, int patch, const float* idr_rgb, const float* sg_rgb, const float* rgb_gt,
                       const float* normal, const float* sdf_output, const uint8_t* net_mask, const uint8_t* obj_mask,
                       int loss_type, int env_loss_type, float alpha, const float* terms, const float* g_terms,
                       float* g_idr_rgb, float* g_sg_rgb, float* g_normal, float* g_sdf_output);

#ifdef __cplusplus
}
#endif
#endif /* NEFII_B200_H_ */
