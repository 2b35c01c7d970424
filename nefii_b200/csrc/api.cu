// extern "C" surface of libnefii_b200.so -- see include/nefii_b200.h for the contract.
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "../../include/nefii_b200.h"
#include "mlp_gemm.cuh"
#include "sdf_mlp.cuh"
#include "tracer.cuh"
#include "loss.cuh"
#include "dense_stack.cuh"

namespace nefii {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
void add_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sg_render_fwd(cudaStream_t, int, int, int, const float*, const float*, const float*, const float*, const float*,
                  const float*, const float*, float*, float*, float*);
int background_sg_fwd(cudaStream_t, int, int, const float*, const float*, float*);
int sg_render_bwd(cudaStream_t, int, int, int, const float*, const float*, const float*, const float*, const float*, const float*,
                  const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*,
                  float*, float*);
int mis_sample(cudaStream_t, int, int, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, float*);
int mis_shade_fwd(cudaStream_t, int, int, const float*, const float*, int, const float*, const float*, const float*, const float*,
                  const float*, const float*, const float*, const unsigned char*, const float*, float*, float*, float*, float*);
int mis_shade_bwd(cudaStream_t, int, int, const float*, const float*, int, const float*, const float*, const float*, const float*,
                  const float*, const float*, const float*, const unsigned char*, const float*, const float*, const float*,
                  const float*, const float*, float*, float*, float*, float*, float*, float*);
int background_sg_bwd(cudaStream_t, int, int, const float*, const float*, const float*, float*);
int sg_param_grad(cudaStream_t, int, const float*, const float*, float, float*, int);
int assemble_input(cudaStream_t, int, int, const float* const*, const int*, const int*, __nv_bfloat16*, __nv_bfloat16*, int, int);
int transpose_planes(cudaStream_t, const __nv_bfloat16*, const __nv_bfloat16*, int, int, int, __nv_bfloat16*, __nv_bfloat16*, int, int, int, float*);
int last_layer_bwd(cudaStream_t, int, int, int, int, const float*, const float*, const __nv_bfloat16*, const __nv_bfloat16*, int,
                   __nv_bfloat16*, __nv_bfloat16*, int, float*, float*);
int reduce_splits(cudaStream_t, const float*, int, long long, int, int, int, float*);
int probe_fp32(cudaStream_t, int, int, float*);
int camera_rays(cudaStream_t, int, int, const float*, const float*, const float*, int, float*, float*);
int sample_network_fwd(cudaStream_t, int, const float*, const float*, const float*, const float*, const float*, const float*, float*);
int sample_network_bwd(cudaStream_t, int, const float*, const float*, const float*, const float*, const float*, const float*, float*,
                       float*, float*, float*, float*, float*);

}  // namespace nefii

extern "C" {

const char* nefii_last_error(void) { return nefii::error_buffer(); }
int nefii_abi_version(void) { return 2; }
int64_t nefii_launch_count(void) { return (int64_t)nefii::launches(); }

int nefii_sg_render_fwd(void* stream, int n_rays, int n_sg, int n_mat, const float* lgt_sgs, const float* specular,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* blending, float* out_rgb, float* out_specular, float* out_diffuse) {
  return nefii::sg_render_fwd((cudaStream_t)stream, n_rays, n_sg, n_mat, lgt_sgs, specular, roughness, albedo, normal,
                              view, blending, out_rgb, out_specular, out_diffuse);
}

int nefii_background_sg_fwd(void* stream, int n_rays, int n_sg, const float* lgt_sgs, const float* dirs,
                            float* out_rgb) {
  return nefii::background_sg_fwd((cudaStream_t)stream, n_rays, n_sg, lgt_sgs, dirs, out_rgb);
}

int nefii_gemm_split_bf16(void* stream, nefii_gemm_desc* d) {
  if (!d) return nefii::set_error(NEFII_ERR_ARG, "nefii_gemm_split_bf16: null descriptor");
  nefii::GemmProblem p;
  p.a_hi = (const __nv_bfloat16*)d->a_hi; p.a_lo = (const __nv_bfloat16*)d->a_lo; p.a_ld = d->a_ld; p.rows_cap = d->rows_cap;
  p.b_hi = (const __nv_bfloat16*)d->b_hi; p.b_lo = (const __nv_bfloat16*)d->b_lo; p.b_ld = d->b_ld; p.n_pad = d->n_pad;
  p.k_pad = d->k_pad; p.count = d->count;
  nefii::GemmEpilogue& e = p.epi;
  e.mode = d->mode; e.act = d->act; e.n_valid = d->n_valid; e.bias = d->bias; e.out_scale = d->out_scale;
  e.dst.hi = (__nv_bfloat16*)d->dst_hi; e.dst.lo = (__nv_bfloat16*)d->dst_lo; e.dst.ld = d->dst_ld;
  e.dst_col0 = d->dst_col0; e.dst_ncols = d->dst_ncols; e.dst_zero_to = d->dst_zero_to;
  e.dst_f32 = d->dst_f32; e.f32_ld = d->f32_ld; e.f32_begin = d->f32_begin; e.f32_end = d->f32_end;
  e.w_last = d->w_last; e.b_last = d->b_last; e.n_last = d->n_last; e.w_last_ld = d->w_last_ld; e.dst_last = d->dst_last;
  e.seed.hi = (__nv_bfloat16*)d->seed_hi; e.seed.lo = (__nv_bfloat16*)d->seed_lo; e.seed.ld = d->seed_ld;
  e.sav_hi = (const __nv_bfloat16*)d->sav_hi; e.sav_lo = (const __nv_bfloat16*)d->sav_lo; e.sav_ld = d->sav_ld;
  e.sav_ncols = d->sav_ncols; e.sav_scale = d->sav_scale;
  p.k_splits = d->k_splits; p.f32_split_stride = d->f32_split_stride;
  p.k_flush = d->k_flush; e.dst_pad_ok = d->dst_pad_ok; e.fmt = d->fmt;
  p.pe.x = d->pe_x; p.pe.n_freqs = d->pe_n_freqs;
  p.pe.side.hi = (__nv_bfloat16*)d->pe_side_hi; p.pe.side.lo = (__nv_bfloat16*)d->pe_side_lo; p.pe.side.ld = d->pe_side_ld;
  p.pe.side_col0 = d->pe_side_col0; p.pe.side_scale = d->pe_side_scale;
  int used = 1;
  p.k_splits_used = &used;
  int rc = nefii::gemm_split_bf16((cudaStream_t)stream, p);
  d->k_splits_used = used;
  return rc;
}

int nefii_split_to_planes(void* stream, const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                          void* dst_hi, void* dst_lo, int rows_pad, int cols_pad) {
  return nefii::split_to_planes((cudaStream_t)stream, src, rows, cols, ld_src, transpose, scale, (__nv_bfloat16*)dst_hi,
                                (__nv_bfloat16*)dst_lo, rows_pad, cols_pad);
}

int nefii_split_to_planes_fmt(void* stream, const float* src, int rows, int cols, int ld_src, int transpose, float scale,
                              void* dst_hi, void* dst_lo, int rows_pad, int cols_pad, int fmt) {
  return nefii::split_to_planes((cudaStream_t)stream, src, rows, cols, ld_src, transpose, scale, (__nv_bfloat16*)dst_hi,
                                (__nv_bfloat16*)dst_lo, rows_pad, cols_pad, fmt);
}

int nefii_sdf_create(void** handle, const nefii_sdf_config* cfg) {
  if (!handle || !cfg) return nefii::set_error(NEFII_ERR_ARG, "nefii_sdf_create: null argument");
  nefii::SdfConfig c;
  c.d_in = cfg->d_in; c.n_freqs = cfg->n_freqs; c.width = cfg->width; c.n_hidden = cfg->n_hidden;
  c.skip_layer = cfg->skip_layer; c.d_out = cfg->d_out; c.d_feat = cfg->d_feat;
  nefii::SdfNet* net = new nefii::SdfNet();
  int rc = net->init(c);
  if (rc) { delete net; return rc; }
  *handle = net;
  return NEFII_OK;
}
int nefii_sdf_destroy(void* handle) {
  nefii::trace_graph_clear();     // cached trace graphs point into the handle's packed weights
  delete static_cast<nefii::SdfNet*>(handle);
  return NEFII_OK;
}
int nefii_sdf_set_format(void* handle, int fmt) {
  if (!handle) return nefii::set_error(NEFII_ERR_ARG, "nefii_sdf_set_format: null handle");
  nefii::trace_graph_clear();     // cached trace graphs carry the plane format in their kernel arguments
  return static_cast<nefii::SdfNet*>(handle)->set_format(fmt);
}
int nefii_sdf_set_pe_prologue(int on) {
  nefii::trace_graph_clear();     // cached trace graphs hold the launch sequence of an evaluation
  return nefii::sdf_set_pe_prologue(on);
}
int nefii_sdf_get_format(void* handle) {
  if (!handle) return -1;
  return static_cast<nefii::SdfNet*>(handle)->format();
}
int nefii_sdf_set_weights(void* handle, void* stream, const float* const* weights, const float* const* biases) {
  if (!handle || !weights || !biases) return nefii::set_error(NEFII_ERR_ARG, "nefii_sdf_set_weights: null argument");
  return static_cast<nefii::SdfNet*>(handle)->set_weights((cudaStream_t)stream, weights, biases);
}
int64_t nefii_sdf_workspace_bytes(void* handle, int rows_cap, int with_grad) {
  if (!handle) return -1;
  return (int64_t)static_cast<nefii::SdfNet*>(handle)->workspace_bytes(rows_cap, with_grad != 0);
}
int nefii_sdf_eval(void* handle, void* stream, int rows_cap, const int32_t* count, const float* x, void* workspace,
                   int64_t workspace_bytes, float* sdf, float* feat, float* grad, int k_flush) {
  if (!handle) return nefii::set_error(NEFII_ERR_ARG, "nefii_sdf_eval: null handle");
  return static_cast<nefii::SdfNet*>(handle)->eval((cudaStream_t)stream, rows_cap, count, x, workspace,
                                                   (size_t)workspace_bytes, sdf, feat, grad, k_flush);
}

static nefii::SdfSource make_source(int sdf_kind, const void* sdf, int n_prims) {
  nefii::SdfSource src;
  if (sdf_kind == 0) src.net = static_cast<const nefii::SdfNet*>(sdf);
  else { src.prims = static_cast<const float*>(sdf); src.n_prims = n_prims; }
  return src;
}
int64_t nefii_trace_workspace_bytes(int sdf_kind, const void* sdf, int n_rays, int n_steps) {
  return (int64_t)nefii::trace_workspace_bytes(make_source(sdf_kind, sdf, 1), n_rays, n_steps);
}
int nefii_ray_trace(void* stream, const nefii_trace_config* cfg, int sdf_kind, const void* sdf, int n_prims, int n_batch,
                    int n_pix, const float* cam_loc, const float* ray_dirs, const uint8_t* object_mask, int flags,
                    const float* linspace, const float* uniforms, void* workspace, int64_t workspace_bytes, float* points,
                    uint8_t* hit, float* dists, int64_t* stats) {
  if (!cfg || !sdf) return nefii::set_error(NEFII_ERR_ARG, "nefii_ray_trace: null config / sdf source");
  nefii::TraceConfig c;
  c.radius = cfg->object_bounding_sphere; c.sdf_threshold = cfg->sdf_threshold; c.line_search_step = cfg->line_search_step;
  c.line_step_iters = cfg->line_step_iters; c.sphere_tracing_iters = cfg->sphere_tracing_iters; c.n_steps = cfg->n_steps;
  c.n_rootfind_steps = cfg->n_rootfind_steps;
  return nefii::ray_trace((cudaStream_t)stream, c, make_source(sdf_kind, sdf, n_prims), n_batch, n_pix, cam_loc, ray_dirs,
                          object_mask, flags, linspace, uniforms, workspace, (size_t)workspace_bytes, points, hit, dists,
                          (long long*)stats);
}
int nefii_trace_set_tiers(int march_flush, int bulk_flush) { return nefii::trace_set_tiers(march_flush, bulk_flush); }
int nefii_trace_set_graph_mode(int mode) { return nefii::trace_set_graph_mode(mode); }
int nefii_trace_graph_mode(void) { return nefii::trace_graph_mode(); }
int64_t nefii_trace_graph_captures(void) { return (int64_t)nefii::trace_graph_captures(); }
int nefii_trace_set_quad_rows(int rows) { return nefii::trace_set_quad_rows(rows); }
int nefii_trace_set_bisect_depth(int depth) { return nefii::trace_set_bisect_depth(depth); }
int nefii_trace_graph_clear(void) { return nefii::trace_graph_clear(); }
int nefii_analytic_sdf_eval(void* stream, const float* prims, int n_prims, int n, const float* x, float* sdf) {
  return nefii::analytic_sdf_eval((cudaStream_t)stream, prims, n_prims, n, nullptr, x, sdf);
}

int nefii_mis_sample(void* stream, int n, int n_sg, const float* lgt_sgs, const float* roughness, const float* normal,
                     const float* view, const float* u, float* wi, float* pdf, float* weight, float* pdf_matrix) {
  return nefii::mis_sample((cudaStream_t)stream, n, n_sg, lgt_sgs, roughness, normal, view, u, wi, pdf, weight, pdf_matrix);
}
int nefii_mis_shade_fwd(void* stream, int n, int n_sg, const float* lgt_sgs, const float* specular, int spec_per_point,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* wi, const float* pdf, const float* weight, const uint8_t* hit, const float* indirect,
                        float* out_rgb, float* out_specular, float* out_diffuse, float* light) {
  return nefii::mis_shade_fwd((cudaStream_t)stream, n, n_sg, lgt_sgs, specular, spec_per_point, roughness, albedo, normal,
                              view, wi, pdf, weight, hit, indirect, out_rgb, out_specular, out_diffuse, light);
}
int nefii_mis_shade_bwd(void* stream, int n, int n_sg, const float* lgt_sgs, const float* specular, int spec_per_point,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* wi, const float* pdf, const float* weight, const uint8_t* hit, const float* indirect,
                        const float* light, const float* g_rgb, const float* g_specular, const float* g_diffuse,
                        float* g_roughness, float* g_albedo, float* g_specular_refl, float* g_indirect, float* g_lgt_acc,
                        float* g_normal) {
  return nefii::mis_shade_bwd((cudaStream_t)stream, n, n_sg, lgt_sgs, specular, spec_per_point, roughness, albedo, normal,
                              view, wi, pdf, weight, hit, indirect, light, g_rgb, g_specular, g_diffuse, g_roughness,
                              g_albedo, g_specular_refl, g_indirect, g_lgt_acc, g_normal);
}
int nefii_background_sg_bwd(void* stream, int n_rays, int n_sg, const float* lgt_sgs, const float* dirs, const float* g_out,
                            float* g_lgt_acc) {
  return nefii::background_sg_bwd((cudaStream_t)stream, n_rays, n_sg, lgt_sgs, dirs, g_out, g_lgt_acc);
}
int nefii_sg_param_grad(void* stream, int n_sg, const float* lgt_sgs, const float* acc, float eps, float* g_lgt, int accumulate) {
  return nefii::sg_param_grad((cudaStream_t)stream, n_sg, lgt_sgs, acc, eps, g_lgt, accumulate);
}

int nefii_assemble_input(void* stream, int rows, int n_seg, const float* const* src, const int32_t* width,
                         const int32_t* n_freqs, void* dst_hi, void* dst_lo, int ld, int k_pad) {
  if (!src || !width || !n_freqs) return nefii::set_error(NEFII_ERR_ARG, "nefii_assemble_input: null argument");
  return nefii::assemble_input((cudaStream_t)stream, rows, n_seg, src, width, n_freqs, (__nv_bfloat16*)dst_hi,
                               (__nv_bfloat16*)dst_lo, ld, k_pad);
}
int nefii_transpose_planes(void* stream, const void* src_hi, const void* src_lo, int ld_src, int rows, int cols, void* dst_hi,
                           void* dst_lo, int ld_dst, int rows_pad, int cols_pad, float* col_sum) {
  return nefii::transpose_planes((cudaStream_t)stream, (const __nv_bfloat16*)src_hi, (const __nv_bfloat16*)src_lo, ld_src, rows,
                                 cols, (__nv_bfloat16*)dst_hi, (__nv_bfloat16*)dst_lo, ld_dst, rows_pad, cols_pad, col_sum);
}
int nefii_last_layer_bwd(void* stream, int act, int rows, int width, int n_out, const float* gy, const float* w_last,
                         const void* h_hi, const void* h_lo, int h_ld, void* g_hi, void* g_lo, int g_ld, float* gw_last,
                         float* gb_last) {
  return nefii::last_layer_bwd((cudaStream_t)stream, act, rows, width, n_out, gy, w_last, (const __nv_bfloat16*)h_hi,
                               (const __nv_bfloat16*)h_lo, h_ld, (__nv_bfloat16*)g_hi, (__nv_bfloat16*)g_lo, g_ld, gw_last, gb_last);
}
int nefii_reduce_splits(void* stream, const float* partial, int n_splits, int64_t stride, int rows, int ld_src, int cols,
                        float* out) {
  return nefii::reduce_splits((cudaStream_t)stream, partial, n_splits, (long long)stride, rows, ld_src, cols, out);
}

int nefii_camera_rays(void* stream, int n_batch, int n_pix, const float* uv, const float* pose, const float* intrinsics, int order,
                      float* dirs, float* cam_loc) {
  return nefii::camera_rays((cudaStream_t)stream, n_batch, n_pix, uv, pose, intrinsics, order, dirs, cam_loc);
}
int nefii_sample_network_fwd(void* stream, int n, const float* surface_output, const float* surface_sdf_values,
                             const float* surface_points_grad, const float* surface_dists, const float* surface_cam_loc,
                             const float* surface_ray_dirs, float* out_points) {
  return nefii::sample_network_fwd((cudaStream_t)stream, n, surface_output, surface_sdf_values, surface_points_grad, surface_dists,
                                   surface_cam_loc, surface_ray_dirs, out_points);
}
int nefii_sample_network_bwd(void* stream, int n, const float* surface_output, const float* surface_sdf_values,
                             const float* surface_points_grad, const float* surface_dists, const float* surface_ray_dirs,
                             const float* g_points, float* g_output, float* g_sdf_values, float* g_dists, float* g_cam_loc,
                             float* g_ray_dirs, float* g_points_grad) {
  return nefii::sample_network_bwd((cudaStream_t)stream, n, surface_output, surface_sdf_values, surface_points_grad, surface_dists,
                                   surface_ray_dirs, g_points, g_output, g_sdf_values, g_dists, g_cam_loc, g_ray_dirs, g_points_grad);
}
int nefii_probe_fp32(void* stream, int blocks, int iters, float* sink) { return nefii::probe_fp32((cudaStream_t)stream, blocks, iters, sink); }
int nefii_gemm_profile_enable(int on) { return nefii::gemm_profile_enable(on); }
int nefii_gemm_set_cluster(int cl) { return nefii::gemm_set_cluster(cl); }
int nefii_gemm_set_debug(int mask) { return nefii::gemm_set_debug(mask); }
int nefii_gemm_set_pdl(int on) { return nefii::gemm_set_pdl(on); }
int nefii_gemm_set_grid_cap(int sms) { return nefii::gemm_set_grid_cap(sms); }
int nefii_gemm_set_k_flush(int k) { return nefii::gemm_set_k_flush(k); }
int nefii_gemm_set_k_flush_head(int k) { return nefii::gemm_set_k_flush_head(k); }
int nefii_gemm_set_trunc_comp(int k_blocks, float rho) { return nefii::gemm_set_trunc_comp(k_blocks, rho); }
int nefii_gemm_set_trunc_comp_fmt(int fmt, int k_blocks, float rho) { return nefii::gemm_set_trunc_comp(k_blocks, rho, fmt); }
int nefii_gemm_profile_fetch(double* out3) { return nefii::gemm_profile_fetch(out3); }

int nefii_sg_render_bwd(void* stream, int n_rays, int n_sg, int n_mat, const float* lgt_sgs, const float* specular,
                        const float* roughness, const float* albedo, const float* normal, const float* view,
                        const float* out_specular, const float* out_diffuse, const float* g_rgb, const float* g_specular,
                        const float* g_diffuse, float* g_lgt_acc, float* g_roughness, float* g_specular_refl, float* g_albedo,
                        float* g_normal, const float* blending, float* g_blending) {
  return nefii::sg_render_bwd((cudaStream_t)stream, n_rays, n_sg, n_mat, lgt_sgs, specular, roughness, albedo, normal, view,
                              blending, out_specular, out_diffuse, g_rgb, g_specular, g_diffuse, g_lgt_acc, g_roughness,
                              g_specular_refl, g_albedo, g_normal, g_blending);
}

int nefii_idr_loss_fwd(void* stream, int n, int patch, const float* idr_rgb, const float* sg_rgb, const float* rgb_gt,
                       const float* normal, const float* sdf_output, const uint8_t* net_mask, const uint8_t* obj_mask,
                       int loss_type, int env_loss_type, float alpha, float* terms) {
  return nefii::idr_loss_fwd((cudaStream_t)stream, n, patch, idr_rgb, sg_rgb, rgb_gt, normal, sdf_output, net_mask, obj_mask,
                             loss_type, env_loss_type, alpha, terms);
}
int nefii_idr_loss_bwd(void* stream, int n, int patch, const float* idr_rgb, const float* sg_rgb, const float* rgb_gt,
                       const float* normal, const float* sdf_output, const uint8_t* net_mask, const uint8_t* obj_mask,
                       int loss_type, int env_loss_type, float alpha, const float* terms, const float* g_terms,
                       float* g_idr_rgb, float* g_sg_rgb, float* g_normal, float* g_sdf_output) {
  return nefii::idr_loss_bwd((cudaStream_t)stream, n, patch, idr_rgb, sg_rgb, rgb_gt, normal, sdf_output, net_mask, obj_mask,
                             loss_type, env_loss_type, alpha, terms, g_terms, g_idr_rgb, g_sg_rgb, g_normal, g_sdf_output);
}

namespace {
int fill_stack(const nefii_dense_stack_desc* c, bool with_ptrs, nefii::DenseStack& d) {
  if (!c || !c->dim_in || !c->dim_out) return nefii::set_error(NEFII_ERR_ARG, "nefii_dense_stack: null descriptor");
  if (c->n_hidden < 1 || c->n_hidden >= nefii::kDenseMaxLayers || c->n_seg < 1 || c->n_seg > 4)
    return nefii::set_error(NEFII_ERR_ARG, "nefii_dense_stack: n_hidden %d / n_seg %d out of range", c->n_hidden, c->n_seg);
  if (with_ptrs && (!c->weights || !c->biases)) return nefii::set_error(NEFII_ERR_ARG, "nefii_dense_stack: null weights / biases");
  d.rows = c->rows; d.n_hidden = c->n_hidden; d.act = c->act; d.n_seg = c->n_seg;
  for (int s = 0; s < c->n_seg; ++s) { d.seg_src[s] = c->seg_src[s]; d.seg_width[s] = c->seg_width[s]; d.seg_freqs[s] = c->seg_freqs[s]; }
  for (int l = 0; l <= c->n_hidden; ++l) {
    d.dim_in[l] = c->dim_in[l]; d.dim_out[l] = c->dim_out[l];
    if (with_ptrs) { d.weights[l] = c->weights[l]; d.biases[l] = c->biases[l]; }
    if (with_ptrs && c->grad_w && c->grad_b) { d.grad_w[l] = c->grad_w[l]; d.grad_b[l] = c->grad_b[l]; }
  }
  d.need_grad = c->need_grad; d.workspace = c->workspace; d.workspace_bytes = c->workspace_bytes; d.y = c->y; d.gy = c->gy;
  return NEFII_OK;
}
}  // namespace
int64_t nefii_dense_stack_workspace_bytes(const nefii_dense_stack_desc* desc) {
  nefii::DenseStack d;
  if (fill_stack(desc, false, d)) return -1;
  return nefii::dense_stack_workspace_bytes(d);
}
int nefii_dense_stack_fwd(void* stream, const nefii_dense_stack_desc* desc) {
  nefii::DenseStack d;
  if (int rc = fill_stack(desc, true, d)) return rc;
  return nefii::dense_stack_fwd((cudaStream_t)stream, d);
}
int nefii_dense_stack_bwd(void* stream, const nefii_dense_stack_desc* desc) {
  nefii::DenseStack d;
  if (int rc = fill_stack(desc, true, d)) return rc;
  return nefii::dense_stack_bwd((cudaStream_t)stream, d);
}

}  // extern "C"
