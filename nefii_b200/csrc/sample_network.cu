// Two small per-ray operators of the hot path.
//
// 1. camera_rays -- rend_util.get_camera_params + lift (reference code/utils/rend_util.py:90-142): pixel uv -> unit world ray
//    directions, one thread per ray, written straight into the tracer's direction buffer (the reference: ~15 tensor ops and a
//    materialised [B,4,S] homogeneous point array).
// 2. sample_network_fwd / _bwd -- SampleNetwork.forward (reference code/model/sample_network.py:10-24, IDR eq. 3):
//    x(theta) = c + (t0 - (s(x; theta) - s0) / (grad s . v0)) v, |grad s . v0| < 1e-8 -> 1e-8.  The forward VALUE is c + t0 v;
//    the operator exists for its backward: dx/ds = -v / (grad s . v0), which is what carries the image loss into the geometry.
//
// Built with -fmad=false -prec-div=true -prec-sqrt=true (per-op float32 rounding as torch evaluates the same expressions).
#include "common.cuh"
#include "tracer_math.cuh"

namespace nefii {

namespace {

constexpr int kBlock = 256;

// pose [B,4,4] row-major camera-to-world, K [B,4,4]; depth plane z = 1.  `order` selects how the 4-term products of
// torch.bmm(pose, [x_lift, y_lift, z, 1]) are summed (cuBLAS's order is not documented; tools/diag_gpu.py camrays measures
// which one reproduces it bit for bit): 0 = FMA chain k = 0..3, 1 = separate multiplies and adds k = 0..3.
__global__ void __launch_bounds__(kBlock)
camera_rays_kernel(int n_batch, int n_pix, const float* __restrict__ uv, const float* __restrict__ pose,
                   const float* __restrict__ intr, int order, float* __restrict__ dirs, float* __restrict__ cam_loc) {
  const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (i < n_batch && cam_loc) {
    cam_loc[i * 3 + 0] = pose[i * 16 + 3]; cam_loc[i * 3 + 1] = pose[i * 16 + 7]; cam_loc[i * 3 + 2] = pose[i * 16 + 11];
  }
  if (i >= (long long)n_batch * n_pix) return;
  const int b = (int)(i / n_pix);
  const float* P = pose + (size_t)b * 16;
  const float* K = intr + (size_t)b * 16;
  const float u = uv[i * 2 + 0], v = uv[i * 2 + 1];
  const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
  const float z = 1.0f;
  // lift (rend_util.py:129-142): (x - cx + cy*sk/fy - sk*y/fy) / fx * z ; (y - cy) / fy * z
  const float x_lift = (((u - cx) + (cy * sk) / fy) - (sk * v) / fy) / fx * z;
  const float y_lift = (v - cy) / fy * z;
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* Pr = P + 4 * r;
    if (order == 0) w[r] = __fmaf_rn(Pr[3], 1.0f, __fmaf_rn(Pr[2], z, __fmaf_rn(Pr[1], y_lift, Pr[0] * x_lift)));
    else w[r] = ((Pr[0] * x_lift + Pr[1] * y_lift) + Pr[2] * z) + Pr[3] * 1.0f;
  }
  const float c0 = w[0] - P[3], c1 = w[1] - P[7], c2 = w[2] - P[11];
  // F.normalize: v / max(||v||_2, 1e-12); torch's reduction over a contiguous innermost axis of 3 is (x0 + x2) + x1
  const float nrm = sqrtf((c0 * c0 + c2 * c2) + c1 * c1);
  const float den = fmaxf(nrm, 1e-12f);
  dirs[i * 3 + 0] = c0 / den; dirs[i * 3 + 1] = c1 / den; dirs[i * 3 + 2] = c2 / den;
}

__device__ __forceinline__ float guarded_dot(const float* g, const float* v) {
  float d = trm::dot3(g, v);
  if (fabsf(d) < 1e-8f) d = 1e-8f;     // sample_network.py:17
  return d;
}

__global__ void __launch_bounds__(kBlock)
sample_network_fwd_kernel(int n, const float* __restrict__ s, const float* __restrict__ s0, const float* __restrict__ grad,
                          const float* __restrict__ t0, const float* __restrict__ cam, const float* __restrict__ dirs,
                          float* __restrict__ out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const float g[3] = {grad[i * 3], grad[i * 3 + 1], grad[i * 3 + 2]};
  const float v[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
  const float d = guarded_dot(g, v);
  const float t = t0[i] - (s[i] - s0[i]) / d;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[i * 3 + c] = cam[i * 3 + c] + t * v[c];
}

// g_out [n,3] -> gradients w.r.t. surface_output, surface_sdf_values, surface_dists [n], surface_cam_loc, surface_ray_dirs,
// surface_points_grad [n,3] (any output pointer may be null).  The dot product uses the DETACHED directions (v0), so
// surface_ray_dirs only receives the gradient of the final `c + t(theta) v`.
__global__ void __launch_bounds__(kBlock)
sample_network_bwd_kernel(int n, const float* __restrict__ s, const float* __restrict__ s0, const float* __restrict__ grad,
                          const float* __restrict__ t0, const float* __restrict__ dirs, const float* __restrict__ g_out,
                          float* __restrict__ g_s, float* __restrict__ g_s0, float* __restrict__ g_t0, float* __restrict__ g_cam,
                          float* __restrict__ g_dirs, float* __restrict__ g_grad) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const float g[3] = {grad[i * 3], grad[i * 3 + 1], grad[i * 3 + 2]};
  const float v[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
  const float go[3] = {g_out[i * 3], g_out[i * 3 + 1], g_out[i * 3 + 2]};
  const float raw = trm::dot3(g, v);
  const bool clamped = fabsf(raw) < 1e-8f;
  const float d = clamped ? 1e-8f : raw;
  const float ds = s[i] - s0[i];
  const float t = t0[i] - ds / d;
  const float gt = trm::dot3(go, v);            // dL/dt(theta)
  if (g_s) g_s[i] = -gt / d;
  if (g_s0) g_s0[i] = gt / d;
  if (g_t0) g_t0[i] = gt;
  if (g_cam) { g_cam[i * 3] = go[0]; g_cam[i * 3 + 1] = go[1]; g_cam[i * 3 + 2] = go[2]; }
  if (g_dirs) { g_dirs[i * 3] = t * go[0]; g_dirs[i * 3 + 1] = t * go[1]; g_dirs[i * 3 + 2] = t * go[2]; }
  if (g_grad) {
    // d t / d dot = ds / dot^2 ; d dot / d grad = v0 (zero where the guard replaced the dot product)
    const float k = clamped ? 0.f : gt * ds / (d * d);
    g_grad[i * 3] = k * v[0]; g_grad[i * 3 + 1] = k * v[1]; g_grad[i * 3 + 2] = k * v[2];
  }
}

}  // namespace

int camera_rays(cudaStream_t stream, int n_batch, int n_pix, const float* uv, const float* pose, const float* intrinsics, int order,
                float* dirs, float* cam_loc) {
  NEFII_CHECK_ARG(n_batch >= 0 && n_pix >= 0, "camera_rays: bad shape");
  const long long n = (long long)n_batch * n_pix;
  if (n_batch == 0) return NEFII_OK;
  NEFII_CHECK_ARG(uv && pose && intrinsics && dirs, "camera_rays: null pointer");
  const long long threads = n > n_batch ? n : n_batch;
  camera_rays_kernel<<<ceil_div(threads, kBlock), kBlock, 0, stream>>>(n_batch, n_pix, uv, pose, intrinsics, order, dirs, cam_loc);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int sample_network_fwd(cudaStream_t stream, int n, const float* s, const float* s0, const float* grad, const float* t0,
                       const float* cam, const float* dirs, float* out) {
  if (n <= 0) return NEFII_OK;
  NEFII_CHECK_ARG(s && s0 && grad && t0 && cam && dirs && out, "sample_network_fwd: null pointer");
  sample_network_fwd_kernel<<<ceil_div(n, kBlock), kBlock, 0, stream>>>(n, s, s0, grad, t0, cam, dirs, out);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int sample_network_bwd(cudaStream_t stream, int n, const float* s, const float* s0, const float* grad, const float* t0,
                       const float* dirs, const float* g_out, float* g_s, float* g_s0, float* g_t0, float* g_cam, float* g_dirs,
                       float* g_grad) {
  if (n <= 0) return NEFII_OK;
  NEFII_CHECK_ARG(s && s0 && grad && t0 && dirs && g_out, "sample_network_bwd: null pointer");
  sample_network_bwd_kernel<<<ceil_div(n, kBlock), kBlock, 0, stream>>>(n, s, s0, grad, t0, dirs, g_out, g_s, g_s0, g_t0, g_cam,
                                                                        g_dirs, g_grad);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
