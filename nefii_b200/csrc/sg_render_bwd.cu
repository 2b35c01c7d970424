// render_with_sg backward (reference: autograd through code/model/sg_render.py:164-295) as one fused FP32 kernel.
// Its own translation unit: unlike the forward (sg_render.cu, compiled without FMA contraction and with IEEE division so that
// its values follow the reference's rounding), the gradient only has to meet the rel 1e-3 tolerance, so this unit is compiled
// with FMA contraction and the fast division / square root (build.py: FAST_FP32) -- the adjoint sweep is instruction-bound.
#include "common.cuh"
#include "sg_math.cuh"
#include "sg_adjoint_math.cuh"

namespace nefii {

using sgm::LightSG;
using sgm::BrdfLobe;

constexpr int kMaxMaterials = 8;

// Backward of render_with_sg: the hand-derived adjoint of sg_adjoint_math.cuh (one reverse sweep per (ray, light SG, material)
// term gives the gradients w.r.t. the light SG, the warped BRDF lobe and the normal; the lobe's dependence on normal /
// roughness / specular reflectance is swept once per (ray, material)).
// Layout: one ray per LANE (its BRDF lobe, upstream gradients and adjoint accumulators stay in registers), the warp walks the
// light SGs together (unpacked once per CTA into shared memory, broadcast reads).  The 7 light-SG adjoints of a term are
// summed over the warp's 32 rays with a shuffle butterfly and added to the WARP's private accumulator rows in shared memory
// (no atomics); CTAs flush to global memory once, in the unit parametrisation {d axis, d sharpness, d amplitude} that
// nefii_sg_param_grad converts.  Roughness / specular-reflectance gradients (tiny [K,*] tensors) likewise.
constexpr int kSgBwdThreads = 128;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

__global__ void __launch_bounds__(kSgBwdThreads, 4)
sg_render_bwd_kernel(int n_rays, int n_sg, int n_mat, const float* __restrict__ lgt, const float* __restrict__ spec,
                     const float* __restrict__ rough, const float* __restrict__ albedo, const float* __restrict__ normal,
                     const float* __restrict__ view, const float* __restrict__ blend, const float* __restrict__ out_spec,
                     const float* __restrict__ out_diff, const float* __restrict__ g_rgb, const float* __restrict__ g_spec,
                     const float* __restrict__ g_diff, float* __restrict__ g_lgt_acc, float* __restrict__ g_rough,
                     float* __restrict__ g_specrefl, float* __restrict__ g_albedo, float* __restrict__ g_normal,
                     float* __restrict__ g_blend) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int kWarps = kSgBwdThreads / 32;
  float* sL = reinterpret_cast<float*>(smem_raw);                    // [n_sg][8]: unit axis, sharpness, amplitude
  float* sAcc = sL + n_sg * 8;                                        // [kWarps][n_sg][7]
  float* sMat = sAcc + kWarps * n_sg * 7;                             // [kWarps][n_mat][4]: d rough, d spec rgb
  for (int m = threadIdx.x; m < n_sg; m += blockDim.x) {
    float raw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) raw[i] = lgt[m * 7 + i];
    LightSG<float> L;
    sgm::load_light(raw, L);
    sL[m * 8 + 0] = L.axis[0]; sL[m * 8 + 1] = L.axis[1]; sL[m * 8 + 2] = L.axis[2]; sL[m * 8 + 3] = L.sharp;
    sL[m * 8 + 4] = L.amp[0]; sL[m * 8 + 5] = L.amp[1]; sL[m * 8 + 6] = L.amp[2]; sL[m * 8 + 7] = 0.f;
  }
  for (int j = threadIdx.x; j < kWarps * (n_sg * 7 + n_mat * 4); j += blockDim.x) sAcc[j] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wAcc = sAcc + warp * n_sg * 7;
  float* wMat = sMat + warp * n_mat * 4;
  const float inv_pi = 1.0f / sgm::K<float>::pi();
  for (long long base = ((long long)blockIdx.x * kWarps + warp) * 32; base < n_rays; base += (long long)gridDim.x * kSgBwdThreads) {
    const long long ray = base + lane;
    const bool live = ray < n_rays;
    float n[3] = {0.f, 0.f, 1.f}, v[3] = {0.f, 0.f, 1.f}, al_pi[3] = {0.f, 0.f, 0.f}, gs[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        n[c] = normal[ray * 3 + c]; v[c] = view[ray * 3 + c]; al_pi[c] = albedo[ray * 3 + c] * inv_pi;
        const float gr = g_rgb ? g_rgb[ray * 3 + c] : 0.f;
        // torch.clamp(min=0) passes the gradient where the summed radiance is positive; the reference sums the diffuse term
        // over the K axis
        gs[c] = out_spec[ray * 3 + c] > 0.f ? gr + (g_spec ? g_spec[ray * 3 + c] : 0.f) : 0.f;
        gd[c] = out_diff[ray * 3 + c] > 0.f ? (gr + (g_diff ? g_diff[ray * 3 + c] : 0.f)) * (float)n_mat : 0.f;
      }
    }
    float n_bar[3] = {0.f, 0.f, 0.f}, ga[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < n_mat; ++k) {
      BrdfLobe<float> B;
      const float sp[3] = {spec[k * 3 + 0], spec[k * 3 + 1], spec[k * 3 + 2]};
      sgm::make_brdf_lobe(n, v, rough[k], sp, B);
      float b_bar[3] = {0.f, 0.f, 0.f}, beta_bar = 0.f, nu_bar[3] = {0.f, 0.f, 0.f};
      // per-point blending weight of this base material (sg_render.py:254-256): the specular sum of material k enters scaled by it
      const float wk = (blend != nullptr && live) ? blend[ray * n_mat + k] : 1.0f;
      const float gsw[3] = {gs[0] * wk, gs[1] * wk, gs[2] * wk};
      float sk[3] = {0.f, 0.f, 0.f};      // this material's un-weighted specular sum (for d / d blending weight)
      for (int m = 0; m < n_sg; ++m) {
        const float4 l0 = *reinterpret_cast<const float4*>(sL + m * 8), l1 = *reinterpret_cast<const float4*>(sL + m * 8 + 4);
        const float a[3] = {l0.x, l0.y, l0.z}, mu[3] = {l1.x, l1.y, l1.z};
        const float lambda = l0.w;
        float t[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // a_bar[3], lambda_bar, mu_bar[3] of this light from this ray
        {
          const float w = (gsw[0] * mu[0]) * B.amp[0] + (gsw[1] * mu[1]) * B.amp[1] + (gsw[2] * mu[2]) * B.amp[2];
          const float phi = sga::specular_phi_vjp(n, a, lambda, B.axis, B.sharp, w, t, t[3], b_bar, beta_bar, n_bar);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            t[4 + c] += (gsw[c] * B.amp[c]) * phi;
            nu_bar[c] += (gsw[c] * mu[c]) * phi;
            sk[c] += (mu[c] * B.amp[c]) * phi;
          }
        }
        if (k == 0) {   // the diffuse term of the same light
          const float w = (gd[0] * mu[0]) * al_pi[0] + (gd[1] * mu[1]) * al_pi[1] + (gd[2] * mu[2]) * al_pi[2];
          const float psi = sga::psi_vjp(n, a, lambda, w, n_bar, t, t[3]);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            t[4 + c] += (gd[c] * al_pi[c]) * psi;
            ga[c] += (gd[c] * mu[c]) * (psi * inv_pi);
          }
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) t[i] = warp_sum(t[i]);
        float mine = t[0];
#pragma unroll
        for (int i = 1; i < 7; ++i) mine = (lane == i) ? t[i] : mine;
        if (lane < 7) wAcc[m * 7 + lane] += mine;
      }
      float g_r = 0.f, g_s[3] = {0.f, 0.f, 0.f};
      sga::brdf_lobe_vjp(n, v, rough[k], sp, b_bar, beta_bar, nu_bar, n_bar, g_r, g_s);
      g_r = warp_sum(g_r); g_s[0] = warp_sum(g_s[0]); g_s[1] = warp_sum(g_s[1]); g_s[2] = warp_sum(g_s[2]);
      if (lane == 0) { wMat[k * 4 + 0] += g_r; wMat[k * 4 + 1] += g_s[0]; wMat[k * 4 + 2] += g_s[1]; wMat[k * 4 + 3] += g_s[2]; }
      if (g_blend != nullptr && live) g_blend[ray * n_mat + k] = (gs[0] * sk[0] + gs[1] * sk[1]) + gs[2] * sk[2];
    }
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        g_albedo[ray * 3 + c] = ga[c];
        if (g_normal) g_normal[ray * 3 + c] = n_bar[c];
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_sg * 7; j += blockDim.x) {
    float x = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) x += sAcc[w * n_sg * 7 + j];
    atomicAdd(&g_lgt_acc[j], x);
  }
  for (int j = threadIdx.x; j < n_mat * 4; j += blockDim.x) {
    float x = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) x += sMat[w * n_mat * 4 + j];
    if ((j & 3) == 0) atomicAdd(&g_rough[j >> 2], x);
    else atomicAdd(&g_specrefl[(j >> 2) * 3 + (j & 3) - 1], x);
  }
}

int sg_render_bwd(cudaStream_t stream, int n_rays, int n_sg, int n_mat, const float* lgt, const float* spec, const float* rough,
                  const float* albedo, const float* normal, const float* view, const float* blend, const float* out_spec,
                  const float* out_diff, const float* g_rgb, const float* g_spec, const float* g_diff, float* g_lgt_acc,
                  float* g_rough, float* g_specrefl, float* g_albedo, float* g_normal, float* g_blend) {
  NEFII_CHECK_ARG(n_rays >= 0 && n_sg > 0 && n_mat > 0 && n_mat <= kMaxMaterials, "sg_render_bwd: bad sizes");
  if (n_rays == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && spec && rough && albedo && normal && view && out_spec && out_diff && g_lgt_acc && g_rough && g_specrefl &&
                      g_albedo,
                  "sg_render_bwd: null pointer");
  constexpr int kWarps = kSgBwdThreads / 32;
  const size_t smem = sizeof(float) * ((size_t)n_sg * 8 + (size_t)kWarps * ((size_t)n_sg * 7 + (size_t)n_mat * 4));
  NEFII_CHECK_ARG(smem <= 48 * 1024, "sg_render_bwd: too many light SGs (%d)", n_sg);
  int blocks = ceil_div(n_rays, kSgBwdThreads);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  sg_render_bwd_kernel<<<blocks, kSgBwdThreads, smem, stream>>>(n_rays, n_sg, n_mat, lgt, spec, rough, albedo, normal, view, blend,
                                                                out_spec, out_diff, g_rgb, g_spec, g_diff, g_lgt_acc, g_rough,
                                                                g_specrefl, g_albedo, g_normal, g_blend);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
