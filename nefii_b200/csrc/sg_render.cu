// render_with_sg forward (reference code/model/sg_render.py:164-295) and the background
// environment lookup (implicit_differentiable_renderer.py:646-663) as fused FP32 kernels.
//
// Layout: one ray is owned by LANES consecutive lanes of a warp; lane j integrates light SGs
// j, j+LANES, ... and the partial RGB sums are combined with warp shuffles.  The M light SGs are
// unpacked once per CTA into shared memory (axis, sharpness, amplitude + the part of
// hemisphere_int that depends on the sharpness only).  Compiled with -fmad=false: see sg_math.cuh.
#include "common.cuh"
#include "sg_math.cuh"

namespace nefii {

using sgm::LightSG;
using sgm::BrdfLobe;

constexpr int kSgThreads = 128;
constexpr int kMaxMaterials = 8;

template <int LANES>
__global__ void __launch_bounds__(kSgThreads)
sg_render_fwd_kernel(int n_rays, int n_sg, int n_mat,
                     const float* __restrict__ lgt, const float* __restrict__ spec,
                     const float* __restrict__ rough, const float* __restrict__ albedo,
                     const float* __restrict__ normal, const float* __restrict__ view,
                     const float* __restrict__ blend,
                     float* __restrict__ out_rgb, float* __restrict__ out_spec, float* __restrict__ out_diff) {
  extern __shared__ unsigned char smem_raw[];
  LightSG<float>* sL = reinterpret_cast<LightSG<float>*>(smem_raw);
  for (int m = threadIdx.x; m < n_sg; m += blockDim.x) {
    float raw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) raw[i] = lgt[m * 7 + i];
    sgm::load_light(raw, sL[m]);
  }
  __syncthreads();

  constexpr int RAYS_PER_BLOCK = kSgThreads / LANES;
  const int sub = threadIdx.x % LANES;
  const int ray_in_block = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * RAYS_PER_BLOCK; base < n_rays;
       base += (long long)gridDim.x * RAYS_PER_BLOCK) {
    const long long ray = base + ray_in_block;
    const bool live = ray < n_rays;
    float n[3] = {0.f, 0.f, 1.f}, v[3] = {0.f, 0.f, 1.f}, a_pi[3] = {0.f, 0.f, 0.f};
    if (live) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        n[c] = normal[ray * 3 + c];
        v[c] = view[ray * 3 + c];
        a_pi[c] = albedo[ray * 3 + c] * (1.0f / sgm::K<float>::pi());
      }
    }
    float s_acc[3] = {0.f, 0.f, 0.f};
    float d_acc[3] = {0.f, 0.f, 0.f};
    if (live) {
      for (int k = 0; k < n_mat; ++k) {
        BrdfLobe<float> B;
        float sp[3] = {spec[k * 3 + 0], spec[k * 3 + 1], spec[k * 3 + 2]};
        sgm::make_brdf_lobe(n, v, rough[k], sp, B);
        float sk[3] = {0.f, 0.f, 0.f};
        for (int m = sub; m < n_sg; m += LANES) {
          float t[3];
          sgm::specular_term(n, sL[m], B, t);
          sk[0] += t[0]; sk[1] += t[1]; sk[2] += t[2];
        }
        const float wk = blend ? blend[ray * n_mat + k] : 1.0f;
        s_acc[0] += sk[0] * wk; s_acc[1] += sk[1] * wk; s_acc[2] += sk[2] * wk;
      }
      for (int m = sub; m < n_sg; m += LANES) {
        float t[3];
        sgm::diffuse_term(n, sL[m], a_pi, t);
        d_acc[0] += t[0]; d_acc[1] += t[1]; d_acc[2] += t[2];
      }
    }
#pragma unroll
    for (int off = LANES / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s_acc[c] += __shfl_xor_sync(0xffffffffu, s_acc[c], off);
        d_acc[c] += __shfl_xor_sync(0xffffffffu, d_acc[c], off);
      }
    }
    if (live && sub == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // the reference broadcasts the diffuse term over the K axis and sums it (sg_render.py:279-286)
        float d = d_acc[c] * (float)n_mat;
        float s = sgm::clamp_min(s_acc[c], 0.f);
        d = sgm::clamp_min(d, 0.f);
        out_spec[ray * 3 + c] = s;
        out_diff[ray * 3 + c] = d;
        out_rgb[ray * 3 + c] = s + d;
      }
    }
  }
}

// sum_m amp_m * exp(sharp_m * (d . axis_m - 1)) for miss rays; lobes normalised with +1e-8
__global__ void __launch_bounds__(kSgThreads)
background_sg_fwd_kernel(int n_rays, int n_sg, const float* __restrict__ lgt,
                         const float* __restrict__ dirs, float* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  float* sL = reinterpret_cast<float*>(smem_raw);  // [M][7] unpacked
  for (int m = threadIdx.x; m < n_sg; m += blockDim.x) {
    float raw[3] = {lgt[m * 7 + 0], lgt[m * 7 + 1], lgt[m * 7 + 2]};
    float ax[3];
    sgm::unit3(raw, ax, 1e-8f);
    sL[m * 7 + 0] = ax[0]; sL[m * 7 + 1] = ax[1]; sL[m * 7 + 2] = ax[2];
    sL[m * 7 + 3] = fabsf(lgt[m * 7 + 3]);
    sL[m * 7 + 4] = fabsf(lgt[m * 7 + 4]); sL[m * 7 + 5] = fabsf(lgt[m * 7 + 5]); sL[m * 7 + 6] = fabsf(lgt[m * 7 + 6]);
  }
  __syncthreads();
  for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays;
       ray += (long long)gridDim.x * blockDim.x) {
    float d[3] = {dirs[ray * 3 + 0], dirs[ray * 3 + 1], dirs[ray * 3 + 2]};
    float acc[3] = {0.f, 0.f, 0.f};
    for (int m = 0; m < n_sg; ++m) {
      const float* L = sL + m * 7;
      float e = expf(L[3] * (sgm::dot3(d, L) - 1.0f));
      acc[0] += L[4] * e; acc[1] += L[5] * e; acc[2] += L[6] * e;
    }
    out[ray * 3 + 0] = acc[0]; out[ray * 3 + 1] = acc[1]; out[ray * 3 + 2] = acc[2];
  }
}

int sg_render_fwd(cudaStream_t stream, int n_rays, int n_sg, int n_mat, const float* lgt, const float* spec,
                  const float* rough, const float* albedo, const float* normal, const float* view,
                  const float* blend, float* out_rgb, float* out_spec, float* out_diff) {
  NEFII_CHECK_ARG(n_rays >= 0 && n_sg > 0 && n_mat > 0 && n_mat <= kMaxMaterials,
                  "sg_render_fwd: bad sizes n_rays=%d n_sg=%d n_mat=%d", n_rays, n_sg, n_mat);
  if (n_rays == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && spec && rough && albedo && normal && view && out_rgb && out_spec && out_diff,
                  "sg_render_fwd: null pointer");
  const size_t smem = sizeof(LightSG<float>) * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 200 * 1024, "sg_render_fwd: too many light SGs (%d)", n_sg);
  // Small batches: a full warp per ray keeps all SMs busy; large batches: 4 lanes per ray so the
  // per-ray BRDF-lobe setup is replicated 4x instead of 32x.
  const long long work_wide = (long long)n_rays * 32;
  const bool wide = work_wide <= (long long)kNumSMs * 2048 * 2;
  if (wide) {
    auto kern = sg_render_fwd_kernel<32>;
    if (smem > 48 * 1024) NEFII_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = ceil_div(n_rays, kSgThreads / 32);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    kern<<<blocks, kSgThreads, smem, stream>>>(n_rays, n_sg, n_mat, lgt, spec, rough, albedo, normal, view, blend,
                                               out_rgb, out_spec, out_diff);
  } else {
    auto kern = sg_render_fwd_kernel<4>;
    if (smem > 48 * 1024) NEFII_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = ceil_div(n_rays, kSgThreads / 4);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    kern<<<blocks, kSgThreads, smem, stream>>>(n_rays, n_sg, n_mat, lgt, spec, rough, albedo, normal, view, blend,
                                               out_rgb, out_spec, out_diff);
  }
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int background_sg_fwd(cudaStream_t stream, int n_rays, int n_sg, const float* lgt, const float* dirs, float* out) {
  NEFII_CHECK_ARG(n_rays >= 0 && n_sg > 0, "background_sg_fwd: bad sizes");
  if (n_rays == 0) return NEFII_OK;
  const size_t smem = sizeof(float) * 7 * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 48 * 1024, "background_sg_fwd: too many light SGs (%d)", n_sg);
  int blocks = ceil_div(n_rays, kSgThreads);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  background_sg_fwd_kernel<<<blocks, kSgThreads, smem, stream>>>(n_rays, n_sg, lgt, dirs, out);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
