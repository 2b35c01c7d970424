// Near-field indirect-illumination integrator (reference pt_render_indirect_mlp,
// code/model/path_tracing_render.py:1255-1487): importance sampling + MIS weights, and the GGX /
// Lambert shading of the three samples, forward and backward.  One thread per surface point; the M
// light SGs are unpacked once per CTA into shared memory.  Built with -fmad=false (mis_math.cuh).
#include "common.cuh"
#include "mis_math.cuh"

namespace nefii {

using mism::MixLobe;
using mism::ShadeGeom;

namespace {

constexpr int kBlock = 128;

__device__ __forceinline__ void load_lobes(const float* __restrict__ lgt, int n_sg, MixLobe<float>* sL) {
  for (int m = threadIdx.x; m < n_sg; m += blockDim.x) {
    float raw[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) raw[i] = lgt[m * 7 + i];
    mism::load_mix_lobe(raw, sL[m]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kBlock)
mis_sample_kernel(int n, int n_sg, const float* __restrict__ lgt, const float* __restrict__ rough,
                  const float* __restrict__ normal, const float* __restrict__ view, const float* __restrict__ u,
                  float* __restrict__ wi, float* __restrict__ pdf, float* __restrict__ weight, float* __restrict__ mat) {
  extern __shared__ unsigned char smem_raw[];
  MixLobe<float>* sL = reinterpret_cast<MixLobe<float>*>(smem_raw);
  load_lobes(lgt, n_sg, sL);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    float nn[3], vv[3], uu[7];
#pragma unroll
    for (int c = 0; c < 3; ++c) { nn[c] = normal[i * 3 + c]; vv[c] = view[i * 3 + c]; }
#pragma unroll
    for (int c = 0; c < 7; ++c) uu[c] = u[i * 7 + c];
    float w[3][3], p[3], m[3][3], wt[3];
    mism::sample_point(sL, n_sg, nn, vv, rough[i], uu, w, p, m, wt);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int c = 0; c < 3; ++c) wi[((size_t)s * n + i) * 3 + c] = w[s][c];
      pdf[(size_t)s * n + i] = p[s];
      weight[(size_t)s * n + i] = wt[s];
      if (mat) {
#pragma unroll
        for (int j = 0; j < 3; ++j) mat[((size_t)s * 3 + j) * n + i] = m[s][j];
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock)
mis_shade_fwd_kernel(int n, int n_sg, const float* __restrict__ lgt, const float* __restrict__ spec, int spec_stride,
                     const float* __restrict__ rough, const float* __restrict__ albedo, const float* __restrict__ normal,
                     const float* __restrict__ view, const float* __restrict__ wi, const float* __restrict__ pdf,
                     const float* __restrict__ weight, const unsigned char* __restrict__ hit,
                     const float* __restrict__ indirect, float* __restrict__ out_rgb, float* __restrict__ out_spec,
                     float* __restrict__ out_diff, float* __restrict__ light_out) {
  extern __shared__ unsigned char smem_raw[];
  MixLobe<float>* sL = reinterpret_cast<MixLobe<float>*>(smem_raw);
  load_lobes(lgt, n_sg, sL);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    float nn[3], vv[3], al[3], sr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      nn[c] = normal[i * 3 + c]; vv[c] = view[i * 3 + c]; al[c] = albedo[i * 3 + c];
      sr[c] = spec[(size_t)i * spec_stride + c];
    }
    const float r = rough[i];
    float rgb[3] = {0.f, 0.f, 0.f}, st[3] = {0.f, 0.f, 0.f}, dt[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
      const size_t si = (size_t)s * n + i;
      float w[3] = {wi[si * 3 + 0], wi[si * 3 + 1], wi[si * 3 + 2]};
      float ind[3] = {indirect[si * 3 + 0], indirect[si * 3 + 1], indirect[si * 3 + 2]};
      const float vis = 1.0f - (hit[si] ? 1.0f : 0.0f);
      float light[3];
      mism::env_light(sL, n_sg, w, light);
      if (light_out) { light_out[si * 3 + 0] = light[0]; light_out[si * 3 + 1] = light[1]; light_out[si * 3 + 2] = light[2]; }
      ShadeGeom<float> g;
      mism::shade_geom(nn, vv, w, g);
      float sp[3], df[3];
      mism::shade_sample(g, r, sr, al, light, vis, ind, weight[si], pdf[si], sp, df);
#pragma unroll
      for (int c = 0; c < 3; ++c) { st[c] += sp[c]; dt[c] += df[c]; rgb[c] += sp[c] + df[c]; }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { out_rgb[i * 3 + c] = rgb[c]; out_spec[i * 3 + c] = st[c]; out_diff[i * 3 + c] = dt[c]; }
  }
}

// Backward.  Per-point gradients are written directly; the light-SG gradients are accumulated in the
// "unit" parametrisation acc[k] = {d/d axis (3), d/d sharpness, d/d amplitude (3)}: each lane owns the lobes
// k = lane, lane+32, ... and sweeps over the 32 points of its warp (values broadcast with shuffles), so no
// cross-lane reduction is needed; CTA totals go to shared memory, then one atomicAdd per value per CTA.
__global__ void __launch_bounds__(kBlock)
mis_shade_bwd_kernel(int n, int n_sg, const float* __restrict__ lgt, const float* __restrict__ spec, int spec_stride,
                     const float* __restrict__ rough, const float* __restrict__ albedo, const float* __restrict__ normal,
                     const float* __restrict__ view, const float* __restrict__ wi, const float* __restrict__ pdf,
                     const float* __restrict__ weight, const unsigned char* __restrict__ hit,
                     const float* __restrict__ indirect, const float* __restrict__ light_in,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_spec, const float* __restrict__ g_diff,
                     float* __restrict__ g_rough, float* __restrict__ g_albedo, float* __restrict__ g_specrefl,
                     float* __restrict__ g_indirect, float* __restrict__ g_lgt_acc, float* __restrict__ g_normal) {
  extern __shared__ unsigned char smem_raw[];
  MixLobe<float>* sL = reinterpret_cast<MixLobe<float>*>(smem_raw);
  float* sAcc = reinterpret_cast<float*>(sL + n_sg);   // [n_sg][7]
  for (int j = threadIdx.x; j < n_sg * 7; j += blockDim.x) sAcc[j] = 0.f;
  load_lobes(lgt, n_sg, sL);
  const int lane = threadIdx.x & 31;
  const int lobes_per_lane = (n_sg + 31) / 32;
  for (long long base = (long long)blockIdx.x * kBlock; base < n; base += (long long)gridDim.x * kBlock) {
    const long long i = base + threadIdx.x;
    const bool live = i < n;
    float wv[3][3], gl[3][3];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int c = 0; c < 3; ++c) { wv[s][c] = 0.f; gl[s][c] = 0.f; }
    if (live) {
      float nn[3], vv[3], al[3], sr[3], gs[3], gd[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        nn[c] = normal[i * 3 + c]; vv[c] = view[i * 3 + c]; al[c] = albedo[i * 3 + c];
        sr[c] = spec[(size_t)i * spec_stride + c];
        const float gr = g_rgb ? g_rgb[i * 3 + c] : 0.f;
        gs[c] = gr + (g_spec ? g_spec[i * 3 + c] : 0.f);
        gd[c] = gr + (g_diff ? g_diff[i * 3 + c] : 0.f);
      }
      const float r = rough[i];
      float gr_acc = 0.f, ga[3] = {0.f, 0.f, 0.f}, gsr[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        const size_t si = (size_t)s * n + i;
        float w[3] = {wi[si * 3 + 0], wi[si * 3 + 1], wi[si * 3 + 2]};
        float ind[3] = {indirect[si * 3 + 0], indirect[si * 3 + 1], indirect[si * 3 + 2]};
        float light[3] = {light_in[si * 3 + 0], light_in[si * 3 + 1], light_in[si * 3 + 2]};
        const float vis = 1.0f - (hit[si] ? 1.0f : 0.0f);
        ShadeGeom<float> g;
        mism::shade_geom(nn, vv, w, g);
        float glight[3], gind[3], gdots[3];
        mism::shade_sample_bwd(g, r, sr, al, light, vis, ind, weight[si], pdf[si], gs, gd, gr_acc, ga, gsr, glight, gind,
                               g_normal ? gdots : nullptr);
        if (g_normal) mism::shade_geom_bwd_normal(g, vv, w, gdots, gn);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          g_indirect[si * 3 + c] = gind[c];
          wv[s][c] = w[c];
          gl[s][c] = glight[c];
        }
      }
      g_rough[i] = gr_acc;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        g_albedo[i * 3 + c] = ga[c];
        if (g_specrefl) g_specrefl[i * 3 + c] = gsr[c];
        if (g_normal) g_normal[i * 3 + c] = gn[c];
      }
    }
    if (g_lgt_acc) {
      for (int t = 0; t < lobes_per_lane; ++t) {
        const int k = lane + 32 * t;
        float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        MixLobe<float> L;
        if (k < n_sg) L = sL[k];
        for (int src = 0; src < 32; ++src) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            float w[3], g3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              w[c] = __shfl_sync(0xffffffffu, wv[s][c], src);
              g3[c] = __shfl_sync(0xffffffffu, gl[s][c], src);
            }
            if (k < n_sg) {
              const float dm1 = sgm::dot3(w, L.axis) - 1.0f;
              const float e = expf(L.sharp * dm1);
              const float tsum = (g3[0] * L.amp[0] + g3[1] * L.amp[1] + g3[2] * L.amp[2]) * e;
              acc[0] += tsum * L.sharp * w[0]; acc[1] += tsum * L.sharp * w[1]; acc[2] += tsum * L.sharp * w[2];
              acc[3] += tsum * dm1;
              acc[4] += g3[0] * e; acc[5] += g3[1] * e; acc[6] += g3[2] * e;
            }
          }
        }
        if (k < n_sg) {
#pragma unroll
          for (int j = 0; j < 7; ++j) atomicAdd(&sAcc[k * 7 + j], acc[j]);
        }
      }
    }
  }
  if (g_lgt_acc) {
    __syncthreads();
    for (int j = threadIdx.x; j < n_sg * 7; j += blockDim.x) atomicAdd(&g_lgt_acc[j], sAcc[j]);
  }
}

// unit-parametrisation gradients -> gradients of the raw lgtSGs parameter (abs() and the lobe normalisation)
__global__ void sg_param_grad_kernel(int n_sg, const float* __restrict__ lgt, const float* __restrict__ acc, float eps,
                                     float* __restrict__ g_lgt, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sg) return;
  const float* raw = lgt + k * 7;
  const float* a = acc + k * 7;
  const float len = sqrtf(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2]);
  const float d = len + eps;
  const float ldg = raw[0] * a[0] + raw[1] * a[1] + raw[2] * a[2];
  float out[7];
  for (int c = 0; c < 3; ++c) out[c] = a[c] / d - (len > 0.f ? raw[c] * ldg / (len * d * d) : 0.f);
  out[3] = a[3] * (raw[3] > 0.f ? 1.f : (raw[3] < 0.f ? -1.f : 0.f));
  for (int c = 0; c < 3; ++c) out[4 + c] = a[4 + c] * (raw[4 + c] > 0.f ? 1.f : (raw[4 + c] < 0.f ? -1.f : 0.f));
  for (int c = 0; c < 7; ++c) g_lgt[k * 7 + c] = accumulate ? g_lgt[k * 7 + c] + out[c] : out[c];
}

// backward of the miss-ray environment lookup: acc[k] += {g.amp e sharp d, ..., g e}
__global__ void __launch_bounds__(kBlock)
background_sg_bwd_kernel(int n, int n_sg, const float* __restrict__ lgt, const float* __restrict__ dirs,
                         const float* __restrict__ g_out, float* __restrict__ g_lgt_acc) {
  extern __shared__ unsigned char smem_raw[];
  MixLobe<float>* sL = reinterpret_cast<MixLobe<float>*>(smem_raw);
  float* sAcc = reinterpret_cast<float*>(sL + n_sg);
  for (int j = threadIdx.x; j < n_sg * 7; j += blockDim.x) sAcc[j] = 0.f;
  for (int m = threadIdx.x; m < n_sg; m += blockDim.x) {
    float raw[7];
    for (int i = 0; i < 7; ++i) raw[i] = lgt[m * 7 + i];
    mism::load_mix_lobe(raw, sL[m]);
    float ax[3];
    sgm::unit3(raw, ax, 1e-8f);   // get_background_rgb normalises with +1e-8
    sL[m].axis[0] = ax[0]; sL[m].axis[1] = ax[1]; sL[m].axis[2] = ax[2];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int lobes_per_lane = (n_sg + 31) / 32;
  for (long long base = (long long)blockIdx.x * kBlock; base < n; base += (long long)gridDim.x * kBlock) {
    const long long i = base + threadIdx.x;
    float w[3] = {0.f, 0.f, 0.f}, g3[3] = {0.f, 0.f, 0.f};
    if (i < n) {
      for (int c = 0; c < 3; ++c) { w[c] = dirs[i * 3 + c]; g3[c] = g_out[i * 3 + c]; }
    }
    for (int t = 0; t < lobes_per_lane; ++t) {
      const int k = lane + 32 * t;
      float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      MixLobe<float> L;
      if (k < n_sg) L = sL[k];
      for (int src = 0; src < 32; ++src) {
        float ws[3], gs[3];
        for (int c = 0; c < 3; ++c) { ws[c] = __shfl_sync(0xffffffffu, w[c], src); gs[c] = __shfl_sync(0xffffffffu, g3[c], src); }
        if (k < n_sg) {
          const float dm1 = sgm::dot3(ws, L.axis) - 1.0f;
          const float e = expf(L.sharp * dm1);
          const float tsum = (gs[0] * L.amp[0] + gs[1] * L.amp[1] + gs[2] * L.amp[2]) * e;
          acc[0] += tsum * L.sharp * ws[0]; acc[1] += tsum * L.sharp * ws[1]; acc[2] += tsum * L.sharp * ws[2];
          acc[3] += tsum * dm1;
          acc[4] += gs[0] * e; acc[5] += gs[1] * e; acc[6] += gs[2] * e;
        }
      }
      if (k < n_sg)
        for (int j = 0; j < 7; ++j) atomicAdd(&sAcc[k * 7 + j], acc[j]);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_sg * 7; j += blockDim.x) atomicAdd(&g_lgt_acc[j], sAcc[j]);
}

inline int grid_for(long long n) {
  int b = ceil_div(n, kBlock);
  return b > kNumSMs * 16 ? kNumSMs * 16 : b;
}

}  // namespace

int mis_sample(cudaStream_t stream, int n, int n_sg, const float* lgt, const float* rough, const float* normal,
               const float* view, const float* u, float* wi, float* pdf, float* weight, float* mat) {
  NEFII_CHECK_ARG(n >= 0 && n_sg > 0, "mis_sample: bad sizes");
  if (n == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && rough && normal && view && u && wi && pdf && weight, "mis_sample: null pointer");
  const size_t smem = sizeof(MixLobe<float>) * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 48 * 1024, "mis_sample: too many light SGs (%d)", n_sg);
  mis_sample_kernel<<<grid_for(n), kBlock, smem, stream>>>(n, n_sg, lgt, rough, normal, view, u, wi, pdf, weight, mat);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int mis_shade_fwd(cudaStream_t stream, int n, int n_sg, const float* lgt, const float* spec, int spec_per_point,
                  const float* rough, const float* albedo, const float* normal, const float* view, const float* wi,
                  const float* pdf, const float* weight, const unsigned char* hit, const float* indirect, float* out_rgb,
                  float* out_spec, float* out_diff, float* light_out) {
  NEFII_CHECK_ARG(n >= 0 && n_sg > 0, "mis_shade_fwd: bad sizes");
  if (n == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && spec && rough && albedo && normal && view && wi && pdf && weight && hit && indirect && out_rgb &&
                      out_spec && out_diff,
                  "mis_shade_fwd: null pointer");
  const size_t smem = sizeof(MixLobe<float>) * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 48 * 1024, "mis_shade_fwd: too many light SGs (%d)", n_sg);
  mis_shade_fwd_kernel<<<grid_for(n), kBlock, smem, stream>>>(n, n_sg, lgt, spec, spec_per_point ? 3 : 0, rough, albedo,
                                                              normal, view, wi, pdf, weight, hit, indirect, out_rgb,
                                                              out_spec, out_diff, light_out);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int mis_shade_bwd(cudaStream_t stream, int n, int n_sg, const float* lgt, const float* spec, int spec_per_point,
                  const float* rough, const float* albedo, const float* normal, const float* view, const float* wi,
                  const float* pdf, const float* weight, const unsigned char* hit, const float* indirect,
                  const float* light, const float* g_rgb, const float* g_spec, const float* g_diff, float* g_rough,
                  float* g_albedo, float* g_specrefl, float* g_indirect, float* g_lgt_acc, float* g_normal) {
  NEFII_CHECK_ARG(n >= 0 && n_sg > 0, "mis_shade_bwd: bad sizes");
  if (n == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && spec && rough && albedo && normal && view && wi && pdf && weight && hit && indirect && light &&
                      g_rough && g_albedo && g_indirect,
                  "mis_shade_bwd: null pointer");
  const size_t smem = (sizeof(MixLobe<float>) + 7 * sizeof(float)) * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 48 * 1024, "mis_shade_bwd: too many light SGs (%d)", n_sg);
  int grid = grid_for(n);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  mis_shade_bwd_kernel<<<grid, kBlock, smem, stream>>>(n, n_sg, lgt, spec, spec_per_point ? 3 : 0, rough, albedo, normal,
                                                        view, wi, pdf, weight, hit, indirect, light, g_rgb, g_spec, g_diff,
                                                        g_rough, g_albedo, g_specrefl, g_indirect, g_lgt_acc, g_normal);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int background_sg_bwd(cudaStream_t stream, int n, int n_sg, const float* lgt, const float* dirs, const float* g_out,
                      float* g_lgt_acc) {
  NEFII_CHECK_ARG(n >= 0 && n_sg > 0, "background_sg_bwd: bad sizes");
  if (n == 0) return NEFII_OK;
  NEFII_CHECK_ARG(lgt && dirs && g_out && g_lgt_acc, "background_sg_bwd: null pointer");
  const size_t smem = (sizeof(MixLobe<float>) + 7 * sizeof(float)) * (size_t)n_sg;
  NEFII_CHECK_ARG(smem <= 48 * 1024, "background_sg_bwd: too many light SGs (%d)", n_sg);
  int grid = grid_for(n);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  background_sg_bwd_kernel<<<grid, kBlock, smem, stream>>>(n, n_sg, lgt, dirs, g_out, g_lgt_acc);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int sg_param_grad(cudaStream_t stream, int n_sg, const float* lgt, const float* acc, float eps, float* g_lgt, int accumulate) {
  NEFII_CHECK_ARG(n_sg > 0 && lgt && acc && g_lgt, "sg_param_grad: bad arguments");
  sg_param_grad_kernel<<<ceil_div(n_sg, 128), 128, 0, stream>>>(n_sg, lgt, acc, eps, g_lgt, accumulate);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
