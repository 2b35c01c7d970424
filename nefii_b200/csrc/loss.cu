// Fused IDRLoss terms of the step-2 recipe (reference code/model/loss.py:162-186, 228-235, 255-264): masked image loss on
// idr / sg rgb, background image loss, mask BCE and 2x2 normal-variance in ONE launch, means formed on the device (the
// reference's `mask.sum() == 0` host round trips disappear: an empty mask simply yields 0), and one launch for all the
// input gradients.  The batch is a few thousand pixels, so one CTA does it; fixed-order double reductions keep it
// deterministic.
#include <cstdint>
#include "common.cuh"
#include "loss.cuh"

namespace nefii {

namespace {

constexpr int kLossThreads = 512;

__device__ __forceinline__ float img_loss(float d, int kind) {
  if (kind == LOSS_L1) return fabsf(d);
  if (kind == LOSS_L2) return d * d;
  const float a = fabsf(d);                       // SmoothL1, beta = 1
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float img_loss_grad(float d, int kind) {
  if (kind == LOSS_L1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  if (kind == LOSS_L2) return 2.f * d;
  return fminf(fmaxf(d, -1.f), 1.f);
}

constexpr int kNumSums = 8;   // idr, sg, bg, bce, normal-var, n_m, n_bg, n_patch

__global__ void __launch_bounds__(kLossThreads)
idr_loss_fwd_kernel(int n, int patch, const float* __restrict__ idr, const float* __restrict__ sg, const float* __restrict__ gt,
                    const float* __restrict__ normal, const float* __restrict__ sdf, const uint8_t* __restrict__ net,
                    const uint8_t* __restrict__ obj, int loss_type, int env_type, float alpha, float* __restrict__ terms) {
  double acc[kNumSums] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += kLossThreads) {
    const bool a = net[i] != 0, b = obj[i] != 0;
    if (a && b) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float g = gt[3 * i + c];
        s0 += img_loss(idr[3 * i + c] - g, loss_type);
        s1 += img_loss(sg[3 * i + c] - g, loss_type);
      }
      acc[0] += s0; acc[1] += s1; acc[5] += 1.0;
    } else {
      // mask term over ~(net & obj): BCE with logits z = -alpha sdf against the ground-truth mask
      const float z = -alpha * sdf[i];
      const float y = b ? 1.f : 0.f;
      acc[3] += fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
      if (!a && !b) {
        float s2 = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) s2 += img_loss(sg[3 * i + c] - gt[3 * i + c], env_type);
        acc[2] += s2; acc[6] += 1.0;
      }
    }
  }
  if (patch > 1) {
    const int n_patches = n / patch;
    for (int p = threadIdx.x; p < n_patches; p += kLossThreads) {
      bool all = true;
      for (int j = 0; j < patch; ++j) all = all && net[p * patch + j] != 0 && obj[p * patch + j] != 0;
      if (!all) continue;
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float mean = 0.f;
        for (int j = 0; j < patch; ++j) mean += normal[3 * (p * patch + j) + c];
        mean /= (float)patch;
        float ss = 0.f;
        for (int j = 0; j < patch; ++j) {
          const float d = normal[3 * (p * patch + j) + c] - mean;
          ss += d * d;
        }
        v += ss / (float)(patch - 1);   // torch.var: unbiased
      }
      acc[4] += v; acc[7] += 1.0;
    }
  }
  // fixed-order reduction: lanes, then warps
  __shared__ double s_part[kLossThreads / 32][kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[kNumSums];
    for (int k = 0; k < kNumSums; ++k) {
      double v = 0;
      for (int w = 0; w < kLossThreads / 32; ++w) v += s_part[w][k];
      t[k] = v;
    }
    terms[0] = t[5] > 0 ? (float)(t[0] / (3.0 * t[5])) : 0.f;
    terms[1] = t[5] > 0 ? (float)(t[1] / (3.0 * t[5])) : 0.f;
    terms[2] = t[6] > 0 ? (float)(t[2] / (3.0 * t[6])) : 0.f;
    terms[3] = (float)(t[3] / ((double)alpha * (double)n));
    terms[4] = t[7] > 0 ? (float)(t[4] / (3.0 * t[7])) : 0.f;
    terms[5] = (float)t[5]; terms[6] = (float)t[6]; terms[7] = (float)t[7];
  }
}

__global__ void idr_loss_bwd_kernel(int n, int patch, const float* __restrict__ idr, const float* __restrict__ sg,
                                    const float* __restrict__ gt, const float* __restrict__ normal, const float* __restrict__ sdf,
                                    const uint8_t* __restrict__ net, const uint8_t* __restrict__ obj, int loss_type, int env_type,
                                    float alpha, const float* __restrict__ terms, const float* __restrict__ g_terms,
                                    float* __restrict__ g_idr, float* __restrict__ g_sg, float* __restrict__ g_normal,
                                    float* __restrict__ g_sdf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool a = net[i] != 0, b = obj[i] != 0;
  const float n_m = terms[5], n_bg = terms[6], n_p = terms[7];
  const float w_m0 = n_m > 0.f ? g_terms[0] / (3.f * n_m) : 0.f;
  const float w_m1 = n_m > 0.f ? g_terms[1] / (3.f * n_m) : 0.f;
  const float w_bg = n_bg > 0.f ? g_terms[2] / (3.f * n_bg) : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float g = gt[3 * i + c];
    if (g_idr) g_idr[3 * i + c] = (a && b) ? w_m0 * img_loss_grad(idr[3 * i + c] - g, loss_type) : 0.f;
    if (g_sg) {
      float v = 0.f;
      if (a && b) v = w_m1 * img_loss_grad(sg[3 * i + c] - g, loss_type);
      else if (!a && !b) v = w_bg * img_loss_grad(sg[3 * i + c] - g, env_type);
      g_sg[3 * i + c] = v;
    }
  }
  if (g_sdf) {
    float v = 0.f;
    if (!(a && b)) {
      const float z = -alpha * sdf[i];
      const float sig = 1.f / (1.f + expf(-z));
      v = -g_terms[3] * (sig - (b ? 1.f : 0.f)) / (float)n;   // d/dsdf of (1/alpha) BCE(-alpha sdf) / n
    }
    g_sdf[i] = v;
  }
  if (g_normal) {
    float out[3] = {0.f, 0.f, 0.f};
    if (patch > 1 && i < (n / patch) * patch && n_p > 0.f) {
      const int p0 = (i / patch) * patch;
      bool all = true;
      for (int j = 0; j < patch; ++j) all = all && net[p0 + j] != 0 && obj[p0 + j] != 0;
      if (all) {
        const float w = g_terms[4] * 2.f / ((float)(patch - 1) * 3.f * n_p);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float mean = 0.f;
          for (int j = 0; j < patch; ++j) mean += normal[3 * (p0 + j) + c];
          mean /= (float)patch;
          out[c] = w * (normal[3 * i + c] - mean);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) g_normal[3 * i + c] = out[c];
  }
}

}  // namespace

int idr_loss_fwd(cudaStream_t stream, int n, int patch, const float* idr, const float* sg, const float* gt, const float* normal,
                 const float* sdf, const uint8_t* net, const uint8_t* obj, int loss_type, int env_type, float alpha, float* terms) {
  NEFII_CHECK_ARG(n >= 0 && patch >= 0 && patch <= 64, "idr_loss_fwd: bad sizes n=%d patch=%d", n, patch);
  NEFII_CHECK_ARG(loss_type >= 0 && loss_type <= 2 && env_type >= 0 && env_type <= 1, "idr_loss_fwd: unknown loss type");
  NEFII_CHECK_ARG(alpha > 0.f, "idr_loss_fwd: alpha must be positive");
  NEFII_CHECK_ARG(terms != nullptr, "idr_loss_fwd: null output");
  NEFII_CHECK_ARG(n == 0 || (idr && sg && gt && normal && sdf && net && obj), "idr_loss_fwd: null input");
  NEFII_CHECK_ARG(patch <= 1 || n % patch == 0, "idr_loss_fwd: %d pixels are not whole patches of %d", n, patch);
  idr_loss_fwd_kernel<<<1, kLossThreads, 0, stream>>>(n, patch, idr, sg, gt, normal, sdf, net, obj, loss_type, env_type, alpha, terms);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

int idr_loss_bwd(cudaStream_t stream, int n, int patch, const float* idr, const float* sg, const float* gt, const float* normal,
                 const float* sdf, const uint8_t* net, const uint8_t* obj, int loss_type, int env_type, float alpha,
                 const float* terms, const float* g_terms, float* g_idr, float* g_sg, float* g_normal, float* g_sdf) {
  NEFII_CHECK_ARG(n >= 0 && patch >= 0 && patch <= 64, "idr_loss_bwd: bad sizes n=%d patch=%d", n, patch);
  if (n == 0) return NEFII_OK;
  NEFII_CHECK_ARG(idr && sg && gt && normal && sdf && net && obj && terms && g_terms, "idr_loss_bwd: null input");
  idr_loss_bwd_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(n, patch, idr, sg, gt, normal, sdf, net, obj, loss_type, env_type, alpha,
                                                             terms, g_terms, g_idr, g_sg, g_normal, g_sdf);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
