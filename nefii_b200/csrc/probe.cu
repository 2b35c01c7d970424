// Measurement aid (bench.py): FP32 FMA throughput of the device, the denominator of `fp32_fraction` for the SG shading and
// MIS kernels ("FP32 peak to be micro-benchmarked like the driver's HBM figure").
// Every thread runs 8 independent FMA chains; flops = blocks * 256 * iters * 8 * 2.
#include "common.cuh"

namespace nefii {

namespace {
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float seed, float* __restrict__ sink) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
    a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
  }
  const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678f) sink[0] = s;   // keeps the chains alive; practically never true
}
}  // namespace

int probe_fp32(cudaStream_t stream, int blocks, int iters, float* sink) {
  NEFII_CHECK_ARG(blocks > 0 && iters > 0 && sink, "probe_fp32: bad argument");
  fp32_probe_kernel<<<blocks, 256, 0, stream>>>(iters, 1.0f, sink);
  NEFII_LAUNCH_CHECK();
  return NEFII_OK;
}

}  // namespace nefii
