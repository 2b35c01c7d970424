// CTA-pair (tcgen05 cta_group::2) instantiations of the layer GEMM; see mlp_gemm.cu / mlp_gemm_kernel.cuh.
#include "mlp_gemm_kernel.cuh"

namespace nefii {

void* gemm_pair_kernel(int key) { return reinterpret_cast<void*>(select_gemm_kernel<2>(key)); }

}  // namespace nefii
