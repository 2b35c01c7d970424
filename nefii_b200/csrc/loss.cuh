// Fused IDRLoss terms (reference code/model/loss.py); see loss.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nefii {

enum LossKind { LOSS_L1 = 0, LOSS_L2 = 1, LOSS_L1_SMOOTH = 2 };

int idr_loss_fwd(cudaStream_t stream, int n, int patch, const float* idr, const float* sg, const float* gt, const float* normal,
                 const float* sdf, const uint8_t* net, const uint8_t* obj, int loss_type, int env_type, float alpha, float* terms);
int idr_loss_bwd(cudaStream_t stream, int n, int patch, const float* idr, const float* sg, const float* gt, const float* normal,
                 const float* sdf, const uint8_t* net, const uint8_t* obj, int loss_type, int env_type, float alpha,
                 const float* terms, const float* g_terms, float* g_idr, float* g_sg, float* g_normal, float* g_sdf);

}  // namespace nefii
