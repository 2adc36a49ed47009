// Interface of the tcgen05/TMA pooling kernels (ep_pool_sm100.cu).
#pragma once
#include "ep_common.cuh"

namespace ep {
bool sm100_supported(int x_dtype, int B, int N, int D, int M);
size_t sm100_workspace_bytes(int B, int N, int D, int M);
int sm100_pool_fwd(const void* x, const float* cls, float scale, int B, int N, int D, int M, float* P, float* S,
                   float* rowmax, float* rowsum, float* attn, int round_p, void* ws, cudaStream_t s);
int sm100_pool_bwd(const void* x, const float* S, float scale, int B, int N, int D, int M, const float* rowmax,
                   const float* rowsum, const float* dP, const float* delta, int ndelta, float* d_cls, void* ws,
                   cudaStream_t s);
void* sm100_dphl_ptr(void* ws, int B, int N, int D, int M);   // (B, J, D) bf16 hi/lo rows of dP inside the workspace
int sm100_J(int N, int D, int M);
}  // namespace ep
