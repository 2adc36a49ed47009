// Interface of the tcgen05/TMA pooling kernels (ep_pool_sm100.cu).
#pragma once
#include "ep_common.cuh"

struct CUtensorMap_st;   // = CUtensorMap (cuda.h)

namespace ep {
bool sm100_supported(int x_dtype, int B, int N, int D, int M);
size_t sm100_workspace_bytes(int B, int N, int D, int M);
int sm100_pool_fwd(const void* x, const float* cls, float scale, int B, int N, int D, int M, float* P, float* S,
                   float* rowmax, float* rowsum, float* attn, int round_p, void* ws, cudaStream_t s, int q_ready = 0);
int sm100_pool_bwd(const void* x, const float* S, float scale, int B, int N, int D, int M, const float* rowmax,
                   const float* rowsum, const float* dP, const float* delta, int ndelta, float* d_cls, void* ws,
                   cudaStream_t s);
void* sm100_dphl_ptr(void* ws, int B, int N, int D, int M);   // (B, J, D) bf16 hi/lo rows of dP inside the workspace
int sm100_J(int N, int D, int M);
void* sm100_qhl_ptr(void* ws, int B, int N, int D, int M);    // (J, D) bf16 hi/lo rows of the scaled queries

// bf16 tensor map of `rank` dims (dims[0] contiguous; strides in bytes for dims 1..rank-1), 128-byte swizzle,
// zero fill out of bounds (ep_pool_sm100.cu)
int make_tmap_bf16(::CUtensorMap_st* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box);

// one-pass fused kernels (ep_fused_sm100.cu): tokens cross HBM once per direction, second fetch from L2
bool fused_supported(int N, int D, int M);
int fused_trace_fetch(long long* host_out, int n);
size_t fused_workspace_bytes(int N, int D, int M);
int fused_pool_fwd(const void* x, const void* qhl, int J, int B, int N, int D, int M, float* P, float* S, float* rowmax,
                   float* rowsum, int round_p, void* xws, cudaStream_t s);
int fused_pool_bwd(const void* x, const void* dphl, int J, int B, int N, int D, int M, const float* S, const float* rowmax,
                   const float* rowsum, const float* delta, float* part, int* groups_out, void* xws, cudaStream_t s);
}  // namespace ep
