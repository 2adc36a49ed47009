// Generic strided, batched fp32 GEMM on CUDA cores (exact fp32 accumulate).  The general path for the
// small head GEMMs (value projection of pooled tokens, classifier, and their gradients); the
// tensor-core kernels replace it where the layout allows.
#include "ep_common.cuh"

namespace ep {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_v0_kernel(GemmDesc g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int z = blockIdx.z;
  const float* A = g.A + (long long)z * g.a_z;
  const float* Bp = g.B + (long long)z * g.b_z;
  float* C = g.C + (long long)z * g.c_z;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const bool a_kfast = (g.a_k == 1), b_kfast = (g.b_k == 1);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = threadIdx.x + 256 * r;
      int i, k;
      if (a_kfast) { k = e & (BK - 1); i = e >> 4; } else { i = e & (BM - 1); k = e >> 6; }
      float v = 0.f;
      if (i0 + i < g.I && k0 + k < g.K) v = __ldg(A + (long long)(i0 + i) * g.a_i + (long long)(k0 + k) * g.a_k);
      As[k][i] = v;
      int j;
      if (b_kfast) { k = e & (BK - 1); j = e >> 4; } else { j = e & (BN - 1); k = e >> 6; }
      v = 0.f;
      if (j0 + j < g.J && k0 + k < g.K) v = __ldg(Bp + (long long)(k0 + k) * g.b_k + (long long)(j0 + j) * g.b_j);
      Bs[k][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    __syncthreads();
  }
  const float* bias = g.bias ? g.bias + (long long)z * g.bias_z : nullptr;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + ty * 4 + p;
    if (i >= g.I) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + tx * 4 + q;
      if (j < g.J) C[(long long)i * g.c_i + (long long)j * g.c_j] = acc[p][q] + (bias ? bias[j] : 0.f);
    }
  }
}

int launch_gemm_v0(const GemmDesc& g, cudaStream_t s) {
  if (g.I <= 0 || g.J <= 0 || g.K <= 0 || g.Z <= 0) return EP_ERR_SHAPE;
  dim3 grid((g.J + BN - 1) / BN, (g.I + BM - 1) / BM, g.Z);
  gemm_v0_kernel<<<grid, 256, 0, s>>>(g);
  EP_LAUNCH_CHECK();
  return 0;
}

// out[j] = sum_i a[i][j]   (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ a, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (j < cols)
    for (int i = threadIdx.y; i < rows; i += 8) s += a[(size_t)i * cols + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    out[j] = t;
  }
}
int launch_colsum(const float* a, int rows, int cols, float* out, cudaStream_t s) {
  colsum_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, s>>>(a, rows, cols, out);
  EP_LAUNCH_CHECK();
  return 0;
}

// out[i] = a[i] . b[i], one warp per row   (delta = dP . P)
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long rows, int cols,
                              float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* pa = reinterpret_cast<const float4*>(a + r * cols);
  const float4* pb = reinterpret_cast<const float4*>(b + r * cols);
  float s = 0.f;
  for (int c = lane; c < cols / 4; c += 32) {
    const float4 u = __ldg(pa + c), v = __ldg(pb + c);
    s += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
  }
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}
int launch_rowdot(const float* a, const float* b, long long rows, int cols, float* out, cudaStream_t s) {
  rowdot_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(a, b, rows, cols, out);
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace ep
