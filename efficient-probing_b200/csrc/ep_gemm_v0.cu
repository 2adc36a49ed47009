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

// ------------------------------------------------------------------------------------------------
// C[z][i][j] = sum_k A[z][k][i] * B[z][k][j]  ("TN": both operands have the contraction index as their
// slow dimension -- weight gradients, contraction over the batch).  TF32 mma.sync m16n8k8 with the
// operands rounded to tf32 in registers, fp32 accumulate; tiles are staged through shared memory as
// they lie in global memory, so no transposed copy of the (large) activation operand is needed.
// ------------------------------------------------------------------------------------------------
namespace ep {

constexpr int TN_BI = 64, TN_BJ = 128, TN_BK = 16, TN_LDA = TN_BI + 8, TN_LDB = TN_BJ + 8, TN_STAGES = 4;
constexpr int TN_STAGE_FLOATS = TN_BK * (TN_LDA + TN_LDB);

__device__ __forceinline__ uint32_t f2tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
// 16-byte async copy global -> shared; bytes beyond src_bytes are zero-filled (src_bytes in {0, 16})
__device__ __forceinline__ void cp_async16(float* dst, const float* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

struct GemmTN {
  const float* A; const float* B; float* C;
  int I, J, K;
  long long lda, ldb, ldc, a_z, b_z, c_z;
};

// requires I % 4 == 0, J % 4 == 0, lda/ldb/a_z/b_z % 4 == 0 and 16-byte aligned bases
__global__ void __launch_bounds__(256) gemm_tn_mma_kernel(GemmTN g) {
  extern __shared__ __align__(16) float tn_smem[];
  const int z = blockIdx.z;
  const float* A = g.A + (long long)z * g.a_z;
  const float* B = g.B + (long long)z * g.b_z;
  float* C = g.C + (long long)z * g.c_z;
  const int i0 = blockIdx.y * TN_BI, j0 = blockIdx.x * TN_BJ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = (warp >> 2) * 32, wj = (warp & 3) * 32;      // warp tile origin inside the CTA tile
  const int gq = lane >> 2, tq = lane & 3;
  float acc[2][4][4] = {};

  auto issue_stage = [&](int st, int k0) {
    float* As = tn_smem + st * TN_STAGE_FLOATS;
    float* Bs = As + TN_BK * TN_LDA;
    {  // A chunk: 16 k x 64 i = 256 float4, one per thread
      const int k = threadIdx.x >> 4, i = (threadIdx.x & 15) * 4;
      const bool ok = k0 + k < g.K && i0 + i < g.I;
      cp_async16(As + k * TN_LDA + i, ok ? A + (long long)(k0 + k) * g.lda + i0 + i : A, ok ? 16 : 0);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // B chunk: 16 k x 128 j = 512 float4, two per thread
      const int e = threadIdx.x + 256 * h;
      const int k = e >> 5, j = (e & 31) * 4;
      const bool ok = k0 + k < g.K && j0 + j < g.J;
      cp_async16(Bs + k * TN_LDB + j, ok ? B + (long long)(k0 + k) * g.ldb + j0 + j : B, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nk = (g.K + TN_BK - 1) / TN_BK;
#pragma unroll
  for (int st = 0; st < TN_STAGES - 1; ++st) {
    if (st < nk) issue_stage(st, st * TN_BK);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group %0;" ::"n"(TN_STAGES - 2) : "memory");
    __syncthreads();
    if (kc + TN_STAGES - 1 < nk) issue_stage((kc + TN_STAGES - 1) % TN_STAGES, (kc + TN_STAGES - 1) * TN_BK);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* As = tn_smem + (kc % TN_STAGES) * TN_STAGE_FLOATS;
    const float* Bs = As + TN_BK * TN_LDA;
#pragma unroll
    for (int kk = 0; kk < TN_BK; kk += 8) {
      uint32_t a[2][4], b[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int i = wi + mi * 16 + gq;
        a[mi][0] = f2tf32(As[(kk + tq) * TN_LDA + i]);
        a[mi][1] = f2tf32(As[(kk + tq) * TN_LDA + i + 8]);
        a[mi][2] = f2tf32(As[(kk + tq + 4) * TN_LDA + i]);
        a[mi][3] = f2tf32(As[(kk + tq + 4) * TN_LDA + i + 8]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int j = wj + ni * 8 + gq;
        b[ni][0] = f2tf32(Bs[(kk + tq) * TN_LDB + j]);
        b[ni][1] = f2tf32(Bs[(kk + tq + 4) * TN_LDB + j]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          asm volatile(
              "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
              : "+f"(acc[mi][ni][0]), "+f"(acc[mi][ni][1]), "+f"(acc[mi][ni][2]), "+f"(acc[mi][ni][3])
              : "r"(a[mi][0]), "r"(a[mi][1]), "r"(a[mi][2]), "r"(a[mi][3]), "r"(b[ni][0]), "r"(b[ni][1]));
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = i0 + wi + mi * 16 + gq + h * 8;
        const int j = j0 + wj + ni * 8 + tq * 2;
        if (i < g.I && j < g.J)                                 // J even: j and j + 1 are both in range
          *reinterpret_cast<float2*>(C + (long long)i * g.ldc + j) = make_float2(acc[mi][ni][2 * h], acc[mi][ni][2 * h + 1]);
      }
}

bool gemm_tn_ok(int I, int J, long long lda, long long ldb, long long ldc, long long a_z, long long b_z, long long c_z) {
  return I % 4 == 0 && J % 4 == 0 && ((lda | ldb | a_z | b_z) & 3) == 0 && ((ldc | c_z) & 1) == 0;
}

int launch_gemm_tn(const float* A, const float* B, float* C, int I, int J, int K, int Z, long long lda, long long ldb,
                   long long ldc, long long a_z, long long b_z, long long c_z, cudaStream_t s) {
  if (I <= 0 || J <= 0 || K <= 0 || Z <= 0) return EP_ERR_SHAPE;
  if (!gemm_tn_ok(I, J, lda, ldb, ldc, a_z, b_z, c_z)) return EP_ERR_ALIGN;
  GemmTN g{A, B, C, I, J, K, lda, ldb, ldc, a_z, b_z, c_z};
  const int smem = TN_STAGES * TN_STAGE_FLOATS * (int)sizeof(float);
  static bool attr = false;
  if (!attr) {
    EP_CUDA(cudaFuncSetAttribute(gemm_tn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  dim3 grid((J + TN_BJ - 1) / TN_BJ, (I + TN_BI - 1) / TN_BI, Z);
  gemm_tn_mma_kernel<<<grid, 256, smem, s>>>(g);
  EP_LAUNCH_CHECK();
  return 0;
}

// dst[z][c][r] = tf32(src[z][r][c]): K-major (transposed) copy of a weight block for the tcgen05 GEMMs
__global__ void transpose_round_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc,
                                       long long src_z, long long dst_z) {
  __shared__ float tile[32][33];
  const float* s = src + (long long)blockIdx.z * src_z;
  float* d = dst + (long long)blockIdx.z * dst_z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? s[(long long)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) d[(long long)c * R + r] = round_tf32(tile[threadIdx.x][i]);
  }
}
int launch_transpose_round(const float* src, float* dst, int R, int Cc, int Z, long long src_z, long long dst_z,
                           cudaStream_t s) {
  transpose_round_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32, Z), dim3(32, 8), 0, s>>>(src, dst, R, Cc, src_z, dst_z);
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace ep
