// Generic strided, batched fp32 GEMM on CUDA cores (exact fp32 accumulate).  The general path for the
// small head GEMMs (value projection of pooled tokens, classifier, and their gradients); the
// tensor-core kernels replace it where the layout allows.
#include "ep_common.cuh"

#include <cuda_bf16.h>

#include <algorithm>

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

// out[j] = sum_i a[i][j]   (bias gradients): 8 columns (one 32-byte sector per row) per CTA, 128 row-threads
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ a, int rows, int cols,
                                                      float* __restrict__ out) {
  __shared__ float red[128][9];
  const int j = blockIdx.x * 8 + threadIdx.x;
  float s = 0.f;
  if (j < cols)
    for (int i = threadIdx.y; i < rows; i += 128) s += a[(size_t)i * cols + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  for (int h = 64; h > 0; h >>= 1) {
    if (threadIdx.y < h) red[threadIdx.y][threadIdx.x] += red[threadIdx.y + h][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && j < cols) out[j] = red[0][threadIdx.x];
}
int launch_colsum(const float* a, int rows, int cols, float* out, cudaStream_t s) {
  colsum_kernel<<<(cols + 7) / 8, dim3(8, 128), 0, s>>>(a, rows, cols, out);
  EP_LAUNCH_CHECK();
  return 0;
}

// out[i] = a[i] . b[i], one warp per row   (delta = dP . P)
// delta[b, m] = sum_j g[b, m*c + j] * (out[b, m*c + j] - bias[m*c + j]).  Since out - bias = W_m . P[b, m], this
// equals dP[b, m] . P[b, m] = sum_n A dA (the softmax-backward row term) without touching the (B, M, D) tensors.
__global__ void delta_from_out_kernel(const float* __restrict__ g, const float* __restrict__ out,
                                      const float* __restrict__ bias, long long rows, int M, int c,
                                      float* __restrict__ delta) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // r = b * M + m
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int m = (int)(r % M);
  const float* gp = g + r * c;            // (B, M*c) row-major: element (b, m*c + j) sits at (b*M + m)*c + j
  const float* op = out + r * c;
  float s = 0.f;
  for (int j = lane; j < c; j += 32) s = fmaf(__ldg(gp + j), __ldg(op + j) - (bias ? __ldg(bias + m * c + j) : 0.f), s);
  s = warp_sum(s);
  if (lane == 0) delta[r] = s;
}
int launch_delta_from_out(const float* g, const float* out, const float* bias, long long rows, int M, int c,
                          float* delta, cudaStream_t s) {
  delta_from_out_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(g, out, bias, rows, M, c, delta);
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

constexpr int TN_BI = 64, TN_BJ = 128, TN_BK = 32, TN_LDA = TN_BI + 8, TN_LDB = TN_BJ + 8, TN_STAGES = 4;
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
#pragma unroll
    for (int h = 0; h < TN_BK / 16; ++h) {  // A chunk: BK k x 64 i, 16 k-rows per pass
      const int k = (threadIdx.x >> 4) + 16 * h, i = (threadIdx.x & 15) * 4;
      const bool ok = k0 + k < g.K && i0 + i < g.I;
      cp_async16(As + k * TN_LDA + i, ok ? A + (long long)(k0 + k) * g.lda + i0 + i : A, ok ? 16 : 0);
    }
#pragma unroll
    for (int h = 0; h < TN_BK / 8; ++h) {  // B chunk: BK k x 128 j, 8 k-rows per pass
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
  EP_CUDA(cudaFuncSetAttribute(gemm_tn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((J + TN_BJ - 1) / TN_BJ, (I + TN_BI - 1) / TN_BI, Z);
  gemm_tn_mma_kernel<<<grid, 256, smem, s>>>(g);
  EP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 operand preparation for the tcgen05 GEMMs.  x = big + small with big = tf32(x), small =
// tf32(x - big); A.B ~= A_big.B_big + A_small.B_big + A_big.B_small is evaluated as ONE TF32 GEMM over
// a contraction three times as long:  A' = [big | small | big],  B' = [big | big | small]  (along K).
// Both operands are small (weights, (B, D') activations), so the copies cost a few microseconds and
// the product is accurate to ~1e-6 instead of TF32's 3e-4.
//   kind 0: A-type, kind 1: B-type.  transpose: dst rows are src columns (weights read MN-major).
// ------------------------------------------------------------------------------------------------
template <typename T> struct Split3;
template <> struct Split3<float> {                      // tf32 big / small
  __device__ static void split(float x, float& big, float& small) { big = round_tf32(x); small = round_tf32(x - big); }
};
template <> struct Split3<__nv_bfloat16> {              // bf16 hi / lo
  __device__ static void split(float x, __nv_bfloat16& big, __nv_bfloat16& small) {
    big = __float2bfloat16_rn(x);
    small = __float2bfloat16_rn(x - __bfloat162float(big));
  }
};
// (K here is the segment stride: the copies may pad each third to a multiple of the GEMM's K step)
template <typename T>
__device__ __forceinline__ void split3_store(T* drow, int K, int k, float x, int kind) {
  T big, small;
  Split3<T>::split(x, big, small);
  drow[k] = big;
  drow[K + k] = kind == 0 ? small : big;
  drow[2 * K + k] = kind == 0 ? big : small;
}

// src [R x K] (row stride ld) -> dst [R x 3Kp], each third zero-padded from K to Kp columns
template <typename T>
__global__ void split3_kernel(const float* __restrict__ src, T* __restrict__ dst, long long R, int K, int Kp, long long ld,
                              int kind) {
  const long long total = R * Kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kp;
    const int k = (int)(i - r * Kp);
    split3_store<T>(dst + r * 3 * Kp, Kp, k, k < K ? __ldg(src + r * ld + k) : 0.f, kind);
  }
}
// bf16 copies with K, Kp, ld all multiples of 8: 8 elements per thread, 16-byte stores
__global__ void split3_bf16x8_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long R, int K,
                                     int Kp, long long ld, int kind) {
  const int k8 = Kp >> 3;
  const long long total = R * k8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k8;
    const int k = (int)(i - r * k8) << 3;
    float v[8];
    if (k < K) load8(src + r * ld + k, v);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) Split3<__nv_bfloat16>::split(v[e], hi[e], lo[e]);
    __nv_bfloat16* drow = dst + r * 3 * Kp + k;
    *reinterpret_cast<uint4*>(drow) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(drow + Kp) = *reinterpret_cast<const uint4*>(kind == 0 ? lo : hi);
    *reinterpret_cast<uint4*>(drow + 2 * Kp) = *reinterpret_cast<const uint4*>(kind == 0 ? hi : lo);
  }
}
int launch_split3(const float* src, void* dst, long long R, int K, int Kp, long long ld, int kind, int bf16,
                  cudaStream_t s) {
  if (bf16 && K % 8 == 0 && Kp % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 31) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long total8 = R * (Kp >> 3);
    const unsigned grid = (unsigned)std::min<long long>((total8 + 255) / 256, 8 * kNumSMs);
    split3_bf16x8_kernel<<<grid, 256, 0, s>>>(src, (__nv_bfloat16*)dst, R, K, Kp, ld, kind);
    EP_LAUNCH_CHECK();
    return 0;
  }
  const long long total = R * Kp;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 8 * kNumSMs);
  if (bf16) split3_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, (__nv_bfloat16*)dst, R, K, Kp, ld, kind);
  else split3_kernel<float><<<grid, 256, 0, s>>>(src, (float*)dst, R, K, Kp, ld, kind);
  EP_LAUNCH_CHECK();
  return 0;
}

// src[z] is [K x R] (K rows, row stride R); dst[z] is [R x 3Kp]  (the transposed, K-major copy, thirds padded to Kp)
template <typename T>
__global__ void split3_transpose_kernel(const float* __restrict__ src, T* __restrict__ dst, int K, int Kp, int R,
                                        long long src_z, long long dst_z, int kind) {
  __shared__ float tile[32][33];
  const float* sp = src + (long long)blockIdx.z * src_z;
  T* dp = dst + (long long)blockIdx.z * dst_z;
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, r = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && r < R) ? sp[(long long)k * R + r] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, k = k0 + threadIdx.x;
    if (r < R && k < Kp) split3_store<T>(dp + (long long)r * 3 * Kp, Kp, k, tile[threadIdx.x][i], kind);   // pad = 0
  }
}
int launch_split3_transpose(const float* src, void* dst, int K, int Kp, int R, int Z, long long src_z, long long dst_z,
                            int kind, int bf16, cudaStream_t s) {
  const dim3 grid((R + 31) / 32, (Kp + 31) / 32, Z), blk(32, 8);
  if (bf16) split3_transpose_kernel<__nv_bfloat16><<<grid, blk, 0, s>>>(src, (__nv_bfloat16*)dst, K, Kp, R, src_z, dst_z, kind);
  else split3_transpose_kernel<float><<<grid, blk, 0, s>>>(src, (float*)dst, K, Kp, R, src_z, dst_z, kind);
  EP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// C[z][i][j] = sum_k A[z][i][k] * B[z][j][k] (+ bias[z][j])  ("NT", both K-contiguous) with 3xTF32 done
// in registers: for the one product whose A operand (the pooled tokens P, B*M*D floats) is too large
// to copy.  mma.sync m16n8k8, cp.async pipeline, 64 x 32 tile (the per-query projection is 32 wide).
// ------------------------------------------------------------------------------------------------
constexpr int NT_BI = 128, NT_BJ = 32, NT_BK = 32, NT_LD = NT_BK + 4, NT_STAGES = 3;
constexpr int NT_STAGE_FLOATS = (NT_BI + NT_BJ) * NT_LD;

struct GemmNT {
  const float* A; const float* B; float* C; const float* bias;
  int I, J, K;
  long long lda, ldb, ldc, a_z, b_z, c_z, bias_z;
};

__global__ void __launch_bounds__(128) gemm_nt_3xtf32_kernel(GemmNT g) {
  extern __shared__ __align__(16) float nt_smem[];
  const int z = blockIdx.z;
  const float* A = g.A + (long long)z * g.a_z;
  const float* B = g.B + (long long)z * g.b_z;
  float* C = g.C + (long long)z * g.c_z;
  const int i0 = blockIdx.y * NT_BI, j0 = blockIdx.x * NT_BJ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = warp * 32;                                    // 4 warps x 32 rows (2 m-tiles), all 32 columns
  const int gq = lane >> 2, tq = lane & 3;
  float acc[2][4][4] = {};

  auto issue_stage = [&](int st, int k0) {
    float* As = nt_smem + st * NT_STAGE_FLOATS;
    float* Bs = As + NT_BI * NT_LD;
#pragma unroll
    for (int h = 0; h < 8; ++h) {                              // A: 128 rows x 32 k = 1024 float4
      const int e = threadIdx.x + 128 * h;
      const int i = e >> 3, k = (e & 7) * 4;
      const bool ok = i0 + i < g.I && k0 + k < g.K;
      cp_async16(As + i * NT_LD + k, ok ? A + (long long)(i0 + i) * g.lda + k0 + k : A, ok ? 16 : 0);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {                              // B: 32 rows x 32 k = 256 float4
      const int e = threadIdx.x + 128 * h;
      const int j = e >> 3, k = (e & 7) * 4;
      const bool ok = j0 + j < g.J && k0 + k < g.K;
      cp_async16(Bs + j * NT_LD + k, ok ? B + (long long)(j0 + j) * g.ldb + k0 + k : B, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nk = (g.K + NT_BK - 1) / NT_BK;
#pragma unroll
  for (int st = 0; st < NT_STAGES - 1; ++st) {
    if (st < nk) issue_stage(st, st * NT_BK);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kc = 0; kc < nk; ++kc) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NT_STAGES - 2) : "memory");
    __syncthreads();
    if (kc + NT_STAGES - 1 < nk) issue_stage((kc + NT_STAGES - 1) % NT_STAGES, (kc + NT_STAGES - 1) * NT_BK);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* As = nt_smem + (kc % NT_STAGES) * NT_STAGE_FLOATS;
    const float* Bs = As + NT_BI * NT_LD;
#pragma unroll
    for (int kk = 0; kk < NT_BK; kk += 8) {
      uint32_t ab[2][4], as_[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r0 = wi + mi * 16 + gq;
        const float af[4] = {As[r0 * NT_LD + kk + tq], As[(r0 + 8) * NT_LD + kk + tq], As[r0 * NT_LD + kk + tq + 4],
                             As[(r0 + 8) * NT_LD + kk + tq + 4]};
#pragma unroll
        for (int e = 0; e < 4; ++e) { ab[mi][e] = f2tf32(af[e]); as_[mi][e] = f2tf32(af[e] - __uint_as_float(ab[mi][e])); }
      }
#define EP_MMA_TF32(ACC, AA, BB)                                                                                   \
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
               : "+f"(ACC[0]), "+f"(ACC[1]), "+f"(ACC[2]), "+f"(ACC[3])                                             \
               : "r"(AA[0]), "r"(AA[1]), "r"(AA[2]), "r"(AA[3]), "r"(BB[0]), "r"(BB[1]))
      uint32_t bb[4][2], bs[4][2];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const float bf[2] = {Bs[(ni * 8 + gq) * NT_LD + kk + tq], Bs[(ni * 8 + gq) * NT_LD + kk + tq + 4]};
#pragma unroll
        for (int e = 0; e < 2; ++e) { bb[ni][e] = f2tf32(bf[e]); bs[ni][e] = f2tf32(bf[e] - __uint_as_float(bb[ni][e])); }
      }
      // the three products are issued product-major so that back-to-back MMAs hit different accumulators
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) EP_MMA_TF32(acc[mi][ni], as_[mi], bb[ni]);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) EP_MMA_TF32(acc[mi][ni], ab[mi], bs[ni]);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) EP_MMA_TF32(acc[mi][ni], ab[mi], bb[ni]);
#undef EP_MMA_TF32
    }
  }
  const float* bias = g.bias ? g.bias + (long long)z * g.bias_z : nullptr;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = i0 + wi + mi * 16 + gq + h * 8;
        const int j = j0 + ni * 8 + tq * 2;
        if (i < g.I) {
          if (j < g.J) C[(long long)i * g.ldc + j] = acc[mi][ni][2 * h] + (bias ? bias[j] : 0.f);
          if (j + 1 < g.J) C[(long long)i * g.ldc + j + 1] = acc[mi][ni][2 * h + 1] + (bias ? bias[j + 1] : 0.f);
        }
      }
}

int launch_gemm_nt3(const float* A, const float* B, float* C, const float* bias, int I, int J, int K, int Z,
                    long long lda, long long ldb, long long ldc, long long a_z, long long b_z, long long c_z,
                    long long bias_z, cudaStream_t s) {
  if (I <= 0 || J <= 0 || K <= 0 || Z <= 0) return EP_ERR_SHAPE;
  if (((lda | ldb | a_z | b_z) & 3) || (K & 3)) return EP_ERR_ALIGN;
  GemmNT g{A, B, C, bias, I, J, K, lda, ldb, ldc, a_z, b_z, c_z, bias_z};
  const int smem = NT_STAGES * NT_STAGE_FLOATS * (int)sizeof(float);
  EP_CUDA(cudaFuncSetAttribute(gemm_nt_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((J + NT_BJ - 1) / NT_BJ, (I + NT_BI - 1) / NT_BI, Z);
  gemm_nt_3xtf32_kernel<<<grid, 128, smem, s>>>(g);
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace ep
