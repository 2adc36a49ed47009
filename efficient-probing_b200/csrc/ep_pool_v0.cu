// General (any M <= 64, any N that fits shared memory, D % 8 == 0) CUDA-core kernels for the EP
// pooling: one CTA per sample.  These are the always-available path and the on-device cross-check
// for the tcgen05 kernels; they read each sample twice (second read from L2).
//
//   forward  : S = scale * cls . x^T -> softmax over tokens -> P = A x         (poolings/ep.py:35-44,
//              with the value projection moved after the pooling)
//   backward : recompute A from (rowmax,rowsum), dA = dP . x^T, dS = A (dA - delta), dq = dS^T x
#include "ep_common.cuh"

namespace ep {

constexpr int kThreads = 256;
constexpr int kMG = 8;       // queries per work item in the logit phase

__host__ __device__ inline int s_ld(int M) { return ((M + 3) & ~3) + 4; }   // row stride of S in smem (floats)
size_t pool_v0_smem_bytes(int N, int M) { return (size_t)N * s_ld(M) * sizeof(float); }

// logits (and optionally dA) for token pairs x query groups; results left in smem S[n][LD].
template <typename XT, bool kBwd>
__device__ __forceinline__ void logit_phase(const XT* __restrict__ xb, const float* __restrict__ cls,
                                            const float* __restrict__ dPb, float scale, int N, int D, int M,
                                            const float* __restrict__ rmax, const float* __restrict__ rsum,
                                            const float* __restrict__ delta, float* S, float* Aout = nullptr) {
  const int LD = s_ld(M);
  const int groups = (M + kMG - 1) / kMG;
  const int pairs = (N + 1) / 2;
  for (int item = threadIdx.x; item < pairs * groups; item += kThreads) {
    const int pr = item / groups, mg = item - pr * groups;
    const int n0 = 2 * pr, n1 = min(2 * pr + 1, N - 1);
    const XT* x0 = xb + (size_t)n0 * D;
    const XT* x1 = xb + (size_t)n1 * D;
    float s0[kMG], s1[kMG], a0[kMG], a1[kMG];
#pragma unroll
    for (int j = 0; j < kMG; ++j) { s0[j] = s1[j] = a0[j] = a1[j] = 0.f; }
    for (int d = 0; d < D; d += 8) {
      float v0[8], v1[8];
      load8(x0 + d, v0);
      load8(x1 + d, v1);
#pragma unroll
      for (int j = 0; j < kMG; ++j) {
        const int m = min(mg * kMG + j, M - 1);
        float q[8];
        load8(cls + (size_t)m * D + d, q);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s0[j] = fmaf(q[e], v0[e], s0[j]); s1[j] = fmaf(q[e], v1[e], s1[j]); }
        if (kBwd) {
          load8(dPb + (size_t)m * D + d, q);
#pragma unroll
          for (int e = 0; e < 8; ++e) { a0[j] = fmaf(q[e], v0[e], a0[j]); a1[j] = fmaf(q[e], v1[e], a1[j]); }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMG; ++j) {
      const int m = mg * kMG + j;
      if (m >= M) break;
      float r0 = s0[j] * scale, r1 = s1[j] * scale;
      if (kBwd) {
        const float mx = rmax[m], inv = 1.f / rsum[m], dl = delta[m];
        const float p0 = expf(r0 - mx) * inv, p1 = expf(r1 - mx) * inv;     // attention, recomputed
        if (Aout) {
          Aout[n0 * LD + m] = p0;
          if (2 * pr + 1 < N) Aout[n1 * LD + m] = p1;
        }
        r0 = p0 * (a0[j] - dl);
        r1 = p1 * (a1[j] - dl);
      }
      S[n0 * LD + m] = r0;
      if (2 * pr + 1 < N) S[n1 * LD + m] = r1;
    }
  }
}

// acc[m][0..1] += sum_n W[n][m] * x[n][dd..dd+1] for a block of <= 32 queries; W in smem.
template <typename XT>
__device__ __forceinline__ void pool_phase(const XT* __restrict__ xb, const float* S, int N, int D, int M,
                                           int mb, int dd, float (&acc)[32][2]) {
  const int LD = s_ld(M);
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j][0] = acc[j][1] = 0.f;
  const int mcount = min(32, M - mb);
  const int quads = (mcount + 3) / 4;
  for (int n = 0; n < N; ++n) {
    const float2 xv = load2(xb + (size_t)n * D + dd);
    const float4* row = reinterpret_cast<const float4*>(S + n * LD + mb);
#pragma unroll
    for (int qd = 0; qd < 8; ++qd) {
      if (qd < quads) {
        const float4 w = row[qd];
        acc[4 * qd + 0][0] = fmaf(w.x, xv.x, acc[4 * qd + 0][0]); acc[4 * qd + 0][1] = fmaf(w.x, xv.y, acc[4 * qd + 0][1]);
        acc[4 * qd + 1][0] = fmaf(w.y, xv.x, acc[4 * qd + 1][0]); acc[4 * qd + 1][1] = fmaf(w.y, xv.y, acc[4 * qd + 1][1]);
        acc[4 * qd + 2][0] = fmaf(w.z, xv.x, acc[4 * qd + 2][0]); acc[4 * qd + 2][1] = fmaf(w.z, xv.y, acc[4 * qd + 2][1]);
        acc[4 * qd + 3][0] = fmaf(w.w, xv.x, acc[4 * qd + 3][0]); acc[4 * qd + 3][1] = fmaf(w.w, xv.y, acc[4 * qd + 3][1]);
      }
    }
  }
}

template <typename XT>
__global__ void __launch_bounds__(kThreads)
pool_fwd_v0_kernel(const XT* __restrict__ x, const float* __restrict__ cls, long long cls_z, float scale, int N, int D,
                   int M, float* __restrict__ P, float* __restrict__ S_out, float* __restrict__ rowmax,
                   float* __restrict__ rowsum, float* __restrict__ attn, int round_p) {
  extern __shared__ __align__(16) float S[];
  const int b = blockIdx.x, LD = s_ld(M);
  const XT* xb = x + (size_t)b * N * D;
  cls += (size_t)b * cls_z;                                  // per-sample queries (forward(x, cls=...), ep.py:32-33)
  // zero the pad columns so the float4 reads of the pooling phase never see garbage
  for (int i = threadIdx.x; i < N * LD; i += kThreads) S[i] = 0.f;
  __syncthreads();
  logit_phase<XT, false>(xb, cls, nullptr, scale, N, D, M, nullptr, nullptr, nullptr, S);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += kThreads / 32) {
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) {
      const float v = S[n * LD + m];
      mx = fmaxf(mx, v);
      if (S_out) S_out[((size_t)b * M + m) * N + n] = v;         // logits, saved for the backward pass
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) { const float e = expf(S[n * LD + m] - mx); S[n * LD + m] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int n = lane; n < N; n += 32) {
      const float a = S[n * LD + m] * inv;
      S[n * LD + m] = a;
      if (attn) attn[((size_t)b * M + m) * N + n] = a;
    }
    if (lane == 0) { rowmax[(size_t)b * M + m] = mx; rowsum[(size_t)b * M + m] = sum; }
  }
  __syncthreads();
  if (P == nullptr) return;
  for (int mb = 0; mb < M; mb += 32) {
    for (int dd = threadIdx.x * 2; dd < D; dd += 2 * kThreads) {
      float acc[32][2];
      pool_phase<XT>(xb, S, N, D, M, mb, dd, acc);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (mb + j < M)
          *reinterpret_cast<float2*>(P + ((size_t)b * M + mb + j) * D + dd) =
              round_p ? make_float2(round_tf32(acc[j][0]), round_tf32(acc[j][1]))   // P feeds TF32 GEMMs
                      : make_float2(acc[j][0], acc[j][1]);
    }
  }
}

// dx[n, d] = sum_m ( dS[n, m] * scale * cls[m, d] + A[n, m] * dP[m, d] )   for one sample; dS and A in smem.
// Thread <-> channel d (coalesced rows), the sample's query-side rows (scaled queries, dP) in registers.
template <typename XT, int kMB>
__device__ __forceinline__ void dx_phase(XT* __restrict__ dxb, const float* __restrict__ cls,
                                         const float* __restrict__ dPb, float scale, int N, int D, int M,
                                         const float* dS, const float* A) {
  const int LD = s_ld(M);
  for (int d = threadIdx.x; d < D; d += kThreads) {
    float q[kMB], pv[kMB];
#pragma unroll
    for (int m = 0; m < kMB; ++m) {
      q[m] = m < M ? scale * __ldg(cls + (size_t)m * D + d) : 0.f;
      pv[m] = m < M ? __ldg(dPb + (size_t)m * D + d) : 0.f;
    }
    for (int n = 0; n < N; ++n) {
      const float4* rs = reinterpret_cast<const float4*>(dS + n * LD);
      const float4* ra = reinterpret_cast<const float4*>(A + n * LD);
      float acc = 0.f;
#pragma unroll
      for (int m4 = 0; m4 < kMB / 4; ++m4) {
        if (4 * m4 < M) {                                    // (pad columns of the smem rows are zero)
          const float4 s4 = rs[m4], a4 = ra[m4];
          acc = fmaf(s4.x, q[4 * m4 + 0], acc); acc = fmaf(a4.x, pv[4 * m4 + 0], acc);
          acc = fmaf(s4.y, q[4 * m4 + 1], acc); acc = fmaf(a4.y, pv[4 * m4 + 1], acc);
          acc = fmaf(s4.z, q[4 * m4 + 2], acc); acc = fmaf(a4.z, pv[4 * m4 + 2], acc);
          acc = fmaf(s4.w, q[4 * m4 + 3], acc); acc = fmaf(a4.w, pv[4 * m4 + 3], acc);
        }
      }
      dxb[(size_t)n * D + d] = (XT)acc;
    }
  }
}

template <typename XT>
__global__ void __launch_bounds__(kThreads)
pool_bwd_v0_kernel(const XT* __restrict__ x, const float* __restrict__ cls, long long cls_z, float scale, int N, int D,
                   int M, const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                   const float* __restrict__ dP, const float* __restrict__ delta,
                   float* __restrict__ dq_slots, int n_slots, float* __restrict__ d_cls_b, XT* __restrict__ dx) {
  extern __shared__ __align__(16) float S[];
  const int b = blockIdx.x, LD = s_ld(M);
  const XT* xb = x + (size_t)b * N * D;
  cls += (size_t)b * cls_z;
  float* A = dx ? S + (size_t)N * LD : nullptr;               // the attention is kept only for the input gradient
  for (int i = threadIdx.x; i < N * LD * (dx ? 2 : 1); i += kThreads) S[i] = 0.f;
  __syncthreads();
  logit_phase<XT, true>(xb, cls, dP + (size_t)b * M * D, scale, N, D, M, rowmax + (size_t)b * M,
                        rowsum + (size_t)b * M, delta + (size_t)b * M, S, A);
  __syncthreads();
  // per-sample queries: their gradient is per sample too (plain stores); shared queries: summed over the batch
  float* slot = d_cls_b ? d_cls_b + (size_t)b * M * D : dq_slots + (size_t)(b % n_slots) * M * D;
  for (int mb = 0; mb < M; mb += 32) {
    for (int dd = threadIdx.x * 2; dd < D; dd += 2 * kThreads) {
      float acc[32][2];
      pool_phase<XT>(xb, S, N, D, M, mb, dd, acc);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (mb + j < M) {
          if (d_cls_b) {
            *reinterpret_cast<float2*>(slot + (size_t)(mb + j) * D + dd) = make_float2(acc[j][0] * scale, acc[j][1] * scale);
          } else {
            atomicAdd(slot + (size_t)(mb + j) * D + dd, acc[j][0]);
            atomicAdd(slot + (size_t)(mb + j) * D + dd + 1, acc[j][1]);
          }
        }
    }
  }
  if (dx) {
    XT* dxb = dx + (size_t)b * N * D;
    const float* dPb = dP + (size_t)b * M * D;
    if (M <= 32) dx_phase<XT, 32>(dxb, cls, dPb, scale, N, D, M, S, A);
    else dx_phase<XT, 64>(dxb, cls, dPb, scale, N, D, M, S, A);
  }
}

__global__ void reduce_slots_kernel(const float* __restrict__ slots, int n_slots, size_t n, float scale,
                                    float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < n_slots; ++k) s += slots[(size_t)k * n + i];
  out[i] = s * scale;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return EP_ERR_UNSUPPORTED;
  if (bytes > 48 * 1024) EP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int pool_fwd_v0(const void* x, int x_dtype, const float* cls, float scale, int B, int N, int D, int M,
                float* P, float* S_out, float* rowmax, float* rowsum, float* attn, int round_p, cudaStream_t s,
                int cls_batched) {
  const long long cls_z = cls_batched ? (long long)M * D : 0;
  if (M > 64) return EP_ERR_UNSUPPORTED;
  const size_t smem = pool_v0_smem_bytes(N, M);
  int rc;
  if (x_dtype == EP_DTYPE_BF16) {
    if ((rc = set_smem(pool_fwd_v0_kernel<__nv_bfloat16>, smem))) return rc;
    pool_fwd_v0_kernel<__nv_bfloat16><<<B, kThreads, smem, s>>>((const __nv_bfloat16*)x, cls, cls_z, scale, N, D, M, P,
                                                                 S_out, rowmax, rowsum, attn, round_p);
  } else {
    if ((rc = set_smem(pool_fwd_v0_kernel<float>, smem))) return rc;
    pool_fwd_v0_kernel<float><<<B, kThreads, smem, s>>>((const float*)x, cls, cls_z, scale, N, D, M, P, S_out, rowmax, rowsum,
                                                        attn, round_p);
  }
  EP_LAUNCH_CHECK();
  return 0;
}

int pool_bwd_v0(const void* x, int x_dtype, const float* cls, float scale, int B, int N, int D, int M,
                const float* rowmax, const float* rowsum, const float* dP, const float* delta,
                float* dq_slots, int n_slots, float* d_cls, cudaStream_t s, int cls_batched, void* dx) {
  if (M > 64) return EP_ERR_UNSUPPORTED;
  const size_t smem = pool_v0_smem_bytes(N, M) * (dx ? 2 : 1);
  const size_t n = (size_t)M * D;
  const long long cls_z = cls_batched ? (long long)M * D : 0;
  float* d_cls_b = cls_batched ? d_cls : nullptr;             // (B, M, D) written directly
  if (!cls_batched) EP_CUDA(cudaMemsetAsync(dq_slots, 0, n * n_slots * sizeof(float), s));
  int rc;
  if (x_dtype == EP_DTYPE_BF16) {
    if ((rc = set_smem(pool_bwd_v0_kernel<__nv_bfloat16>, smem))) return rc;
    pool_bwd_v0_kernel<__nv_bfloat16><<<B, kThreads, smem, s>>>((const __nv_bfloat16*)x, cls, cls_z, scale, N, D, M, rowmax,
                                                                 rowsum, dP, delta, dq_slots, n_slots, d_cls_b,
                                                                 (__nv_bfloat16*)dx);
  } else {
    if ((rc = set_smem(pool_bwd_v0_kernel<float>, smem))) return rc;
    pool_bwd_v0_kernel<float><<<B, kThreads, smem, s>>>((const float*)x, cls, cls_z, scale, N, D, M, rowmax, rowsum, dP,
                                                         delta, dq_slots, n_slots, d_cls_b, (float*)dx);
  }
  EP_LAUNCH_CHECK();
  if (!cls_batched) {
    reduce_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dq_slots, n_slots, n, scale, d_cls);
    EP_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace ep
