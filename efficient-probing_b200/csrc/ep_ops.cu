// Operand copies written by the kernels that produce the data (ABI 2, the *_ops entry points of include/ep_b200.h).
//
// The tcgen05 GEMMs of the head take fp32 tensors as bf16 hi/lo copies; made on demand by every consumer they were
// eight launches (~43 us) of a 0.6 ms training step.  Here:
//   refresh_kernel  every weight-derived copy in ONE launch, once per step right after the optimizer: the scaled
//                   queries as hi/lo rows (one-pass forward), v.weight as [hi|hi|lo] rows (projection) and per query
//                   transposed (dP = g . W_m), fc.weight as [hi|hi|lo] rows (logits) and transposed (dy);
//   g_ops_kernel    everything the projection backward needs from g = dL/d out in one pass over g and out:
//                   delta[b, m] = g[b, m] . (out[b, m] - bias[m])  (the softmax-backward row term, ep_api.cu),
//                   g as [hi|lo|hi] rows per (sample, query) (A operand of dP) and g^T as [hi|hi|lo] rows (B operand
//                   of dW_v, contraction over the batch).
// The layouts are the ones launch_split3 / launch_split3_transpose / split_hilo_kernel write (ep_gemm_v0.cu,
// ep_pool_sm100.cu): consumers cannot tell who made a copy.
#include "ep_common.cuh"

#include <algorithm>

namespace ep {

namespace {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ void split_tf32(float x, float& big, float& small) {
  big = round_tf32(x);
  small = round_tf32(x - big);
}
// thirds of a 3-term operand row: kind 0 (A side) = [big | small | big], kind 1 (B side) = [big | big | small]
template <typename T>
__device__ __forceinline__ void store3(T* drow, int Kp, int k, T big, T small, int kind) {
  drow[k] = big;
  drow[Kp + k] = kind == 0 ? small : big;
  drow[2 * Kp + k] = kind == 0 ? big : small;
}

// ---- job bodies: `vb` of `nvb` virtual blocks of 256 threads ------------------------------------------------

// src [R x K] (row stride ld) -> dst [R x 3Kp] bf16, thirds zero-padded to Kp; K, Kp, ld multiples of 8
__device__ void job_rows_bf16x8(const RefreshJob& j, int vb, int nvb) {
  const int k8 = j.Kp >> 3;
  const long long total = (long long)j.R * k8;
  __nv_bfloat16* dst = (__nv_bfloat16*)j.dst;
  for (long long i = (long long)vb * 256 + threadIdx.x; i < total; i += (long long)nvb * 256) {
    const long long r = i / k8;
    const int k = (int)(i - r * k8) << 3;
    float v[8];
    if (k < j.K) load8(j.src + r * j.ld + k, v);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
    __nv_bfloat16* drow = dst + r * 3 * j.Kp + k;
    *reinterpret_cast<uint4*>(drow) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(drow + j.Kp) = *reinterpret_cast<const uint4*>(j.kind == 0 ? lo : hi);
    *reinterpret_cast<uint4*>(drow + 2 * j.Kp) = *reinterpret_cast<const uint4*>(j.kind == 0 ? hi : lo);
  }
}

// src[z] [K x R] (row stride R) -> dst[z] [R x 3Kp]: one 32 x 32 tile per virtual block, threads as (32, 8)
template <typename T>
__device__ void job_transpose(const RefreshJob& j, int vb, float (*tile)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nrx = (j.R + 31) / 32, nky = (j.Kp + 31) / 32;
  const int bx = vb % nrx, by = (vb / nrx) % nky, bz = vb / (nrx * nky);
  const float* sp = j.src + (long long)bz * j.src_z;
  T* dp = (T*)j.dst + (long long)bz * j.dst_z;
  const int r0 = bx * 32, k0 = by * 32;
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, r = r0 + tx;
    tile[i][tx] = (k < j.K && r < j.R) ? sp[(long long)k * j.R + r] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, k = k0 + tx;
    if (r < j.R && k < j.Kp) {
      const float x = tile[tx][i];                                // zero in the padding
      T* drow = dp + (long long)r * 3 * j.Kp;
      if constexpr (sizeof(T) == 2) {
        __nv_bfloat16 hi, lo;
        split_bf16(x, hi, lo);
        store3<__nv_bfloat16>(drow, j.Kp, k, hi, lo, j.kind);
      } else {
        float big, small;
        split_tf32(x, big, small);
        store3<float>(drow, j.Kp, k, big, small, j.kind);
      }
    }
  }
}

// dst[(2m + {0, 1}) * D + d] = hi / lo of scale * src[m * D + d]; rows 2M .. J-1 zero   (split_hilo_kernel's layout)
__device__ void job_hilo(const RefreshJob& j, int vb, int nvb) {
  const int pairs = j.J / 2, D = j.D;
  __nv_bfloat16* dst = (__nv_bfloat16*)j.dst;
  for (size_t i = (size_t)vb * 256 + threadIdx.x; i < (size_t)pairs * D / 4; i += (size_t)nvb * 256) {
    const int m = (int)(i / (D / 4)), d = (int)(i % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < j.M) v = *reinterpret_cast<const float4*>(j.src + (size_t)m * D + d);
    const float f[4] = {v.x * j.scale, v.y * j.scale, v.z * j.scale, v.w * j.scale};
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(f[e], hi[e], lo[e]);
    __nv_bfloat16* ph = dst + (size_t)(2 * m) * D + d;
    *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(ph + D) = *reinterpret_cast<const uint2*>(lo);
  }
}

struct RefreshArgs {
  RefreshJob job[kMaxRefreshJobs];
  int begin[kMaxRefreshJobs + 1];      // first block of each job; begin[n] = grid size
  int n;
};

__global__ void __launch_bounds__(256) refresh_kernel(const RefreshArgs a) {
  __shared__ float tile[32][33];
  pdl_trigger();
  pdl_wait();
  int q = 0;
  while (q + 1 < a.n && (int)blockIdx.x >= a.begin[q + 1]) ++q;
  const RefreshJob& j = a.job[q];
  const int vb = (int)blockIdx.x - a.begin[q], nvb = a.begin[q + 1] - a.begin[q];
  switch (j.type) {
    case REFRESH_ROWS: job_rows_bf16x8(j, vb, nvb); break;
    case REFRESH_TRANSPOSE:
      if (j.bf16) job_transpose<__nv_bfloat16>(j, vb, tile);
      else job_transpose<float>(j, vb, tile);
      break;
    default: job_hilo(j, vb, nvb); break;
  }
}

// ---- g_ops ---------------------------------------------------------------------------------------------------
// One CTA per (32 samples, query m); 8 warps, warp w owns samples w, w + 8, w + 16, w + 24 of the tile; the query's
// c channels are walked in chunks of 32 (lane = channel).  g3 row r = b * M + m holds [hi | lo | hi] thirds of width
// c (bf16, or tf32 big/small as fp32 when c % 8 != 0); g3t row f = m * c + j holds [hi | hi | lo] thirds of width B.
__global__ void __launch_bounds__(256)
g_ops_kernel(const float* __restrict__ g, const float* __restrict__ out, const float* __restrict__ bias, int B, int M, int c,
             float* __restrict__ delta, void* __restrict__ g3, int g3_bf16, __nv_bfloat16* __restrict__ g3t) {
  __shared__ float tile[32][33];                                   // [sample][channel] of the current chunk
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b0 = blockIdx.x * 32, m = blockIdx.y;
  const int Dp = M * c;
  float dl[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < c; j0 += 32) {
    const int j = j0 + lane;
    const bool jok = j < c;
    const float bj = (bias && jok) ? __ldg(bias + m * c + j) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int bl = w + 8 * q, b = b0 + bl;
      float gv = 0.f, ov = 0.f;
      if (jok && b < B) {
        gv = __ldg(g + (size_t)b * Dp + m * c + j);
        ov = __ldg(out + (size_t)b * Dp + m * c + j);
      }
      dl[q] = fmaf(gv, ov - bj, dl[q]);
      tile[bl][lane] = gv;
      if (g3 && jok && b < B) {
        const size_t r = (size_t)b * M + m;
        if (g3_bf16) {
          __nv_bfloat16 hi, lo;
          split_bf16(gv, hi, lo);
          store3<__nv_bfloat16>((__nv_bfloat16*)g3 + r * 3 * c, c, j, hi, lo, 0);
        } else {
          float big, small;
          split_tf32(gv, big, small);
          store3<float>((float*)g3 + r * 3 * c, c, j, big, small, 0);
        }
      }
    }
    if (g3t) {
      __syncthreads();
      // transposed: warp w owns channels w, w + 8, ... of the chunk, lane = sample -> 64 contiguous bytes per third
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int jl = w + 8 * q, jj = j0 + jl, b = b0 + lane;
        if (jj < c && b < B) {
          __nv_bfloat16 hi, lo;
          split_bf16(tile[lane][jl], hi, lo);
          store3<__nv_bfloat16>(g3t + (size_t)(m * c + jj) * 3 * B, B, b, hi, lo, 1);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float s = warp_sum(dl[q]);
    const int b = b0 + w + 8 * q;
    if (lane == 0 && b < B) delta[(size_t)b * M + m] = s;
  }
}

}  // namespace

int launch_refresh(const RefreshJob* jobs, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  if (n > kMaxRefreshJobs) return EP_ERR_SHAPE;
  RefreshArgs a;
  a.n = n;
  int total = 0;
  for (int q = 0; q < n; ++q) {
    const RefreshJob& j = jobs[q];
    a.job[q] = j;
    a.begin[q] = total;
    long long nb;
    if (j.type == REFRESH_ROWS) {
      if (!j.bf16 || j.K % 8 || j.Kp % 8 || j.ld % 8 || (reinterpret_cast<uintptr_t>(j.src) & 15) ||
          (reinterpret_cast<uintptr_t>(j.dst) & 15))
        return EP_ERR_ALIGN;
      nb = std::min<long long>(((long long)j.R * (j.Kp >> 3) + 255) / 256, 2 * kNumSMs);
    } else if (j.type == REFRESH_TRANSPOSE) {
      nb = (long long)((j.R + 31) / 32) * ((j.Kp + 31) / 32) * j.Z;
    } else {
      nb = std::min<long long>(((long long)(j.J / 2) * j.D / 4 + 255) / 256, kNumSMs);
    }
    total += (int)std::max<long long>(1, nb);
  }
  a.begin[n] = total;
  EP_CUDA(launch_pdl(refresh_kernel, dim3(total), dim3(256), 0, s, a));
  EP_LAUNCH_CHECK();
  return 0;
}

int launch_g_ops(const float* g, const float* out, const float* bias, int B, int M, int c, float* delta, void* g3,
                 int g3_bf16, void* g3t, cudaStream_t s) {
  EP_CUDA(launch_pdl(g_ops_kernel, dim3((B + 31) / 32, M), dim3(256), 0, s, g, out, bias, B, M, c, delta, g3, g3_bf16,
                     (__nv_bfloat16*)g3t));
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace ep
