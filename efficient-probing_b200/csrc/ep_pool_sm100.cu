// tcgen05 / TMA kernels for the EP pooling (sm_100a).
//
// The pooling is two contractions over the token tensor x (B, N, D) bf16:
//   logit-type   Z[b,n,j] = sum_d x[b,n,d] * W[b?][j,d]        (S = q.x^T in forward, dA = dP.x^T in backward)
//   pool-type    Y[b?][d,j] = sum_n x[b,n,d] * V[b,j,n]        (P = A x in forward,  dq = dS^T x in backward)
// Both stream x from HBM exactly once through a TMA -> shared-memory ring and feed it to the tensor
// core as the 128-row operand of tcgen05.mma (cta_group::1, M=128, bf16 in, fp32 accumulate in TMEM):
//   ks_kernel : x chunk [128 tokens x 64 d] is the K-major A operand, W chunk [J x 64 d] the K-major B
//               operand; the accumulators of ALL token tiles of a sample live in TMEM while the kernel
//               walks d, so W is streamed once per sample and nothing large is resident.
//   kp_kernel : the SAME 128-byte-swizzled x bytes are read as an MN-major A operand [128 d x 16 tokens]
//               (instruction-descriptor transpose bit); V arrives as operand-ready blocks [J x 64 tokens]
//               (bf16 hi/lo rows, 128B-swizzle baked in) that the upstream kernel wrote to global memory,
//               copied into the ring with one cp.async.bulk each; the (d x J) accumulators stay in TMEM
//               per sample (forward) or for the whole launch (backward, gradients summed over the batch).
// fp32 operands (queries, probabilities, gradients) enter as bf16 hi/lo pairs in adjacent operand
// rows (j = 2m: hi, 2m+1: lo), so every product is exact and the sum carries ~16 mantissa bits; the
// two accumulator columns are added when the result leaves TMEM.
// Warp roles (one CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (two warps per TMEM lane quadrant).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ep_ptx.cuh"
#include "ep_sm100.cuh"

namespace ep {
using namespace ptx;
int g_sm_limit = 0;
int g_debug = 0;        // developer knobs (ep_set_debug, see include/ep_b200.h)

constexpr int kTileRows = 128;    // ks: token rows per MMA tile (UMMA M)
constexpr int kChunkD = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kXChunkBytes = kTileRows * 128;   // 16 KB
constexpr int kTokBlock = 64;     // kp: tokens per operand block (4 MMA K-steps)
constexpr int kBrickBytes = 2 * kTokBlock * 128;   // [64 tokens x 128 d] = two swizzled halves, 16 KB
constexpr int kSmemBudget = 220 * 1024;

struct KSParams {
  int B, N, D, M, J, ntiles, G, ngroups, nchunks, nstages, w_batched, nbuf, bufcols, tmem_cols, nkb, ndelta;
  int tail_rows;         // > 0: the last token tile of a sample is loaded as this many rows only (single tile group)
  float* out;            // mode 0, 2: logits (B, M, N)
  uint8_t* blocks;       // operand blocks [B][nkb][J rows x 64 tokens] bf16 hi/lo, swizzled: dS (mode 1), exp(S - max) (mode 2)
  const float* S;        // mode 1: saved logits
  const float* rmax;     // mode 1
  const float* rsum;     // mode 1
  const float* delta;    // mode 1: ndelta partial sums of delta, each (B, M)
  float* rmax_out;       // mode 2: softmax row statistics (B, M)
  float* rsum_out;       // mode 2
};

struct KPParams {
  int B, N, D, M, J, nkb, nsl, xslots, wslots, nbuf, bufcols, tmem_cols, round_out;
  const uint8_t* blocks; // operand blocks [B][nkb][J x 64 tokens]: exp(S - rowmax) (mode 0) or dS (mode 1)
  const float* rsum;     // mode 0
  float* out;            // mode 0: P (B, M, D) fp32, or bf16 hi/lo rows (B, M, 2, D) when round_out; mode 1: partial dq
};

constexpr int kEpiWarps = 8;       // warps 4..11: two per TMEM lane quadrant
constexpr int kThreads = 32 * (4 + kEpiWarps);

// 1-D bulk copy global -> shared, completion counted on an mbarrier (operand blocks are contiguous)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// byte offset of (operand row j, token t) inside a [J x 64 tokens] bf16 block (128-byte rows, 128B swizzle)
__host__ __device__ __forceinline__ uint32_t block_offset(uint32_t j, uint32_t t) {
  return j * 128u + (((t >> 3) ^ (j & 7u)) << 4) + (t & 7u) * 2u;
}
__device__ __forceinline__ void store_hilo(uint8_t* blk, int m, int t, float e) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(e);
  const __nv_bfloat16 lo = __float2bfloat16_rn(e - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(blk + block_offset(2 * m, t)) = hi;
  *reinterpret_cast<__nv_bfloat16*>(blk + block_offset(2 * m + 1, t)) = lo;
}

// ------------------------------------------------------------------------------------------------
// logit-type kernel
// ------------------------------------------------------------------------------------------------
// 8 values per lane reduced over the 32 lanes of a warp in 9 shuffles: at each of the first three steps a lane
// hands half of its values to the partner and keeps the other half.  Returns value ((lane >> 2) & 7 in bit order
// 4,3,2) reduced over all lanes (replicated on 4 lanes).
template <bool kMax>
__device__ __forceinline__ float reduce8(const float (&v)[8], int lane) {
  auto op = [](float a, float b) { return kMax ? fmaxf(a, b) : a + b; };
  const bool u1 = lane & 16, u2 = lane & 8, u3 = lane & 4;
  float a[4], bq[2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    a[i] = op(u1 ? v[i + 4] : v[i], __shfl_xor_sync(0xffffffffu, u1 ? v[i] : v[i + 4], 16));
#pragma unroll
  for (int i = 0; i < 2; ++i)
    bq[i] = op(u2 ? a[i + 2] : a[i], __shfl_xor_sync(0xffffffffu, u2 ? a[i] : a[i + 2], 8));
  float c = op(u3 ? bq[1] : bq[0], __shfl_xor_sync(0xffffffffu, u3 ? bq[0] : bq[1], 4));
  c = op(c, __shfl_xor_sync(0xffffffffu, c, 2));
  return op(c, __shfl_xor_sync(0xffffffffu, c, 1));
}
constexpr int kUB = 3;                                      // mode 2: accumulator units fetched per TMEM wait

// kMode 0: logits S out.  kMode 1: dS operand blocks out (backward).  kMode 2: forward with the softmax fused into
// the epilogue -- needs the whole sample in one accumulator buffer (ngroups == 1): writes S, rowmax, rowsum and
// exp(S - rowmax) as the pool-type kernel's operand blocks, so no separate pass over S exists.
constexpr int kStatBytes = 2 * kEpiWarps * 64 * 4;          // mode 2: per-warp partial max / sum of 64 queries

template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
ks_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_xt,
          const __grid_constant__ CUtensorMap tm_w, const KSParams p) {
  // smem: [barriers, 1 KB][stage 0][stage 1]...[16 KB slack].  A stage = operand chunk + the x chunks of the
  // sample's token tiles; with tail_rows the last tile holds only its valid rows (rounded to 8): the MMA still
  // reads 128 rows there -- the rows past the tail come from whatever follows in shared memory (next stage or
  // the slack) and only reach accumulator rows n >= N, which the epilogue never uses.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t bar_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_base = bar_base + (kMode == 2 ? 1024u + (uint32_t)kStatBytes : 1024u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_bytes = (uint32_t)p.J * 128u;
  const uint32_t tail_bytes = p.tail_rows ? (uint32_t)p.tail_rows * 128u : (uint32_t)kXChunkBytes;
  const uint32_t stage_bytes = w_bytes + (uint32_t)(p.G - 1) * kXChunkBytes + tail_bytes;
  // barriers: full[nstages], empty[nstages], tmem_full[2], tmem_empty[2], then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.nstages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * p.nstages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * p.nstages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.nstages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 32 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_xt); prefetch_tmap(&tm_w); }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nitems = p.B * p.ngroups;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_x = policy_evict_first();
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int b = item / p.ngroups, g = item - b * p.ngroups;
        const int t0 = g * p.G, gt = min(p.G, p.ntiles - t0);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
          const bool tail = p.tail_rows != 0;                 // (single group: the last tile of the item is the tail)
          mbar_arrive_expect_tx(full_bar(s), w_bytes + (uint32_t)(gt - 1) * kXChunkBytes +
                                                 (tail ? tail_bytes : (uint32_t)kXChunkBytes));
          tma_load_3d(dst, &tm_w, full_bar(s), c * kChunkD, 0, p.w_batched ? b : 0);
          for (int t = 0; t < gt; ++t)
            tma_load_3d_hint(dst + w_bytes + (uint32_t)t * kXChunkBytes, (tail && t == gt - 1) ? &tm_xt : &tm_x,
                             full_bar(s), c * kChunkD, (t0 + t) * kTileRows, b, pol_x);
          if (++s == p.nstages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, p.J, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int g = item % p.ngroups;
        const int gt = min(p.G, p.ntiles - g * p.G);
        const int buf = it % p.nbuf;
        mbar_wait(tempty_bar(buf), (((uint32_t)(it / p.nbuf)) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.bufcols);
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t wsm = smem_base + (uint32_t)s * stage_bytes;
          for (int t = 0; t < gt; ++t) {
            const uint32_t xsm = wsm + w_bytes + (uint32_t)t * kXChunkBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(acc + (uint32_t)(t * p.J), smem_desc_sw128(xsm + 32u * k, 16, 1024),
                       smem_desc_sw128(wsm + 32u * k, 16, 1024), idesc, (uint32_t)((c | k) != 0));
          }
          umma_commit(empty_bar(s));
          if (++s == p.nstages) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else if (warp >= 4) {
    // epilogue: warp (4 + 4h + q) reads TMEM lanes [32q, 32q+32) and takes every other (tile, 16-column)
    // unit: units u = t * (J/16) + j0/16 with u % 2 == h
    const int wq = (warp - 4) & 3, eh = (warp - 4) >> 2;
    const int upt = p.J >> 4;
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int b = item / p.ngroups, g = item - b * p.ngroups;
      const int t0 = g * p.G, gt = min(p.G, p.ntiles - t0);
      const int buf = it % p.nbuf;
      // mode 1: lane l keeps the row statistics of queries l and 32 + l of this sample
      float st_mx[2] = {0.f, 0.f}, st_inv[2] = {0.f, 0.f}, st_dl[2] = {0.f, 0.f};
      if (kMode == 1) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = h * 32 + lane;
          if (m < p.M) {
            const size_t bm = (size_t)b * p.M + m;
            st_mx[h] = __ldg(p.rmax + bm);
            st_inv[h] = 1.f / __ldg(p.rsum + bm);
            float dl = 0.f;
            for (int q = 0; q < p.ndelta; ++q) dl += __ldg(p.delta + (size_t)q * p.B * p.M + bm);
            st_dl[h] = dl;
          }
        }
      }
      mbar_wait(tfull_bar(buf), ((uint32_t)(it / p.nbuf)) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * p.bufcols);
      if constexpr (kMode == 2) {
        // ---- fused softmax: two passes over the accumulators (TMEM reads are cheap), no atomics -- every warp
        // reduces into its own row of the partial tables and the rows are combined in a fixed order.  The
        // epilogue has to finish inside the next sample's streaming time, so TMEM is read four units per wait
        // and the 8 queries of a unit are reduced over the 32 token lanes with a transposing butterfly
        // (9 shuffles; lane l ends with query ((l >> 2) & 7)'s value).
        float* pmax = reinterpret_cast<float*>(smem_raw + (bar_base + 1024u - smem_u32(smem_raw)));   // [8][64]
        float* psum = pmax + kEpiWarps * 64;                                                          // [8][64]
        const int ew = warp - 4;
        const int nunits = gt * upt;
        const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        pmax[ew * 64 + lane] = -INFINITY;
        pmax[ew * 64 + 32 + lane] = -INFINITY;
        __syncwarp();
        // this warp's units u = eh, eh + 2, ... as (tile, column unit), advanced without divisions
        const int t_first = eh / upt, j_first = eh - t_first * upt;
        auto advance = [&](int& t, int& jj) {
          jj += 2;
          if (jj >= upt) { jj -= upt; ++t; }
          if (jj >= upt) { jj -= upt; ++t; }
        };
        {
          int t = t_first, jj = j_first;
          for (int u0 = eh; u0 < nunits; u0 += 2 * kUB) {
            uint32_t r[kUB][16];
            int tk[kUB], jk[kUB];
#pragma unroll
            for (int k = 0; k < kUB; ++k) {
              tk[k] = t; jk[k] = jj << 4;
              if (u0 + 2 * k < nunits) tmem_ld16(acc + (uint32_t)(t * p.J + (jj << 4)), r[k]);
              advance(t, jj);
            }
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < kUB; ++k) {
              if (u0 + 2 * k < nunits) {
                const bool valid = (t0 + tk[k]) * kTileRows + wq * 32 + lane < p.N;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  v[i] = valid ? __uint_as_float(r[k][2 * i]) + __uint_as_float(r[k][2 * i + 1]) : -INFINITY;
                const float red = reduce8<true>(v, lane);
                if ((lane & 3) == 0) {
                  float* slot = pmax + ew * 64 + (jk[k] >> 1) + ridx;
                  *slot = fmaxf(*slot, red);
                }
                __syncwarp();
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        float mx_lo = -INFINITY, mx_hi = -INFINITY;           // lane l: queries l and 32 + l
#pragma unroll
        for (int w = 0; w < kEpiWarps; ++w) {
          mx_lo = fmaxf(mx_lo, pmax[w * 64 + lane]);
          mx_hi = fmaxf(mx_hi, pmax[w * 64 + 32 + lane]);
        }
        psum[ew * 64 + lane] = 0.f;
        psum[ew * 64 + 32 + lane] = 0.f;
        __syncwarp();
        {
          int t = t_first, jj = j_first;
          for (int u0 = eh; u0 < nunits; u0 += 2 * kUB) {
            uint32_t r[kUB][16];
            int tk[kUB], jk[kUB];
#pragma unroll
            for (int k = 0; k < kUB; ++k) {
              tk[k] = t; jk[k] = jj << 4;
              if (u0 + 2 * k < nunits) tmem_ld16(acc + (uint32_t)(t * p.J + (jj << 4)), r[k]);
              advance(t, jj);
            }
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < kUB; ++k) {
              if (u0 + 2 * k < nunits) {
                const int j0 = jk[k];
                const int n = (t0 + tk[k]) * kTileRows + wq * 32 + lane;
                const bool valid = n < p.N;
                const int kb = n >> 6, tt = n & 63;
                uint8_t* blk = kb < p.nkb ? p.blocks + ((size_t)b * p.nkb + kb) * ((size_t)p.J * 128) : nullptr;
                float* srow = p.out + ((size_t)b * p.M + (j0 >> 1)) * p.N + n;
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int m = (j0 >> 1) + i;                // warp-uniform
                  const float v = __uint_as_float(r[k][2 * i]) + __uint_as_float(r[k][2 * i + 1]);
                  const float mx = __shfl_sync(0xffffffffu, m < 32 ? mx_lo : mx_hi, m & 31);
                  const bool on = valid && m < p.M;
                  e[i] = on ? __expf(v - mx) : 0.f;           // rows past 2M, tokens past N: zeros
                  if (on) srow[(size_t)i * p.N] = v;
                  if (blk) store_hilo(blk, m, tt, e[i]);
                }
                const float red = reduce8<false>(e, lane);
                if ((lane & 3) == 0) psum[ew * 64 + (j0 >> 1) + ridx] += red;
                __syncwarp();
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(buf));
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (ew == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = 32 * h + lane;
            float su = 0.f;
#pragma unroll
            for (int w = 0; w < kEpiWarps; ++w) su += psum[w * 64 + m];
            if (m < p.M) {
              p.rmax_out[(size_t)b * p.M + m] = h ? mx_hi : mx_lo;
              p.rsum_out[(size_t)b * p.M + m] = su;
            }
          }
        }
        continue;
      }
      for (int u = eh; u < gt * upt; u += 2) {
        const int t = u / upt, j0 = (u - t * upt) << 4;
        const int n = (t0 + t) * kTileRows + wq * 32 + lane;
        uint32_t r[16];
        tmem_ld16(acc + (uint32_t)(t * p.J + j0), r);
        // the global loads of this 8-query batch are issued before the first use (they are independent;
        // otherwise the load -> exp -> store chain serialises on memory latency)
        float sv[8];
        if (kMode == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = min((j0 >> 1) + i, p.M - 1);
            sv[i] = (n < p.N) ? __ldg(p.S + ((size_t)b * p.M + m) * p.N + n) : 0.f;
          }
        }
        tmem_ld_wait();
        uint8_t* blk = nullptr;
        const int kb = n >> 6, tt = n & 63;
        if (kMode == 1 && kb < p.nkb) blk = p.blocks + ((size_t)b * p.nkb + kb) * ((size_t)p.J * 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = (j0 >> 1) + i;                       // warp-uniform
          float v = __uint_as_float(r[2 * i]) + __uint_as_float(r[2 * i + 1]);
          if (kMode == 0) {
            if (m < p.M && n < p.N) p.out[((size_t)b * p.M + m) * p.N + n] = v;
          } else {
            // dS = A (dA - delta), written straight into the pool-type kernel's operand block; tokens
            // past N and operand rows past 2M are written as zeros so the block needs no memset
            float ds = 0.f;
            if (m < p.M) {
              const bool h = m >= 32;                          // (selects, not indexing: the arrays stay in registers)
              const int src = m & 31;
              const float mx = __shfl_sync(0xffffffffu, h ? st_mx[1] : st_mx[0], src);
              const float inv = __shfl_sync(0xffffffffu, h ? st_inv[1] : st_inv[0], src);
              const float dl = __shfl_sync(0xffffffffu, h ? st_dl[1] : st_dl[0], src);
              if (n < p.N) ds = __expf(sv[i] - mx) * inv * (v - dl);
            }
            if (blk) store_hilo(blk, m, tt, ds);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// pool-type kernel
// ------------------------------------------------------------------------------------------------
template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
kp_kernel(const __grid_constant__ CUtensorMap tm_x, const KPParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wt_bytes = (uint32_t)p.J * 128u;
  const uint32_t wt_base = smem_base + (uint32_t)p.xslots * kBrickBytes;
  const uint32_t bar_base = wt_base + (uint32_t)p.wslots * wt_bytes;
  auto xfull = [&](int s) { return bar_base + 8u * s; };
  auto xempty = [&](int s) { return bar_base + 8u * (p.xslots + s); };
  auto wfull = [&](int s) { return bar_base + 8u * (2 * p.xslots + s); };
  auto wempty = [&](int s) { return bar_base + 8u * (2 * p.xslots + p.wslots + s); };
  auto afull = [&](int b) { return bar_base + 8u * (2 * p.xslots + 2 * p.wslots + b); };
  auto aempty = [&](int b) { return bar_base + 8u * (2 * p.xslots + 2 * p.wslots + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.xslots + 2 * p.wslots + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.xslots; ++s) { mbar_init(xfull(s), 1); mbar_init(xempty(s), 1); }
    for (int s = 0; s < p.wslots; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(afull(b), 1); mbar_init(aempty(b), 32 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&tm_x);
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int d0 = blockIdx.y * p.nsl * 128;
  const int nsl = min(p.nsl, p.D / 128 - blockIdx.y * p.nsl);

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_x = policy_evict_first();
      int s = 0, ws = 0;
      uint32_t ph = 0, wph = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(wempty(ws), wph ^ 1u);                    // operand block of this token block
          mbar_arrive_expect_tx(wfull(ws), wt_bytes);
          bulk_load(wt_base + (uint32_t)ws * wt_bytes, p.blocks + ((size_t)b * p.nkb + kb) * wt_bytes, wt_bytes,
                    wfull(ws));
          if (++ws == p.wslots) { ws = 0; wph ^= 1u; }
          for (int sl = 0; sl < nsl; ++sl) {
            mbar_wait(xempty(s), ph ^ 1u);
            const uint32_t dst = smem_base + (uint32_t)s * kBrickBytes;
            mbar_arrive_expect_tx(xfull(s), kBrickBytes);
            tma_load_3d_hint(dst, &tm_x, xfull(s), d0 + sl * 128, kb * kTokBlock, b, pol_x);
            tma_load_3d_hint(dst + kBrickBytes / 2, &tm_x, xfull(s), d0 + sl * 128 + 64, kb * kTokBlock, b, pol_x);
            if (++s == p.xslots) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(128, p.J, 1, 0);       // A (tokens as K) is MN-major
      int xs = 0, ws = 0, it = 0;
      uint32_t xph = 0, wph = 0;
      bool first = true;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const int buf = (kMode == 0) ? it % p.nbuf : 0;
        if (kMode == 0) {
          mbar_wait(aempty(buf), (((uint32_t)(it / p.nbuf)) & 1u) ^ 1u);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.bufcols);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(wfull(ws), wph);
          tc_fence_after();
          const uint32_t wsm = wt_base + (uint32_t)ws * wt_bytes;
          for (int sl = 0; sl < nsl; ++sl) {
            mbar_wait(xfull(xs), xph);
            tc_fence_after();
            const uint32_t xsm = smem_base + (uint32_t)xs * kBrickBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accum = (kMode == 0) ? (uint32_t)((kb | k) != 0) : (uint32_t)(!(first && kb == 0 && k == 0));
              umma_f16(acc + (uint32_t)(sl * p.J), smem_desc_sw128(xsm + 2048u * k, kBrickBytes / 2, 1024),
                       smem_desc_sw128(wsm + 32u * k, 16, 1024), idesc, accum);
            }
            umma_commit(xempty(xs));
            if (++xs == p.xslots) { xs = 0; xph ^= 1u; }
          }
          umma_commit(wempty(ws));
          if (++ws == p.wslots) { ws = 0; wph ^= 1u; }
        }
        first = false;
        if (kMode == 0) umma_commit(afull(buf));
      }
      if (kMode == 1) umma_commit(afull(0));
    }
  } else if (warp >= 4) {
    // epilogue: warp (4 + 4h + q) reads TMEM lanes [32q, 32q+32) and every other (slice, 16-column) unit
    const int wq = (warp - 4) & 3, eh = (warp - 4) >> 2;
    const int upt = p.J >> 4;
    auto drain = [&](int buf, int b) {
      const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * p.bufcols);
      float invl[2] = {1.f, 1.f};                            // lane l keeps 1/rowsum of queries l and 32 + l
      if (kMode == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (h * 32 + lane < p.M) invl[h] = 1.f / __ldg(p.rsum + (size_t)b * p.M + h * 32 + lane);
      }
      for (int u = eh; u < nsl * upt; u += 2) {
        const int sl = u / upt, j0 = (u - sl * upt) << 4;
        const int d = d0 + sl * 128 + wq * 32 + lane;
        uint32_t r[16];
        tmem_ld16(acc + (uint32_t)(sl * p.J + j0), r);
        tmem_ld_wait();
        if (kMode == 0 && p.round_out) {
          // P as bf16 hi/lo rows (b, m, {hi, lo}, d): same bytes as fp32.  A lane holds one channel of 8 queries; an
          // 8 x 8 transpose among the 8 lanes of a group (hi | lo packed per query, 3 butterfly stages) leaves lane j
          // with query j of the group's 8 consecutive channels: two 16-byte stores instead of sixteen 2-byte ones
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const int m0 = (j0 >> 1) + i;                    // warp-uniform
            const float v0 = (__uint_as_float(r[2 * i]) + __uint_as_float(r[2 * i + 1])) * __shfl_sync(0xffffffffu, invl[(m0 >> 5) & 1], m0 & 31);
            const float v1 = (__uint_as_float(r[2 * i + 2]) + __uint_as_float(r[2 * i + 3])) * __shfl_sync(0xffffffffu, invl[((m0 + 1) >> 5) & 1], (m0 + 1) & 31);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xffff0000u));
            const uint32_t lb = *reinterpret_cast<const uint32_t*>(&l2);
            w[i] = __byte_perm(hb, lb, 0x5410);              // (hi, lo) of query i
            w[i + 1] = __byte_perm(hb, lb, 0x7632);
          }
#pragma unroll
          for (int st = 4; st >= 1; st >>= 1) {
            const bool up = (lane & st) != 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (q & st) continue;
              const uint32_t got = __shfl_xor_sync(0xffffffffu, up ? w[q] : w[q | st], st);
              if (up) w[q] = got; else w[q | st] = got;
            }
          }
          const int m = (j0 >> 1) + (lane & 7);
          if (m < p.M) {
            const uint4 oh = make_uint4(__byte_perm(w[0], w[1], 0x5410), __byte_perm(w[2], w[3], 0x5410),
                                        __byte_perm(w[4], w[5], 0x5410), __byte_perm(w[6], w[7], 0x5410));
            const uint4 ol = make_uint4(__byte_perm(w[0], w[1], 0x7632), __byte_perm(w[2], w[3], 0x7632),
                                        __byte_perm(w[4], w[5], 0x7632), __byte_perm(w[6], w[7], 0x7632));
            unsigned short* pr = reinterpret_cast<unsigned short*>(p.out) + (((size_t)b * p.M + m) * 2) * p.D + (d - (lane & 7));
            *reinterpret_cast<uint4*>(pr) = oh;
            *reinterpret_cast<uint4*>(pr + p.D) = ol;
          }
          continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = (j0 >> 1) + i;                       // warp-uniform
          if (m < p.M) {
            float v = __uint_as_float(r[2 * i]) + __uint_as_float(r[2 * i + 1]);
            if (kMode == 0) {
              v *= __shfl_sync(0xffffffffu, invl[m >> 5], m & 31);
              if (p.round_out) {                             // P as bf16 hi/lo rows (b, m, {hi, lo}, d): same bytes as fp32
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
                __nv_bfloat16* pr = reinterpret_cast<__nv_bfloat16*>(p.out) + (((size_t)b * p.M + m) * 2) * p.D + d;
                pr[0] = hi;
                pr[p.D] = lo;
              } else {
                p.out[((size_t)b * p.M + m) * p.D + d] = v;
              }
            } else {
              p.out[((size_t)blockIdx.x * p.M + m) * p.D + d] = v;
            }
          }
        }
      }
    };
    if (kMode == 0) {
      int it = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const int buf = it % p.nbuf;
        mbar_wait(afull(buf), ((uint32_t)(it / p.nbuf)) & 1u);
        tc_fence_after();
        drain(buf, b);
        tc_fence_before();
        mbar_arrive(aempty(buf));
      }
    } else {
      mbar_wait(afull(0), 0);
      tc_fence_after();
      drain(0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// small helpers: hi/lo split of fp32 operands, row statistics of the logits, partial reduction
// ------------------------------------------------------------------------------------------------
// dst[(z*J + 2m + {0,1}) * D + d] = hi/lo(scale * src[(z*M + m) * D + d]); rows 2M..J-1 zero.
__global__ void split_hilo_kernel(const float* __restrict__ src, float scale, int M, int J, int D,
                                  __nv_bfloat16* __restrict__ dst) {
  const int z = blockIdx.y;
  const int pairs = J / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)pairs * D / 4; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / (D / 4)), d = (int)(i % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) v = *reinterpret_cast<const float4*>(src + ((size_t)z * M + m) * D + d);
    const float f[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = __float2bfloat16_rn(f[e]);
      lo[e] = __float2bfloat16_rn(f[e] - __bfloat162float(hi[e]));
    }
    __nv_bfloat16* ph = dst + ((size_t)z * J + 2 * m) * D + d;
    *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(ph + D) = *reinterpret_cast<uint2*>(lo);
  }
}

// one warp per (b, j/2) operand row pair of the logits: rowmax, rowsum = sum exp(S - rowmax), optional
// attention map, and (blocks != nullptr) exp(S - rowmax) as bf16 hi/lo operand blocks for the pool-type
// kernel -- tokens past N and row pairs past M are written as zeros, so the blocks need no memset.
// kRegs > 0: the row (N <= 32 * kRegs) is held in registers -- one read, one exp per element.
template <int kRegs>
__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ S, int B, int M, int J, int N, int nkb,
                                                       float* __restrict__ rmax, float* __restrict__ rsum,
                                                       float* __restrict__ attn, uint8_t* __restrict__ blocks) {
  const int pairs = J >> 1;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= (long long)B * pairs) return;
  const int b = (int)(r / pairs), m = (int)(r % pairs);
  const int lane = threadIdx.x & 31;
  const float* row = S + ((size_t)b * M + min(m, M - 1)) * N;
  const size_t blk_bytes = (size_t)J * 128;
  if (kRegs > 0) {
    float v[kRegs > 0 ? kRegs : 1];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kRegs; ++k) {
      const int n = lane + 32 * k;
      v[k] = (m < M && n < N) ? __ldg(row + n) : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < kRegs; ++k) {
      v[k] = (m < M && lane + 32 * k < N) ? __expf(v[k] - mx) : 0.f;
      sum += v[k];
    }
    sum = warp_sum(sum);
    if (m < M) {
      if (lane == 0) { rmax[(size_t)b * M + m] = mx; rsum[(size_t)b * M + m] = sum; }
      if (attn) {
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < kRegs; ++k)
          if (lane + 32 * k < N) attn[((size_t)b * M + m) * N + lane + 32 * k] = v[k] * inv;
      }
    }
    if (blocks) {
#pragma unroll
      for (int k = 0; k < kRegs; ++k) {
        const int n = lane + 32 * k;
        if (n < nkb * kTokBlock) store_hilo(blocks + ((size_t)b * nkb + (n >> 6)) * blk_bytes, m, n & 63, v[k]);
      }
    }
    return;
  }
  float mx = 0.f, inv = 0.f;
  if (m < M) {
    mx = -INFINITY;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, row[n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) sum += __expf(row[n] - mx);
    sum = warp_sum(sum);
    inv = 1.f / sum;
    if (lane == 0) { rmax[(size_t)b * M + m] = mx; rsum[(size_t)b * M + m] = sum; }
    if (attn)
      for (int n = lane; n < N; n += 32) attn[((size_t)b * M + m) * N + n] = __expf(row[n] - mx) * inv;
  }
  if (blocks) {
    for (int n = lane; n < nkb * kTokBlock; n += 32) {
      const float e = (m < M && n < N) ? __expf(row[n] - mx) : 0.f;
      store_hilo(blocks + ((size_t)b * nkb + (n >> 6)) * blk_bytes, m, n & 63, e);
    }
  }
}

// out[i] = scale * sum_k part[k][i]: 32 outputs per block, the partials dealt over 8 thread rows and combined in
// a fixed order (deterministic); 1024 blocks keep enough loads in flight for a ~19 MB read
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int nparts, size_t n,
                                                              float scale, float* __restrict__ out) {
  __shared__ float sm[8][33];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t i = (size_t)blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < n)
    for (int k = ty; k < nparts; k += 8) s += __ldg(part + (size_t)k * n + i);
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sm[r][tx];
    out[i] = t * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 tensor (d2, d1, d0) row-major, box (1, rows, 64 elements), 128-byte swizzle, zero fill out of bounds
int make_tmap(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return EP_ERR_DEVICE;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : EP_ERR_UNSUPPORTED;
}

}  // namespace
int make_tmap_bf16(::CUtensorMap_st* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || rank < 1 || rank > 5) return EP_ERR_DEVICE;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(m), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d,
                  st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : EP_ERR_UNSUPPORTED;
}
namespace {

int round16(int v) { return (v + 15) / 16 * 16; }
int pow2_cols(int c) { int v = 32; while (v < c) v <<= 1; return v; }

struct Plan {
  bool ok = false;
  int J, ntiles, G, ngroups, ks_stages, ks_nbuf, nkb, nsl_fwd, nsl_bwd, ysplit_fwd, ysplit_bwd, xslots, wslots, tail_rows;
  size_t ks_smem, kp_smem;
};

Plan make_plan(int N, int D, int M) {
  Plan pl;
  if (D % 128 != 0 || M < 1 || M > 64) return pl;
  pl.J = round16(2 * M);
  pl.ntiles = (N + kTileRows - 1) / kTileRows;
  // tiles of one sample handled together: bounded by TMEM columns (512) and by two smem stages
  int G = std::min(pl.ntiles, 512 / pl.J);
  while (G > 1 && 2 * ((size_t)pl.J * 128 + (size_t)G * kXChunkBytes) > (size_t)kSmemBudget) --G;
  if (G < 1 || 2 * ((size_t)pl.J * 128 + (size_t)G * kXChunkBytes) > (size_t)kSmemBudget) return pl;
  pl.ngroups = (pl.ntiles + G - 1) / G;
  G = (pl.ntiles + pl.ngroups - 1) / pl.ngroups;            // balance the groups
  pl.G = G;
  // ragged last tile: load only its valid rows (multiple of 8 = one swizzle atom) when one group covers the sample
  const int rem = N - (pl.ntiles - 1) * kTileRows;
  pl.tail_rows = (pl.ngroups == 1 && rem < kTileRows) ? (rem + 7) / 8 * 8 : 0;
  const size_t stage = (size_t)pl.J * 128 + (size_t)(G - 1) * kXChunkBytes +
                       (pl.tail_rows ? (size_t)pl.tail_rows * 128 : (size_t)kXChunkBytes);
  pl.ks_stages = (int)std::min<size_t>(8, (kSmemBudget - kXChunkBytes) / stage);
  pl.ks_nbuf = (2 * G * pl.J <= 512) ? 2 : 1;
  pl.ks_smem = 1024 /*alignment*/ + 1024 /*barriers*/ + pl.ks_stages * stage + kXChunkBytes /*slack*/;
  pl.nkb = (N + kTokBlock - 1) / kTokBlock;
  const int slices = D / 128;
  const int max_sl_fwd = std::max(1, 256 / pl.J), max_sl_bwd = std::max(1, 512 / pl.J);   // fwd double-buffers TMEM
  pl.ysplit_fwd = (slices + max_sl_fwd - 1) / max_sl_fwd;
  pl.nsl_fwd = (slices + pl.ysplit_fwd - 1) / pl.ysplit_fwd;
  pl.ysplit_bwd = (slices + max_sl_bwd - 1) / max_sl_bwd;
  pl.nsl_bwd = (slices + pl.ysplit_bwd - 1) / pl.ysplit_bwd;
  pl.wslots = 4;
  pl.xslots = (int)std::min<size_t>(10, (kSmemBudget - (size_t)pl.wslots * pl.J * 128) / kBrickBytes);
  pl.kp_smem = (size_t)pl.xslots * kBrickBytes + (size_t)pl.wslots * pl.J * 128 + 1024 + 512;
  pl.ok = pl.xslots >= 3 && pl.ks_stages >= 2;
  return pl;
}

struct Ws100 {
  size_t qhl, S_unused, dphl, dS, part, fused, total;
};
Ws100 carve100(int B, int N, int D, int M, const Plan& pl) {
  Ws100 w;
  size_t off = 0;
  w.qhl = off;  off += align_up((size_t)pl.J * D * 2, 1024);
  w.dphl = off; off += align_up((size_t)B * pl.J * D * 2, 1024);
  w.dS = off;   off += align_up((size_t)B * pl.nkb * pl.J * 128, 1024);   // operand blocks (exp(S - max) / dS)
  w.part = off; off += align_up((size_t)kNumSMs * M * D * 4, 256);
  w.fused = off; off += align_up(fused_workspace_bytes(N, D, M), 256);   // pair-mode scratch of the one-pass kernels
  w.S_unused = 0;
  w.total = off;
  return w;
}

template <typename K>
int set_dyn_smem(K kernel, size_t bytes) {
  EP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int kMode>
int launch_ks(const void* x, const void* w, int w_batched, int B, int N, int D, int M, const Plan& pl, float* out,
              uint8_t* blocks, const float* S, const float* rmax, const float* rsum, const float* delta, int ndelta,
              cudaStream_t s, float* rmax_out = nullptr, float* rsum_out = nullptr) {
  CUtensorMap tm_x, tm_xt, tm_w;
  int rc;
  if ((rc = make_tmap(&tm_x, x, D, N, B, kTileRows))) return rc;
  if ((rc = make_tmap(&tm_xt, x, D, N, B, pl.tail_rows ? pl.tail_rows : kTileRows))) return rc;
  if ((rc = make_tmap(&tm_w, w, D, pl.J, w_batched ? B : 1, pl.J))) return rc;
  KSParams p{};
  p.B = B; p.N = N; p.D = D; p.M = M; p.J = pl.J; p.ntiles = pl.ntiles; p.G = pl.G; p.ngroups = pl.ngroups;
  p.nchunks = D / kChunkD; p.nstages = pl.ks_stages; p.w_batched = w_batched; p.nbuf = pl.ks_nbuf;
  p.bufcols = pl.G * pl.J; p.tmem_cols = pow2_cols(p.nbuf * p.bufcols);
  p.nkb = pl.nkb; p.ndelta = ndelta; p.tail_rows = pl.tail_rows;
  p.out = out; p.blocks = blocks; p.S = S; p.rmax = rmax; p.rsum = rsum; p.delta = delta;
  p.rmax_out = rmax_out; p.rsum_out = rsum_out;
  const size_t smem = pl.ks_smem + (kMode == 2 ? kStatBytes : 0);
  if (smem > 227 * 1024) return EP_ERR_UNSUPPORTED;
  if ((rc = set_dyn_smem(ks_kernel<kMode>, smem))) return rc;
  const int grid = std::min(B * pl.ngroups, stream_sms());
  ks_kernel<kMode><<<grid, kThreads, smem, s>>>(tm_x, tm_xt, tm_w, p);
  EP_LAUNCH_CHECK();
  return 0;
}

template <int kMode>
int launch_kp(const void* x, int B, int N, int D, int M, const Plan& pl, const uint8_t* blocks, const float* rsum,
              float* out, int* groups_out, int round_out, cudaStream_t s) {
  CUtensorMap tm_x;
  int rc;
  if ((rc = make_tmap(&tm_x, x, D, N, B, kTokBlock))) return rc;
  KPParams p{};
  p.B = B; p.N = N; p.D = D; p.M = M; p.J = pl.J; p.nkb = pl.nkb;
  p.nsl = kMode == 0 ? pl.nsl_fwd : pl.nsl_bwd;
  const int ysplit = kMode == 0 ? pl.ysplit_fwd : pl.ysplit_bwd;
  p.xslots = pl.xslots; p.wslots = pl.wslots;
  p.bufcols = p.nsl * pl.J;
  p.nbuf = (kMode == 0 && 2 * p.bufcols <= 512) ? 2 : 1;
  p.tmem_cols = pow2_cols(p.nbuf * p.bufcols);
  p.blocks = blocks; p.rsum = rsum; p.out = out;
  p.round_out = round_out;
  if ((rc = set_dyn_smem(kp_kernel<kMode>, pl.kp_smem))) return rc;
  const int gx = std::max(1, std::min(B, stream_sms() / ysplit));
  if (groups_out) *groups_out = gx;
  kp_kernel<kMode><<<dim3(gx, ysplit), kThreads, pl.kp_smem, s>>>(tm_x, p);
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool sm100_supported(int x_dtype, int B, int N, int D, int M) {
  if (x_dtype != EP_DTYPE_BF16 || B < 1 || N < 1) return false;
  return make_plan(N, D, M).ok && encode_fn() != nullptr;
}

size_t sm100_workspace_bytes(int B, int N, int D, int M) {
  Plan pl = make_plan(N, D, M);
  if (!pl.ok) return 0;
  return carve100(B, N, D, M, pl).total;
}

int sm100_pool_fwd(const void* x, const float* cls, float scale, int B, int N, int D, int M, float* P, float* S,
                   float* rowmax, float* rowsum, float* attn, int round_p, void* ws, cudaStream_t s, int q_ready) {
  const Plan pl = make_plan(N, D, M);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  const Ws100 w = carve100(B, N, D, M, pl);
  __nv_bfloat16* qhl = (__nv_bfloat16*)((char*)ws + w.qhl);
  uint8_t* blocks = (uint8_t*)ws + w.dS;
  int rc;
  StageTimer tm(s);
  if (!q_ready) {          // else: ep_refresh_operands wrote the scaled hi/lo query rows after the last parameter update
    split_hilo_kernel<<<dim3(std::max(1, pl.J * D / 8 / 256), 1), 256, 0, s>>>(cls, scale, M, pl.J, D, qhl);
    EP_LAUNCH_CHECK();
    tm.mark("split_q");
  }
  // one-pass kernel: logits, softmax and pooled tokens of a sample without leaving the SM (ep_fused_sm100.cu)
  if (attn == nullptr && P != nullptr && !(g_debug & 1024) && fused_supported(N, D, M)) {
    rc = fused_pool_fwd(x, qhl, pl.J, B, N, D, M, P, S, rowmax, rowsum, round_p, (char*)ws + w.fused, s);
    tm.mark("fused fwd");
    return rc;
  }
  // the softmax rides in the logit kernel's epilogue when one accumulator buffer holds the whole sample
  const bool fused = pl.ngroups == 1 && attn == nullptr && P != nullptr && !(g_debug & 512) &&
                     pl.ks_smem + kStatBytes <= 227 * 1024;
  if (fused) {
    if ((rc = launch_ks<2>(x, qhl, 0, B, N, D, M, pl, S, blocks, nullptr, nullptr, nullptr, nullptr, 0, s, rowmax, rowsum)))
      return rc;
    tm.mark("ks<2> logits+softmax");
  } else {
  if ((rc = launch_ks<0>(x, qhl, 0, B, N, D, M, pl, S, nullptr, nullptr, nullptr, nullptr, nullptr, 0, s))) return rc;
  tm.mark("ks<0> logits");
  const long long rows = (long long)B * (pl.J / 2);
  {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    uint8_t* blk = P ? blocks : nullptr;
    const int span = pl.nkb * kTokBlock;                       // tokens covered incl. the zero padding
    if (span <= 32 * 10) rowstats_kernel<10><<<grid, 256, 0, s>>>(S, B, M, pl.J, N, pl.nkb, rowmax, rowsum, attn, blk);
    else if (span <= 32 * 24) rowstats_kernel<24><<<grid, 256, 0, s>>>(S, B, M, pl.J, N, pl.nkb, rowmax, rowsum, attn, blk);
    else rowstats_kernel<0><<<grid, 256, 0, s>>>(S, B, M, pl.J, N, pl.nkb, rowmax, rowsum, attn, blk);
  }
  EP_LAUNCH_CHECK();
  tm.mark("rowstats");
  }
  if (P == nullptr) return 0;
  rc = launch_kp<0>(x, B, N, D, M, pl, blocks, rowsum, P, nullptr, round_p, s);
  tm.mark("kp<0> pool");
  return rc;
}

// dP arrives as bf16 hi/lo rows (B, J, D) in the workspace (written by the projection backward) together
// with `ndelta` partial sums of delta = dP . P
void* sm100_dphl_ptr(void* ws, int B, int N, int D, int M) {
  const Plan pl = make_plan(N, D, M);
  return (char*)ws + carve100(B, N, D, M, pl).dphl;
}
int sm100_J(int N, int D, int M) { return make_plan(N, D, M).J; }
void* sm100_qhl_ptr(void* ws, int B, int N, int D, int M) {
  const Plan pl = make_plan(N, D, M);
  return (char*)ws + carve100(B, N, D, M, pl).qhl;
}

int sm100_pool_bwd(const void* x, const float* S, float scale, int B, int N, int D, int M, const float* rowmax,
                   const float* rowsum, const float* dP, const float* delta, int ndelta, float* d_cls, void* ws,
                   cudaStream_t s) {
  const Plan pl = make_plan(N, D, M);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  const Ws100 w = carve100(B, N, D, M, pl);
  __nv_bfloat16* dphl = (__nv_bfloat16*)((char*)ws + w.dphl);
  uint8_t* blocks = (uint8_t*)ws + w.dS;
  float* part = (float*)((char*)ws + w.part);
  int rc;
  StageTimer tm(s);
  if (dP) {   // fp32 dP (B, M, D) given: split it here; nullptr = the hi/lo rows are already in the workspace
    split_hilo_kernel<<<dim3(std::max(1, std::min(64, pl.J * D / 8 / 256)), B), 256, 0, s>>>(dP, 1.f, M, pl.J, D, dphl);
    EP_LAUNCH_CHECK();
    tm.mark("split_dP");
  }
  int groups = 0;
  if (ndelta == 1 && !(g_debug & 1024) && fused_supported(N, D, M)) {
    // one-pass kernel: dA, dS and the query gradient of a sample without leaving the SM (ep_fused_sm100.cu)
    if ((rc = fused_pool_bwd(x, dphl, pl.J, B, N, D, M, S, rowmax, rowsum, delta, part, &groups, (char*)ws + w.fused, s))) return rc;
    tm.mark("fused bwd");
  } else {
    if ((rc = launch_ks<1>(x, dphl, 1, B, N, D, M, pl, nullptr, blocks, S, rowmax, rowsum, delta, ndelta, s))) return rc;
    tm.mark("ks<1> dS");
    if ((rc = launch_kp<1>(x, B, N, D, M, pl, blocks, nullptr, part, &groups, 0, s))) return rc;
    tm.mark("kp<1> dq");
  }
  const size_t n = (size_t)M * D;
  EP_CUDA(launch_pdl(reduce_partials_kernel, dim3((unsigned)((n + 31) / 32)), dim3(256), 0, s, (const float*)part, groups, n, scale,
                     d_cls));
  EP_LAUNCH_CHECK();
  tm.mark("reduce");
  return 0;
}

}  // namespace ep
