// One-pass fused EP pooling kernels (sm_100a): the tokens of a sample cross HBM ONCE per direction.
//
//   forward   (poolings/ep.py:39-45)   S = x q^T  ->  softmax over tokens  ->  P = A x           in one kernel
//   backward      dA = x dP^T -> dS = A (dA - delta)  ->  dq += dS^T x      in one kernel
//
// A sample (N x D bf16: 526 KB at N=257, D=1024) fits neither shared memory nor, with the pooled accumulators,
// a one-shot TMEM plan, and the softmax needs all of a sample's logits before the first pooled product.  So a
// persistent CTA walks its samples and fetches each one twice, back to back, through ONE TMA ring:
//   L(i)  d-chunks [128 tokens x 64 d] (K-major A operand)  -> logits / dA of all token tiles accumulate in TMEM
//   P(i)  bricks  [64 tokens x 128 d] (the same swizzled bytes read as the MN-major A operand) -> pooled sums
// The second fetch finds the sample in L2 (148 CTAs x 526 KB = 78 MB of the 126 MB; first fetch evict_last, second
// evict_first): measured with tools/dev_l2_probe.cu this order streams c2 at 108 us against 89 us for a single
// pass and 2 x 89 us for two kernels.  Between the two phases nothing touches global memory: exp(S - max) (forward)
// or dS (backward) go from the TMEM epilogue straight into shared memory as the UMMA B operand of the second phase.
// To cover the epilogue's latency the first `lead` chunks of the NEXT sample are fetched and multiplied before the
// pooled phase of the current one (double-buffered logit accumulators).
//
// fp32 operands (queries, probabilities, dP, dS) are bf16 hi/lo pairs as in ep_pool_sm100.cu, but the pair is two
// K-steps into ONE accumulator column (hi rows and lo rows are separate B operands) instead of two columns: the
// accumulators need half the TMEM -- 2 x 96 (logits, double-buffered) + 256 (pooled, D = 1024) = 448 columns at
// M = 32 -- which is what lets the whole sample live in one CTA.
//
// Warp roles (384 threads, one CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-11 epilogue (two per TMEM lane quadrant).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "ep_ptx.cuh"
#include "ep_sm100.cuh"

namespace ep {
using namespace ptx;

namespace fused {
constexpr int kSlotBytes = 16384;      // ring slot: [128 tokens x 64 d] chunk tile, or [64 tokens x 128 d] brick
constexpr int kEpiWarpsF = 8;
constexpr int kThreadsF = 32 * (4 + kEpiWarpsF);
constexpr int kStatFloats = 2 * kEpiWarpsF * 64;   // per-warp partial max / sum of up to 64 queries

struct FParams {
  int B, N, D, M, Mp;                  // Mp = M rounded up to 16: UMMA N, accumulator columns per tile / slice
  int ntiles, nfull, tail_rows;        // token tiles of 128; nfull of them loaded as full boxes; tail_rows > 0: the last
                                       // tile is a short box sharing its slot with the query chunk
  int nchunks, nkb, nsl, nslots, lead, nbuf, bufcols, pcol0, tmem_cols, w_batched, qoff;
  float* S;                            // fwd: logits out (B, M, N);  bwd: saved logits in
  float* rmax; float* rsum;            // fwd: out;  bwd: in
  const float* delta;                  // bwd: (B, M)
  float* out;                          // fwd: P as bf16 hi/lo rows (B, M, 2, D) or fp32 (B, M, D);  bwd: partial dq [grid][M][D]
  int round_out;
};

__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}

// 16 values per lane reduced over the 32 lanes of a warp in 16 shuffles (transposing butterfly): lane l ends with
// value ((l >> 1) & 15 read as bits 16,8,4,2 -> 8,4,2,1) reduced over all lanes, replicated on 2 lanes.
template <bool kMax>
__device__ __forceinline__ float reduce16(const float (&v)[16], int lane) {
  auto op = [](float a, float b) { return kMax ? fmaxf(a, b) : a + b; };
  const bool u1 = lane & 16, u2 = lane & 8, u3 = lane & 4, u4 = lane & 2;
  float a[8], b[4], c[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = op(u1 ? v[i + 8] : v[i], __shfl_xor_sync(0xffffffffu, u1 ? v[i] : v[i + 8], 16));
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = op(u2 ? a[i + 4] : a[i], __shfl_xor_sync(0xffffffffu, u2 ? a[i] : a[i + 4], 8));
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i] = op(u3 ? b[i + 2] : b[i], __shfl_xor_sync(0xffffffffu, u3 ? b[i] : b[i + 2], 4));
  float d = op(u4 ? c[1] : c[0], __shfl_xor_sync(0xffffffffu, u4 ? c[0] : c[1], 2));
  return op(d, __shfl_xor_sync(0xffffffffu, d, 1));
}
__device__ __forceinline__ int reduce16_index(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// byte offset of (operand row m, token t) inside a K-major [Mp rows x 64 tokens] bf16 block (128-byte rows, 128B swizzle)
__device__ __forceinline__ uint32_t blk_off(uint32_t m, uint32_t t) {
  return m * 128u + (((t >> 3) ^ (m & 7u)) << 4) + (t & 7u) * 2u;
}

// kBwd = false: forward (logits, softmax, pooled tokens).  kBwd = true: backward (dA, dS, query gradient).
template <bool kBwd>
__global__ void __launch_bounds__(kThreadsF, 1)
fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_xt,
             const __grid_constant__ CUtensorMap tm_xb, const __grid_constant__ CUtensorMap tm_w, const FParams p) {
  // smem: [ring: nslots x 16 KB][operand blocks: nkb x (hi [Mp x 128 B], lo [Mp x 128 B])][stats][barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (ring - smem_u32(smem_raw));
  const uint32_t half_bytes = (uint32_t)p.Mp * 128u;            // one hi or lo operand block
  const uint32_t blk_base = ring + (uint32_t)p.nslots * kSlotBytes;
  const uint32_t blk_bytes = 2u * half_bytes * (uint32_t)p.nkb;
  const uint32_t stat_base = blk_base + blk_bytes;
  const uint32_t bar_base = stat_base + kStatFloats * 4u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.nslots + s); };
  const uint32_t misc = bar_base + 16u * p.nslots;
  auto tfull_bar = [&](int b) { return misc + 8u * b; };         // logit accumulators of buffer b complete
  const uint32_t eready_bar = misc + 16u;                         // operand blocks of the current sample written (and
                                                                  // its logit accumulators read: the buffer is free)
  const uint32_t pdone_bar = misc + 24u;                          // pooled MMAs of the current sample complete
  const uint32_t pfree_bar = misc + 32u;                          // (fwd) pooled accumulators drained
  const uint32_t tmem_slot = misc + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - ring));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nslots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) mbar_init(tfull_bar(b), 1);
    mbar_init(eready_bar, kEpiWarpsF);
    mbar_init(pdone_bar, 1);
    mbar_init(pfree_bar, kEpiWarpsF);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_xt); prefetch_tmap(&tm_xb); prefetch_tmap(&tm_w); }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // operand blocks start as zeros: rows m >= M and tokens that no epilogue thread owns stay zero for the whole launch
  for (uint32_t o = threadIdx.x * 16u; o < blk_bytes; o += kThreadsF * 16u)
    *reinterpret_cast<uint4*>(gen + (blk_base - ring) + o) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  int nmine = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) ++nmine;
  const uint32_t w_bytes = 2u * half_bytes;                       // query / dP chunk: hi rows then lo rows
  const bool mixed_tail = p.tail_rows > 0;                        // the short last tile shares the query chunk's slot

  if (warp == 0) {
    if (lane == 0 && nmine > 0) {
      const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
      int s = 0;
      uint32_t ph = 0;
      auto acquire = [&]() -> uint32_t {
        mbar_wait(empty_bar(s), ph ^ 1u);
        return ring + (uint32_t)s * kSlotBytes;
      };
      auto advance = [&]() { if (++s == p.nslots) { s = 0; ph ^= 1u; } };
      // L: chunk c of sample b = [query chunk (+ short tail tile)] slot, then one slot per full token tile
      auto load_L = [&](int b, int c0, int c1) {
        for (int c = c0; c < c1; ++c) {
          uint32_t dst = acquire();
          mbar_arrive_expect_tx(full_bar(s), w_bytes + (mixed_tail ? (uint32_t)p.tail_rows * 128u : 0u));
          tma_load_4d_hint(dst + (uint32_t)p.qoff, &tm_w, full_bar(s), c * 64, 0, 0, p.w_batched ? b : 0, pol_last);
          tma_load_4d_hint(dst + (uint32_t)p.qoff + half_bytes, &tm_w, full_bar(s), c * 64, 1, 0, p.w_batched ? b : 0, pol_last);
          if (mixed_tail) tma_load_3d_hint(dst, &tm_xt, full_bar(s), c * 64, p.nfull * 128, b, pol_last);
          advance();
          for (int t = 0; t < p.nfull; ++t) {
            dst = acquire();
            mbar_arrive_expect_tx(full_bar(s), kSlotBytes);
            tma_load_3d_hint(dst, &tm_x, full_bar(s), c * 64, t * 128, b, pol_last);
            advance();
          }
        }
      };
      auto load_P = [&](int b) {
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int sl = 0; sl < p.nsl; ++sl) {
            const uint32_t dst = acquire();
            mbar_arrive_expect_tx(full_bar(s), kSlotBytes);
            tma_load_3d_hint(dst, &tm_xb, full_bar(s), sl * 128, kb * 64, b, pol_first);
            tma_load_3d_hint(dst + kSlotBytes / 2, &tm_xb, full_bar(s), sl * 128 + 64, kb * 64, b, pol_first);
            advance();
          }
      };
      load_L(blockIdx.x, 0, p.nchunks);
      for (int i = 0; i < nmine; ++i) {
        const int b = blockIdx.x + i * gridDim.x;
        if (i + 1 < nmine) load_L(b + gridDim.x, 0, p.lead);
        load_P(b);
        if (i + 1 < nmine) load_L(b + gridDim.x, p.lead, p.nchunks);
      }
    }
  } else if (warp == 1) {
    // MMA issue.  The whole warp walks the schedule (uniform control flow, descriptors in uniform registers) and one
    // elected lane issues: a tcgen05.mma with N <= 64 occupies the tensor core for only ~40-48 cycles
    // (tools/dev_umma_probe.cu), so per-instruction address arithmetic in a single divergent thread would be the limit.
    if (nmine > 0) {
      const bool leader = elect_one();
      const uint32_t idesc_L = idesc_bf16(128, 2 * p.Mp, 0, 0);   // logits: N = hi rows + lo rows of the query chunk
      const uint32_t idesc_P = idesc_bf16(128, p.Mp, 1, 0);       // pooled: A (tokens as K) is MN-major
      const uint64_t dK = smem_desc_sw128(0, 16, 1024);           // + (address >> 4)
      const uint64_t dMN = smem_desc_sw128(0, kSlotBytes / 2, 1024);
      const uint32_t ncolL = 2u * (uint32_t)p.Mp;
      int s = 0;
      uint32_t ph = 0;
      auto advance = [&]() { if (++s == p.nslots) { s = 0; ph ^= 1u; } };
      // logits of sample ordinal j, chunks [c0, c1), into accumulator buffer j % nbuf
      auto mma_L = [&](int j, int c0, int c1) {
        const uint32_t acc = tmem_base + (uint32_t)((j % p.nbuf) * p.bufcols);
        for (int c = c0; c < c1; ++c) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const int ws = s;
          const uint32_t wslot = ring + (uint32_t)s * kSlotBytes;
          const uint64_t bd = dK + (uint64_t)((wslot + (uint32_t)p.qoff) >> 4);
          advance();
          for (int t = 0; t < p.nfull; ++t) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint64_t ad = dK + (uint64_t)((ring + (uint32_t)s * kSlotBytes) >> 4);
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(acc + (uint32_t)t * ncolL, ad + 2u * k, bd + 2u * k, idesc_L, (uint32_t)((c | k) != 0));
              umma_commit(empty_bar(s));
            }
            __syncwarp();
            advance();
          }
          if (leader) {
            if (mixed_tail) {
              const uint64_t ad = dK + (uint64_t)(wslot >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(acc + (uint32_t)p.nfull * ncolL, ad + 2u * k, bd + 2u * k, idesc_L, (uint32_t)((c | k) != 0));
            }
            umma_commit(empty_bar(ws));
          }
          __syncwarp();
        }
      };
      mma_L(0, 0, p.nchunks);
      if (leader) umma_commit(tfull_bar(0));
      __syncwarp();
      const uint32_t pacc = tmem_base + (uint32_t)p.pcol0;
      for (int i = 0; i < nmine; ++i) {
        const int j = i + 1;
        // (lead > 0 needs nbuf == 2: buffer j % 2 was read by the epilogue of sample j - 2, which finished before
        //  eready of sample j - 2 completed -- waited for two iterations ago)
        if (j < nmine && p.lead > 0) mma_L(j, 0, p.lead);
        mbar_wait(eready_bar, (uint32_t)(i & 1));                  // operand blocks of sample i are in shared memory
        if (!kBwd) mbar_wait(pfree_bar, ((uint32_t)(i & 1)) ^ 1u); // pooled accumulators of sample i - 1 drained
        tc_fence_after();
        for (int kb = 0; kb < p.nkb; ++kb) {
          const uint64_t bd_hi = dK + (uint64_t)((blk_base + (uint32_t)kb * 2u * half_bytes) >> 4);
          const uint64_t bd_lo = bd_hi + (uint64_t)(half_bytes >> 4);
          for (int sl = 0; sl < p.nsl; ++sl) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint64_t ad = dMN + (uint64_t)((ring + (uint32_t)s * kSlotBytes) >> 4);
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t first = kBwd ? (uint32_t)((i | kb | k) != 0) : (uint32_t)((kb | k) != 0);
                umma_f16(pacc + (uint32_t)(sl * p.Mp), ad + 128u * k, bd_hi + 2u * k, idesc_P, first);
                umma_f16(pacc + (uint32_t)(sl * p.Mp), ad + 128u * k, bd_lo + 2u * k, idesc_P, 1u);
              }
              umma_commit(empty_bar(s));
            }
            __syncwarp();
            advance();
          }
        }
        if (leader) umma_commit(pdone_bar);
        __syncwarp();
        if (j < nmine) {
          mma_L(j, p.lead, p.nchunks);
          if (leader) umma_commit(tfull_bar(j % p.nbuf));
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, wq = ew & 3, eh = ew >> 2;
    const int upt = p.Mp >> 4;                                     // 16-query units per tile / slice
    float* pmax = reinterpret_cast<float*>(gen + (stat_base - ring));   // [8][64]
    float* psum = pmax + kEpiWarpsF * 64;                               // [8][64]
    uint8_t* blk_gen = gen + (blk_base - ring);
    const int ridx = reduce16_index(lane);
    const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
    const int nunits = p.ntiles * upt;

    // pooled accumulators -> global (fwd: normalised P of sample b; bwd: this CTA's partial dq)
    auto drain = [&](int b) {
      const uint32_t acc = tmem_base + lane_base + (uint32_t)p.pcol0;
      float invl[2] = {1.f, 1.f};                                  // lane l keeps 1/rowsum of queries l and 32 + l
      if (!kBwd) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (h * 32 + lane < p.M) invl[h] = 1.f / psum[h * 32 + lane];      // row sums of this sample (table row 0 = totals)
      }
      for (int u = eh; u < p.nsl * upt; u += 2) {
        const int sl = u / upt, j0 = (u - sl * upt) << 4;
        const int d = sl * 128 + wq * 32 + lane;
        uint32_t r[16];
        tmem_ld16(acc + (uint32_t)(sl * p.Mp + j0), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int m = j0 + i;                                    // warp-uniform
          if (m < p.M) {
            float v = __uint_as_float(r[i]);
            if (!kBwd) {
              v *= __shfl_sync(0xffffffffu, m < 32 ? invl[0] : invl[1], m & 31);
              if (p.round_out) {                                   // P as bf16 hi/lo rows (b, m, {hi, lo}, d)
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
    // logits (hi column + lo column) of unit u = (tile, 16 queries) for this lane's token
    auto load_unit = [&](uint32_t acc, int t, int j0, float (&v)[16]) {
      uint32_t rh[16], rl[16];
      tmem_ld16(acc + (uint32_t)(t * 2 * p.Mp + j0), rh);
      tmem_ld16(acc + (uint32_t)(t * 2 * p.Mp + p.Mp + j0), rl);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(rh[q]) + __uint_as_float(rl[q]);
    };
    auto store_hilo = [&](uint8_t* blk, int m, int tt, float e) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(e);
      const __nv_bfloat16 lo = __float2bfloat16_rn(e - __bfloat162float(hi));
      *reinterpret_cast<__nv_bfloat16*>(blk + blk_off(m, tt)) = hi;
      *reinterpret_cast<__nv_bfloat16*>(blk + half_bytes + blk_off(m, tt)) = lo;
    };

    for (int i = 0; i < nmine; ++i) {
      const int b = blockIdx.x + i * gridDim.x;
      const int buf = i % p.nbuf;
      if (i > 0) {                                                 // pooled MMAs of sample i - 1 complete: blocks reusable
        mbar_wait(pdone_bar, (uint32_t)((i - 1) & 1));
        tc_fence_after();
        if (!kBwd) {
          drain(b - gridDim.x);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pfree_bar);
        }
      }
      // bwd: lane l keeps the row statistics of queries l and 32 + l of this sample
      float st_mx[2] = {0.f, 0.f}, st_inv[2] = {0.f, 0.f}, st_dl[2] = {0.f, 0.f};
      if (kBwd) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = h * 32 + lane;
          if (m < p.M) {
            const size_t bm = (size_t)b * p.M + m;
            st_mx[h] = __ldg(p.rmax + bm);
            st_inv[h] = 1.f / __ldg(p.rsum + bm);
            st_dl[h] = __ldg(p.delta + bm);
          }
        }
      }
      mbar_wait(tfull_bar(buf), ((uint32_t)(i / p.nbuf)) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + (uint32_t)(buf * p.bufcols);

      if (!kBwd) {
        // ---- pass 1: per-query maximum over the tokens (per-warp partial rows, combined in a fixed order); the
        // first kCache units of a warp stay in registers for pass 2
        constexpr int kCache = 3;
        float cache[kCache][16];
        pmax[ew * 64 + lane] = -INFINITY;
        pmax[ew * 64 + 32 + lane] = -INFINITY;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kCache; ++k) {
          const int u = eh + 2 * k;
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (u < nunits && t * 128 + wq * 32 < p.N) {             // (warp-uniform) some lane of this warp holds a token
            load_unit(acc, t, j0, cache[k]);
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = (t * 128 + wq * 32 + lane < p.N) ? cache[k][q] : -INFINITY;
            const float red = reduce16<true>(v, lane);
            if ((lane & 1) == 0) {
              float* slot = pmax + ew * 64 + j0 + ridx;
              *slot = fmaxf(*slot, red);
            }
            __syncwarp();
          }
        }
        for (int u = eh + 2 * kCache; u < nunits; u += 2) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float v[16];
          load_unit(acc, t, j0, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = (t * 128 + wq * 32 + lane < p.N) ? v[q] : -INFINITY;
          const float red = reduce16<true>(v, lane);
          if ((lane & 1) == 0) {
            float* slot = pmax + ew * 64 + j0 + ridx;
            *slot = fmaxf(*slot, red);
          }
          __syncwarp();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarpsF) : "memory");
        float mx_lo = -INFINITY, mx_hi = -INFINITY;                // lane l: queries l and 32 + l
#pragma unroll
        for (int w = 0; w < kEpiWarpsF; ++w) {
          mx_lo = fmaxf(mx_lo, pmax[w * 64 + lane]);
          mx_hi = fmaxf(mx_hi, pmax[w * 64 + 32 + lane]);
        }
        // (every warp's drain of sample i - 1 -- the last reader of the totals in psum row 0 -- precedes the barrier above)
        psum[ew * 64 + lane] = 0.f;
        psum[ew * 64 + 32 + lane] = 0.f;
        __syncwarp();
        // ---- pass 2: exp, row sums, saved logits, operand blocks
        auto emit = [&](int t, int j0, const float (&v)[16]) {
          const int n = t * 128 + wq * 32 + lane;
          const bool valid = n < p.N;
          const int kb = n >> 6, tt = n & 63;
          uint8_t* blk = kb < p.nkb ? blk_gen + (size_t)kb * 2u * half_bytes : nullptr;
          float* srow = p.S + ((size_t)b * p.M + j0) * p.N + n;
          float e[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int m = j0 + q;                                  // warp-uniform
            const float mx = __shfl_sync(0xffffffffu, m < 32 ? mx_lo : mx_hi, m & 31);
            const bool on = valid && m < p.M;
            e[q] = on ? __expf(v[q] - mx) : 0.f;
            if (on) srow[(size_t)q * p.N] = v[q];
            if (blk && m < p.M) store_hilo(blk, m, tt, e[q]);
          }
          const float red = reduce16<false>(e, lane);
          if ((lane & 1) == 0) psum[ew * 64 + j0 + ridx] += red;
          __syncwarp();
        };
#pragma unroll
        for (int k = 0; k < kCache; ++k) {
          const int u = eh + 2 * k;
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (u < nunits && t * 128 + wq * 32 < p.N) emit(t, j0, cache[k]);
        }
        for (int u = eh + 2 * kCache; u < nunits; u += 2) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float v[16];
          load_unit(acc, t, j0, v);
          emit(t, j0, v);
        }
        fence_proxy_async();                                       // generic-proxy block writes -> visible to the MMAs
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarpsF) : "memory");
        // totals into row 0 of psum (read by the next drain) and the saved statistics
        if (ew == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = 32 * h + lane;
            float su = 0.f;
#pragma unroll
            for (int w = 0; w < kEpiWarpsF; ++w) su += psum[w * 64 + m];
            psum[m] = su;
            if (m < p.M) {
              p.rmax[(size_t)b * p.M + m] = h ? mx_hi : mx_lo;
              p.rsum[(size_t)b * p.M + m] = su;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarpsF) : "memory");
      } else {
        // ---- backward: dS = A (dA - delta), A recomputed from the saved logits and row statistics
        for (int u = eh; u < nunits; u += 2) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          const int n = t * 128 + wq * 32 + lane;
          const bool valid = n < p.N;
          float sv[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {                           // issued before the TMEM wait: independent loads
            const int m = min(j0 + q, p.M - 1);
            sv[q] = valid ? __ldg(p.S + ((size_t)b * p.M + m) * p.N + n) : 0.f;
          }
          float v[16];
          load_unit(acc, t, j0, v);
          const int kb = n >> 6, tt = n & 63;
          uint8_t* blk = kb < p.nkb ? blk_gen + (size_t)kb * 2u * half_bytes : nullptr;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int m = j0 + q;                                  // warp-uniform
            if (m < p.M) {
              const bool h = m >= 32;
              const float mx = __shfl_sync(0xffffffffu, h ? st_mx[1] : st_mx[0], m & 31);
              const float inv = __shfl_sync(0xffffffffu, h ? st_inv[1] : st_inv[0], m & 31);
              const float dl = __shfl_sync(0xffffffffu, h ? st_dl[1] : st_dl[0], m & 31);
              const float ds = valid ? __expf(sv[q] - mx) * inv * (v[q] - dl) : 0.f;
              if (blk) store_hilo(blk, m, tt, ds);
            }
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar);
      }
    }
    if (nmine > 0) {
      mbar_wait(pdone_bar, (uint32_t)((nmine - 1) & 1));
      tc_fence_after();
      drain(blockIdx.x + (nmine - 1) * gridDim.x);
    } else if (kBwd) {
      // a CTA without samples still owns a partial-gradient slice: zeros
      for (size_t o = (size_t)(warp - 4) * 32 + lane; o < (size_t)p.M * p.D; o += 32 * kEpiWarpsF)
        p.out[(size_t)blockIdx.x * p.M * p.D + o] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int round16(int v) { return (v + 15) / 16 * 16; }
int pow2_cols(int c) { int v = 32; while (v < c) v <<= 1; return v; }

struct FPlan {
  bool ok = false;
  int Mp, ntiles, nfull, tail_rows, nchunks, nkb, nsl, nslots, lead, nbuf, bufcols, pcol0, tmem_cols, qoff;
  size_t smem;
};

// L2 budget for the samples in flight (every CTA holds one sample plus the lead of the next between its two fetches)
constexpr size_t kL2Budget = 100ull << 20;

FPlan make_fplan(int N, int D, int M, int ctas) {
  FPlan pl;
  if (D % 128 != 0 || M < 1 || M > 64 || N < 1) return pl;
  pl.Mp = round16(M);
  pl.ntiles = (N + 127) / 128;
  const int rem = N - (pl.ntiles - 1) * 128;                      // rows of the last tile
  const int w_bytes = 2 * pl.Mp * 128;
  // the short last tile rides in the query chunk's slot when both fit in 16 KB
  pl.tail_rows = (rem < 128 && (rem + 7) / 8 * 8 * 128 + w_bytes <= kSlotBytes) ? (rem + 7) / 8 * 8 : 0;
  pl.nfull = pl.tail_rows ? pl.ntiles - 1 : pl.ntiles;
  pl.qoff = kSlotBytes - w_bytes;
  pl.nchunks = D / 64;
  pl.nkb = (N + 63) / 64;
  pl.nsl = D / 128;
  // TMEM: logit accumulators (hi and lo query rows are separate columns: ntiles x 2 Mp), double-buffered when
  // they fit twice beside the pooled accumulators (hi/lo operand blocks are two K-steps into ONE column: nsl x Mp)
  pl.bufcols = pl.ntiles * 2 * pl.Mp;
  pl.nbuf = (2 * pl.bufcols + pl.nsl * pl.Mp <= 512) ? 2 : 1;
  pl.pcol0 = pl.nbuf * pl.bufcols;
  const int cols = pl.pcol0 + pl.nsl * pl.Mp;
  if (cols > 512) return pl;
  pl.tmem_cols = pow2_cols(cols);
  const size_t fixed = 1024 /*alignment*/ + (size_t)2 * pl.Mp * 128 * pl.nkb + kStatFloats * 4 + 1024 /*barriers*/;
  const size_t avail = 227 * 1024;
  if (fixed + 6 * (size_t)kSlotBytes > avail) return pl;
  pl.nslots = (int)std::min<size_t>(12, (avail - fixed) / kSlotBytes);
  pl.smem = fixed + (size_t)pl.nslots * kSlotBytes;
  // chunks of the next sample fetched before the pooled phase: enough to cover the epilogue, bounded by the L2 budget
  const size_t sample = (size_t)N * D * 2;
  pl.lead = std::min(((g_debug >> 16) & 15) ? ((g_debug >> 16) & 15) - 1 : 2, pl.nchunks / 2);   // dev knob: bits 16-19 = lead + 1
  if (pl.nbuf == 1) pl.lead = 0;                                  // a lead needs the second logit buffer
  while (pl.lead > 0 && sample * ctas * (pl.nchunks + pl.lead) / pl.nchunks > kL2Budget) --pl.lead;
  if (sample * ctas > kL2Budget) return pl;
  pl.ok = true;
  return pl;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  EP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

// operand rows (Z, J, D) bf16 with hi/lo interleaved (row 2m: hi, 2m + 1: lo) seen as (D, 2, J/2, Z): one box =
// Mp rows of one kind x 64 d, rows past J/2 zero-filled
int make_w_tmap(CUtensorMap* m, const void* base, int D, int J, int Z, int Mp) {
  const uint64_t dims[4] = {(uint64_t)D, 2, (uint64_t)(J / 2), (uint64_t)Z};
  const uint64_t strides[3] = {(uint64_t)D * 2, (uint64_t)D * 4, (uint64_t)J * D * 2};
  const uint32_t box[4] = {64, 1, (uint32_t)Mp, 1};
  return make_tmap_bf16(m, base, 4, dims, strides, box);
}
int make_x_tmap(CUtensorMap* m, const void* x, int B, int N, int D, int rows) {
  const uint64_t dims[3] = {(uint64_t)D, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)D * 2, (uint64_t)D * N * 2};
  const uint32_t box[3] = {64, (uint32_t)rows, 1};
  return make_tmap_bf16(m, x, 3, dims, strides, box);
}

template <bool kBwd>
int launch_fused(const void* x, const void* w, int w_batched, int J, int B, int N, int D, int M, const FPlan& pl, FParams p,
                 int grid, cudaStream_t s) {
  CUtensorMap tm_x, tm_xt, tm_xb, tm_w;
  int rc;
  if ((rc = make_x_tmap(&tm_x, x, B, N, D, 128))) return rc;
  if ((rc = make_x_tmap(&tm_xt, x, B, N, D, pl.tail_rows ? pl.tail_rows : 128))) return rc;
  if ((rc = make_x_tmap(&tm_xb, x, B, N, D, 64))) return rc;
  if ((rc = make_w_tmap(&tm_w, w, D, J, w_batched ? B : 1, pl.Mp))) return rc;
  p.B = B; p.N = N; p.D = D; p.M = M; p.Mp = pl.Mp;
  p.ntiles = pl.ntiles; p.nfull = pl.nfull; p.tail_rows = pl.tail_rows;
  p.nchunks = pl.nchunks; p.nkb = pl.nkb; p.nsl = pl.nsl; p.nslots = pl.nslots; p.lead = pl.lead; p.nbuf = pl.nbuf;
  p.bufcols = pl.bufcols; p.pcol0 = pl.pcol0; p.tmem_cols = pl.tmem_cols; p.w_batched = w_batched; p.qoff = pl.qoff;
  if ((rc = set_smem(fused_kernel<kBwd>, pl.smem))) return rc;
  fused_kernel<kBwd><<<grid, kThreadsF, pl.smem, s>>>(tm_x, tm_xt, tm_xb, tm_w, p);
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fused
using namespace fused;

bool fused_supported(int N, int D, int M) { return make_fplan(N, D, M, stream_sms()).ok; }

// qhl: (J, D) bf16 hi/lo rows of the scaled queries
int fused_pool_fwd(const void* x, const void* qhl, int J, int B, int N, int D, int M, float* P, float* S, float* rowmax,
                   float* rowsum, int round_p, cudaStream_t s) {
  const int grid = std::min(B, stream_sms());
  const FPlan pl = make_fplan(N, D, M, grid);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  FParams p{};
  p.S = S; p.rmax = rowmax; p.rsum = rowsum; p.out = P; p.round_out = round_p;
  return launch_fused<false>(x, qhl, 0, J, B, N, D, M, pl, p, grid, s);
}

// dphl: (B, J, D) bf16 hi/lo rows of dP; part: [grid][M][D] partial query gradients (*groups_out = grid)
int fused_pool_bwd(const void* x, const void* dphl, int J, int B, int N, int D, int M, const float* S, const float* rowmax,
                   const float* rowsum, const float* delta, float* part, int* groups_out, cudaStream_t s) {
  const int grid = std::min(B, stream_sms());
  const FPlan pl = make_fplan(N, D, M, grid);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  FParams p{};
  p.S = const_cast<float*>(S); p.rmax = const_cast<float*>(rowmax); p.rsum = const_cast<float*>(rowsum);
  p.delta = delta; p.out = part;
  *groups_out = grid;
  return launch_fused<true>(x, dphl, 1, J, B, N, D, M, pl, p, grid, s);
}

}  // namespace ep
