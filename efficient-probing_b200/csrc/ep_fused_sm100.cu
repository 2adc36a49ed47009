// One-pass fused EP pooling kernels (sm_100a): the tokens of a sample cross HBM ONCE per direction.
//
//   forward   (poolings/ep.py:39-45)   S = x q^T  ->  softmax over tokens  ->  P = A x           in one kernel
//   backward      dA = x dP^T -> dS = A (dA - delta)  ->  dq += dS^T x      in one kernel
//
// A sample (N x D bf16: 526 KB at N=257, D=1024) does not fit shared memory, and the softmax needs all of a
// sample's logits before the first pooled product.  So a persistent CTA walks its samples and fetches each one
// twice through TMA:
//   L(i)  d-chunks [128 tokens x 64 d] (K-major A operand)  -> logits / dA of all token tiles accumulate in TMEM
//   P(i)  bricks  [128 tokens x 128 d] (the same swizzled bytes read as the MN-major A operand) -> pooled sums
// The second fetch finds the sample in L2 (148 CTAs x 526 KB = 78 MB of the 126 MB; first fetch evict_last, second
// evict_first; tools/dev_l2_probe.cu: 108 us for c2 in this order against 89 us for a single pass and 2 x 89 us for
// two kernels).  Between the phases nothing touches global memory: exp(S - max) (forward) or dS (backward) go from
// the TMEM epilogue straight into shared memory as the UMMA B operand of the pooled phase.
//
// Schedule.  HBM only stays busy if first fetches never stop, so the bricks of P(i) run CONCURRENTLY with the chunks of
// L(i+1) (the first `lead` chunks may go ahead of the first brick and cover the softmax epilogue of sample i; after
// that L(i+1) advances in step with P(i), which bounds the L2 footprint to one sample plus the lead per CTA).  That
// needs the logit accumulators double-buffered, which decides the operand layouts:
// fp32 operands (queries, probabilities, dP, dS) are bf16 hi/lo pairs as in ep_pool_sm100.cu; in the logit phase the
// pair is either two columns of one N = 2 Mp MMA (added when leaving TMEM) or two K-steps into ONE column (hi rows
// and lo rows as separate B operands, twice the MMAs), whichever lets TMEM hold two logit buffers next to the pooled
// accumulators (make_fplan; c2 forward: 2 x 3 tiles x 64 columns + 2 x 64).  Forward pooled phase: hi and lo are
// separate columns of one N = 2 Mp MMA, D is walked in groups of slices whose accumulators ping-pong between two
// TMEM buffers, so a group drains while the next accumulates.  Backward: the query gradient accumulates over the
// whole launch, all of D resident (8 slices x 32 columns, hi/lo as two K-steps).  A tcgen05.mma with N <= 64
// occupies the tensor core for 40-48 cycles whatever N (tools/dev_umma_probe.cu: 39/40/48/64/128 cycles at
// N = 16/32/64/128/256), so the MMA warp issues from uniform registers with the whole warp converged.
//
// Pipeline bookkeeping is sized by what the primitives cost when issued from one thread (tools/dev_issue_probe.cu, SM
// cycles): a TMA load 120-170, mbarrier.try_wait on a completed phase 170, tcgen05.commit 45.  With a barrier pair per
// 16 KB slot one producer thread and the MMA warp each spent ~20 us per sample on these alone.  So there are two
// rings with their own producer thread and few, large stages: the L ring holds whole d-chunks (query chunk + short
// tail tile + the full token tiles, the tiles one TMA instruction), the P ring "tall bricks" [128 tokens x 128 d] (two
// token blocks of a d-slice, one TMA instruction), and each ring has its own MMA-issuing warp, so that one warp's
// barrier bookkeeping overlaps the other's MMAs (issue blocks while the tensor core's short queue is full).  The two
// streams run independently; what orders them are monotonic progress counters in shared memory: the L warp starts
// chunk c of sample i+1 only when the P warp has issued the matching share of sample i's bricks (beyond the first
// `lead` chunks) and the epilogue has released the logit buffer.  The L producer also prefetches into L2 a few chunks
// ahead (cp.async.bulk.prefetch.tensor), so that the two chunk stages shared memory has room for cover L2 latency.
//
// What was measured (profiles/r02_*, DESIGN.md sections 5-6):
//  * L2.  With 148 CTAs the live tokens (148 x (one sample + lead) = 80-100 MB at c2) only partly survive in L2
//    between their two fetches: dram__bytes_read is 0.78 GB (forward) / 0.94 GB (backward, incl. 0.13 GB of dP) per
//    pass at c2 against 0.54 GB of tokens; the miss rate falls gradually with fewer CTAs (0.68 GB at 104, 0.55 GB at
//    74) but every CTA taken away costs more than its misses, so all SMs are used.
//  * tcgen05.mma issue is effectively synchronous for the issuing thread (tools/dev_mix_probe.cu: 48 cycles per
//    N = 64 MMA whether one warp issues or two interleave, and any extra latency in the issuing loop shows up in
//    full), so every barrier round trip of an MMA warp idles the tensor core unless the other MMA warp is issuing.
//  * The epilogue is bound by three narrow pipes (tools/dev_epi_probe.cu, cycles per 16-query unit with 16 warps):
//    the 16-lane XU pipe shared by ex2 and single-value F2F conversions (exp 540, + hi/lo split with F2F 1580, with
//    packed F2FP 835), the load/store unit on the saved logits (rows of N floats, 4-byte aligned only: 1440), and TMEM
//    reads (64 B / clock).  Hence: packed conversions, 16-byte pooled-token stores through a register transpose, the
//    logits of a warp's first units kept in registers between the softmax passes, and the saved logits written after
//    the pooled phase has been released (by the warps that have no drain work).
//  * Code size: the kernel is executed phase by phase by 20 warps in five roles; with every developer knob compiled
//    in it was 145 KB of SASS and its instruction-cache misses cost ~10 % -- the knobs now live in a separate
//    instantiation (kDev) and each epilogue phase exists once (62 KB forward, 90 KB backward).
//  c2, M = 32: forward 249 -> 213 us, backward 227 -> 192 us (the two kernels each replaces: 243 / 251 us).
//
// Pair mode (developer knob, ep_set_debug bit 29).  CTAs 2k and 2k+1 share sample k, 74 + k, ...: each takes half of
// the d-slices in BOTH phases (half the chunks, half the bricks), the partial logits are exchanged through a global
// scratch buffer that lives in L2 (written, fenced, flag released; the partner polls the flag with an acquire load --
// all 148 CTAs are co-resident, one per SM), and both CTAs run the same softmax on the sum, so each has every
// probability for its own slices.  Live tokens: 74 samples instead of 148 -- DRAM reads fall to 0.73 GB at c2 -- but
// the softmax epilogue is now done twice per sample and the pass takes 340 us against 245 us: not the default.
//
// Warp roles (one CTA per SM): warp 0 L producer, warp 1 L MMA issuer, warp 2 TMEM allocator + P MMA issuer, warp 3 P
// producer, warps 4.. epilogue (kEW warps, kEW / 4 per TMEM lane quadrant).
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "ep_common.cuh"
#include "ep_ptx.cuh"
#include "ep_sm100.cuh"

namespace ep {
using namespace ptx;

namespace fused {
constexpr int kSlotBytes = 16384;      // ring slot: [128 tokens x 64 d] chunk tile, or [64 tokens x 128 d] brick
constexpr int kEW = 16;                // epilogue warps
constexpr int kSub = kEW / 4;          // ... per TMEM lane quadrant
constexpr int kThreadsF = 32 * (4 + kEW);
__host__ __device__ constexpr int stat_floats(int Mp) { return 2 * kEW * Mp + 128; }   // per-warp partial max / sum [kEW][Mp] + [2][64] row sums
constexpr int kCache = 2;              // logit units a warp keeps in registers between the two softmax passes

// developer trace (ep_set_debug bit 11): clock64 stamps of CTA 0's MMA warp and epilogue warp 4, 16 stamps x 8 samples
__device__ long long g_trace[128];
#define EP_TRACE(i, k) do { if (trace && blockIdx.x == 0 && (i) < 8 && lane == 0) g_trace[(i) * 16 + (k)] = clock64(); } while (0)

struct FParams {
  int B, N, D, M, Mp;                  // Mp = M rounded up to 16: UMMA N, accumulator columns per tile / slice
  int ntiles, nfull, tail_rows;        // token tiles of 128; nfull of them loaded as full boxes; tail_rows > 0: the last
                                       // tile is a short box sharing its slot with the query chunk
  int nchunks, nkb, nsl, nL, nP, lbytes, lead, nbuf, bufcols, pcol0, tmem_cols, w_batched, qoff, toff;   // nL/nP ring stages; chunk stage = [tail tile @0][query chunk @qoff][token tiles @toff]
  int nkp, pf;                         // tall bricks per slice (= ceil(nkb / 2)); L2 prefetch distance in chunks
  int nstream;                         // first chunks of a sample fetched evict_first: they are re-read last and would not
                                       // survive in L2 anyway; keeping them out protects the rest (see make_fplan)
  int nslg, G, pbufcols;               // pooled phase: G groups of <= nslg slices, accumulator buffers of pbufcols columns
  int lsplit;                          // logit phase: 1 = hi/lo query rows as two K-steps into one column (N = Mp),
                                       //              0 = as separate columns of one N = 2 Mp MMA (added in the epilogue)
  int lcolw;                           // logit accumulator columns per token tile (Mp or 2 Mp)
  int kl, last_rows;                   // k-steps (of 16 tokens) and token rows of the last tall brick of a slice
  float* S;                            // fwd: logits out (B, M, N);  bwd: saved logits in
  float* rmax; float* rsum;            // fwd: out;  bwd: in
  const float* delta;                  // bwd: (B, M)
  float* out;                          // fwd: P as bf16 hi/lo rows (B, M, 2, D) or fp32 (B, M, D);  bwd: partial dq [grid][M][D]
  int round_out;
  int pair;                            // 1: two CTAs (2k, 2k+1) share a sample, each takes half of D (see header)
  int nslA;                            // pair mode: d-slices of CTA 2k (the rest belong to CTA 2k+1)
  float* xchg;                         // pair mode: partial-logit exchange [cta][2][ntiles][Mp][128] fp32
  int* flags;                          // pair mode: per CTA, samples whose partial logits are published (zeroed before launch)
  int trace;
  int noepi;                           // developer: epilogue warps only run the barrier protocol
  int nomma;                           // developer: skip the MMA instructions (timing experiments; results are garbage)
  int skip;                            // developer: bit 0 skip the operand-block stores, bit 1 the row sums, bit 2 the exp (garbage results)
};

__device__ __forceinline__ void tma_load_4d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                 int c3, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile.L2::cache_hint [%0, {%1, %2, %3}], %4;"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
               : "memory");
}

// 16 values per lane reduced over the 32 lanes of a warp in 16 shuffles (transposing butterfly): lane l ends with
// value reduce16_index(l) reduced over all lanes, replicated on 2 lanes.
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

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kPBytes = 2 * kSlotBytes;   // P ring stage: a tall brick [128 tokens x 128 d] as two 64-d halves

// kBwd = false: forward (logits, softmax, pooled tokens).  kBwd = true: backward (dA, dS, query gradient).
// kDev = true compiles the developer knobs in (trace stamps, noepi / nomma timing experiments, pair mode)
template <bool kBwd, bool kDev>
__global__ void __launch_bounds__(kThreadsF, 1)
fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x1,
             const __grid_constant__ CUtensorMap tm_xt, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_bl,
             const __grid_constant__ CUtensorMap tm_w, const FParams p) {
  // smem: [L ring: nL x lbytes][P ring: nP x 32 KB][operand blocks: nkb x (hi [Mp x 128 B], lo [Mp x 128 B])][stats][barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (ring - smem_u32(smem_raw));
  const uint32_t half_bytes = (uint32_t)p.Mp * 128u;            // one hi or lo operand block
  const uint32_t pring = ring + (uint32_t)(p.nL * p.lbytes);
  const uint32_t blk_base = pring + (uint32_t)p.nP * kPBytes;
  const uint32_t blk_bytes = 2u * half_bytes * (uint32_t)p.nkb;
  const uint32_t stat_base = blk_base + blk_bytes;
  const uint32_t bar_base = stat_base + (uint32_t)stat_floats(p.Mp) * 4u;
  auto lfull_bar = [&](int s) { return bar_base + 8u * s; };              // L ring stage s loaded / consumed
  auto lempty_bar = [&](int s) { return bar_base + 32u + 8u * s; };
  auto pfull_bar = [&](int s) { return bar_base + 64u + 8u * s; };         // P ring
  auto pempty_bar = [&](int s) { return bar_base + 96u + 8u * s; };
  const uint32_t misc = bar_base + 128u;
  auto tfull_bar = [&](int b) { return misc + 8u * b; };         // logit accumulators of buffer b complete
  const uint32_t eready_bar = misc + 16u;                         // operand blocks of the current sample written (and
                                                                  // its logit accumulators read: that buffer is free)
  auto pdone_bar = [&](int b) { return misc + 24u + 8u * b; };   // pooled MMAs into accumulator buffer b complete
  auto pfree_bar = [&](int b) { return misc + 40u + 8u * b; };   // (fwd) accumulator buffer b drained
  const uint32_t tmem_slot = misc + 56u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - ring));
  // monotonic progress counters (written by one thread, polled by another): tall bricks issued by the P warp, samples
  // whose logit accumulators the epilogue has finished reading
  volatile int* p_issued = reinterpret_cast<volatile int*>(gen + (misc + 64u - ring));
  volatile int* e_done = reinterpret_cast<volatile int*>(gen + (misc + 68u - ring));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool trace = kDev && p.trace, noepi = kDev && p.noepi, nomma = kDev && p.nomma, pair = kDev && p.pair;

  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) { mbar_init(lfull_bar(q), 1); mbar_init(lempty_bar(q), 1); mbar_init(pfull_bar(q), 1); mbar_init(pempty_bar(q), 1); }
    for (int b = 0; b < 2; ++b) mbar_init(tfull_bar(b), 1);
    mbar_init(eready_bar, kEW);
    for (int b = 0; b < 2; ++b) { mbar_init(pdone_bar(b), 1); mbar_init(pfree_bar(b), kEW); }
    *p_issued = 0;
    *e_done = 0;
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_x1); prefetch_tmap(&tm_xt); prefetch_tmap(&tm_w); }
  if (warp == 3 && lane == 0) { prefetch_tmap(&tm_b); prefetch_tmap(&tm_bl); }
  if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  // operand blocks start as zeros: rows m >= M and tokens that no epilogue thread owns stay zero for the whole launch
  for (uint32_t o = threadIdx.x * 16u; o < blk_bytes; o += kThreadsF * 16u)
    *reinterpret_cast<uint4*>(gen + (blk_base - ring) + o) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();       // everything above is CTA-local set-up; the tokens' companions (queries, dP, statistics) come from the previous kernel
  const uint32_t tmem_base = *tmem_slot_ptr;
  // this CTA's samples, d-chunks [c_lo, c_lo + nch) and d-slices [sl_lo, sl_lo + nslh)
  const int cta = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ncta = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int half = pair ? (int)(blockIdx.x & 1) : 0;
  const int sl_lo = half ? p.nslA : 0, nslh = pair ? (half ? p.nsl - p.nslA : p.nslA) : p.nsl;
  const int c_lo = 2 * sl_lo, nch = 2 * nslh;
  const int G = (nslh + p.nslg - 1) / p.nslg;                     // pooled groups of this CTA (fwd; bwd plans one group)
  int nmine = 0;
  for (int b = cta; b < p.B; b += ncta) ++nmine;
  const uint32_t w_bytes = 2u * half_bytes;                       // query / dP chunk: hi rows then lo rows
  const bool mixed_tail = p.tail_rows > 0;                        // the short last tile shares the query chunk's slot

  // tiles of a chunk are loaded by boxes of up to 256 token rows
  const int tile_rows = p.nfull * 128;

  if (warp == 0) {
    // ---- L producer: chunk after chunk, sample after sample, into the L ring; stage = [short tail tile][query chunk at
    // qoff][nfull x 16 KB token tiles at toff]
    if (lane == 0 && nmine > 0) {
      const uint64_t pol_last = policy_evict_last();
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_w = p.w_batched ? pol_stream : pol_last;           // per-sample dP rows are read once
      const uint32_t tx = w_bytes + (mixed_tail ? (uint32_t)p.tail_rows * 128u : 0u) + (uint32_t)p.nfull * kSlotBytes;
      int s = 0;
      uint32_t ph = 0;
      long long t_wait = 0, t_begin = clock64();
      for (int i = 0; i < nmine; ++i) {
        const int b = cta + i * ncta;
        const int zb = p.w_batched ? b : 0;
        for (int c = c_lo; c < c_lo + nch; ++c) {
          const long long t0 = trace ? clock64() : 0;
          mbar_wait(lempty_bar(s), ph ^ 1u);
          if (trace) t_wait += clock64() - t0;
          const uint32_t bar = lfull_bar(s), dst = ring + (uint32_t)(s * p.lbytes);
          mbar_arrive_expect_tx(bar, tx);
          const uint64_t pol_x = (c - c_lo) < p.nstream ? pol_stream : pol_last;
          for (int r = 0; r < tile_rows; r += 256)                 // two tiles per instruction, a last odd one alone
            tma_load_3d_hint(dst + (uint32_t)p.toff + (uint32_t)r * 128u, tile_rows - r >= 256 ? &tm_x : &tm_x1, bar, c * 64, r, b, pol_x);
          tma_load_4d_hint(dst + (uint32_t)p.qoff, &tm_w, bar, c * 64, 0, 0, zb, pol_w);
          tma_load_4d_hint(dst + (uint32_t)p.qoff + half_bytes, &tm_w, bar, c * 64, 1, 0, zb, pol_w);
          if (mixed_tail) tma_load_3d_hint(dst, &tm_xt, bar, c * 64, tile_rows, b, pol_x);
          if (p.pf > 0) {                                          // the chunk pf ahead (maybe of the next sample) -> L2
            int cp = c + p.pf, bp = b;
            if (cp >= c_lo + nch) { cp -= nch; bp += ncta; }
            if (bp < p.B) {
              for (int r = 0; r < tile_rows; r += 256) tma_prefetch_3d(tile_rows - r >= 256 ? &tm_x : &tm_x1, cp * 64, r, bp, pol_last);
              if (mixed_tail) tma_prefetch_3d(&tm_xt, cp * 64, tile_rows, bp, pol_last);
            }
          }
          if (++s == p.nL) { s = 0; ph ^= 1u; }
        }
      }
      if (trace && blockIdx.x == 0) { g_trace[120] = t_wait; g_trace[121] = clock64() - t_begin; }   // L producer: ring-full wait, total
    }
  } else if (warp == 3) {
    // ---- P producer: the tall bricks of sample after sample, slice after slice (the MMA warp's group order is slice
    // order), into the P ring
    if (lane == 0 && nmine > 0) {
      const uint64_t pol_first = policy_evict_first();
      const uint32_t last_bytes = (uint32_t)p.last_rows * 256u;
      int s = 0;
      uint32_t ph = 0;
      long long t_wait = 0, t_begin = clock64();
      for (int i = 0; i < nmine; ++i) {
        const int b = cta + i * ncta;
        // slices in DESCENDING order: the logit phase fetched d ascending, so the most recently fetched lines -- the
        // ones most likely still in L2 -- are re-read first (under an LRU-like policy, re-reading in fetch order
        // would miss everything as soon as the live set exceeds the capacity)
        for (int sl = sl_lo + nslh - 1; sl >= sl_lo; --sl)
          for (int kp = 0; kp < p.nkp; ++kp) {
            const long long t0 = trace ? clock64() : 0;
            mbar_wait(pempty_bar(s), ph ^ 1u);
            if (trace) t_wait += clock64() - t0;
            const bool last = kp == p.nkp - 1 && p.last_rows < 128;
            mbar_arrive_expect_tx(pfull_bar(s), last ? last_bytes : (uint32_t)kPBytes);
            tma_load_4d_hint(pring + (uint32_t)s * kPBytes, last ? &tm_bl : &tm_b, pfull_bar(s), 0, kp * 128, 2 * sl, b, pol_first);
            if (++s == p.nP) { s = 0; ph ^= 1u; }
          }
      }
      if (trace && blockIdx.x == 0) { g_trace[125] = t_wait; g_trace[126] = clock64() - t_begin; }   // P producer
    }
  } else if (warp == 1) {
    // ---- L MMA issue: the whole warp walks the chunks (uniform control flow, descriptors in uniform registers), one
    // elected lane issues
    if (nmine > 0) {
      const bool leader = elect_one();
      const uint32_t idesc_L = idesc_bf16(128, p.lsplit ? p.Mp : 2 * p.Mp, 0, 0);
      const uint64_t dK = smem_desc_sw128(0, 16, 1024);           // + (address >> 4)
      const uint64_t lo_off = (uint64_t)(half_bytes >> 4);
      const int nst = nslh * p.nkp;                                // tall bricks per sample (of this CTA)
      int ls = 0;
      uint32_t lph = 0;
      long long t_wait = 0, t_gate = 0, t_issue = 0, t_begin = clock64();
      for (int j = 0; j < nmine; ++j) {
        const uint32_t acc = tmem_base + (uint32_t)((j % p.nbuf) * p.bufcols);
        for (int c = 0; c < nch; ++c) {
          // order against the other stream: the logit buffer is free (its previous sample's epilogue has read it), and
          // chunk c runs at most `lead` chunks ahead of the matching share of the previous sample's bricks
          const long long tg = trace ? clock64() : 0;
          if (c == 0 && j >= p.nbuf) {
            while (*e_done < j - p.nbuf + 1) __nanosleep(64);
            tc_fence_after();
          }
          if (j > 0) {
            const int need = (j - 1) * nst + (c > p.lead ? (c - p.lead) * nst / nch : 0);
            while (*p_issued < need) __nanosleep(32);
          }
          if (trace) t_gate += clock64() - tg;
          const long long t0 = trace ? clock64() : 0;
          mbar_wait(lfull_bar(ls), lph);
          if (trace) t_wait += clock64() - t0;
          const long long ti = trace ? clock64() : 0;
          const uint32_t st0 = ring + (uint32_t)(ls * p.lbytes);
          const uint64_t bd = dK + (uint64_t)((st0 + (uint32_t)p.qoff) >> 4);
          if (leader && !nomma) {
            for (int t = 0; t < p.ntiles; ++t) {
              // full tiles at toff, the short tail tile at the start of the stage (the MMA reads 128 rows there: the
              // rows past the tail are the query chunk / first tile and only reach accumulator rows n >= N)
              const uint64_t ad = dK + (uint64_t)((t < p.nfull ? st0 + (uint32_t)p.toff + (uint32_t)t * kSlotBytes : st0) >> 4);
              const uint32_t d = acc + (uint32_t)(t * p.lcolw);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_f16(d, ad + 2u * k, bd + 2u * k, idesc_L, (uint32_t)((c | k) != 0));
                if (p.lsplit) umma_f16(d, ad + 2u * k, bd + lo_off + 2u * k, idesc_L, 1u);
              }
            }
          }
          if (leader) umma_commit(lempty_bar(ls));
          __syncwarp();
          if (trace) t_issue += clock64() - ti;
          if (++ls == p.nL) { ls = 0; lph ^= 1u; }
        }
        if (leader) umma_commit(tfull_bar(j % p.nbuf));
        __syncwarp();
        EP_TRACE(j, 2);                                            // L warp: logits of sample j issued
      }
      if (trace && blockIdx.x == 0 && lane == 0) {               // L warp: load-starved wait, ordering wait, total
        g_trace[122] = t_wait; g_trace[123] = t_gate; g_trace[124] = clock64() - t_begin; g_trace[112] = t_issue;
      }
    }
  } else if (warp == 2) {
    // ---- P MMA issue (this warp also owns the TMEM allocation)
    if (nmine > 0) {
      const bool leader = elect_one();
      const uint32_t idesc_P = idesc_bf16(128, kBwd ? p.Mp : 2 * p.Mp, 1, 0);   // pooled: A (tokens as K) is MN-major
      const uint64_t dK = smem_desc_sw128(0, 16, 1024);
      const uint64_t dMN = smem_desc_sw128(0, kPBytes / 2, 1024);  // halves of a tall brick are 128 rows apart
      const uint64_t dMN_last = smem_desc_sw128(0, (uint32_t)p.last_rows * 128u, 1024);
      const uint64_t lo_off = (uint64_t)(half_bytes >> 4);
      const uint32_t pcolw = kBwd ? (uint32_t)p.Mp : 2u * (uint32_t)p.Mp;      // accumulator columns per slice
      int ps = 0, issued = 0;
      uint32_t pph = 0;
      long long t_wait = 0, t_gate = 0, t_free = 0, t_issue = 0, t_begin = clock64();
      for (int i = 0; i < nmine; ++i) {
        EP_TRACE(i, 0);                                            // P warp: start waiting for the operand blocks
        const long long tg = trace ? clock64() : 0;
        mbar_wait(eready_bar, (uint32_t)(i & 1));                  // operand blocks of sample i are in shared memory
        tc_fence_after();
        if (trace) t_gate += clock64() - tg;
        EP_TRACE(i, 1);                                            // P warp: pooled phase starts
        for (int g = 0; g < (kBwd ? 1 : G); ++g) {
          const int gg = kBwd ? 2 * i : i * G + g;                 // bwd: one group per sample, always buffer 0
          if (!kBwd) {                                             // this buffer's previous group has been drained
            const long long tf = trace ? clock64() : 0;
            mbar_wait(pfree_bar(gg & 1), (((uint32_t)(gg >> 1)) & 1u) ^ 1u);
            tc_fence_after();
            if (trace) t_free += clock64() - tf;
          }
          const int gs = kBwd ? nslh : min(p.nslg, nslh - g * p.nslg);
          for (int slg = 0; slg < gs; ++slg) {
            const uint32_t pacc = tmem_base + (uint32_t)p.pcol0 + (kBwd ? 0u : (uint32_t)((gg & 1) * p.pbufcols)) + (uint32_t)slg * pcolw;
            for (int kp = 0; kp < p.nkp; ++kp) {
              const long long t0 = trace ? clock64() : 0;
              mbar_wait(pfull_bar(ps), pph);
              if (trace) t_wait += clock64() - t0;
              const long long ti = trace ? clock64() : 0;
              const bool last = kp == p.nkp - 1 && p.last_rows < 128;
              const int ks = last ? p.kl : 8;
              const uint64_t ad = (last ? dMN_last : dMN) + (uint64_t)((pring + (uint32_t)ps * kPBytes) >> 4);
              const uint64_t bd0 = dK + (uint64_t)((blk_base + (uint32_t)(2 * kp) * 2u * half_bytes) >> 4);
              if (leader && !nomma) {
                for (int k = 0; k < ks; ++k) {
                  const uint64_t bd = bd0 + (uint64_t)((k >> 2) * (2u * half_bytes >> 4)) + 2u * (k & 3);
                  if (kBwd) {
                    umma_f16(pacc, ad + 128u * k, bd, idesc_P, (uint32_t)((i | kp | k) != 0));
                    umma_f16(pacc, ad + 128u * k, bd + lo_off, idesc_P, 1u);
                  } else {
                    umma_f16(pacc, ad + 128u * k, bd, idesc_P, (uint32_t)((kp | k) != 0));
                  }
                }
              }
              ++issued;
              if (leader) { umma_commit(pempty_bar(ps)); *p_issued = issued; }
              __syncwarp();
              if (trace) t_issue += clock64() - ti;
              if (++ps == p.nP) { ps = 0; pph ^= 1u; }
            }
          }
          if (leader) umma_commit(pdone_bar(gg & 1));
          __syncwarp();
        }
        EP_TRACE(i, 3);                                            // P warp: sample i pooled
      }
      if (trace && blockIdx.x == 0 && lane == 0) {               // P warp: load-starved wait, operand-block wait, total
        g_trace[117] = t_wait; g_trace[118] = t_gate; g_trace[119] = clock64() - t_begin; g_trace[113] = t_free; g_trace[114] = t_issue;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, wq = ew & 3, es = ew >> 2;            // quadrant, sub-warp within the quadrant
    const int upt = p.Mp >> 4;                                     // 16-query units per tile / slice
    const int SW = p.Mp;                                                // table row stride
    float* pmax = reinterpret_cast<float*>(gen + (stat_base - ring));   // [kEW][Mp]
    float* psum = pmax + kEW * SW;                                      // [kEW][Mp]
    float* tot = psum + kEW * SW;                                       // [2][64] softmax row sums, by sample parity
    const bool has_lo = lane < SW, has_hi = 32 + lane < SW;             // this lane's table columns l and 32 + l exist
    uint8_t* blk_gen = gen + (blk_base - ring);
    const int ridx = reduce16_index(lane);
    const uint32_t lane_base = ((uint32_t)(wq * 32)) << 16;
    const int nunits = p.ntiles * upt;

    // fwd: group g of sample ordinal i (accumulator buffer gg & 1) -> normalised P rows in global memory
    auto drain_group = [&](int b, int i, int g) {
      const int gg = i * G + g;
      mbar_wait(pdone_bar(gg & 1), ((uint32_t)(gg >> 1)) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + (uint32_t)(p.pcol0 + (gg & 1) * p.pbufcols);
      const int gs = min(p.nslg, nslh - g * p.nslg);
      float invl[2] = {1.f, 1.f};                                  // lane l keeps 1/rowsum of queries l and 32 + l
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (h * 32 + lane < p.M) invl[h] = 1.f / tot[(i & 1) * 64 + h * 32 + lane];
      for (int u = es; u < gs * upt; u += kSub) {
        const int slg = u / upt, j0 = (u - slg * upt) << 4;
        const int d = (sl_lo + nslh - 1 - (g * p.nslg + slg)) * 128 + wq * 32 + lane;   // pooled phase walks the slices downwards
        uint32_t rh[16], rl[16];
        tmem_ld16(acc + (uint32_t)(slg * 2 * p.Mp + j0), rh);
        tmem_ld16(acc + (uint32_t)(slg * 2 * p.Mp + p.Mp + j0), rl);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int m = j0 + q;                                    // warp-uniform
          v[q] = (__uint_as_float(rh[q]) + __uint_as_float(rl[q])) * __shfl_sync(0xffffffffu, m < 32 ? invl[0] : invl[1], m & 31);
        }
        if (p.round_out) {
          // P as bf16 hi/lo rows (b, m, {hi, lo}, d).  A lane holds one channel of 16 queries; an 8 x 8 transpose of
          // query pairs among the 8 lanes of a group (3 butterfly stages) leaves lane j with queries 2j, 2j + 1 of the
          // group's 8 consecutive channels: 16-byte streaming stores instead of 2-byte ones
          uint32_t hp[8], lp[8];
#pragma unroll
          for (int pr = 0; pr < 8; ++pr) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * pr], v[2 * pr + 1]);
            const float2 hf = __bfloat1622float2(h);
            const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * pr] - hf.x, v[2 * pr + 1] - hf.y);
            hp[pr] = *reinterpret_cast<const uint32_t*>(&h);
            lp[pr] = *reinterpret_cast<const uint32_t*>(&l);
          }
#pragma unroll
          for (int st = 4; st >= 1; st >>= 1) {
            const bool up = (lane & st) != 0;
#pragma unroll
            for (int pr = 0; pr < 8; ++pr) {
              if (pr & st) continue;
              const uint32_t sh = __shfl_xor_sync(0xffffffffu, up ? hp[pr] : hp[pr | st], st);
              const uint32_t sl = __shfl_xor_sync(0xffffffffu, up ? lp[pr] : lp[pr | st], st);
              if (up) { hp[pr] = sh; lp[pr] = sl; } else { hp[pr | st] = sh; lp[pr | st] = sl; }
            }
          }
          // streaming stores (evict-first in L2): the outputs must not displace the tokens waiting for their second fetch
          const int d8 = d - (lane & 7);                           // first of the group's 8 channels
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const int m = j0 + 2 * (lane & 7) + hq;
            if (m < p.M) {
              const uint32_t sel = hq ? 0x7632u : 0x5410u;
              const uint4 oh = make_uint4(__byte_perm(hp[0], hp[1], sel), __byte_perm(hp[2], hp[3], sel),
                                          __byte_perm(hp[4], hp[5], sel), __byte_perm(hp[6], hp[7], sel));
              const uint4 ol = make_uint4(__byte_perm(lp[0], lp[1], sel), __byte_perm(lp[2], lp[3], sel),
                                          __byte_perm(lp[4], lp[5], sel), __byte_perm(lp[6], lp[7], sel));
              unsigned short* prow = reinterpret_cast<unsigned short*>(p.out) + (((size_t)b * p.M + m) * 2) * p.D + d8;
              __stcs(reinterpret_cast<uint4*>(prow), oh);
              __stcs(reinterpret_cast<uint4*>(prow + p.D), ol);
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (j0 + q < p.M) __stcs(p.out + ((size_t)b * p.M + j0 + q) * p.D + d, v[q]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfree_bar(gg & 1));
    };
    auto load_unit = [&](uint32_t acc, int t, int j0, float (&v)[16]) {
      uint32_t r[16];
      tmem_ld16(acc + (uint32_t)(t * p.lcolw + j0), r);
      if (p.lsplit) {
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
      } else {                                                     // hi and lo query rows are separate columns
        uint32_t r2[16];
        tmem_ld16(acc + (uint32_t)(t * p.lcolw + p.Mp + j0), r2);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]) + __uint_as_float(r2[q]);
      }
    };
    // 16 queries (j0 ..) of this lane's token n -> the bf16 hi / lo operand rows of its 64-token block.  Conversions
    // are the packed kind (two values per instruction on the FMA pipe; the single-value F2F runs on the 16-lane XU
    // pipe).  2-byte shared-memory stores were the slowest part of the epilogue (32 per unit and lane), so the values
    // go through an 8 x 8 transpose of query pairs among the 8 lanes of a group first: lane j ends with queries 2j,
    // 2j + 1 of the group's 8 consecutive tokens = one 16-byte swizzle segment per row.  Rows m >= M stay zero.
    auto store_rows = [&](int n0, int n, int j0, const float (&x)[16]) {
      (void)n;
      uint32_t hp[8], lp[8];
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * pr], x[2 * pr + 1]);
        hp[pr] = *reinterpret_cast<const uint32_t*>(&h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[2 * pr] - __uint_as_float(hp[pr] << 16), x[2 * pr + 1] - __uint_as_float(hp[pr] & 0xffff0000u));
        lp[pr] = *reinterpret_cast<const uint32_t*>(&l2);
      }
#pragma unroll
      for (int st = 4; st >= 1; st >>= 1) {
        const bool up = (lane & st) != 0;
#pragma unroll
        for (int pr = 0; pr < 8; ++pr) {
          if (pr & st) continue;
          const uint32_t sh = __shfl_xor_sync(0xffffffffu, up ? hp[pr] : hp[pr | st], st);
          const uint32_t sl = __shfl_xor_sync(0xffffffffu, up ? lp[pr] : lp[pr | st], st);
          if (up) { hp[pr] = sh; lp[pr] = sl; } else { hp[pr | st] = sh; lp[pr | st] = sl; }
        }
      }
      uint8_t* blk = blk_gen + (size_t)(n0 >> 6) * 2u * half_bytes;
      const uint32_t c3 = (((uint32_t)n0 & 63u) >> 3) + (uint32_t)(lane >> 3);   // the group's 16-byte segment of a row
#pragma unroll
      for (int hq = 0; hq < 2; ++hq) {
        const uint32_t m = (uint32_t)j0 + 2u * (uint32_t)(lane & 7) + (uint32_t)hq;
        if ((int)m < p.M) {
          const uint32_t sel = hq ? 0x7632u : 0x5410u;
          const uint32_t off = m * 128u + ((c3 ^ (m & 7u)) << 4);
          *reinterpret_cast<uint4*>(blk + off) = make_uint4(__byte_perm(hp[0], hp[1], sel), __byte_perm(hp[2], hp[3], sel),
                                                            __byte_perm(hp[4], hp[5], sel), __byte_perm(hp[6], hp[7], sel));
          *reinterpret_cast<uint4*>(blk + half_bytes + off) = make_uint4(__byte_perm(lp[0], lp[1], sel), __byte_perm(lp[2], lp[3], sel),
                                                                         __byte_perm(lp[4], lp[5], sel), __byte_perm(lp[6], lp[7], sel));
        }
      }
    };

    // fwd runs one extra iteration: the drain of the last sample's last group
    for (int i = 0; i < nmine + (kBwd ? 0 : 1); ++i) {
      const int b = cta + i * ncta;
      if (i < nmine) {
      const int buf = i % p.nbuf;
      if (warp == 4) EP_TRACE(i, 8);                               // epilogue: sample i begins
      // bwd: lane l keeps the row statistics of queries l and 32 + l of this sample
      float st_mx[2] = {0.f, 0.f}, st_inv[2] = {0.f, 0.f}, st_dl[2] = {0.f, 0.f};
      float cache[kCache][16];
      // bwd: A = exp(S - max) / sum of one unit from the saved logits (straight-line, loads issued together)
      auto load_probs = [&](int t, int j0, float (&a)[16]) {
        const int n = t * 128 + wq * 32 + lane;
        const bool valid = n < p.N;
        const float* srow = p.S + ((size_t)b * p.M + j0) * p.N + min(n, p.N - 1);
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = __ldcs(srow + (size_t)min(q, p.M - 1 - j0) * p.N);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int m = min(j0 + q, p.M - 1);                      // warp-uniform
          const float mx = __shfl_sync(0xffffffffu, m >= 32 ? st_mx[1] : st_mx[0], m & 31);
          const float inv = __shfl_sync(0xffffffffu, m >= 32 ? st_inv[1] : st_inv[0], m & 31);
          a[q] = valid ? ex2_ftz((a[q] - mx) * kLog2e) * inv : 0.f;
        }
      };
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
        if (!noepi) {
          // the probabilities of this warp's first kCache units, recomputed from the saved logits while the dA MMAs
          // of the sample still run (the loads do not depend on them)
#pragma unroll
          for (int k = 0; k < kCache; ++k) {
            const int u = es + kSub * k;
            const int t = u / upt, j0 = (u - t * upt) << 4;
            if (u < nunits && t * 128 + wq * 32 < p.N) load_probs(t, j0, cache[k]);
          }
        }
      }
      mbar_wait(tfull_bar(buf), ((uint32_t)(i / p.nbuf)) & 1u);
      tc_fence_after();
      if (warp == 4) EP_TRACE(i, 9);                               // epilogue: logits complete
      const uint32_t acc = tmem_base + lane_base + (uint32_t)(buf * p.bufcols);
      // pair mode: publish this CTA's partial logits (its half of D), then wait for the partner's; from here on
      // fetch_unit() returns the sum.  Scratch slot i & 1: the partner read slot i & 1 of sample i - 2 before it
      // published sample i - 1, which this CTA has already consumed.
      const float* peer = nullptr;
      if (pair && !noepi) {
        const size_t slot_floats = (size_t)p.ntiles * p.Mp * 128;
        float* mine = p.xchg + ((size_t)blockIdx.x * 2 + (i & 1)) * slot_floats;
        peer = p.xchg + ((size_t)(blockIdx.x ^ 1) * 2 + (i & 1)) * slot_floats;
        for (int u = es; u < nunits; u += kSub) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float v[16];
          load_unit(acc, t, j0, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) __stcg(mine + ((size_t)t * p.Mp + j0 + q) * 128 + wq * 32 + lane, v[q]);
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        if (ew == 0 && lane == 0) {
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.flags + blockIdx.x), "r"(i + 1) : "memory");
          int seen;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.flags + (blockIdx.x ^ 1)) : "memory");
          } while (seen < i + 1);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
      }
      auto fetch_unit = [&](int t, int j0, float (&v)[16]) {
        load_unit(acc, t, j0, v);
        if (pair && peer) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] += __ldcg(peer + ((size_t)t * p.Mp + j0 + q) * 128 + wq * 32 + lane);
        }
      };
      // the operand blocks may be rewritten once every pooled MMA of sample i - 1 has completed
      auto wait_blocks_free = [&]() {
        if (i > 0) {
          const int gl = kBwd ? 2 * (i - 1) : i * G - 1;           // its last group
          mbar_wait(pdone_bar(gl & 1), ((uint32_t)(gl >> 1)) & 1u);
          tc_fence_after();
          if (warp == 4) EP_TRACE(i, 10);                          // epilogue: pooled MMAs of sample i - 1 complete
        }
      };

      if (noepi) {
        wait_blocks_free();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        if (ew == 0 && lane == 0) *e_done = i + 1;
        if (!kBwd) {
          if (i > 0) { const int gg = i * G - 1; mbar_wait(pdone_bar(gg & 1), ((uint32_t)(gg >> 1)) & 1u); __syncwarp(); if (lane == 0) mbar_arrive(pfree_bar(gg & 1)); }
          for (int g = 0; g + 1 < G; ++g) { const int gg = i * G + g; mbar_wait(pdone_bar(gg & 1), ((uint32_t)(gg >> 1)) & 1u); __syncwarp(); if (lane == 0) mbar_arrive(pfree_bar(gg & 1)); }
        }
      } else if (!kBwd) {
        // ---- pass 1: per-query maximum over the tokens (per-warp partial rows, combined in a fixed order).  TMEM reads
        // run at 64 B / clock per SM, so the first kCache units of a warp stay in registers for the later passes
        if (has_lo) pmax[ew * SW + lane] = -INFINITY;
        if (has_hi) pmax[ew * SW + 32 + lane] = -INFINITY;
        __syncwarp();
        auto unit_max = [&](int t, int j0, const float (&x)[16]) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = (t * 128 + wq * 32 + lane < p.N) ? x[q] : -INFINITY;
          const float red = reduce16<true>(v, lane);
          if ((lane & 1) == 0) {
            float* slot = pmax + ew * SW + j0 + ridx;
            *slot = fmaxf(*slot, red);
          }
          __syncwarp();
        };
#pragma unroll
        for (int k = 0; k < kCache; ++k) {
          const int u = es + kSub * k;
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (u < nunits && t * 128 + wq * 32 < p.N) {             // (warp-uniform) some lane of this warp holds a token
            fetch_unit(t, j0, cache[k]);
            unit_max(t, j0, cache[k]);
          }
        }
        for (int u = es + kSub * kCache; u < nunits; u += kSub) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float v[16];
          fetch_unit(t, j0, v);
          unit_max(t, j0, v);
        }
        if (warp == 4) EP_TRACE(i, 4);                             // epilogue: pass 1 done
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        float mx_lo = -INFINITY, mx_hi = -INFINITY;                // lane l: queries l and 32 + l
#pragma unroll
        for (int w = 0; w < kEW; ++w) {
          if (has_lo) mx_lo = fmaxf(mx_lo, pmax[w * SW + lane]);
          if (has_hi) mx_hi = fmaxf(mx_hi, pmax[w * SW + 32 + lane]);
        }
        if (has_lo) psum[ew * SW + lane] = 0.f;
        if (has_hi) psum[ew * SW + 32 + lane] = 0.f;
        __syncwarp();
        wait_blocks_free();
        if (warp == 4) EP_TRACE(i, 5);                             // (dev) pass 2 begins
        // ---- pass 2: exp, row sums and the operand blocks of the pooled phase: what the P warp waits for.  Straight-line
        // code, the 16 queries of a unit are independent chains
        auto unit_exp = [&](int t, int j0, const float (&x)[16]) {
          const int n0 = t * 128 + wq * 32;                        // warp-uniform; its 32 tokens lie in one 64-token block
          const int n = n0 + lane;
          const bool valid = n < p.N;
          float e[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int m = j0 + q;                                  // warp-uniform
            const float mxs = __shfl_sync(0xffffffffu, m < 32 ? mx_lo : mx_hi, m & 31) * kLog2e;
            const float ex = ex2_ftz(fmaf(x[q], kLog2e, -mxs));
            e[q] = (valid && m < p.M) ? ex : 0.f;
          }
          if (!(kDev && (p.skip & 1))) store_rows(n0, n, j0, e);
          if (!(kDev && (p.skip & 2))) {
            const float red = reduce16<false>(e, lane);
            if ((lane & 1) == 0) psum[ew * SW + j0 + ridx] += red;
          } else if (lane == 0) psum[ew * SW + j0] += e[0];
          __syncwarp();
        };
#pragma unroll
        for (int k = 0; k < kCache; ++k) {
          const int u = es + kSub * k;
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (u < nunits && t * 128 + wq * 32 < p.N) unit_exp(t, j0, cache[k]);
        }
        for (int u = es + kSub * kCache; u < nunits; u += kSub) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float v[16];
          fetch_unit(t, j0, v);
          unit_exp(t, j0, v);
        }
        fence_proxy_async();                                       // generic-proxy block writes -> visible to the MMAs
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar);
        if (warp == 4) EP_TRACE(i, 12);                            // epilogue: operand blocks written
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        // row sums of this sample (for its drains) and the saved statistics
        if (ew == 0) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = 32 * h + lane;
            float su = 0.f;
            if (m < SW)
              for (int w = 0; w < kEW; ++w) su += psum[w * SW + m];
            tot[(i & 1) * 64 + m] = su;
            if (m < p.M && half == 0) {
              p.rmax[(size_t)b * p.M + m] = h ? mx_hi : mx_lo;
              p.rsum[(size_t)b * p.M + m] = su;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        // ---- pass 3, off the critical path (the pooled MMAs of the sample are running): the saved logits.  Their
        // stores are what the load/store unit is slowest at (rows of N floats are not 16-byte aligned in general), so
        // the upper half of each quadrant's warps writes them, re-reading TMEM, while the lower half starts draining;
        // the logit accumulators of the sample are released when they are through
        if (es >= kSub / 2) {
          if (half == 0) {
            for (int u = es - kSub / 2; u < nunits; u += kSub / 2) {
              const int t = u / upt, j0 = (u - t * upt) << 4;
              if (t * 128 + wq * 32 >= p.N) continue;
              float v[16];
              fetch_unit(t, j0, v);
              const int n = t * 128 + wq * 32 + lane;
              if (n < p.N) {
                float* srow = p.S + ((size_t)b * p.M + j0) * p.N + n;
#pragma unroll
                for (int q = 0; q < 16; ++q)
                  if (j0 + q < p.M) __stcs(srow + (size_t)q * p.N, v[q]);
              }
            }
          }
          tc_fence_before();
          asm volatile("bar.sync 2, %0;" ::"n"(32 * kEW / 2) : "memory");
          if (ew == kEW / 2 && lane == 0) *e_done = i + 1;         // the logit accumulators of sample i are free
        }
      } else {
        // ---- backward: dS = A (dA - delta), A recomputed from the saved logits and row statistics
        wait_blocks_free();
        // dS of one unit -> operand blocks (straight-line: independent chains per query)
        auto emit_ds = [&](int t, int j0, const float (&a)[16], const float (&v)[16]) {
          const int n0 = t * 128 + wq * 32;                        // warp-uniform; its 32 tokens lie in one 64-token block
          const int n = n0 + lane;
          float ds[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int m = min(j0 + q, p.M - 1);                    // warp-uniform
            const float dl = __shfl_sync(0xffffffffu, m >= 32 ? st_dl[1] : st_dl[0], m & 31);
            ds[q] = n < p.N ? a[q] * (v[q] - dl) : 0.f;
          }
          store_rows(n0, n, j0, ds);
        };
#pragma unroll
        for (int k = 0; k < kCache; ++k) {                         // units whose A is already in registers
          const int u = es + kSub * k;
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (u < nunits && t * 128 + wq * 32 < p.N) {
            float v[16];
            fetch_unit(t, j0, v);
            emit_ds(t, j0, cache[k], v);
          }
        }
        for (int u = es + kSub * kCache; u < nunits; u += kSub) {
          const int t = u / upt, j0 = (u - t * upt) << 4;
          if (t * 128 + wq * 32 >= p.N) continue;
          float a[16];
          load_probs(t, j0, a);
          float v[16];
          fetch_unit(t, j0, v);
          emit_ds(t, j0, a, v);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(eready_bar);
        if (warp == 4) EP_TRACE(i, 12);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEW) : "memory");
        if (ew == 0 && lane == 0) *e_done = i + 1;                 // every warp has read the dA accumulators of sample i
      }
      }  // i < nmine
      if (!kBwd && !noepi) {
        // drains: the last group of the previous sample, then this sample's groups but the last (which completes
        // together with the next sample's logits)
        const int g_hi = i < nmine ? G - 1 : 0;
        for (int dg = i > 0 ? -1 : 0; dg < g_hi; ++dg)
          drain_group(dg < 0 ? b - ncta : b, dg < 0 ? i - 1 : i, dg < 0 ? G - 1 : dg);
        if (warp == 4 && i < nmine) EP_TRACE(i, 13);               // epilogue: drains done
      }
    }
    if (nmine > 0 && kBwd) {
      // bwd: the query-gradient partial of this CTA, all slices
      const int gl = 2 * (nmine - 1);
      mbar_wait(pdone_bar(0), ((uint32_t)(gl >> 1)) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + (uint32_t)p.pcol0;
      for (int u = es; u < nslh * upt; u += kSub) {
        const int sl = u / upt, j0 = (u - sl * upt) << 4;
        const int d = (sl_lo + nslh - 1 - sl) * 128 + wq * 32 + lane;   // accumulator sl holds the sl-th slice from the top
        uint32_t r[16];
        tmem_ld16(acc + (uint32_t)(sl * p.Mp + j0), r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (j0 + q < p.M) p.out[((size_t)cta * p.M + j0 + q) * p.D + d] = __uint_as_float(r[q]);
      }
    } else if (kBwd) {
      // a CTA without samples still owns a partial-gradient slice: zeros
      if (half == 0)
        for (size_t o = (size_t)(warp - 4) * 32 + lane; o < (size_t)p.M * p.D; o += 32 * kEW)
          p.out[(size_t)cta * p.M * p.D + o] = 0.f;
    }
  }
  // dependents may be scheduled from here on, NOT from the start: a dependent grid's CTAs would sit on every SM (or slot)
  // this persistent kernel leaves free for the whole launch and keep kernels of other streams -- the overlapped gradient
  // all-reduce -- from starting (measured: 2 GPUs 0.631 -> 0.659 ms/step with the trigger at the top)
  tc_fence_before();
  __syncthreads();
  pdl_trigger();        // (after the CTA-wide barrier: the idle lanes of the producer warps reach this point at once)
  if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int round16(int v) { return (v + 15) / 16 * 16; }
int pow2_cols(int c) { int v = 32; while (v < c) v <<= 1; return v; }

struct FPlan {
  bool ok = false;
  int Mp, ntiles, nfull, tail_rows, nchunks, nkb, nsl, nL, nP, lbytes, nkp, pf, lead, nbuf, bufcols, pcol0, tmem_cols, qoff, toff, nslg, G, pbufcols;
  int lsplit, lcolw, kl, last_rows, pair, nslA, nstream;
  size_t smem;
};

// L2 budget for the tokens alive between their two fetches (one sample plus the lead of the next per CTA or CTA pair):
// measured, about half of the 126 MB is usable for this pattern (see "Pair mode" above); it only sizes the lead
constexpr size_t kL2Budget = 58ull << 20;

FPlan make_fplan(int N, int D, int M, int ctas, bool bwd) {
  FPlan pl;
  if (D % 128 != 0 || M < 1 || M > 64 || N < 1) return pl;
  pl.Mp = round16(M);
  pl.ntiles = (N + 127) / 128;
  const int rem = N - (pl.ntiles - 1) * 128;                      // rows of the last tile
  const int w_bytes = 2 * pl.Mp * 128;
  // a last tile of at most 64 tokens is loaded as a short box in front of the query chunk
  pl.tail_rows = rem <= 64 ? (rem + 7) / 8 * 8 : 0;
  pl.nfull = pl.tail_rows ? pl.ntiles - 1 : pl.ntiles;
  pl.qoff = pl.tail_rows * 128;
  pl.toff = pl.qoff + w_bytes;
  pl.nchunks = D / 64;
  pl.nkb = (N + 63) / 64;
  pl.nsl = D / 128;
  // one CTA per sample, or a pair splitting D when that many whole samples would not stay in L2
  const size_t sample = (size_t)N * D * 2;
  // (pair mode is measured slower than one CTA per sample -- both CTAs run the whole softmax -- and stays behind
  //  the developer knob: ep_set_debug bit 29)
  const int fp = (g_debug >> 28) & 3;
  pl.pair = fp == 2 && pl.nsl >= 2 && ctas >= 2;
  pl.nslA = pl.pair ? (pl.nsl + 1) / 2 : pl.nsl;                  // d-slices of a CTA (the larger half)
  const int live = pl.pair ? ctas / 2 : ctas;
  const int nslc = pl.nslA, nchc = 2 * pl.nslA;
  pl.nkp = (pl.nkb + 1) / 2;                                      // tall bricks (128 tokens) per d-slice
  pl.kl = (N - (pl.nkp - 1) * 128 + 15) / 16;                     // 16-token k-steps of the last tall brick
  pl.last_rows = pl.kl * 16;
  // TMEM plan.  Logit accumulators per token tile: 2 Mp columns (hi and lo query rows side by side: one MMA per
  // k-step, half the shared-memory operand reads) or Mp (two K-steps into one column); one or two buffers.  Pooled
  // accumulators: backward all slices resident (nsl x Mp, hi/lo as K-steps), forward two ping-pong buffers of nslg
  // slices x 2 Mp columns.  Preference: two logit buffers (the next sample's first chunks overlap the softmax), then
  // the fewest pooled groups, then side-by-side columns.
  const int force = (g_debug >> 20) & 3;                          // dev knob: 1 = one logit buffer, 2 = K-split logits
  pl.nbuf = 0;
  for (int cand = 0; cand < 4; ++cand) {
    const int nbuf = cand < 2 ? 2 : 1, lsplit = cand & 1;
    if ((force == 1 && nbuf == 2) || (force == 2 && !lsplit)) continue;
    if (pl.nbuf > nbuf) break;                                    // a plan with two logit buffers exists
    const int lcolw = lsplit ? pl.Mp : 2 * pl.Mp;
    const int left = 512 - nbuf * pl.ntiles * lcolw;
    int nslg, G, pbufcols;
    if (bwd) {
      if (left < nslc * pl.Mp) continue;
      nslg = nslc; G = 1; pbufcols = nslc * pl.Mp;
    } else {
      if (left < 2 * 2 * pl.Mp) continue;
      nslg = std::min(nslc, left / (2 * 2 * pl.Mp));
      G = (nslc + nslg - 1) / nslg;
      nslg = (nslc + G - 1) / G;                                  // balance the groups
      pbufcols = nslg * 2 * pl.Mp;
    }
    // forward: K-split logits are preferred when they leave room for larger pooled groups (fewer drains and half the
    // TMEM bytes read by the softmax: c2 / c3 forward 213 -> 210 us)
    if (pl.nbuf == nbuf && G >= pl.G) continue;
    pl.nbuf = nbuf; pl.lsplit = lsplit; pl.lcolw = lcolw; pl.nslg = nslg; pl.G = G; pl.pbufcols = pbufcols;
  }
  if (!pl.nbuf) return pl;
  pl.bufcols = pl.ntiles * pl.lcolw;
  pl.pcol0 = pl.nbuf * pl.bufcols;
  pl.tmem_cols = pow2_cols(pl.pcol0 + (bwd ? 1 : 2) * pl.pbufcols);
  const size_t fixed = 1024 /*alignment*/ + (size_t)2 * pl.Mp * 128 * pl.nkb + (size_t)stat_floats(pl.Mp) * 4 + 1024 /*barriers*/;
  const size_t avail = 227 * 1024;
  // rings: two or three tall bricks, the rest chunk stages (at least two)
  pl.lbytes = pl.toff + pl.nfull * kSlotBytes;
  pl.nP = 2;
  if (fixed + (size_t)pl.nP * kPBytes + 2 * (size_t)pl.lbytes > avail) return pl;
  pl.nL = (int)std::min<size_t>(4, (avail - fixed - (size_t)pl.nP * kPBytes) / pl.lbytes);
  if (fixed + (size_t)(pl.nP + 1) * kPBytes + (size_t)pl.nL * pl.lbytes <= avail) pl.nP = 3;
  pl.smem = fixed + (size_t)pl.nP * kPBytes + (size_t)pl.nL * pl.lbytes;
  pl.pf = ((g_debug >> 22) & 7) ? ((g_debug >> 22) & 7) - 1 : 0;  // dev knob: bits 22-24 = L2 prefetch distance + 1
  // chunks of the next sample fetched before the first brick: enough to cover the epilogue, bounded by the L2 budget
  pl.lead = std::min(((g_debug >> 16) & 15) ? ((g_debug >> 16) & 15) - 1 : 4, nchc / 2);   // dev knob: bits 16-19 = lead + 1
  if (pl.nbuf == 1) pl.lead = 0;                                  // a lead needs the second logit buffer
  const size_t budget = kL2Budget + ((size_t)((g_debug >> 25) & 7) * 10 << 20);   // dev knob: bits 25-27 = +10 MB each
  while (pl.lead > 1 && sample * live * (nchc + pl.lead) / nchc > budget) --pl.lead;
  // the first quarter of a sample's chunks is re-read last (the pooled phase walks D downwards) and is what L2 loses
  // when the live set does not fit: fetched evict_first, it leaves the room to the rest (c2: backward 192 -> 188 us,
  // c3: 211 -> 203 us; forward within 1 %).  dev knob: bits 0-3 = chunks + 1
  pl.nstream = (g_debug & 15) ? std::min(nchc, (g_debug & 15) - 1) : nchc / 4;
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

// tokens seen as (64 d, N, D / 64, B): a box of (64, rows, 2 * nb, 1) is nb bricks [rows tokens x 128 d], each as two
// 128-byte-swizzled halves of 64 d
int make_brick_tmap(CUtensorMap* m, const void* x, int B, int N, int D, int rows, int nb) {
  const uint64_t dims[4] = {64, (uint64_t)N, (uint64_t)(D / 64), (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)D * 2, 128, (uint64_t)D * N * 2};
  const uint32_t box[4] = {64, (uint32_t)rows, (uint32_t)(2 * nb), 1};
  return make_tmap_bf16(m, x, 4, dims, strides, box);
}

template <bool kBwd>
int launch_fused(const void* x, const void* w, int w_batched, int J, int B, int N, int D, int M, const FPlan& pl, FParams p,
                 int grid, cudaStream_t s) {
  CUtensorMap tm_x, tm_x1, tm_xt, tm_b, tm_bl, tm_w;
  int rc;
  if ((rc = make_x_tmap(&tm_x, x, B, N, D, pl.nfull >= 2 ? 256 : 128))) return rc;
  if ((rc = make_x_tmap(&tm_x1, x, B, N, D, 128))) return rc;
  if ((rc = make_x_tmap(&tm_xt, x, B, N, D, pl.tail_rows ? pl.tail_rows : 128))) return rc;
  if ((rc = make_brick_tmap(&tm_b, x, B, N, D, 128, 1))) return rc;
  if ((rc = make_brick_tmap(&tm_bl, x, B, N, D, pl.last_rows, 1))) return rc;
  if ((rc = make_w_tmap(&tm_w, w, D, J, w_batched ? B : 1, pl.Mp))) return rc;
  p.B = B; p.N = N; p.D = D; p.M = M; p.Mp = pl.Mp;
  p.ntiles = pl.ntiles; p.nfull = pl.nfull; p.tail_rows = pl.tail_rows;
  p.nchunks = pl.nchunks; p.nkb = pl.nkb; p.nsl = pl.nsl; p.nL = pl.nL; p.nP = pl.nP; p.lbytes = pl.lbytes; p.nkp = pl.nkp; p.pf = pl.pf;
  p.lead = pl.lead; p.nbuf = pl.nbuf; p.nstream = pl.nstream;
  p.nslg = pl.nslg; p.G = pl.G; p.pbufcols = pl.pbufcols; p.lsplit = pl.lsplit; p.lcolw = pl.lcolw; p.kl = pl.kl;
  p.last_rows = pl.last_rows; p.pair = pl.pair; p.nslA = pl.nslA;
  p.trace = (g_debug & 2048) ? 1 : 0; p.nomma = (g_debug & 4096) ? 1 : 0; p.noepi = (g_debug & 8192) ? 1 : 0; p.skip = (g_debug >> 14) & 3;
  p.bufcols = pl.bufcols; p.pcol0 = pl.pcol0; p.tmem_cols = pl.tmem_cols; p.w_batched = w_batched; p.qoff = pl.qoff; p.toff = pl.toff;
  if (p.trace || p.nomma || p.noepi || p.pair || p.skip) {                  // developer knobs: the instrumented instantiation
    if ((rc = set_smem(fused_kernel<kBwd, true>, pl.smem))) return rc;
    EP_CUDA(launch_pdl(fused_kernel<kBwd, true>, dim3(grid), dim3(kThreadsF), pl.smem, s, tm_x, tm_x1, tm_xt, tm_b, tm_bl, tm_w, p));
  } else {
    if ((rc = set_smem(fused_kernel<kBwd, false>, pl.smem))) return rc;
    EP_CUDA(launch_pdl(fused_kernel<kBwd, false>, dim3(grid), dim3(kThreadsF), pl.smem, s, tm_x, tm_x1, tm_xt, tm_b, tm_bl, tm_w, p));
  }
  EP_LAUNCH_CHECK();
  return 0;
}

}  // namespace fused
using namespace fused;

int fused_trace_fetch(long long* host_out, int n) {
  if (n > 128) n = 128;
  EP_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * sizeof(long long)));
  return 0;
}

bool fused_supported(int N, int D, int M) {
  return make_fplan(N, D, M, stream_sms(), false).ok && make_fplan(N, D, M, stream_sms(), true).ok;
}

// pair mode scratch: partial-logit exchange slots + flags (zero when the shape runs one CTA per sample)
size_t fused_workspace_bytes(int N, int D, int M) {
  const FPlan pl = make_fplan(N, D, M, kNumSMs, false);
  if (!pl.ok) return 0;
  return align_up((size_t)kNumSMs * 2 * pl.ntiles * pl.Mp * 128 * sizeof(float), 256) + align_up(kNumSMs * sizeof(int), 256);
}

namespace {
// grid and pair-mode scratch of one launch
int setup_launch(const FPlan& pl, int B, int N, int D, int M, void* xws, FParams* p, int* grid, cudaStream_t s) {
  const int sms = stream_sms();
  *grid = pl.pair ? 2 * std::min(B, sms / 2) : std::min(B, sms);
  if (pl.pair) {
    if (!xws) return EP_ERR_WORKSPACE;
    const size_t xb = align_up((size_t)kNumSMs * 2 * pl.ntiles * pl.Mp * 128 * sizeof(float), 256);
    p->xchg = (float*)xws;
    p->flags = (int*)((char*)xws + xb);
    EP_CUDA(cudaMemsetAsync(p->flags, 0, kNumSMs * sizeof(int), s));
  }
  return 0;
}
}  // namespace

// qhl: (J, D) bf16 hi/lo rows of the scaled queries; xws: fused_workspace_bytes() of scratch
int fused_pool_fwd(const void* x, const void* qhl, int J, int B, int N, int D, int M, float* P, float* S, float* rowmax,
                   float* rowsum, int round_p, void* xws, cudaStream_t s) {
  const FPlan pl = make_fplan(N, D, M, stream_sms(), false);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  FParams p{};
  p.S = S; p.rmax = rowmax; p.rsum = rowsum; p.out = P; p.round_out = round_p;
  int grid, rc;
  if ((rc = setup_launch(pl, B, N, D, M, xws, &p, &grid, s))) return rc;
  return launch_fused<false>(x, qhl, 0, J, B, N, D, M, pl, p, grid, s);
}

// dphl: (B, J, D) bf16 hi/lo rows of dP; part: [groups][M][D] partial query gradients (*groups_out = CTAs or CTA pairs)
int fused_pool_bwd(const void* x, const void* dphl, int J, int B, int N, int D, int M, const float* S, const float* rowmax,
                   const float* rowsum, const float* delta, float* part, int* groups_out, void* xws, cudaStream_t s) {
  const FPlan pl = make_fplan(N, D, M, stream_sms(), true);
  if (!pl.ok) return EP_ERR_UNSUPPORTED;
  FParams p{};
  p.S = const_cast<float*>(S); p.rmax = const_cast<float*>(rowmax); p.rsum = const_cast<float*>(rowsum);
  p.delta = delta; p.out = part;
  int grid, rc;
  if ((rc = setup_launch(pl, B, N, D, M, xws, &p, &grid, s))) return rc;
  *groups_out = pl.pair ? grid / 2 : grid;
  return launch_fused<true>(x, dphl, 1, J, B, N, D, M, pl, p, grid, s);
}

}  // namespace ep
