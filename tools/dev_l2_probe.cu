// Developer probe (not part of the library): how fast can 148 persistent CTAs stream c2-shaped samples through
// TMA when every sample is fetched TWICE by the same CTA (second fetch expected to hit L2)?  No tensor work: a
// consumer thread just releases the ring slots.  Decides whether a fused logits+pool kernel may re-read x from L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/dev_l2_probe.cu -o tools/dev_l2_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../efficient-probing_b200/csrc/ep_ptx.cuh"
using namespace ep::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t make_policy(int kind) {
  uint64_t p = 0;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

struct Params { int B, ntiles, nchunks, slots, passes, pol1, pol2, lag; };

// lag = 0: pass 2 of sample i directly after pass 1 of sample i.  lag = 1: order is pass1(i+1), pass2(i) -- what a
// software-pipelined fused kernel does.
__global__ void __launch_bounds__(64, 1) probe_kernel(const __grid_constant__ CUtensorMap tm, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + (uint32_t)p.slots * 16384u;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (p.slots + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = p.ntiles * p.nchunks;
  int nmine = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) ++nmine;
  // sequence of (sample ordinal, pass) this CTA walks
  auto walk = [&](auto&& fn) {
    if (p.passes == 1) { for (int i = 0; i < nmine; ++i) fn(i, 0); return; }
    if (p.lag == 0) { for (int i = 0; i < nmine; ++i) { fn(i, 0); fn(i, 1); } return; }
    fn(0, 0);
    for (int i = 0; i < nmine; ++i) { if (i + 1 < nmine) fn(i + 1, 0); fn(i, 1); }
  };
  if (warp == 0 && lane == 0) {
    const uint64_t pol[2] = {make_policy(p.pol1), make_policy(p.pol2)};
    int s = 0; uint32_t ph = 0;
    walk([&](int i, int pass) {
      const int b = blockIdx.x + i * gridDim.x;
      for (int u = 0; u < per; ++u) {
        const int c = u / p.ntiles, t = u - c * p.ntiles;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(full(s), 16384u);
        tma_load_3d_hint(base + (uint32_t)s * 16384u, &tm, full(s), c * 64, t * 128, b, pol[pass]);
        if (++s == p.slots) { s = 0; ph ^= 1u; }
      }
    });
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    walk([&](int, int) {
      for (int u = 0; u < per; ++u) {
        mbar_wait(full(s), ph);
        mbar_arrive(empty(s));
        if (++s == p.slots) { s = 0; ph ^= 1u; }
      }
    });
  }
}

// The fused forward's planned load order on one in-order ring of 16 KB slots:
//   L(0);  for i: La(i+1) [first `lead` d-chunks of the next sample], P(i) [second fetch of sample i as 64-token x
//   128-d bricks], Lb(i+1) [its remaining chunks].  L chunk c = query chunk (8 KB) + the sample's token tiles
//   [128 x 64 d] (the ragged last tile as an 8-row-granular box).
struct Params2 { int B, N, D, slots, lead, pol1, pol2, tail_rows; };
__global__ void __launch_bounds__(64, 1) fused_order_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_tail,
                                                            const __grid_constant__ CUtensorMap tm_brick, const __grid_constant__ CUtensorMap tm_q,
                                                            const Params2 p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + (uint32_t)p.slots * 16384u;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (p.slots + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (p.N + 127) / 128, nchunks = p.D / 64, nkb = (p.N + 63) / 64, nsl = p.D / 128;
  int nmine = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) ++nmine;
  // walk(fn): fn(kind, sample ordinal, a, b): kind 0 = q chunk a; 1 = x tile b of chunk a; 2 = brick (kb = a, slice = b)
  auto walk = [&](auto&& fn) {
    auto Lpart = [&](int i, int c0, int c1) {
      for (int c = c0; c < c1; ++c) { fn(0, i, c, 0); for (int t = 0; t < ntiles; ++t) fn(1, i, c, t); }
    };
    auto Ppart = [&](int i) { for (int kb = 0; kb < nkb; ++kb) for (int sl = 0; sl < nsl; ++sl) fn(2, i, kb, sl); };
    if (nmine == 0) return;
    Lpart(0, 0, nchunks);
    for (int i = 0; i < nmine; ++i) {
      if (i + 1 < nmine) Lpart(i + 1, 0, p.lead);
      Ppart(i);
      if (i + 1 < nmine) Lpart(i + 1, p.lead, nchunks);
    }
  };
  if (warp == 0 && lane == 0) {
    const uint64_t pol[2] = {make_policy(p.pol1), make_policy(p.pol2)};
    const uint64_t polq = make_policy(2);
    int s = 0; uint32_t ph = 0;
    walk([&](int kind, int i, int a, int bb) {
      const int b = blockIdx.x + i * gridDim.x;
      mbar_wait(empty(s), ph ^ 1u);
      const uint32_t dst = base + (uint32_t)s * 16384u;
      if (kind == 0) {
        mbar_arrive_expect_tx(full(s), 8192u);
        tma_load_3d_hint(dst, &tm_q, full(s), a * 64, 0, 0, polq);
      } else if (kind == 1) {
        const bool tail = p.tail_rows && bb == ntiles - 1;
        mbar_arrive_expect_tx(full(s), tail ? (uint32_t)p.tail_rows * 128u : 16384u);
        tma_load_3d_hint(dst, tail ? &tm_tail : &tm, full(s), a * 64, bb * 128, b, pol[0]);
      } else {
        mbar_arrive_expect_tx(full(s), 16384u);
        tma_load_3d_hint(dst, &tm_brick, full(s), bb * 128, a * 64, b, pol[1]);
        tma_load_3d_hint(dst + 8192u, &tm_brick, full(s), bb * 128 + 64, a * 64, b, pol[1]);
      }
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    });
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    walk([&](int, int, int, int) {
      mbar_wait(full(s), ph);
      mbar_arrive(empty(s));
      if (++s == p.slots) { s = 0; ph ^= 1u; }
    });
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 257, D = argc > 2 ? atoi(argv[2]) : 1024, B = 1024, NBUF = 3;
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const size_t bytes = (size_t)B * N * D * 2;
  std::vector<void*> bufs(NBUF);
  std::vector<CUtensorMap> maps(NBUF);
  for (int i = 0; i < NBUF; ++i) {
    CK(cudaMalloc(&bufs[i], bytes));
    CK(cudaMemset(bufs[i], i + 1, bytes));
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)D * N * 2};
    cuuint32_t box[3] = {64, 128, 1}, estr[3] = {1, 1, 1};
    CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, bufs[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  // extra maps for the fused-order probe
  const int rem = N - (N - 1) / 128 * 128;
  const int tail_rows = rem < 128 ? (rem + 7) / 8 * 8 : 0;
  void* qbuf; CK(cudaMalloc(&qbuf, (size_t)64 * D * 2)); CK(cudaMemset(qbuf, 0, (size_t)64 * D * 2));
  std::vector<CUtensorMap> maps_tail(NBUF), maps_brick(NBUF);
  CUtensorMap map_q;
  auto mk = [&](CUtensorMap* m, void* base, int d0, int d1, int d2, int rows) {
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * d1 * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)rows, 1}, estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  for (int i = 0; i < NBUF; ++i) { mk(&maps_tail[i], bufs[i], D, N, B, tail_rows ? tail_rows : 128); mk(&maps_brick[i], bufs[i], D, N, B, 64); }
  mk(&map_q, qbuf, D, 64, 1, 64);
  const int slots = 12;
  const size_t smem = (size_t)slots * 16384 + 2048;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, int Buse, int grid, int passes, int pol1, int pol2, int lag, bool rotate) {
    Params p{Buse, (N + 127) / 128, D / 64, slots, passes, pol1, pol2, lag};
    float best = 1e9f, sum = 0.f; const int reps = 8;
    for (int it = 0; it < reps + 2; ++it) {
      const int bi = rotate ? it % NBUF : 0;
      CK(cudaEventRecord(e0));
      probe_kernel<<<grid, 64, smem>>>(maps[bi], p);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    const double alg = (double)Buse * N * D * 2;
    printf("%-46s B=%4d grid=%3d passes=%d pol=%d/%d lag=%d : avg %7.1f us best %7.1f us  -> %6.0f GB/s algorithmic (x%d fetched)\n",
           name, Buse, grid, passes, pol1, pol2, lag, sum / reps * 1e3, best * 1e3, alg / (sum / reps * 1e-3) / 1e9, passes);
  };
  CK(cudaFuncSetAttribute(fused_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  auto run2 = [&](const char* name, int grid, int nslots, int lead, int pol1, int pol2, int tail) {
    Params2 p{B, N, D, nslots, lead, pol1, pol2, tail};
    float best = 1e9f, sum = 0.f; const int reps = 8;
    for (int it = 0; it < reps + 2; ++it) {
      const int bi = it % NBUF;
      CK(cudaEventRecord(e0));
      fused_order_kernel<<<grid, 64, smem>>>(maps[bi], maps_tail[bi], maps_brick[bi], map_q, p);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    const double alg = (double)B * N * D * 2;
    printf("fused order %-24s grid=%3d slots=%2d lead=%2d pol=%d/%d tail=%3d : avg %7.1f us best %7.1f us -> %6.0f GB/s algorithmic\n",
           name, grid, nslots, lead, pol1, pol2, tail, sum / reps * 1e3, best * 1e3, alg / (sum / reps * 1e-3) / 1e9);
  };
  printf("N=%d D=%d sample=%.0f KB batch=%.0f MB\n", N, D, N * D * 2 / 1024.0, bytes / 1e6);
  run("single pass from HBM, evict_first", B, 148, 1, 1, 1, 0, true);
  run("single pass from HBM, evict_normal", B, 148, 1, 0, 0, 0, true);
  for (int lag = 0; lag < 2; ++lag) {
    run("two passes, normal then evict_first", B, 148, 2, 0, 1, lag, true);
    run("two passes, evict_last then evict_first", B, 148, 2, 2, 1, lag, true);
    run("two passes, normal then normal", B, 148, 2, 0, 0, lag, true);
    run("two passes, evict_first both", B, 148, 2, 1, 1, lag, true);
  }
  run("two passes on 74 CTAs (footprint 39 MB)", B, 74, 2, 0, 1, 1, true);
  run("L2-resident single pass (B=64, 34 MB, no rotate)", 64, 64, 1, 0, 0, 0, false);
  run("L2-resident single pass (B=128, 67 MB, no rotate)", 128, 128, 1, 0, 0, 0, false);
  run("L2-resident single pass (B=148, 78 MB, no rotate)", 148, 148, 1, 0, 0, 0, false);
  run("L2-resident 4 samples/CTA (B=148 x2 passes)", 148, 148, 2, 0, 0, 0, false);
  for (int lead : {0, 1, 2, 3, 4, 6, 8}) run2("last/first", 148, 12, lead, 2, 1, tail_rows);
  for (int lead : {0, 2, 4}) run2("normal/first", 148, 12, lead, 0, 1, tail_rows);
  for (int lead : {0, 2, 4}) run2("first/first", 148, 12, lead, 1, 1, tail_rows);
  for (int lead : {0, 2}) run2("last/first full tail", 148, 12, lead, 2, 1, 0);
  for (int lead : {0, 2}) run2("last/first 8 slots", 148, 8, lead, 2, 1, tail_rows);
  for (int lead : {0, 2}) run2("last/first 6 slots", 148, 6, lead, 2, 1, tail_rows);
  for (int lead : {2}) run2("last/first 132 CTAs", 132, 12, lead, 2, 1, tail_rows);
  return 0;
}
