// Developer probe (not part of the library): what do two interleaved tcgen05.mma streams cost when they share one
// tensor core, and how much does concurrent TMA traffic into shared memory slow them?
//   stream L: K-major A [128 tokens x 16 d], N = nL     (the logit phase of the fused kernels)
//   stream P: MN-major A [128 d x 16 tokens], N = nP    (the pooled phase)
//   TMA:      32 KB boxes streamed from global memory into a 2-slot ring by a third thread
// Modes (bit mask): 1 = L warp, 2 = P warp, 4 = TMA stream, 8 = one warp alternates L and P MMAs (instead of 1|2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/dev_mix_probe.cu -o tools/dev_mix_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "../efficient-probing_b200/csrc/ep_ptx.cuh"
using namespace ep::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P { int mode, nL, nP, count, nload, rows; long long* out; int cevery; };

__global__ void __launch_bounds__(128, 1) mix_probe(const __grid_constant__ CUtensorMap tm, const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_store[8];
  __shared__ uint32_t tmem_slot;
  auto bar = [&](int i) { return smem_u32(&bar_store[i]); };   // 0: L done, 1: P done, 2-3: full, 4-5: empty
  for (uint32_t o = threadIdx.x * 16u; o < 144u * 1024u; o += blockDim.x * 16u)
    *reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)) + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar(i), 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t bsm = base + 128u * 1024u;                    // B operand: 16 KB
  const uint32_t ring = base + 144u * 1024u;                   // TMA ring: 2 x 32 KB
  const uint32_t idL = idesc_bf16(128, p.nL, 0, 0), idP = idesc_bf16(128, p.nP, 1, 0);
  const uint64_t aL0 = smem_desc_sw128(base, 16, 1024), aP0 = smem_desc_sw128(base, 16384, 1024);
  const uint64_t bd0 = smem_desc_sw128(bsm, 16, 1024);
  long long t0 = clock64(), t1 = 0;
  if (warp == 1 && (p.mode & 1)) {
    const bool leader = elect_one();
    for (int it = 0; it < p.count; it += 4) {
      const uint64_t ad = aL0 + (uint64_t)(((it >> 2) & 7) * (16384 >> 4));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (leader) umma_f16(tmem + (uint32_t)(((it >> 2) & 1) * p.nL), ad + 2u * u, bd0 + 2ull * u, idL, 1u);
      if (p.cevery && ((it + 4) % p.cevery) == 0 && leader) umma_commit(bar(6));
      __syncwarp();
    }
    if (leader) umma_commit(bar(0));
    __syncwarp();
    mbar_wait(bar(0), 0);
    t1 = clock64();
    if (blockIdx.x == 0 && leader) p.out[0] = t1 - t0;
  } else if (warp == 2 && (p.mode & 2)) {
    const bool leader = elect_one();
    for (int it = 0; it < p.count; it += 4) {
      const uint64_t ad = aP0 + (uint64_t)(((it >> 3) & 3) * (32768 >> 4)) + (uint64_t)(((it >> 2) & 1) * 4) * 128u;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (leader) umma_f16(tmem + 256u + (uint32_t)(((it >> 2) & 1) * p.nP), ad + 128u * u, bd0 + 2ull * u, idP, 1u);
      if (p.cevery && ((it + 4) % p.cevery) == 0 && leader) umma_commit(bar(7));
      __syncwarp();
    }
    if (leader) umma_commit(bar(1));
    __syncwarp();
    mbar_wait(bar(1), 0);
    t1 = clock64();
    if (blockIdx.x == 0 && leader) p.out[1] = t1 - t0;
  } else if (warp == 1 && (p.mode & 8)) {
    const bool leader = elect_one();
    for (int it = 0; it < p.count; it += 4) {
      const uint64_t adl = aL0 + (uint64_t)(((it >> 2) & 7) * (16384 >> 4));
      const uint64_t adp = aP0 + (uint64_t)(((it >> 3) & 3) * (32768 >> 4)) + (uint64_t)(((it >> 2) & 1) * 4) * 128u;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (leader) {
          umma_f16(tmem + (uint32_t)(((it >> 2) & 1) * p.nL), adl + 2u * u, bd0 + 2ull * u, idL, 1u);
          umma_f16(tmem + 256u + (uint32_t)(((it >> 2) & 1) * p.nP), adp + 128u * u, bd0 + 2ull * u, idP, 1u);
        }
      __syncwarp();
    }
    if (leader) umma_commit(bar(0));
    __syncwarp();
    mbar_wait(bar(0), 0);
    t1 = clock64();
    if (blockIdx.x == 0 && leader) p.out[0] = t1 - t0;
  } else if (warp == 0 && threadIdx.x == 0 && (p.mode & 4)) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.nload; ++i) {
      mbar_wait(bar(4 + s), ph ^ 1u);
      mbar_arrive_expect_tx(bar(2 + s), 32768u);
      tma_load_3d(ring + (uint32_t)s * 32768u, &tm, bar(2 + s), 0, (i * 256) % p.rows, blockIdx.x);
      if (++s == 2) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 3 && (threadIdx.x & 31) == 0 && (p.mode & 4)) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < p.nload; ++i) {
      mbar_wait(bar(2 + s), ph);
      mbar_arrive(bar(4 + s));
      if (++s == 2) { s = 0; ph ^= 1u; }
    }
    if (blockIdx.x == 0) p.out[2] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int rows = 32768;                                       // per CTA: 32768 rows x 128 B = 4 MB; 148 CTAs: 620 MB
  void* buf; CK(cudaMalloc(&buf, (size_t)148 * rows * 128)); CK(cudaMemset(buf, 0, (size_t)148 * rows * 128));
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, 148};
  cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
  cuuint32_t box[3] = {64, 256, 1}, estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n"); return 1;
  }
  long long* out; CK(cudaMalloc(&out, 32));
  const size_t smem = 209 * 1024 + 1024;
  CK(cudaFuncSetAttribute(mix_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int count = 4800;
  printf("%5s %4s %4s | %12s %12s %14s\n", "mode", "nL", "nP", "L cyc/MMA", "P cyc/MMA", "TMA B/cyc");
  for (int cevery : {0, 12, 8, 4})
  for (int nn : {64})
    for (int mode : {1, 2, 3, 7}) {
      const int nload = ((mode & 3) == 3 || (mode & 8)) ? 640 : 320;   // 10 / 20 MB per CTA
      P p{mode, nn, nn, count, nload, rows, out, cevery};
      if (mode == 1) printf("commit every %d MMAs\n", cevery);
      CK(cudaMemset(out, 0, 32));
      for (int rep = 0; rep < 2; ++rep) { mix_probe<<<148, 128, smem>>>(tm, p); CK(cudaDeviceSynchronize()); }
      long long h[4]; CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
      const int nmma = (mode & 8) ? 2 * count : count;
      printf("%5d %4d %4d | %12.1f %12.1f %14.1f\n", mode, nn, nn, h[0] ? (double)h[0] / nmma : 0.0, h[1] ? (double)h[1] / count : 0.0,
             h[2] ? (double)nload * 32768 / (double)h[2] : 0.0);
    }
  return 0;
}
