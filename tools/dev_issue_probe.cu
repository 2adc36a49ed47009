// Developer probe (not part of the library): issue cost, in SM cycles, of the pipeline primitives a warp-specialised
// tcgen05 kernel spends its producer / consumer threads on -- TMA tensor loads (one elected thread), mbarrier
// try_wait on a completed phase, tcgen05.commit, tcgen05.fence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/dev_issue_probe.cu -o tools/dev_issue_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "../efficient-probing_b200/csrc/ep_ptx.cuh"
using namespace ep::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) issue_probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm4,
                                                      long long* out, int B) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[24];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) { for (int i = 0; i < 24; ++i) mbar_init(bar0 + 8u * i, i < 8 ? 8 : 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x % B;
  long long r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (warp == 1 && lane == 0) {
    // (0) 64 TMA loads of 16 KB (3-D box 64 x 128) into 8 slots, each with expect_tx, no waiting in between
    const uint64_t pol = policy_evict_first();
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i) {
      const int s = i & 7;
      mbar_arrive_expect_tx(bar0 + 8u * s, 16384u);
      tma_load_3d_hint(base + (uint32_t)s * 16384u, &tm, bar0 + 8u * s, (i & 15) * 64, 0, b, pol);
    }
    r[0] = clock64() - t0;
    for (int s = 0; s < 8; ++s) mbar_wait(bar0 + 8u * s, 0u);   // 8 arrivals + 8 x 16 KB per barrier = one phase
    // (1) 64 TMA loads of 16 KB as one 4-D box (64 x 64 x 2): a [64 tokens x 128 d] brick in one instruction
    t0 = clock64();
    for (int i = 0; i < 64; ++i) {
      const int s = i & 7;
      mbar_arrive_expect_tx(bar0 + 8u * s, 16384u);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
          ::"r"(base + (uint32_t)s * 16384u), "l"(reinterpret_cast<uint64_t>(&tm4)), "r"(bar0 + 8u * s), "r"(0), "r"(0), "r"((i & 7) * 2), "r"(b), "l"(pol)
          : "memory");
    }
    r[1] = clock64() - t0;
    for (int s = 0; s < 8; ++s) mbar_wait(bar0 + 8u * s, 1u);
    // (2) try_wait on a completed phase, 64 times (phases 0 and 1 are complete: parity 1 passes)
    t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_wait(bar0 + 8u * (i & 7), 1u);
    r[2] = clock64() - t0;
    // (3) tcgen05.commit with nothing pending onto barriers 8..15, 64 times
    t0 = clock64();
    for (int i = 0; i < 64; ++i) umma_commit(bar0 + 64u + 8u * (i & 7));
    r[3] = clock64() - t0;
    // (4) tcgen05.fence::after_thread_sync, 64 times
    t0 = clock64();
    for (int i = 0; i < 64; ++i) tc_fence_after();
    r[4] = clock64() - t0;
    // (5) mbarrier.arrive (plain), 64 times
    t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_arrive(bar0 + 64u + 8u * (i & 7));
    r[5] = clock64() - t0;
    if (blockIdx.x == 0) for (int i = 0; i < 8; ++i) out[i] = r[i];
  }
  if (warp == 2) {
    // (6) whole warp waits on barriers 16..19 (never armed: parity 1 = the phase before the first passes), 64 times
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_wait(bar0 + 128u + 8u * (i & 3), 1u);
    long long d = clock64() - t0;
    if (blockIdx.x == 0 && lane == 0) out[6] = d;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_slot, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 257, D = 1024, B = 148;
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  void* buf; CK(cudaMalloc(&buf, (size_t)B * N * D * 2)); CK(cudaMemset(buf, 0, (size_t)B * N * D * 2));
  CUtensorMap tm, tm4;
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)D * N * 2};
    cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc3 failed\n"); return 1; }
  }
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)N, (cuuint64_t)(D / 64), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)D * 2, 128, (cuuint64_t)D * N * 2};
    cuuint32_t box[4] = {64, 64, 2, 1}, es[4] = {1, 1, 1, 1};
    if (enc(&tm4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("enc4 failed\n"); return 1; }
  }
  long long* out; CK(cudaMalloc(&out, 64)); CK(cudaMemset(out, 0, 64));
  const size_t smem = 8 * 16384 + 2048;
  CK(cudaFuncSetAttribute(issue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int grid : {1, 148}) {
    for (int rep = 0; rep < 2; ++rep) { issue_probe<<<grid, 128, smem>>>(tm, tm4, out, B); CK(cudaDeviceSynchronize()); }
    long long h[8]; CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
    printf("grid=%3d  cycles per op:  TMA 3-D 16 KB %.0f | TMA 4-D brick 16 KB %.0f | try_wait(done) %.0f | tcgen05.commit %.0f | tcgen05.fence %.0f | "
           "mbarrier.arrive %.0f | warp-wide try_wait(done) %.0f\n", grid, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0);
  }
  return 0;
}
