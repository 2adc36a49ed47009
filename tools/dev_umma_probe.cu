// Developer probe (not part of the library): issue cost of tcgen05.mma kind::f16 (bf16, M = 128, K = 16, cta_group::1)
// as a function of N, of accumulator dependence (same TMEM columns back to back vs round-robin over several) and of
// the A operand's major-ness / how many distinct A tiles are touched.  Operands are whatever shared memory holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/dev_umma_probe.cu -o tools/dev_umma_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "../efficient-probing_b200/csrc/ep_ptx.cuh"
using namespace ep::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct P { int N, nacc, a_mn, count, a_tiles, same_a; long long* out; int warp_issue; };

__global__ void __launch_bounds__(128, 1) umma_probe(const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_store[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar = smem_u32(&bar_store[0]);
  // zero the operand area (8 A tiles of 16 KB + B of 32 KB)
  for (uint32_t o = threadIdx.x * 16u; o < 160u * 1024u; o += blockDim.x * 16u)
    *reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)) + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (p.warp_issue && (threadIdx.x >> 5) == 1) {
    // whole warp walks the loop (uniform control flow and operands), one elected lane issues
    const uint32_t idesc = idesc_bf16(128, p.N, p.a_mn, 0);
    const uint32_t bsm = base + 128u * 1024u;
    const uint64_t ad0 = p.a_mn ? smem_desc_sw128(base, 8192, 1024) : smem_desc_sw128(base, 16, 1024);
    const uint64_t bd0 = smem_desc_sw128(bsm, 16, 1024);
    const uint64_t astep = p.a_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t accstep = (uint32_t)p.N;
    const uint32_t m1 = p.nacc > 1 ? 1u : 0u, m2 = p.nacc > 2 ? 2u : 0u;
    const bool leader = elect_one();
    long long t0 = clock64();
    for (int it = 0; it < p.count; it += 8) {
      const uint64_t ad = ad0 + (p.same_a ? 0ull : (uint64_t)(((it >> 3) & 7) * (16384 >> 4)));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t acc = tmem + (((u & 1) * m1) + ((u & 2) * (m2 >> 1))) * accstep;
        if (leader) umma_f16(acc, ad + astep * (u & 3), bd0 + 2ull * (u & 3), idesc, 1u);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (leader) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && leader) { p.out[0] = t1 - t0; p.out[1] = t2 - t0; }
  } else if (!p.warp_issue && threadIdx.x == 32) {
    const uint32_t idesc = idesc_bf16(128, p.N, p.a_mn, 0);
    const uint32_t bsm = base + 128u * 1024u;
    // descriptors precomputed; the loop body is 8 unrolled MMAs whose descriptors differ by immediates
    const uint64_t ad0 = p.a_mn ? smem_desc_sw128(base, 8192, 1024) : smem_desc_sw128(base, 16, 1024);
    const uint64_t bd0 = smem_desc_sw128(bsm, 16, 1024);
    const uint64_t astep = p.a_mn ? (2048u >> 4) : (32u >> 4);
    const uint32_t accstep = (uint32_t)p.N;
    const uint32_t m1 = p.nacc > 1 ? 1u : 0u, m2 = p.nacc > 2 ? 2u : 0u;
    long long t0 = clock64();
    for (int it = 0; it < p.count; it += 8) {
      const uint64_t ad = ad0 + (p.same_a ? 0ull : (uint64_t)(((it >> 3) & 7) * (16384 >> 4)));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t acc = tmem + (((u & 1) * m1) + ((u & 2) * (m2 >> 1))) * accstep;
        umma_f16(acc, ad + astep * (u & 3), bd0 + 2ull * (u & 3), idesc, 1u);
      }
    }
    long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { p.out[0] = t1 - t0; p.out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; CK(cudaMalloc(&out, 16));
  const size_t smem = 161 * 1024 + 1024;
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("%4s %5s %5s %7s %7s | %10s %10s\n", "N", "nacc", "A_mn", "a_tiles", "same_a", "issue cyc", "done cyc/MMA");
  const int count = 2048;
  for (int wi = 0; wi < 2; ++wi)
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int N : {16, 32, 64, 128})
      for (int nacc : {1, 4})
        for (int same_a : {0}) {
          if (nacc * N > 512) continue;
          P p{N, nacc, a_mn, count, 8, same_a, out, wi};
          if (nacc == 1 && N == 16 && a_mn == 0) printf("warp_issue=%d\n", wi);
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            umma_probe<<<148, 128, smem>>>(p);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
          printf("%4d %5d %5d %7d %7d | %10.1f %10.1f\n", N, nacc, a_mn, 8, same_a, (double)h[0] / count, (double)h[1] / count);
        }
  return 0;
}
