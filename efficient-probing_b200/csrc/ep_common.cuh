// Shared helpers for the libep_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>
#include "../../include/ep_b200.h"

// every kernel launch in the library is followed by this: error check + launch accounting
#define EP_LAUNCH_CHECK()                                  \
  do {                                                     \
    ++ep::g_launch_count;                                  \
    cudaError_t e__ = cudaGetLastError();                  \
    if (e__ != cudaSuccess) return (int)e__;               \
  } while (0)

#define EP_CUDA(call)                                      \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return (int)e__;               \
  } while (0)

namespace ep {

extern unsigned long long g_launch_count;     // kernels launched by this library in this process
constexpr int kNumSMs = 148;           // B200; upper bound used for workspace sizing
extern int g_sm_limit;                 // ep_set_sm_limit: CTAs the persistent token-streaming kernels may launch (0 = all SMs)
// SMs of the current device (queried once), never more than the workspace bound
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v < kNumSMs ? v : kNumSMs;
    else
      n = kNumSMs;
  }
  return n;
}
inline int stream_sms() { const int n = num_sms(); return g_sm_limit > 0 && g_sm_limit < n ? g_sm_limit : n; }

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// The kernels of a training step form one dependent chain on one stream; launched with the programmatic-stream-
// serialization attribute a kernel may be scheduled while its predecessor is still draining, run its prologue (barrier
// init, TMEM allocation, tensor-map prefetch, shared-memory clears) and then block in pdl_wait() until the predecessor
// grid has completed and its writes are visible.  Every kernel launched through launch_pdl() calls pdl_wait() before
// its first global-memory access.  pdl_trigger() (dependents may be scheduled once every CTA of this grid has called
// it) sits at the start of the short kernels and at the END of the persistent ones (one-pass pooling kernels, GEMMs):
// triggered early, a dependent's CTAs would occupy whatever SMs a long kernel leaves free and starve other streams.
// Without the attribute both instructions are no-ops.  EP_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa): operands of the TF32 tensor-core GEMMs are stored
// pre-rounded so that the hardware's truncation of the low bits is exact (no bias)
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 consecutive token channels as floats, from bf16 (one 128-bit load) or fp32 (two).
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ float2 load2(const __nv_bfloat16* p) {
  uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float2 load2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

extern int g_debug;
// developer aid (ep_set_debug bit 5): CUDA-event time of every kernel of one call, printed to stderr
struct TimingRecord { char name[24]; float us; };
extern TimingRecord g_timings[512];
extern int g_ntimings;
struct StageTimer {
  bool on; cudaStream_t s; cudaEvent_t ev[12]; const char* name[12]; int n = 0;
  StageTimer(cudaStream_t st) : on((g_debug & 32) != 0), s(st) { if (on) mark("start"); }
  void mark(const char* nm) {
    if (!on || n >= 12) return;
    cudaEventCreate(&ev[n]); cudaEventRecord(ev[n], s); name[n++] = nm;
  }
  ~StageTimer() {
    if (!on) return;
    cudaStreamSynchronize(s);
    for (int i = 1; i < n; ++i) {
      float ms = 0.f; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      if (g_debug & 256) fprintf(stderr, "[ep timing] %-18s %8.1f us\n", name[i], ms * 1e3f);
      if (g_ntimings < 512) {
        snprintf(g_timings[g_ntimings].name, sizeof(g_timings[0].name), "%s", name[i]);
        g_timings[g_ntimings++].us = ms * 1e3f;
      }
    }
    for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
  }
};

// ---- internal launchers shared between translation units (all return ep_status / cudaError) ----
struct GemmDesc {      // C[z][i][j] = sum_k A[z][i][k] * B[z][k][j] (+ bias[z][j]); element strides
  const float* A; const float* B; float* C; const float* bias;
  int I, J, K, Z;
  long long a_i, a_k, a_z, b_k, b_j, b_z, c_i, c_j, c_z, bias_z;
};
int launch_gemm_v0(const GemmDesc& g, cudaStream_t s);
int launch_colsum(const float* a, int rows, int cols, float* out, cudaStream_t s);            // out[j] = sum_i a[i][j]
int launch_delta_from_out(const float* g, const float* out, const float* bias, long long rows, int M, int c,
                          float* delta, cudaStream_t s);

// TF32 tensor-core GEMM (ep_gemm_sm100.cu)
struct GemmTC;
bool gemm_tc_available();
// C[z][i][j] = sum_k A.. B..; see ep_api.cu for the operand descriptions
enum TcOperand { TC_KMAJOR = 0, TC_MNMAJOR = 1 };
struct TcSide {            // one operand as a 3-D fp32 tensor (d0 contiguous) and how tiles index it
  const void* base;
  unsigned long long d0, d1, d2, s1, s2;   // sizes and element strides of dims 1, 2
  int mn_major;            // 0: d0 is the contraction index; 1: d0 is the output (row/col) index
  int swap;                // 0: dim1 = rows-or-k, dim2 = batch; 1: dim1 = batch, dim2 = rows-or-k
  int zdiv;                // batch coordinate = z / zdiv
  int bf16;                // elements are bf16 (both operands of a GEMM must agree); 0 = fp32 read as tf32
  // hi/lo operand pair (bf16, 3-term product, see GemmTC::x3): the lo tile is the hi tile's coordinates with
  // batch = z * zmul + lo_z (A side only) and k + lo_k; all zero = plain operand
  int zmul, lo_z, lo_k;
  int lo_mn;               // MN-major operand with the lo copy in the same row: lo tile = output coordinate + lo_mn
};
int tc_gemm(const TcSide& A, const TcSide& B, int I, int J, int K, int Z, int NT, float* C, long long c_row,
            long long c_col, long long c_z, const float* bias, long long bias_z, int round_out, cudaStream_t s);

int launch_gemm_tn(const float* A, const float* B, float* C, int I, int J, int K, int Z, long long lda, long long ldb,
                   long long ldc, long long a_z, long long b_z, long long c_z, cudaStream_t s);
bool gemm_tn_ok(int I, int J, long long lda, long long ldb, long long ldc, long long a_z, long long b_z, long long c_z);
// 3-term split copies for the tcgen05 GEMMs; bf16 != 0: dst is bf16 (hi/lo split), else fp32 (tf32 big/small)
int launch_split3(const float* src, void* dst, long long R, int K, int Kp, long long ld, int kind, int bf16, cudaStream_t s);
int launch_split3_transpose(const float* src, void* dst, int K, int Kp, int R, int Z, long long src_z, long long dst_z,
                            int kind, int bf16, cudaStream_t s);
int launch_gemm_nt3(const float* A, const float* B, float* C, const float* bias, int I, int J, int K, int Z,
                    long long lda, long long ldb, long long ldc, long long a_z, long long b_z, long long c_z,
                    long long bias_z, cudaStream_t s);

// projection backward with the fused epilogue: dP written as bf16 hi/lo operand rows (B, Jrows, D)
int tc_gemm_dp(const TcSide& A, const TcSide& B, int I, int J, int K, int Z, int NT, void* hl, int Jrows, cudaStream_t s);

// operand copies written by their producers (ep_ops.cu)
enum { REFRESH_ROWS = 0, REFRESH_TRANSPOSE = 1, REFRESH_HILO = 2 };
constexpr int kMaxRefreshJobs = 6;
struct RefreshJob {
  int type;
  const float* src; void* dst;
  int K, Kp, R, Z;           // ROWS: src [R x K] -> dst [R x 3Kp];  TRANSPOSE: src[z] [K x R] -> dst[z] [R x 3Kp]
  long long ld, src_z, dst_z;
  int kind, bf16;            // kind 0: [big|small|big], 1: [big|big|small];  bf16 0: tf32 big/small stored as fp32
  float scale; int M, J, D;  // HILO: dst [J x D] bf16 rows (2m: hi, 2m + 1: lo) of scale * src [M x D]
};
int launch_refresh(const RefreshJob* jobs, int n, cudaStream_t s);
// delta[b, m] = g[b, m] . (out[b, m] - bias[m]);  g3 (nullable): [(b, m)][3c] A-side copy (bf16, or fp32 tf32 pairs);
// g3t (nullable): [M c][3B] bf16 B-side copy of g^T
int launch_g_ops(const float* g, const float* out, const float* bias, int B, int M, int c, float* delta, void* g3,
                 int g3_bf16, void* g3t, cudaStream_t s);

// head kernels with the optional operand copies of their outputs (ep_epilogue.cu)
int launch_bn_fwd(const float* h, int B, int F, float eps, float momentum, int training, float* running_mean,
                  float* running_var, long long* nbt, float* y, float* save_mean, float* save_invstd, void* y3, int Fp,
                  cudaStream_t s);
int launch_bn_bwd(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh, cudaStream_t s);
int launch_ce(const float* logits, const long long* targets, int B, int K, float loss_scale, float grad_scale,
              float* loss_sum, float* dlogits, int* correct, void* d3, int Kp, float* scratch, float* loss_acc,
              cudaStream_t s);

int pool_fwd_v0(const void* x, int x_dtype, const float* cls, float scale, int B, int N, int D, int M,
                float* P, float* S_out, float* rowmax, float* rowsum, float* attn, int round_p, cudaStream_t s,
                int cls_batched = 0);
int pool_bwd_v0(const void* x, int x_dtype, const float* cls, float scale, int B, int N, int D, int M,
                const float* rowmax, const float* rowsum, const float* dP, const float* delta,
                float* dq_slots, int n_slots, float* d_cls, cudaStream_t s, int cls_batched = 0, void* dx = nullptr);
size_t pool_v0_smem_bytes(int N, int M);
constexpr int kDqSlots = 16;

}  // namespace ep
