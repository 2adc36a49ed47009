// extern "C" entry points of libep_b200.so (declared in include/ep_b200.h): argument checking,
// workspace carving and kernel-family dispatch.  No allocation, no host synchronisation.
#include "ep_common.cuh"
#include "ep_sm100.cuh"

#include <algorithm>
#include <cstdlib>

using namespace ep;

unsigned long long ep::g_launch_count = 0;
namespace ep { extern int g_debug; }
static int g_gemm_mode = 0;                   // 0 = tensor cores (3-term bf16 / TF32 products), 1 = fp32 CUDA cores
static int g_kernel_mode = 0;                 // 0 auto, 1 general, 2 tcgen05
static thread_local int t_last_family = 0;
static thread_local int t_fp32_gemm = 0;      // EP_OPS_FP32 of the *_ops call running on this thread
namespace {
struct OpsScope {                             // per-call, per-thread GEMM mode of the *_ops entry points
  int saved;
  explicit OpsScope(int ops) : saved(t_fp32_gemm) { if (ops & EP_OPS_FP32) t_fp32_gemm = 1; }
  ~OpsScope() { t_fp32_gemm = saved; }
};
}  // namespace

bool ep::pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("EP_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

extern "C" int ep_abi_version(void) { return EP_ABI_VERSION; }

extern "C" const char* ep_strerror(int code) {
  switch (code) {
    case EP_OK: return "ok";
    case EP_ERR_NULL: return "required pointer is NULL";
    case EP_ERR_SHAPE: return "invalid shape (sizes must be positive and D divisible by d_out*num_queries)";
    case EP_ERR_ALIGN: return "D must be a multiple of 8 and pointers 16-byte aligned";
    case EP_ERR_DTYPE: return "token dtype must be bf16 (0) or fp32 (1)";
    case EP_ERR_WORKSPACE: return "workspace smaller than ep_workspace_bytes()";
    case EP_ERR_UNSUPPORTED: return "shape not supported by the selected kernel family";
    case EP_ERR_DEVICE: return "current CUDA device is not compute capability 10.x (B200)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown ep_status";
  }
}

extern "C" int ep_device_check(void) {
  int dev = 0, major = 0;
  EP_CUDA(cudaGetDevice(&dev));
  EP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? 0 : EP_ERR_DEVICE;
}

extern "C" int ep_set_kernel_mode(int mode) {
  if (mode < 0 || mode > 2) return EP_ERR_SHAPE;
  g_kernel_mode = mode;
  return 0;
}
extern "C" int ep_last_kernel_family(void) { return t_last_family; }
extern "C" int ep_set_gemm_mode(int mode) {
  if (mode < 0 || mode > 1) return EP_ERR_SHAPE;
  g_gemm_mode = mode;
  return 0;
}
extern "C" int ep_set_debug(int flags) { ep::g_debug = flags; return 0; }
ep::TimingRecord ep::g_timings[512];
int ep::g_ntimings = 0;
extern "C" int ep_timing_count(void) { return ep::g_ntimings; }
extern "C" int ep_timing_get(int i, char* name, int name_len, float* us) {
  if (i < 0 || i >= ep::g_ntimings || !name || !us || name_len < 1) return EP_ERR_SHAPE;
  snprintf(name, (size_t)name_len, "%s", ep::g_timings[i].name);
  *us = ep::g_timings[i].us;
  return 0;
}
extern "C" int ep_timing_reset(void) { ep::g_ntimings = 0; return 0; }
extern "C" int ep_debug_trace(long long* host_out, int n) {
  if (!host_out || n < 1) return EP_ERR_NULL;
  return fused_trace_fetch(host_out, n);
}
extern "C" int ep_set_sm_limit(int n) {
  if (n < 0) return EP_ERR_SHAPE;
  ep::g_sm_limit = n;
  return 0;
}
extern "C" int ep_kernel_family_for(int x_dtype, int B, int N, int D, int M) {
  if (g_kernel_mode == 1) return 1;
  return sm100_supported(x_dtype, B, N, D, M) ? 2 : (g_kernel_mode == 2 ? 0 : 1);
}
extern "C" unsigned long long ep_launch_count(void) { return ep::g_launch_count; }

namespace {
struct Ws {                      // workspace layout
  size_t dP, delta, slots, sm100, w_t, w_p, g_r, g_t, total;
};
// The small GEMMs run on the tensor cores in TF32 unless the developer knob (bit 7) asks for the fp32
// CUDA-core GEMM; operands are rounded to tf32 (nearest) first so the hardware's truncation is exact.
// 3-term operand copies of the K-concatenated GEMMs: bf16 hi/lo (kind::f16, half the bytes, twice the MMA
// rate) rather than tf32 big/small; both give ~1e-5.  bf16 rows need K % 8 == 0 for 16-byte TMA strides.
constexpr int kSplitBf16 = 1;
bool use_tc() { return gemm_tc_available() && g_gemm_mode == 0 && !t_fp32_gemm; }
int round_nt(int c) { return std::min(256, (c + 31) / 32 * 32); }
// column tile of the classifier GEMMs: the narrowest that still gives every SM a tile
int lin_nt(int rows, int cols) {
  const int rt = (rows + 127) / 128;
  for (int nt = 128; nt >= 64; nt >>= 1)
    if (rt * ((cols + nt - 1) / nt) >= kNumSMs / 2 || nt == 64) return nt;
  return 64;
}
int dp_nt(int D) { return std::min(256, (D + 31) / 32 * 32); }     // column tile of the fused dP GEMM
Ws carve(int B, int N, int D, int M) {
  Ws w;
  size_t off = 0;
  w.w_t = off;   off += align_up((size_t)3 * D * D * sizeof(float), 256);   // 3xTF32 copy of v_w^T per query
  w.w_p = off;   off += align_up((size_t)3 * D * D * 2, 256);               // bf16 [hi|hi|lo] copy of v_w (projection)
  w.g_r = off;   off += align_up((size_t)3 * B * D * sizeof(float), 256);   // 3xTF32 copy of g_out
  w.g_t = off;   off += align_up((size_t)3 * B * D * 2, 256);               // bf16 [hi|hi|lo] copy of g_out^T
  w.dP = off;    off += align_up((size_t)B * M * D * sizeof(float), 256);
  w.delta = off; off += align_up((size_t)B * M * sizeof(float), 256);
  w.slots = off; off += align_up((size_t)kDqSlots * M * D * sizeof(float), 256);
  w.sm100 = off; off += align_up(sm100_workspace_bytes(B, N, D, M), 256);
  w.total = off;
  return w;
}
int check_common(const void* x, int x_dtype, const float* cls, int B, int N, int D, int M, int d_out) {
  if (!x || !cls) return EP_ERR_NULL;
  if (B <= 0 || N <= 0 || D <= 0 || M <= 0 || d_out <= 0) return EP_ERR_SHAPE;
  if (D % (d_out * M) != 0) return EP_ERR_SHAPE;
  if (D % 8 != 0) return EP_ERR_ALIGN;
  if (x_dtype != EP_DTYPE_BF16 && x_dtype != EP_DTYPE_F32) return EP_ERR_DTYPE;
  if (((uintptr_t)x & 15) || ((uintptr_t)cls & 15)) return EP_ERR_ALIGN;
  return 0;
}
bool use_sm100(int x_dtype, int B, int N, int D, int M, int* rc);
// P (the saved pooled tokens) is kept as bf16 hi/lo rows (b, m, {hi, lo}, d) -- byte-for-byte the size of the fp32
// tensor -- when the tcgen05 kernels produce and consume it: the projection and its weight gradient then read
// it in place as the 3-term bf16 operand ([hi|lo|hi] along the contraction) of tcgen05 GEMMs.
bool p_hilo(int x_dtype, int B, int N, int D, int M, int d_out) {
  int rc = 0;
  const int c = D / d_out / M;
  return use_tc() && kSplitBf16 && c % 4 == 0 && D % 64 == 0 && B % 64 == 0 && use_sm100(x_dtype, B, N, D, M, &rc);
}
bool use_sm100(int x_dtype, int B, int N, int D, int M, int* rc) {
  *rc = 0;
  if (g_kernel_mode == 1) return false;
  const bool ok = sm100_supported(x_dtype, B, N, D, M);
  if (g_kernel_mode == 2 && !ok) *rc = EP_ERR_UNSUPPORTED;
  return ok;
}
}  // namespace

extern "C" size_t ep_workspace_bytes(int B, int N, int D, int M, int d_out) {
  (void)d_out;
  if (B <= 0 || N <= 0 || D <= 0 || M <= 0) return 0;
  return carve(B, N, D, M).total;
}

extern "C" int ep_pooled_layout(int x_dtype, int B, int N, int D, int M, int d_out) {
  if (B <= 0 || N <= 0 || D <= 0 || M <= 0 || d_out <= 0 || D % d_out || (D / d_out) % M) return EP_ERR_SHAPE;
  return p_hilo(x_dtype, B, N, D, M, d_out) ? 1 : 0;
}

namespace {
// general = 1: the general kernel family with fp32 P whatever the mode (ep_fwd_ex)
int fwd_impl(const void* x, int x_dtype, const float* cls_token, int cls_batched, int general, const float* v_w,
             const float* v_b, float scale, int B, int N, int D, int M, int d_out, float* out, float* S, float* rowmax,
             float* rowsum, float* P, float* attn, void* workspace, size_t workspace_bytes, void* stream, int ops = 0) {
  OpsScope scope(ops);
  int rc = check_common(x, x_dtype, cls_token, B, N, D, M, d_out);
  if (rc) return rc;
  if (!v_w || !out || !S || !rowmax || !rowsum || !P) return EP_ERR_NULL;
  const Ws w = carve(B, N, D, M);
  if (w.total > 0 && (!workspace || workspace_bytes < w.total)) return EP_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const int round_p = !general && p_hilo(x_dtype, B, N, D, M, d_out) ? 1 : 0;   // P as bf16 hi/lo rows (see p_hilo)
  if (!general && use_sm100(x_dtype, B, N, D, M, &rc)) {
    t_last_family = 2;
    rc = sm100_pool_fwd(x, cls_token, scale, B, N, D, M, P, S, rowmax, rowsum, attn, round_p,
                        (char*)workspace + w.sm100, s, (ops & EP_OPS_WEIGHTS) ? 1 : 0);
  } else {
    if (rc) return rc;
    t_last_family = 1;
    rc = pool_fwd_v0(x, x_dtype, cls_token, scale, B, N, D, M, P, S, rowmax, rowsum, attn, round_p, s, cls_batched);
  }
  if (rc) return rc;
  // out[b, m*c + j] = v_w[m*c + j, :] . P[b, m, :] (+ v_b)     -- batched over the M queries
  const int Dp = D / d_out, c = Dp / M;
  if (round_p) {
    // tcgen05 3-term bf16 GEMM: A = P's hi/lo rows read in place, B = [W_hi | W_hi | W_lo] (a 6*D'*D-byte copy
    // whose first and last thirds are the hi/lo pair)
    void* w3f = (char*)workspace + w.w_p;
    StageTimer tm(s);
    if (!(ops & EP_OPS_WEIGHTS)) {
      if ((rc = launch_split3(v_w, w3f, Dp, D, D, D, 1, 1, s))) return rc;
      tm.mark("proj split3 W");
    }
    const unsigned long long D3 = 3ull * D;
    TcSide A{P, (unsigned long long)D, 2ull * M, (unsigned long long)B, (unsigned long long)D, 2ull * M * D, TC_KMAJOR, 1, 1,
             1, 2, 1, 0};
    TcSide Bm{w3f, D3, (unsigned long long)c, (unsigned long long)M, D3, D3 * c, TC_KMAJOR, 0, 1, 1, 0, 0, 2 * D};
    rc = tc_gemm(A, Bm, B, c, D, M, round_nt(c), out, Dp, 1, c, v_b, c, 0, s);
    tm.mark("proj tc-gemm");
    return rc;
  }
  if (use_tc() && D % 4 == 0)       // fp32 P: 3xTF32 in registers (mma.sync), P read once
    return launch_gemm_nt3(P, v_w, out, v_b, B, c, D, M, (long long)M * D, D, Dp, D, (long long)c * D, c, c, s);
  GemmDesc g{};
  g.A = P; g.B = v_w; g.C = out; g.bias = v_b;
  g.I = B; g.J = c; g.K = D; g.Z = M;
  g.a_i = (long long)M * D; g.a_k = 1; g.a_z = D;
  g.b_k = 1; g.b_j = D; g.b_z = (long long)c * D;
  g.c_i = Dp; g.c_j = 1; g.c_z = c; g.bias_z = c;
  return launch_gemm_v0(g, s);
}
}  // namespace

extern "C" int ep_fwd(const void* x, int x_dtype, const float* cls_token, const float* v_w, const float* v_b,
                      float scale, int B, int N, int D, int M, int d_out, float* out, float* S, float* rowmax,
                      float* rowsum, float* P, float* attn, void* workspace, size_t workspace_bytes, void* stream) {
  return fwd_impl(x, x_dtype, cls_token, 0, 0, v_w, v_b, scale, B, N, D, M, d_out, out, S, rowmax, rowsum, P, attn,
                  workspace, workspace_bytes, stream);
}

extern "C" int ep_fwd_ops(const void* x, int x_dtype, const float* cls_token, const float* v_w, const float* v_b,
                          float scale, int B, int N, int D, int M, int d_out, float* out, float* S, float* rowmax,
                          float* rowsum, float* P, float* attn, void* workspace, size_t workspace_bytes, int ops,
                          void* stream) {
  return fwd_impl(x, x_dtype, cls_token, 0, 0, v_w, v_b, scale, B, N, D, M, d_out, out, S, rowmax, rowsum, P, attn,
                  workspace, workspace_bytes, stream, ops);
}

extern "C" int ep_fwd_ex(const void* x, int x_dtype, const float* cls_token, int cls_batched, const float* v_w,
                         const float* v_b, float scale, int B, int N, int D, int M, int d_out, float* out, float* S,
                         float* rowmax, float* rowsum, float* P, float* attn, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return fwd_impl(x, x_dtype, cls_token, cls_batched ? 1 : 0, 1, v_w, v_b, scale, B, N, D, M, d_out, out, S, rowmax,
                  rowsum, P, attn, workspace, workspace_bytes, stream);
}

namespace {
int bwd_proj_impl(const float* g_out, const float* P, int p_layout, int general_dp, const float* out, const float* v_w,
                  const float* v_b, int x_dtype, int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                  void* workspace, size_t workspace_bytes, void* stream, int ops = 0);
}
extern "C" int ep_bwd_proj(const float* g_out, const float* P, const float* out, const float* v_w, const float* v_b,
                           int x_dtype, int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_proj_impl(g_out, P, -1, 0, out, v_w, v_b, x_dtype, B, N, D, M, d_out, d_v_w, d_v_b, workspace,
                       workspace_bytes, stream);
}
extern "C" int ep_bwd_proj_ops(const float* g_out, const float* P, const float* out, const float* v_w, const float* v_b,
                               int x_dtype, int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                               void* workspace, size_t workspace_bytes, int ops, void* stream) {
  return bwd_proj_impl(g_out, P, -1, 0, out, v_w, v_b, x_dtype, B, N, D, M, d_out, d_v_w, d_v_b, workspace,
                       workspace_bytes, stream, ops);
}

namespace {
// p_layout: -1 = what ep_fwd produces for this shape and mode, else 0 (fp32) / 1 (bf16 hi/lo rows);
// general_dp = 1: dP is written as fp32 (B, M, D) for the general pooling kernels whatever the family
int bwd_proj_impl(const float* g_out, const float* P, int p_layout, int general_dp, const float* out, const float* v_w,
                  const float* v_b, int x_dtype, int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                  void* workspace, size_t workspace_bytes, void* stream, int ops) {
  OpsScope scope(ops);
  if (!g_out || !P || !out || !v_w || !d_v_w) return EP_ERR_NULL;
  if (B <= 0 || N <= 0 || D <= 0 || M <= 0 || d_out <= 0 || D % (d_out * M) != 0) return EP_ERR_SHAPE;
  const Ws w = carve(B, N, D, M);
  if (!workspace || workspace_bytes < w.total) return EP_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  float* dP = (float*)((char*)workspace + w.dP);
  float* delta = (float*)((char*)workspace + w.delta);
  const int Dp = D / d_out, c = Dp / M;
  int rc;
  StageTimer tm(s);
  // delta[b, m] = sum_n A dA = dP[b, m] . P[b, m] = g[b, m, :] . (out[b, m, :] - bias[m, :])
  // (EP_OPS_INPUT: ep_bn_bwd_ops left delta and the copies of g_out in the workspace -- g_ops_kernel, ep_ops.cu)
  const bool g_ready = (ops & EP_OPS_INPUT) != 0, w_ready = (ops & EP_OPS_WEIGHTS) != 0;
  // the weight-gradient half and the dP half are independent: EP_OPS_ONLY_DW / EP_OPS_NO_DW let a caller run them
  // as two calls on two streams (134 MB of P read next to 134 MB of dP written)
  const bool do_dw = !(ops & EP_OPS_NO_DW), do_dp = !(ops & EP_OPS_ONLY_DW);
  if ((ops & EP_OPS_NO_DW) && (ops & EP_OPS_ONLY_DW)) return EP_ERR_SHAPE;
  if (!g_ready && do_dp && (rc = launch_delta_from_out(g_out, out, v_b, (long long)B * M, M, c, delta, s))) return rc;
  const bool hilo = p_layout < 0 ? p_hilo(x_dtype, B, N, D, M, d_out) : p_layout == 1;
  if (hilo && !(use_tc() && c % 4 == 0 && B % 64 == 0 && D % 64 == 0)) return EP_ERR_UNSUPPORTED;   // mode changed since ep_fwd
  if (use_tc() && c % 4 == 0) {
    // d_v_w[m*c + j, d] = sum_b g[b, m*c + j] * P[b, m, d]: contraction over the batch
    if (!do_dw) {
    } else if (hilo) {
      // tcgen05 3-term bf16 GEMM: rows d, cols j, contraction over b.  A = P's hi/lo rows in place, MN-major
      // (channels contiguous); B = g^T as [g_hi | g_hi | g_lo] (K-major copy, 6*B*D' bytes)
      void* g3t = (char*)workspace + w.g_t;
      if (!g_ready) {
        if ((rc = launch_split3_transpose(g_out, g3t, B, B, Dp, 1, 0, 0, 1, 1, s))) return rc;     // [Dp][3B]
        tm.mark("dW split3t g");
      }
      const unsigned long long B3 = 3ull * B;
      TcSide A{P, (unsigned long long)D, 2ull * M, (unsigned long long)B, (unsigned long long)D, 2ull * M * D, TC_MNMAJOR, 1,
               1, 1, 2, 1, 0};
      TcSide Bm{g3t, B3, (unsigned long long)c, (unsigned long long)M, B3, B3 * c, TC_KMAJOR, 0, 1, 1, 0, 0, 2 * B};
      if ((rc = tc_gemm(A, Bm, D, c, B, M, round_nt(c), d_v_w, 1, D, (long long)c * D, nullptr, 0, 0, s))) return rc;
      tm.mark("dW tc-gemm");
    } else {
      // fp32 P, both operands batch-major -> TF32 mma.sync "TN" kernel reading them as they lie
      if ((rc = launch_gemm_tn(g_out, P, d_v_w, c, D, B, M, Dp, (long long)M * D, D, c, D, (long long)c * D, s))) return rc;
      tm.mark("dW tn-gemm");
    }
    if (do_dw && d_v_b && (rc = launch_colsum(g_out, B, Dp, d_v_b, s))) return rc;
    if (do_dp) {  // dP[b, m, d] = sum_j g[b, m*c + j] * v_w[m*c + j, d]: tcgen05, 3xTF32 (dP drives the query gradient)
      float* g3 = (float*)((char*)workspace + w.g_r);              // [(b, m)][3c]  = [big | small | big]
      float* w3 = (float*)((char*)workspace + w.w_t);              // [m][d][3c]    = [big | big | small]
      const int bf = kSplitBf16 && c % 8 == 0;               // bf16 rows need 16-byte strides: 3c * 2 B
      if (!g_ready && (rc = launch_split3(g_out, g3, (long long)B * M, c, c, c, 0, bf, s))) return rc;
      if (!w_ready && (rc = launch_split3_transpose(v_w, w3, c, c, D, M, (long long)c * D, (long long)3 * c * D, 1, bf, s))) return rc;
      tm.mark("split3 g, W");
      const unsigned long long c3 = 3ull * c;
      TcSide A{g3, c3, (unsigned long long)M, (unsigned long long)B, c3, c3 * M, TC_KMAJOR, 1, 1, bf};
      TcSide Bm{w3, c3, (unsigned long long)D, (unsigned long long)M, c3, c3 * D, TC_KMAJOR, 0, 1, bf};
      int fam_rc = 0;
      if (!general_dp && use_sm100(x_dtype, B, N, D, M, &fam_rc)) {
        // tcgen05 pooling kernels follow: the GEMM epilogue emits what they consume -- dP as bf16 hi/lo operand
        // rows -- and fp32 dP is never written
        void* sm = (char*)workspace + w.sm100;
        const int J = sm100_J(N, D, M);
        void* hl = sm100_dphl_ptr(sm, B, N, D, M);
        if (J != 2 * M) EP_CUDA(cudaMemsetAsync(hl, 0, (size_t)B * J * D * 2, s));
        if ((rc = tc_gemm_dp(A, Bm, B, D, 3 * c, M, dp_nt(D), hl, J, s))) return rc;
        tm.mark("dP tc-gemm+hl");
        return 0;
      }
      if (fam_rc && !general_dp) return fam_rc;
      if ((rc = tc_gemm(A, Bm, B, D, 3 * c, M, round_nt(D), dP, (long long)M * D, 1, D, nullptr, 0, 0, s))) return rc;
      tm.mark("dP tc-gemm");
    }
    return 0;
  }
  if (do_dw) {  // d_v_w[m*c + j, d] = sum_b g[b, m*c + j] * P[b, m, d]
    GemmDesc g{};
    g.A = g_out; g.B = P; g.C = d_v_w;
    g.I = c; g.J = D; g.K = B; g.Z = M;
    g.a_i = 1; g.a_k = Dp; g.a_z = c;
    g.b_k = (long long)M * D; g.b_j = 1; g.b_z = D;
    g.c_i = D; g.c_j = 1; g.c_z = (long long)c * D;
    if ((rc = launch_gemm_v0(g, s))) return rc;
  }
  if (do_dw && d_v_b && (rc = launch_colsum(g_out, B, Dp, d_v_b, s))) return rc;
  if (do_dp) {  // dP[b, m, d] = sum_j g[b, m*c + j] * v_w[m*c + j, d]
    GemmDesc g{};
    g.A = g_out; g.B = v_w; g.C = dP;
    g.I = B; g.J = D; g.K = c; g.Z = M;
    g.a_i = Dp; g.a_k = 1; g.a_z = c;
    g.b_k = D; g.b_j = 1; g.b_z = (long long)c * D;
    g.c_i = (long long)M * D; g.c_j = 1; g.c_z = D;
    if ((rc = launch_gemm_v0(g, s))) return rc;
  }
  return 0;
}
}  // namespace

extern "C" int ep_bwd_ex(const void* x, int x_dtype, const float* cls_token, int cls_batched, const float* v_w,
                         float scale, int B, int N, int D, int M, int d_out, const float* S, const float* rowmax,
                         const float* rowsum, const float* P, int p_layout, const float* out, const float* v_b,
                         const float* g_out, float* d_cls_token, float* d_v_w, float* d_v_b, void* dx,
                         void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, x_dtype, cls_token, B, N, D, M, d_out);
  if (rc) return rc;
  if (!v_w || !S || !rowmax || !rowsum || !P || !out || !g_out || !d_cls_token || !d_v_w) return EP_ERR_NULL;
  if (p_layout != 0 && p_layout != 1) return EP_ERR_SHAPE;
  if (M > 64) return EP_ERR_UNSUPPORTED;
  if ((rc = bwd_proj_impl(g_out, P, p_layout, 1, out, v_w, v_b, x_dtype, B, N, D, M, d_out, d_v_w, d_v_b, workspace,
                          workspace_bytes, stream)))
    return rc;
  const Ws w = carve(B, N, D, M);
  t_last_family = 1;
  return pool_bwd_v0(x, x_dtype, cls_token, scale, B, N, D, M, rowmax, rowsum, (const float*)((char*)workspace + w.dP),
                     (const float*)((char*)workspace + w.delta), (float*)((char*)workspace + w.slots), kDqSlots,
                     d_cls_token, (cudaStream_t)stream, cls_batched ? 1 : 0, dx);
}

extern "C" int ep_bwd_pool(const void* x, int x_dtype, const float* cls_token, float scale, int B, int N, int D, int M,
                           int d_out, const float* S, const float* rowmax, const float* rowsum, float* d_cls_token,
                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, x_dtype, cls_token, B, N, D, M, d_out);
  if (rc) return rc;
  if (!S || !rowmax || !rowsum || !d_cls_token) return EP_ERR_NULL;
  const Ws w = carve(B, N, D, M);
  if (!workspace || workspace_bytes < w.total) return EP_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  const float* dP = (const float*)((char*)workspace + w.dP);
  const float* delta = (const float*)((char*)workspace + w.delta);
  float* slots = (float*)((char*)workspace + w.slots);
  if (use_sm100(x_dtype, B, N, D, M, &rc)) {
    t_last_family = 2;
    const int c = D / d_out / M;
    const bool fused = use_tc() && c % 4 == 0;                            // what ep_bwd_proj did (same predicate)
    return sm100_pool_bwd(x, S, scale, B, N, D, M, rowmax, rowsum, fused ? nullptr : dP, delta, 1, d_cls_token,
                          (char*)workspace + w.sm100, s);
  }
  if (rc) return rc;
  t_last_family = 1;
  return pool_bwd_v0(x, x_dtype, cls_token, scale, B, N, D, M, rowmax, rowsum, dP, delta, slots, kDqSlots,
                     d_cls_token, s);
}

extern "C" int ep_bwd(const void* x, int x_dtype, const float* cls_token, const float* v_w, float scale, int B, int N,
                      int D, int M, int d_out, const float* S, const float* rowmax, const float* rowsum,
                      const float* P, const float* out, const float* v_b, const float* g_out, float* d_cls_token,
                      float* d_v_w, float* d_v_b, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, x_dtype, cls_token, B, N, D, M, d_out);
  if (rc) return rc;
  if (!v_w || !S || !rowmax || !rowsum || !P || !out || !g_out || !d_cls_token || !d_v_w) return EP_ERR_NULL;
  if ((rc = ep_bwd_proj(g_out, P, out, v_w, v_b, x_dtype, B, N, D, M, d_out, d_v_w, d_v_b, workspace, workspace_bytes,
                        stream)))
    return rc;
  return ep_bwd_pool(x, x_dtype, cls_token, scale, B, N, D, M, d_out, S, rowmax, rowsum, d_cls_token, workspace,
                     workspace_bytes, stream);
}

extern "C" int ep_attention_maps(const void* x, int x_dtype, const float* cls_token, float scale, int B, int N, int D,
                                 int M, float* attn, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common(x, x_dtype, cls_token, B, N, D, M, 1);
  if (rc == EP_ERR_SHAPE && B > 0 && N > 0 && D > 0 && M > 0) rc = (D % 8) ? EP_ERR_ALIGN : 0;  // no channel split here
  if (rc) return rc;
  if (!attn) return EP_ERR_NULL;
  const size_t need = align_up((size_t)2 * B * M * sizeof(float), 256);
  if (!workspace || workspace_bytes < need) return EP_ERR_WORKSPACE;
  // tcgen05 route (logit kernel + row statistics writing the normalised map) when the shape is covered and the
  // caller's workspace is the full ep_workspace_bytes() one: the logits go to its dP region
  {
    int frc = 0;
    const Ws w = carve(B, N, D, M);
    if (use_sm100(x_dtype, B, N, D, M, &frc) && workspace_bytes >= w.total && (size_t)N <= (size_t)D &&
        (size_t)2 * B * M <= (size_t)kDqSlots * M * D) {
      float* S = (float*)((char*)workspace + w.dP);
      float* stats = (float*)((char*)workspace + w.slots);
      t_last_family = 2;
      return sm100_pool_fwd(x, cls_token, scale, B, N, D, M, nullptr, S, stats, stats + (size_t)B * M, attn, 0,
                            (char*)workspace + w.sm100, (cudaStream_t)stream);
    }
    if (frc) return frc;
  }
  float* rowmax = (float*)workspace;
  float* rowsum = rowmax + (size_t)B * M;
  t_last_family = 1;
  return pool_fwd_v0(x, x_dtype, cls_token, scale, B, N, D, M, nullptr, nullptr, rowmax, rowsum, attn, 0,
                     (cudaStream_t)stream);
}

static size_t pad64(int v) { return ((size_t)v + 63) / 64 * 64; }
extern "C" size_t ep_linear_workspace_bytes(int B, int F, int K) {
  if (B <= 0 || F <= 0 || K <= 0) return 0;
  const size_t Fp = pad64(F), Kp = pad64(K);                  // operand-copy thirds are padded to the GEMM's K step
  return align_up(3 * K * Fp * 4, 256) + align_up(3 * F * Kp * 4, 256) + align_up(3 * B * Fp * 4, 256) +
         align_up(3 * B * Kp * 4, 256);
}
namespace {
struct LinWs { float *w_r, *w_t, *y_r, *d_r; };
bool lin_ws(void* ws, size_t bytes, int B, int F, int K, LinWs* o) {
  if (!ws || bytes < ep_linear_workspace_bytes(B, F, K)) return false;
  char* p = (char*)ws;
  const size_t Fp = pad64(F), Kp = pad64(K);
  o->w_r = (float*)p; p += align_up(3 * K * Fp * 4, 256);
  o->w_t = (float*)p; p += align_up(3 * F * Kp * 4, 256);
  o->y_r = (float*)p; p += align_up(3 * B * Fp * 4, 256);
  o->d_r = (float*)p;
  return true;
}
}  // namespace

namespace {
// the bf16 copies of BOTH classifier GEMM directions exist (what the *_ops producers write and consumers trust)
bool lin_ops_ok(int F, int K) { return use_tc() && kSplitBf16 && F % 8 == 0 && K % 8 == 0; }

int linear_fwd_impl(const float* y, const float* W, const float* b, int B, int F, int K, float* logits, void* workspace,
                    size_t workspace_bytes, int ops, void* stream) {
  OpsScope scope(ops);
  if (!y || !W || !logits) return EP_ERR_NULL;
  if (B <= 0 || F <= 0 || K <= 0) return EP_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  LinWs lw;
  if (use_tc() && F % 4 == 0 && K % 4 == 0 && lin_ws(workspace, workspace_bytes, B, F, K, &lw)) {
    int rc;                                                     // 3xTF32: y' = [big|small|big], W' = [big|big|small]
    StageTimer tm(s);
    // bf16 hi/lo copies [hi|lo|hi] x [hi|hi|lo] with thirds padded to the K step: the GEMM loads the hi and lo
    // thirds of each operand once per stage and issues the three products itself (GemmTC::x3)
    const int bf = kSplitBf16 && F % 8 == 0;
    const int Fp = bf ? (int)pad64(F) : F;
    const bool ready = lin_ops_ok(F, K);                        // else the flags are ignored: self-made copies
    if (!(ready && (ops & EP_OPS_WEIGHTS)) && (rc = launch_split3(W, lw.w_r, K, F, Fp, F, 1, bf, s))) return rc;
    if (!(ready && (ops & EP_OPS_INPUT)) && (rc = launch_split3(y, lw.y_r, B, F, Fp, F, 0, bf, s))) return rc;
    tm.mark("lin split3 W,y");
    const unsigned long long F3 = 3ull * Fp;
    TcSide A{lw.y_r, F3, (unsigned long long)B, 1ull, F3, F3 * B, TC_KMAJOR, 0, 1, bf, 1, 0, bf ? Fp : 0};
    TcSide Bm{lw.w_r, F3, (unsigned long long)K, 1ull, F3, F3 * K, TC_KMAJOR, 0, 1, bf, 0, 0, bf ? 2 * Fp : 0};
    rc = tc_gemm(A, Bm, B, K, bf ? Fp : 3 * F, 1, lin_nt(B, K), logits, K, 1, 0, b, 0, 0, s);
    tm.mark("lin logits gemm");
    return rc;
  }
  GemmDesc g{};
  g.A = y; g.B = W; g.C = logits; g.bias = b;
  g.I = B; g.J = K; g.K = F; g.Z = 1;
  g.a_i = F; g.a_k = 1; g.b_k = 1; g.b_j = F; g.c_i = K; g.c_j = 1;
  return launch_gemm_v0(g, s);
}

int linear_bwd_impl(const float* dlogits, const float* y, const float* W, int B, int F, int K, float* dW, float* db,
                    float* dy, void* workspace, size_t workspace_bytes, int ops, void* stream) {
  OpsScope scope(ops);
  if (!dlogits) return EP_ERR_NULL;
  if (B <= 0 || F <= 0 || K <= 0) return EP_ERR_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  StageTimer tm(s);
  LinWs lw;
  const bool aligned = use_tc() && F % 4 == 0 && K % 4 == 0;
  const bool tc = aligned && lin_ws(workspace, workspace_bytes, B, F, K, &lw);   // dy needs the operand copies
  const bool ready = tc && lin_ops_ok(F, K);
  if (dW) {
    if (!y) return EP_ERR_NULL;
    if (ready && (ops & EP_OPS_INPUT) && !(g_debug & (1 << 30))) {
      // dW[k, f] = sum_b dlogits[b, k] * y[b, f] on tcgen05: both operands are batch-major, i.e. MN-major (output index
      // contiguous), and their [hi | lo | hi] copies already exist (written by ep_ce_fwd_bwd_ops / ep_bn_fwd_ops); the
      // lo tile of a stage is the hi tile's box shifted by one third along the row
      const int Fp = (int)pad64(F), Kp = (int)pad64(K);
      const unsigned long long K3 = 3ull * Kp, F3 = 3ull * Fp;
      TcSide A{lw.d_r, K3, (unsigned long long)B, 1ull, K3, K3 * B, TC_MNMAJOR, 0, 1, 1, 1, 0, 0, Kp};
      TcSide Bm{lw.y_r, F3, (unsigned long long)B, 1ull, F3, F3 * B, TC_MNMAJOR, 0, 1, 1, 0, 0, 0, Fp};
      if ((rc = tc_gemm(A, Bm, K, F, B, 1, 64, dW, F, 1, 0, nullptr, 0, 0, s))) return rc;
      tm.mark("lin dW tc-gemm");
    } else if (aligned) {          // batch-major operands as they lie: mma.sync TF32 "TN" kernel
      if ((rc = launch_gemm_tn(dlogits, y, dW, K, F, B, 1, K, F, F, 0, 0, 0, s))) return rc;
      tm.mark("lin dW tn-gemm");
    } else {
      GemmDesc g{};
      g.A = dlogits; g.B = y; g.C = dW;
      g.I = K; g.J = F; g.K = B; g.Z = 1;
      g.a_i = 1; g.a_k = K; g.b_k = F; g.b_j = 1; g.c_i = F; g.c_j = 1;
      if ((rc = launch_gemm_v0(g, s))) return rc;
    }
  }
  if (db && (rc = launch_colsum(dlogits, B, K, db, s))) return rc;
  if (dy) {
    if (!W) return EP_ERR_NULL;
    if (tc) {                      // dy[b, f] = sum_k dlogits[b, k] * W[k, f]: tcgen05, 3xTF32, transposed weight copy
      const int bf = kSplitBf16 && K % 8 == 0;
      const int Kp = bf ? (int)pad64(K) : K;
      if (!(ready && (ops & EP_OPS_INPUT)) && (rc = launch_split3(dlogits, lw.d_r, B, K, Kp, K, 0, bf, s))) return rc;   // [b][3Kp]
      if (!(ready && (ops & EP_OPS_WEIGHTS)) && (rc = launch_split3_transpose(W, lw.w_t, K, Kp, F, 1, 0, 0, 1, bf, s)))
        return rc;                                                                                                   // [f][3Kp]
      tm.mark("lin split3 dl,Wt");
      const unsigned long long K3 = 3ull * Kp;
      TcSide A{lw.d_r, K3, (unsigned long long)B, 1ull, K3, K3 * B, TC_KMAJOR, 0, 1, bf, 1, 0, bf ? Kp : 0};
      TcSide Bm{lw.w_t, K3, (unsigned long long)F, 1ull, K3, K3 * F, TC_KMAJOR, 0, 1, bf, 0, 0, bf ? 2 * Kp : 0};
      if ((rc = tc_gemm(A, Bm, B, F, bf ? Kp : 3 * K, 1, lin_nt(B, F), dy, F, 1, 0, nullptr, 0, 0, s))) return rc;
      tm.mark("lin dy gemm");
    } else {
      GemmDesc g{};
      g.A = dlogits; g.B = W; g.C = dy;
      g.I = B; g.J = F; g.K = K; g.Z = 1;
      g.a_i = K; g.a_k = 1; g.b_k = F; g.b_j = 1; g.c_i = F; g.c_j = 1;
      if ((rc = launch_gemm_v0(g, s))) return rc;
    }
  }
  return 0;
}
}  // namespace

extern "C" int ep_linear_fwd(const float* y, const float* W, const float* b, int B, int F, int K, float* logits,
                             void* workspace, size_t workspace_bytes, void* stream) {
  return linear_fwd_impl(y, W, b, B, F, K, logits, workspace, workspace_bytes, 0, stream);
}
extern "C" int ep_linear_fwd_ops(const float* y, const float* W, const float* b, int B, int F, int K, float* logits,
                                 void* workspace, size_t workspace_bytes, int ops, void* stream) {
  return linear_fwd_impl(y, W, b, B, F, K, logits, workspace, workspace_bytes, ops, stream);
}
extern "C" int ep_linear_bwd(const float* dlogits, const float* y, const float* W, int B, int F, int K, float* dW,
                             float* db, float* dy, void* workspace, size_t workspace_bytes, void* stream) {
  return linear_bwd_impl(dlogits, y, W, B, F, K, dW, db, dy, workspace, workspace_bytes, 0, stream);
}
extern "C" int ep_linear_bwd_ops(const float* dlogits, const float* y, const float* W, int B, int F, int K, float* dW,
                                 float* db, float* dy, void* workspace, size_t workspace_bytes, int ops, void* stream) {
  return linear_bwd_impl(dlogits, y, W, B, F, K, dW, db, dy, workspace, workspace_bytes, ops, stream);
}

// ---- producers of the operand copies (ABI 2) -------------------------------------------------------------------

extern "C" int ep_bn_fwd_ops(const float* h, int B, int F, float eps, float momentum, int training, float* running_mean,
                             float* running_var, long long* nbt, float* y, float* save_mean, float* save_invstd, int K,
                             void* lin_workspace, size_t lin_workspace_bytes, int ops, void* stream) {
  OpsScope scope(ops);
  if (!h || !y || !running_mean || !running_var) return EP_ERR_NULL;
  if (B <= 0 || F <= 0) return EP_ERR_SHAPE;
  LinWs lw;
  void* y3 = nullptr;
  if (lin_workspace && K > 0 && lin_ops_ok(F, K) && lin_ws(lin_workspace, lin_workspace_bytes, B, F, K, &lw)) y3 = lw.y_r;
  return launch_bn_fwd(h, B, F, eps, momentum, training, running_mean, running_var, nbt, y, save_mean, save_invstd, y3,
                       (int)pad64(F), (cudaStream_t)stream);
}

extern "C" int ep_ce_fwd_bwd_ops(const float* logits, const long long* targets, int B, int K, float loss_scale,
                                 float grad_scale, float* step_loss, float* loss_acc, float* dlogits, int* correct,
                                 float* scratch, int F, void* lin_workspace, size_t lin_workspace_bytes, int ops,
                                 void* stream) {
  OpsScope scope(ops);
  if (!logits || !targets || !scratch) return EP_ERR_NULL;
  if (B <= 0 || K <= 0) return EP_ERR_SHAPE;
  LinWs lw;
  void* d3 = nullptr;
  if (dlogits && lin_workspace && F > 0 && lin_ops_ok(F, K) && lin_ws(lin_workspace, lin_workspace_bytes, B, F, K, &lw))
    d3 = lw.d_r;
  return launch_ce(logits, targets, B, K, loss_scale, grad_scale, step_loss, dlogits, correct, d3, (int)pad64(K), scratch,
                   loss_acc, (cudaStream_t)stream);
}

extern "C" int ep_bn_bwd_ops(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh,
                             const float* out, const float* v_b, int x_dtype, int N, int D, int M, int d_out,
                             void* workspace, size_t workspace_bytes, int ops, void* stream) {
  OpsScope scope(ops);
  if (!dy || !y || !save_invstd || !dh || !out) return EP_ERR_NULL;
  if (B <= 0 || F <= 0 || N <= 0 || D <= 0 || M <= 0 || d_out <= 0 || D % (d_out * M) != 0 || F != D / d_out)
    return EP_ERR_SHAPE;
  const Ws w = carve(B, N, D, M);
  if (!workspace || workspace_bytes < w.total) return EP_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = launch_bn_bwd(dy, y, save_invstd, B, F, dh, s))) return rc;
  // what bwd_proj_impl would derive from dh = g_out, by the same predicates
  const int c = F / M;
  const bool tc = use_tc() && c % 4 == 0;
  const bool hilo = p_hilo(x_dtype, B, N, D, M, d_out);
  const int bf = kSplitBf16 && c % 8 == 0;
  return launch_g_ops(dh, out, v_b, B, M, c, (float*)((char*)workspace + w.delta), tc ? (char*)workspace + w.g_r : nullptr, bf,
                      (tc && hilo) ? (char*)workspace + w.g_t : nullptr, s);
}

extern "C" int ep_refresh_operands(const float* cls_token, const float* v_w, float scale, int x_dtype, int B, int N, int D,
                                   int M, int d_out, void* workspace, size_t workspace_bytes, const float* fc_w, int K,
                                   void* lin_workspace, size_t lin_workspace_bytes, int which, void* stream) {
  if (which == 0) which = EP_REFRESH_ALL;
  if (!cls_token || !v_w) return EP_ERR_NULL;
  if (B <= 0 || N <= 0 || D <= 0 || M <= 0 || d_out <= 0 || D % (d_out * M) != 0) return EP_ERR_SHAPE;
  const Ws w = carve(B, N, D, M);
  if (!workspace || workspace_bytes < w.total) return EP_ERR_WORKSPACE;
  const int Dp = D / d_out, c = Dp / M;
  RefreshJob jobs[kMaxRefreshJobs];
  int n = 0, rc = 0;
  if ((which & EP_REFRESH_QUERIES) && use_sm100(x_dtype, B, N, D, M, &rc)) {           // scaled queries as hi/lo rows (one-pass forward, logit kernels)
    RefreshJob j{};
    j.type = REFRESH_HILO; j.src = cls_token; j.dst = sm100_qhl_ptr((char*)workspace + w.sm100, B, N, D, M);
    j.scale = scale; j.M = M; j.J = sm100_J(N, D, M); j.D = D;
    jobs[n++] = j;
  }
  if ((which & EP_REFRESH_VALUE) && p_hilo(x_dtype, B, N, D, M, d_out)) {            // projection: [hi | hi | lo] rows of v.weight
    RefreshJob j{};
    j.type = REFRESH_ROWS; j.src = v_w; j.dst = (char*)workspace + w.w_p;
    j.R = Dp; j.K = D; j.Kp = D; j.ld = D; j.kind = 1; j.bf16 = 1;
    jobs[n++] = j;
  }
  if ((which & EP_REFRESH_VALUE) && use_tc() && c % 4 == 0) {                        // dP = g . W_m: per-query transposed [hi | hi | lo] rows [m][d][3c]
    RefreshJob j{};
    j.type = REFRESH_TRANSPOSE; j.src = v_w; j.dst = (char*)workspace + w.w_t;
    j.K = c; j.Kp = c; j.R = D; j.Z = M; j.src_z = (long long)c * D; j.dst_z = (long long)3 * c * D;
    j.kind = 1; j.bf16 = kSplitBf16 && c % 8 == 0;
    jobs[n++] = j;
  }
  LinWs lw;
  if ((which & EP_REFRESH_FC) && fc_w && K > 0 && lin_ops_ok(Dp, K) && lin_ws(lin_workspace, lin_workspace_bytes, B, Dp, K, &lw)) {
    const int Fp = (int)pad64(Dp), Kp = (int)pad64(K);
    RefreshJob j{};                                     // logits: [hi | hi | lo] rows of fc.weight, thirds padded to Fp
    j.type = REFRESH_ROWS; j.src = fc_w; j.dst = lw.w_r;
    j.R = K; j.K = Dp; j.Kp = Fp; j.ld = Dp; j.kind = 1; j.bf16 = 1;
    jobs[n++] = j;
    RefreshJob t{};                                     // dy: fc.weight^T as [f][3Kp]
    t.type = REFRESH_TRANSPOSE; t.src = fc_w; t.dst = lw.w_t;
    t.K = K; t.Kp = Kp; t.R = Dp; t.Z = 1; t.kind = 1; t.bf16 = 1;
    jobs[n++] = t;
  }
  return launch_refresh(jobs, n, (cudaStream_t)stream);
}
