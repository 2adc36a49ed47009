// Small head kernels after the pooling: BatchNorm1d(affine=False), cross-entropy forward+backward,
// fused multi-tensor LARS.  All fp32.
#include "ep_common.cuh"

namespace ep {

// ---------------- BatchNorm1d(affine=False, eps) -- probe_heads.py:109-110 ----------------
// one CTA per 8 features (a 32-byte sector per row), 8 x 128 threads: x = feature, y strides the batch.
// F/8 CTAs instead of F/32 keep 128 SMs busy at F = 1024; a warp still reads whole sectors.
constexpr int BN_F = 8, BN_Y = 128, BN_CACHE = 8;
__device__ __forceinline__ float block_colsum(float v, float (*red)[BN_F + 1]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  for (int h = BN_Y / 2; h > 0; h >>= 1) {
    if (threadIdx.y < h) red[threadIdx.y][threadIdx.x] += red[threadIdx.y + h][threadIdx.x];
    __syncthreads();
  }
  const float r = red[0][threadIdx.x];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(BN_F * BN_Y)
bn_fwd_kernel(const float* __restrict__ h, int B, int F, float eps, float momentum, int training,
              float* __restrict__ running_mean, float* __restrict__ running_var, long long* nbt,
              float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_invstd,
              __nv_bfloat16* __restrict__ y3, int Fp) {
  // y3 (nullable): the [hi | lo | hi] bf16 operand copy of y, rows of 3 Fp (thirds zero-padded from F to Fp; the grid
  // then covers Fp features) -- what the tcgen05 classifier GEMM reads (ep_linear_fwd_ops)
  __shared__ float red[BN_Y][BN_F + 1];
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.x * BN_F + threadIdx.x;
  const bool ok = f < F;
  float mean, invstd;
  // a thread's rows stay in registers between the passes when the batch is small enough (B <= 1024: one load, not three)
  const bool cached = B <= BN_Y * BN_CACHE;
  float hv[BN_CACHE];
  if (cached) {
#pragma unroll
    for (int i = 0; i < BN_CACHE; ++i) {
      const int b = threadIdx.y + i * BN_Y;
      hv[i] = (ok && b < B) ? h[(size_t)b * F + f] : 0.f;
    }
  }
  if (training) {
    float s = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < BN_CACHE; ++i) if (threadIdx.y + i * BN_Y < B) s += hv[i];
    } else if (ok) for (int b = threadIdx.y; b < B; b += BN_Y) s += h[(size_t)b * F + f];
    mean = block_colsum(s, red) / B;
    float v = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < BN_CACHE; ++i) if (threadIdx.y + i * BN_Y < B) { const float d = hv[i] - mean; v = fmaf(d, d, v); }
    } else if (ok) for (int b = threadIdx.y; b < B; b += BN_Y) { const float d = h[(size_t)b * F + f] - mean; v = fmaf(d, d, v); }
    const float var = block_colsum(v, red) / B;                    // biased, used to normalise
    invstd = rsqrtf(var + eps);
    if (ok && threadIdx.y == 0) {
      const float unbiased = B > 1 ? var * ((float)B / (float)(B - 1)) : var;
      running_mean[f] = (1.f - momentum) * running_mean[f] + momentum * mean;
      running_var[f] = (1.f - momentum) * running_var[f] + momentum * unbiased;
      if (save_mean) save_mean[f] = mean;
      if (save_invstd) save_invstd[f] = invstd;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && nbt) *nbt += 1;
  } else {
    mean = ok ? running_mean[f] : 0.f;
    invstd = ok ? 1.f / sqrtf(running_var[f] + eps) : 0.f;
  }
  auto emit = [&](int b, float hval) {
    const float v = (hval - mean) * invstd;
    y[(size_t)b * F + f] = v;
    if (y3) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* r = y3 + (size_t)b * 3 * Fp + f;
      r[0] = hi; r[Fp] = lo; r[2 * Fp] = hi;
    }
  };
  if (ok) {
    if (cached) {
#pragma unroll
      for (int i = 0; i < BN_CACHE; ++i) if (threadIdx.y + i * BN_Y < B) emit(threadIdx.y + i * BN_Y, hv[i]);
    } else {
      for (int b = threadIdx.y; b < B; b += BN_Y) emit(b, h[(size_t)b * F + f]);
    }
  } else if (y3 && f < Fp) for (int b = threadIdx.y; b < B; b += BN_Y) {
    __nv_bfloat16* r = y3 + (size_t)b * 3 * Fp + f;
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    r[0] = z; r[Fp] = z; r[2 * Fp] = z;
  }
}

__global__ void __launch_bounds__(BN_F * BN_Y)
bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ invstd, int B, int F,
              float* __restrict__ dh) {
  __shared__ float red[BN_Y][BN_F + 1];
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.x * BN_F + threadIdx.x;
  const bool ok = f < F;
  float s1 = 0.f, s2 = 0.f;
  const bool cached = B <= BN_Y * BN_CACHE;                        // rows of this thread kept in registers
  float gv[BN_CACHE], yv[BN_CACHE];
  if (cached) {
#pragma unroll
    for (int i = 0; i < BN_CACHE; ++i) {
      const int b = threadIdx.y + i * BN_Y;
      const bool in = ok && b < B;
      gv[i] = in ? dy[(size_t)b * F + f] : 0.f;
      yv[i] = in ? y[(size_t)b * F + f] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < BN_CACHE; ++i) if (threadIdx.y + i * BN_Y < B) { s1 += gv[i]; s2 = fmaf(gv[i], yv[i], s2); }
  } else if (ok) for (int b = threadIdx.y; b < B; b += BN_Y) {
    const float g = dy[(size_t)b * F + f];
    s1 += g;
    s2 = fmaf(g, y[(size_t)b * F + f], s2);
  }
  const float m1 = block_colsum(s1, red) / B;
  const float m2 = block_colsum(s2, red) / B;
  if (ok) {
    const float is = invstd[f];
    if (cached) {
#pragma unroll
      for (int i = 0; i < BN_CACHE; ++i) {
        const int b = threadIdx.y + i * BN_Y;
        if (b < B) dh[(size_t)b * F + f] = is * (gv[i] - m1 - yv[i] * m2);
      }
    } else {
      for (int b = threadIdx.y; b < B; b += BN_Y)
        dh[(size_t)b * F + f] = is * (dy[(size_t)b * F + f] - m1 - y[(size_t)b * F + f] * m2);
    }
  }
}

// ---------------- CrossEntropyLoss (mean) fwd+bwd, one CTA per row ----------------
constexpr int CE_CACHE = 4;
__global__ void __launch_bounds__(256)
ce_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, int K, float loss_scale,
          float grad_scale, float* loss_sum, float* __restrict__ dlogits, int* correct,
          __nv_bfloat16* __restrict__ d3, int Kp, float* scratch, float* loss_acc) {
  // d3 (nullable): the [hi | lo | hi] bf16 operand copy of dlogits, rows of 3 Kp, thirds zero-padded from K to Kp
  // (ep_linear_bwd_ops).  scratch (nullable): gridDim.x row losses + a counter word -- the last CTA to finish adds the
  // rows in a fixed order, OVERWRITES loss_sum[0] and adds the value to loss_acc (no zeroing, deterministic)
  __shared__ float redf[8];
  __shared__ int redi[8];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = logits + (size_t)b * K;
  // the row stays in registers between the passes when K <= 1024 (4 values per thread: one load, one exp per element)
  const bool cached = K <= 256 * CE_CACHE;
  float rv[CE_CACHE];
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  if (cached) {
#pragma unroll
    for (int i = 0; i < CE_CACHE; ++i) {
      const int k = threadIdx.x + i * 256;
      rv[i] = k < K ? row[k] : -INFINITY;
      if (rv[i] > mx) { mx = rv[i]; arg = k; }
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 256) {
      const float v = row[k];
      if (v > mx) { mx = v; arg = k; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  if (lane == 0) { redf[warp] = mx; redi[warp] = arg; }
  __syncthreads();
  mx = redf[0]; arg = redi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (redf[w] > mx || (redf[w] == mx && redi[w] < arg)) { mx = redf[w]; arg = redi[w]; }
  __syncthreads();
  float s = 0.f;
  if (cached) {
#pragma unroll
    for (int i = 0; i < CE_CACHE; ++i) {
      rv[i] = threadIdx.x + i * 256 < K ? expf(rv[i] - mx) : 0.f;      // from here on: exp(logit - max)
      s += rv[i];
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 256) s += expf(row[k] - mx);
  }
  s = warp_sum(s);
  if (lane == 0) redf[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += redf[w];
  const int t = (int)targets[b];
  const float inv = 1.f / s;
  auto emit = [&](int k, float e) {                                  // e = exp(logit - max), 0 in the padding
    const float v = k < K ? (e * inv - (k == t ? 1.f : 0.f)) * grad_scale : 0.f;
    if (k < K) dlogits[(size_t)b * K + k] = v;
    if (d3) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* r = d3 + (size_t)b * 3 * Kp + k;
      r[0] = hi; r[Kp] = lo; r[2 * Kp] = hi;
    }
  };
  if (dlogits) {
    const int kend = d3 ? Kp : K;
    if (cached) {
#pragma unroll
      for (int i = 0; i < CE_CACHE; ++i) if (threadIdx.x + i * 256 < kend) emit(threadIdx.x + i * 256, rv[i]);
    } else {
      for (int k = threadIdx.x; k < kend; k += 256) emit(k, k < K ? expf(row[k] - mx) : 0.f);
    }
  }
  const float nll = (logf(s) + mx - row[t]) * loss_scale;
  if (!scratch) {
    if (threadIdx.x == 0) {
      if (loss_sum) atomicAdd(loss_sum, nll);
      if (correct && arg == t) atomicAdd(correct, 1);
    }
    return;
  }
  __shared__ bool last;
  if (threadIdx.x == 0) {
    if (correct && arg == t) atomicAdd(correct, 1);
    scratch[b] = nll;
    __threadfence();
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + gridDim.x);
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float acc = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) acc += __ldcg(scratch + i);   // fixed order: strided, then tree
  acc = warp_sum(acc);
  __syncthreads();
  if (lane == 0) redf[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += redf[w];
    if (loss_sum) loss_sum[0] = tot;
    if (loss_acc) loss_acc[0] += tot;
    *reinterpret_cast<unsigned*>(scratch + gridDim.x) = 0u;
  }
}

// ---------------- LARS (util/lars.py:13-37), all tensors in two launches ----------------
struct LarsArgs {
  float* p[EP_LARS_MAX_TENSORS];
  const float* g[EP_LARS_MAX_TENSORS];
  float* mu[EP_LARS_MAX_TENSORS];
  long long n[EP_LARS_MAX_TENSORS];
  int trust[EP_LARS_MAX_TENSORS];
  int count;
};
// hyper = {lr, weight_decay, momentum, trust_coefficient, grad_scale}

// deterministic norms: every CTA writes its partial sums, the update kernel adds them in a fixed order
__global__ void __launch_bounds__(256) lars_norm_kernel(LarsArgs a, const float* __restrict__ hyper, float* partial, int vec) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.y;
  if (!a.trust[t]) return;
  const float wd = hyper[1], gs = hyper[4];
  float sp = 0.f, su = 0.f;
  // vec: every tensor is 16-byte aligned -- 128-bit loads for the bulk, the last n % 4 elements one by one
  const long long n4 = vec ? a.n[t] >> 2 : 0;
  const float4* p4 = reinterpret_cast<const float4*>(a.p[t]);
  const float4* g4 = reinterpret_cast<const float4*>(a.g[t]);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 p = p4[i], g = g4[i];
    const float pv[4] = {p.x, p.y, p.z, p.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float u = fmaf(wd, pv[e], gv[e] * gs);
      sp = fmaf(pv[e], pv[e], sp);
      su = fmaf(u, u, su);
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < a.n[t]; i += (long long)gridDim.x * 256) {
    const float p = a.p[t][i];
    const float u = fmaf(wd, p, a.g[t][i] * gs);
    sp = fmaf(p, p, sp);
    su = fmaf(u, u, su);
  }
  __shared__ float r0[8], r1[8];
  sp = warp_sum(sp); su = warp_sum(su);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = sp; r1[threadIdx.x >> 5] = su; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { x += r0[w]; y += r1[w]; }
    partial[((size_t)t * gridDim.x + blockIdx.x) * 2] = x;
    partial[((size_t)t * gridDim.x + blockIdx.x) * 2 + 1] = y;
  }
}

__global__ void __launch_bounds__(256) lars_update_kernel(LarsArgs a, const float* __restrict__ hyper,
                                                          const float* __restrict__ partial, int vec) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.y;
  const float lr = hyper[0], wd = hyper[1], mom = hyper[2], tc = hyper[3], gs = hyper[4];
  __shared__ float qs;
  const bool tr = a.trust[t] != 0;
  if (threadIdx.x < 32) {
    float q = 1.f;
    if (tr) {
      float x = 0.f, y = 0.f;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {       // fixed order: lane-strided, then butterfly
        x += partial[((size_t)t * gridDim.x + i) * 2];
        y += partial[((size_t)t * gridDim.x + i) * 2 + 1];
      }
      x = warp_sum(x); y = warp_sum(y);
      const float pn = sqrtf(x), un = sqrtf(y);
      q = (pn > 0.f && un > 0.f) ? tc * pn / un : 1.f;
    }
    if (threadIdx.x == 0) qs = q;
  }
  __syncthreads();
  const float q = qs;
  const long long n4 = vec ? a.n[t] >> 2 : 0;
  float4* p4 = reinterpret_cast<float4*>(a.p[t]);
  float4* m4 = reinterpret_cast<float4*>(a.mu[t]);
  const float4* g4 = reinterpret_cast<const float4*>(a.g[t]);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 p = p4[i], g = g4[i], mu = m4[i];
    float pv[4] = {p.x, p.y, p.z, p.w}, mv[4] = {mu.x, mu.y, mu.z, mu.w};
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float u = gv[e] * gs;
      if (tr) u = fmaf(wd, pv[e], u) * q;
      mv[e] = fmaf(mom, mv[e], u);
      pv[e] = fmaf(-lr, mv[e], pv[e]);
    }
    m4[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
    p4[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < a.n[t]; i += (long long)gridDim.x * 256) {
    const float p = a.p[t][i];
    float u = a.g[t][i] * gs;
    if (tr) u = fmaf(wd, p, u) * q;
    const float m = fmaf(mom, a.mu[t][i], u);
    a.mu[t][i] = m;
    a.p[t][i] = fmaf(-lr, m, p);
  }
}

// ---------------- AdamW / SGD (main_linprobe.py:403-408: {"lars": LARS, "adamw": AdamW}, else SGD) ----------------
struct OptArgs {
  float* p[EP_LARS_MAX_TENSORS];
  const float* g[EP_LARS_MAX_TENSORS];
  float* s1[EP_LARS_MAX_TENSORS];     // AdamW: exp_avg; SGD: momentum_buffer
  float* s2[EP_LARS_MAX_TENSORS];     // AdamW: exp_avg_sq
  long long n[EP_LARS_MAX_TENSORS];
};

// torch.optim.AdamW (decoupled weight decay).  hyper = {lr, beta1, beta2, eps, weight_decay, grad_scale,
// bias_correction1 = 1 - beta1^t, bias_correction2 = 1 - beta2^t}
__global__ void __launch_bounds__(256) adamw_kernel(OptArgs a, const float* __restrict__ hyper) {
  const int t = blockIdx.y;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gs = hyper[5];
  const float step_size = lr / hyper[6], inv_bc2_sqrt = rsqrtf(hyper[7]);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n[t]; i += (long long)gridDim.x * 256) {
    const float g = a.g[t][i] * gs;
    float p = a.p[t][i] * (1.f - lr * wd);
    const float m = b1 * a.s1[t][i] + (1.f - b1) * g;
    const float v = b2 * a.s2[t][i] + (1.f - b2) * g * g;
    a.s1[t][i] = m;
    a.s2[t][i] = v;
    p -= step_size * m / (sqrtf(v) * inv_bc2_sqrt + eps);
    a.p[t][i] = p;
  }
}

// torch.optim.SGD (dampening 0, no nesterov).  hyper = {lr, weight_decay, momentum, grad_scale, first_step}
__global__ void __launch_bounds__(256) sgd_kernel(OptArgs a, const float* __restrict__ hyper) {
  const int t = blockIdx.y;
  const float lr = hyper[0], wd = hyper[1], mom = hyper[2], gs = hyper[3];
  const bool first = hyper[4] != 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < a.n[t]; i += (long long)gridDim.x * 256) {
    const float p = a.p[t][i];
    float d = fmaf(wd, p, a.g[t][i] * gs);
    if (mom != 0.f && a.s1[t]) {
      d = first ? d : fmaf(mom, a.s1[t][i], d);
      a.s1[t][i] = d;
    }
    a.p[t][i] = fmaf(-lr, d, p);
  }
}

}  // namespace ep

using namespace ep;

namespace {
int fill_opt_args(OptArgs* a, int n, float* const* params, const float* const* grads, float* const* s1, float* const* s2,
                  const long long* numels, bool need_s1, bool need_s2, long long* mx) {
  if (n <= 0 || n > EP_LARS_MAX_TENSORS) return EP_ERR_SHAPE;
  if (!params || !grads || !numels || (need_s1 && !s1) || (need_s2 && !s2)) return EP_ERR_NULL;
  *mx = 0;
  for (int i = 0; i < n; ++i) {
    if (!params[i] || !grads[i] || (need_s1 && !s1[i]) || (need_s2 && !s2[i])) return EP_ERR_NULL;
    a->p[i] = params[i]; a->g[i] = grads[i]; a->s1[i] = s1 ? s1[i] : nullptr; a->s2[i] = s2 ? s2[i] : nullptr;
    a->n[i] = numels[i];
    if (numels[i] > *mx) *mx = numels[i];
  }
  return 0;
}
int opt_grid(long long mx) {
  int bx = (int)((mx + 256 * 8 - 1) / (256 * 8));
  return bx < 1 ? 1 : (bx > 2 * kNumSMs ? 2 * kNumSMs : bx);
}
}  // namespace

extern "C" int ep_adamw_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const long long* numels, const float* hyper, void* stream) {
  OptArgs a;
  long long mx;
  if (!hyper) return EP_ERR_NULL;
  int rc = fill_opt_args(&a, n, params, grads, exp_avg, exp_avg_sq, numels, true, true, &mx);
  if (rc) return rc;
  adamw_kernel<<<dim3(opt_grid(mx), n), 256, 0, (cudaStream_t)stream>>>(a, hyper);
  EP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ep_sgd_step(int n, float* const* params, const float* const* grads, float* const* momentum_buf,
                           const long long* numels, const float* hyper, void* stream) {
  OptArgs a;
  long long mx;
  if (!hyper) return EP_ERR_NULL;
  int rc = fill_opt_args(&a, n, params, grads, momentum_buf, nullptr, numels, false, false, &mx);
  if (rc) return rc;
  sgd_kernel<<<dim3(opt_grid(mx), n), 256, 0, (cudaStream_t)stream>>>(a, hyper);
  EP_LAUNCH_CHECK();
  return 0;
}


extern "C" int ep_bn_fwd(const float* h, int B, int F, float eps, float momentum, int training, float* running_mean,
                         float* running_var, long long* nbt, float* y, float* save_mean, float* save_invstd,
                         void* stream) {
  if (!h || !y || !running_mean || !running_var) return EP_ERR_NULL;
  if (B <= 0 || F <= 0) return EP_ERR_SHAPE;
  return launch_bn_fwd(h, B, F, eps, momentum, training, running_mean, running_var, nbt, y, save_mean, save_invstd, nullptr, 0,
                       (cudaStream_t)stream);
}
int ep::launch_bn_fwd(const float* h, int B, int F, float eps, float momentum, int training, float* running_mean,
                      float* running_var, long long* nbt, float* y, float* save_mean, float* save_invstd, void* y3, int Fp,
                      cudaStream_t s) {
  const int cols = y3 ? Fp : F;
  EP_CUDA(launch_pdl(bn_fwd_kernel, dim3((cols + BN_F - 1) / BN_F), dim3(BN_F, BN_Y), 0, s, h, B, F, eps, momentum, training,
                     running_mean, running_var, nbt, y, save_mean, save_invstd, (__nv_bfloat16*)y3, Fp));
  EP_LAUNCH_CHECK();
  return 0;
}
int ep::launch_bn_bwd(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh, cudaStream_t s) {
  EP_CUDA(launch_pdl(bn_bwd_kernel, dim3((F + BN_F - 1) / BN_F), dim3(BN_F, BN_Y), 0, s, dy, y, save_invstd, B, F, dh));
  EP_LAUNCH_CHECK();
  return 0;
}
int ep::launch_ce(const float* logits, const long long* targets, int B, int K, float loss_scale, float grad_scale,
                  float* loss_sum, float* dlogits, int* correct, void* d3, int Kp, float* scratch, float* loss_acc,
                  cudaStream_t s) {
  EP_CUDA(launch_pdl(ce_kernel, dim3(B), dim3(256), 0, s, logits, targets, K, loss_scale, grad_scale, loss_sum, dlogits, correct,
                     (__nv_bfloat16*)d3, Kp, scratch, loss_acc));
  EP_LAUNCH_CHECK();
  return 0;
}

extern "C" int ep_bn_bwd(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh, void* stream) {
  if (!dy || !y || !save_invstd || !dh) return EP_ERR_NULL;
  if (B <= 0 || F <= 0) return EP_ERR_SHAPE;
  return launch_bn_bwd(dy, y, save_invstd, B, F, dh, (cudaStream_t)stream);
}

extern "C" int ep_ce_fwd_bwd(const float* logits, const long long* targets, int B, int K, float loss_scale,
                             float grad_scale, float* loss_sum, float* dlogits, int* correct, void* stream) {
  if (!logits || !targets) return EP_ERR_NULL;
  if (B <= 0 || K <= 0) return EP_ERR_SHAPE;
  return launch_ce(logits, targets, B, K, loss_scale, grad_scale, loss_sum, dlogits, correct, nullptr, 0, nullptr, nullptr,
                   (cudaStream_t)stream);
}

extern "C" int ep_lars_step(int n, float* const* params, const float* const* grads, float* const* mus,
                            const long long* numels, const int* apply_trust, const float* hyper, float* scratch,
                            void* stream) {
  if (n <= 0 || n > EP_LARS_MAX_TENSORS) return EP_ERR_SHAPE;
  if (!params || !grads || !mus || !numels || !apply_trust || !hyper || !scratch) return EP_ERR_NULL;
  LarsArgs a;
  a.count = n;
  long long mx = 0;
  int vec = 1;
  for (int i = 0; i < n; ++i) {
    if (!params[i] || !grads[i] || !mus[i]) return EP_ERR_NULL;
    a.p[i] = params[i]; a.g[i] = grads[i]; a.mu[i] = mus[i]; a.n[i] = numels[i]; a.trust[i] = apply_trust[i];
    if (numels[i] > mx) mx = numels[i];
    if ((reinterpret_cast<uintptr_t>(params[i]) | reinterpret_cast<uintptr_t>(grads[i]) | reinterpret_cast<uintptr_t>(mus[i])) & 15) vec = 0;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int bx = (int)((mx + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  if (bx > 2 * kNumSMs) bx = 2 * kNumSMs;                  // 2 * n * bx <= EP_LARS_SCRATCH_FLOATS
  EP_CUDA(launch_pdl(lars_norm_kernel, dim3(bx, n), dim3(256), 0, s, a, hyper, scratch, vec));
  EP_LAUNCH_CHECK();
  EP_CUDA(launch_pdl(lars_update_kernel, dim3(bx, n), dim3(256), 0, s, a, hyper, (const float*)scratch, vec));
  EP_LAUNCH_CHECK();
  return 0;
}
