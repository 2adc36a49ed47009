/* ep_b200.h -- C ABI of libep_b200.so, the sm_100a implementation of the EP probe-head hot path.
 *
 * Every entry point replaces a piece of the reference's PyTorch-eager path (paths relative to
 * billpsomas/efficient-probing); the Python host side (efficient-probing_b200/) binds them with
 * ctypes, INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named *_host;
 *   - the library never allocates, frees or retains device memory: outputs and workspace are
 *     caller-owned (PyTorch caching allocator on the reference side);
 *   - every compute call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *     synchronisation, is CUDA-graph capturable and keeps no state between calls.  The developer
 *     knobs (ep_set_kernel_mode, ep_set_gemm_mode, ep_set_sm_limit, ep_set_debug) are PROCESS-WIDE
 *     settings read when a call is made: set them before capturing a graph (their value is baked
 *     into it), not concurrently with calls on other threads, and do not change the kernel / GEMM
 *     mode between a forward and its backward (ep_pooled_layout tells which layout of P a mode
 *     writes).  ep_set_debug(32) timers synchronise the stream at the end of a call (developer aid);
 *   - return 0 on success, a NEGATIVE ep_status for rejected arguments, a POSITIVE value for a
 *     cudaError_t raised by a launch.  There is no CPU fallback of any kind.
 *   - tensors are contiguous row-major; token tensors x are (B, N, D) bf16 (EP_DTYPE_BF16, the fast
 *     path) or fp32 (EP_DTYPE_F32, converted on load); parameters, statistics, gradients fp32.
 */
#ifndef EP_B200_H_
#define EP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EP_ABI_VERSION 2

#define EP_DTYPE_BF16 0
#define EP_DTYPE_F32  1

typedef enum {
  EP_OK = 0,
  EP_ERR_NULL = -1,        /* a required pointer is NULL                                   */
  EP_ERR_SHAPE = -2,       /* B,N,D,M,d_out <= 0 or D % (d_out*M) != 0 (ep.py:40 reshape)   */
  EP_ERR_ALIGN = -3,       /* D % 8 != 0 (128-bit token loads) or misaligned pointer        */
  EP_ERR_DTYPE = -4,       /* x_dtype not EP_DTYPE_BF16 / EP_DTYPE_F32                      */
  EP_ERR_WORKSPACE = -5,   /* workspace smaller than ep_workspace_bytes()                   */
  EP_ERR_UNSUPPORTED = -6, /* shape outside what the kernels cover (e.g. N*M too large)     */
  EP_ERR_DEVICE = -7       /* current device is not compute capability 10.x                 */
} ep_status;

int ep_abi_version(void);
const char* ep_strerror(int code);
/* 0 when the current CUDA device is sm_100-class, EP_ERR_DEVICE otherwise. */
int ep_device_check(void);
/* Select the pooling kernel family: 0 = automatic (tcgen05 kernels where the shape allows, else the
 * general kernels), 1 = force the general CUDA-core kernels, 2 = force tcgen05 (error if unsupported). */
int ep_set_kernel_mode(int mode);
/* Precision of the head's small dense contractions (value projection, classifier and their gradients):
 * 0 (default) = tensor cores, fp32 accumulate: three-term bf16 / TF32 products (~5e-6 relative) in the tcgen05
 *     GEMMs, plain TF32 (~3e-4) in the mma.sync weight-gradient kernel;
 * 1 = fp32 FMA on the CUDA cores (evaluation, where top-1 decisions must not move).
 * Like the kernel mode it must not change between a forward and its backward (it decides the layout of P). */
int ep_set_gemm_mode(int mode);
/* Number of CTAs the persistent token-streaming kernels (one CTA per SM) launch from now on; 0 = one per SM of
 * the device (default).  A data-parallel caller lowers it around ep_bwd_pool so that the gradient all-reduce it
 * overlaps with that call (main_linprobe.py:581-583) finds free SMs instead of delaying statically scheduled
 * CTAs.  Process-wide; read at launch (i.e. at capture time for CUDA graphs). */
int ep_set_sm_limit(int n);
/* Which family the last ep_fwd/ep_bwd call on this thread used (1 = general, 2 = tcgen05). */
int ep_last_kernel_family(void);
/* Developer knob: bit 5 (32) makes ep_fwd / ep_bwd_proj / ep_bwd_pool bracket each of their kernels with
 * CUDA events on the call's stream (one host sync per call) and record the durations, read back with
 * ep_timing_*; bit 8 (256) also prints them; bit 9 (512) runs the forward softmax as a separate kernel instead
 * of in the logit kernel's epilogue (same results); bit 10 (1024) selects the two-kernel path per direction
 * instead of the one-pass kernels (same results; cross-check).  Bits 0-3 and 16-27 override the one-pass kernels'
 * plan (same results), bits 11-15 and 28-29 select their instrumented
 * instantiation (pipeline stamps, pair mode, timing experiments whose RESULTS ARE GARBAGE): csrc/ep_fused_sm100.cu.
 * 0 in normal use. */
int ep_set_debug(int flags);
int ep_timing_count(void);
int ep_timing_get(int i, char* name, int name_len, float* microseconds);
int ep_timing_reset(void);
/* Developer aid: with ep_set_debug bit 11 (2048) the one-pass fused kernels stamp clock64 at their pipeline
 * hand-offs (CTA 0, first 8 samples, 16 stamps each); this copies the first n (<= 128) stamps to the host. */
int ep_debug_trace(long long* host_out, int n);
/* Which family ep_fwd/ep_bwd would use for this shape under the current mode (0 = none: forced tcgen05
 * but unsupported). */
int ep_kernel_family_for(int x_dtype, int B, int N, int D, int M);
/* Layout of the saved pooled-token buffer P that ep_fwd writes and ep_bwd / ep_bwd_proj read back, for this
 * shape under the current kernel and GEMM modes (which must not change between a forward and its backward):
 *   0: fp32 (B, M, D);
 *   1: bf16 pairs (B, M, 2, D) -- row 0 = bf16(P), row 1 = bf16(P - row 0) -- the same number of bytes, read in
 *      place as the [hi | lo | hi] operand of the tcgen05 projection and weight-gradient GEMMs.
 * Callers only need this to inspect P; the buffer is always B*M*D*4 bytes.  < 0: EP_ERR_SHAPE. */
int ep_pooled_layout(int x_dtype, int B, int N, int D, int M, int d_out);
/* Number of kernels this library has launched in this process (host-side count; launches replayed by a
 * CUDA graph are not seen here -- count the captured step once and multiply). */
unsigned long long ep_launch_count(void);

/* Bytes of scratch ep_fwd / ep_bwd / ep_attention_maps need for this shape (max over the three). */
size_t ep_workspace_bytes(int B, int N, int D, int M, int d_out);

/* EfficientProbing.forward -- poolings/ep.py:28-47 (num_heads == 1).
 *   S[b,m,n] = scale * cls_token[m] . x[b,n];  A = softmax_n(S);  P[b,m] = sum_n A[b,m,n] x[b,n];
 *   out[b, m*c:(m+1)*c] = v_w[m*c:(m+1)*c] @ P[b,m] (+ v_b),  c = D / (d_out*M).
 * Outputs: out (B, D/d_out); saved for backward: the logits S (B, M, N), rowmax, rowsum (B, M) -- the
 * shift and the normaliser of the softmax (A = exp(S - rowmax) / rowsum) -- and P (B, M, D).
 * attn (B, M, N) is written when non-NULL (the attention maps of tools/ep_attention_maps.py:51-58). */
int ep_fwd(const void* x, int x_dtype, const float* cls_token, const float* v_w, const float* v_b,
           float scale, int B, int N, int D, int M, int d_out,
           float* out, float* S, float* rowmax, float* rowsum, float* P, float* attn,
           void* workspace, size_t workspace_bytes, void* stream);

/* Backward of ep_fwd (replaces autograd through ep.py:35-45; engine_finetune.py:73 -> misc.py:267).
 * g_out = dL/d out (B, D/d_out); out = what ep_fwd returned and v_b the bias it used (nullable) -- the
 * softmax-backward row term sum_n A dA equals g . (out - v_b) per query, so no pass over P is needed for it.
 * Writes d_cls_token (M, D), d_v_w (D/d_out, D), d_v_b (D/d_out, nullable).  The frozen-backbone path needs
 * no dL/dx (main_linprobe.py:393-400; ep_bwd_ex produces it). */
int ep_bwd(const void* x, int x_dtype, const float* cls_token, const float* v_w, float scale,
           int B, int N, int D, int M, int d_out,
           const float* S, const float* rowmax, const float* rowsum, const float* P, const float* out,
           const float* v_b, const float* g_out,
           float* d_cls_token, float* d_v_w, float* d_v_b,
           void* workspace, size_t workspace_bytes, void* stream);

/* The two halves of ep_bwd, for callers that overlap the gradient all-reduce with the token-streaming
 * half (the DDP bucket overlap of main_linprobe.py:581-583, done explicitly):
 *   ep_bwd_proj : needs only g_out, out and the saved P -- writes d_v_w, d_v_b (99% of the EP gradient bytes)
 *                 and leaves dP = g_out . v_w and delta = g_out . (out - v_b) in the workspace;
 *   ep_bwd_pool : streams x once, recomputes A, writes d_cls_token.  Must follow ep_bwd_proj on the
 *                 same stream with the same workspace. */
int ep_bwd_proj(const float* g_out, const float* P, const float* out, const float* v_w, const float* v_b, int x_dtype,
                int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                void* workspace, size_t workspace_bytes, void* stream);
int ep_bwd_pool(const void* x, int x_dtype, const float* cls_token, float scale, int B, int N, int D, int M, int d_out,
                const float* S, const float* rowmax, const float* rowsum, float* d_cls_token,
                void* workspace, size_t workspace_bytes, void* stream);

/* The rest of EfficientProbing.forward's surface (poolings/ep.py:28-33) and the input gradient, on the general
 * (CUDA-core) kernel family -- the paths no reference script exercises while probing a frozen backbone:
 *   cls_batched != 0 : cls_token is (B, M, D), the per-sample external queries of forward(x, cls=...)
 *                      (ep.py:32-33); d_cls_token is then (B, M, D) as well;
 *   dx != NULL       : dL/dx (B, N, D) in x's dtype is written too (--finetuning, main_linprobe.py:152-154):
 *                      dx[b,n] = sum_m  scale * dS[b,m,n] * cls_token[m] + A[b,m,n] * dP[b,m].
 * ep_fwd_ex always saves P as fp32 (layout 0).  ep_bwd_ex takes the layout of the P it is given: 0 after
 * ep_fwd_ex, ep_pooled_layout(...) after ep_fwd (so dx is available behind the fast forward as well). */
int ep_fwd_ex(const void* x, int x_dtype, const float* cls_token, int cls_batched, const float* v_w, const float* v_b,
              float scale, int B, int N, int D, int M, int d_out,
              float* out, float* S, float* rowmax, float* rowsum, float* P, float* attn,
              void* workspace, size_t workspace_bytes, void* stream);
int ep_bwd_ex(const void* x, int x_dtype, const float* cls_token, int cls_batched, const float* v_w, float scale,
              int B, int N, int D, int M, int d_out,
              const float* S, const float* rowmax, const float* rowsum, const float* P, int p_layout,
              const float* out, const float* v_b, const float* g_out,
              float* d_cls_token, float* d_v_w, float* d_v_b, void* dx,
              void* workspace, size_t workspace_bytes, void* stream);

/* tools/ep_attention_maps.py:51-58 for a batch: attn[b] = softmax(scale * cls_token @ x[b]^T), (B, M, N).
 * With a workspace of ep_workspace_bytes() and a shape the tcgen05 kernels cover (bf16 tokens, D % 128 == 0, N <= D)
 * the logits come from the tensor-core kernel; otherwise (or with the small workspace) from the general one. */
int ep_attention_maps(const void* x, int x_dtype, const float* cls_token, float scale,
                      int B, int N, int D, int M, float* attn,
                      void* workspace, size_t workspace_bytes, void* stream);

/* nn.BatchNorm1d(F, affine=False, eps) -- probe_heads.py:109-110.
 * training != 0: batch statistics (biased variance), running stats updated with `momentum`
 * (unbiased variance), *num_batches_tracked += 1; save_mean / save_invstd (F,) kept for backward.
 * training == 0: y = (h - running_mean) / sqrt(running_var + eps). */
int ep_bn_fwd(const float* h, int B, int F, float eps, float momentum, int training,
              float* running_mean, float* running_var, long long* num_batches_tracked,
              float* y, float* save_mean, float* save_invstd, void* stream);
/* dh = invstd * (dy - mean_b(dy) - y * mean_b(dy*y)). */
int ep_bn_bwd(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh, void* stream);

/* nn.Linear(F, K, bias=True) -- probe_heads.py:76.  logits = y @ W^T + b.
 * With a workspace of ep_linear_workspace_bytes() the contraction runs on the tensor cores in TF32
 * (operands rounded to nearest tf32, fp32 accumulate); with workspace == NULL on the CUDA cores in fp32. */
size_t ep_linear_workspace_bytes(int B, int F, int K);
int ep_linear_fwd(const float* y, const float* W, const float* b, int B, int F, int K, float* logits,
                  void* workspace, size_t workspace_bytes, void* stream);
/* dW = dlogits^T @ y (K, F); db = sum_b dlogits (K,); dy = dlogits @ W (B, F).  Any output may be NULL. */
int ep_linear_bwd(const float* dlogits, const float* y, const float* W, int B, int F, int K,
                  float* dW, float* db, float* dy, void* workspace, size_t workspace_bytes, void* stream);

/* nn.CrossEntropyLoss() (mean) forward + backward in one pass -- main_linprobe.py:589, engine_finetune.py:62.
 * loss_sum[0] += sum_b nll_b * loss_scale  (caller zeroes it; loss_scale = 1/B gives the mean);
 * dlogits = (softmax(logits) - onehot) * grad_scale  (grad_scale = 1/B for the mean loss);
 * correct[0] += #(argmax == target) when non-NULL (engine_finetune.py:63 top-1). */
int ep_ce_fwd_bwd(const float* logits, const long long* targets, int B, int K, float loss_scale, float grad_scale,
                  float* loss_sum, float* dlogits, int* correct, void* stream);

/* util/lars.py:13-37 for up to EP_LARS_MAX_TENSORS tensors in one call (pointer tables are HOST arrays
 * of device pointers).  hyper is a DEVICE array {lr, weight_decay, momentum, trust_coefficient, grad_scale}
 * so a captured graph can be replayed with a new learning rate; grad_scale (1/world_size after a
 * sum all-reduce) multiplies every gradient first.  apply_trust_host[i] != 0 for tensors with
 * ndim > 1 (lars.py:21).  scratch: EP_LARS_SCRATCH_FLOATS floats (per-CTA partial norms, summed in a fixed
 * order: the update is bit-reproducible run to run). */
#define EP_LARS_MAX_TENSORS 8
#define EP_LARS_SCRATCH_FLOATS 8192
int ep_lars_step(int n, float* const* params_host, const float* const* grads_host, float* const* mus_host,
                 const long long* numels_host, const int* apply_trust_host, const float* hyper,
                 float* scratch, void* stream);

/* ---- Operand reuse across the calls of one training step (ABI 2) --------------------------------------------
 * The tcgen05 GEMMs of the head read their fp32 operands as bf16 hi/lo copies (3-term products).  Every entry point
 * above is self-contained: it derives the copies it needs from its fp32 arguments, which costs eight small launches
 * per training step.  The *_ops variants below take an `ops` bitmask instead and let the kernel that PRODUCES a tensor
 * write the operand copy its consumer reads, into the same caller-owned workspaces (`workspace` of
 * ep_workspace_bytes(), `lin_workspace` of ep_linear_workspace_bytes(); their layout is private to the library):
 *   EP_OPS_WEIGHTS  the weight-derived copies (scaled queries, v.weight, fc.weight, both orientations) in the
 *                   workspaces are current: ep_refresh_operands() has run on them since the parameters last changed
 *                   (a training loop calls it once per step, right after the optimizer);
 *   EP_OPS_INPUT    the activation-derived copies were written by the producing *_ops call of this step:
 *                   ep_bn_fwd_ops -> y for ep_linear_fwd_ops / the dW of ep_linear_bwd_ops; ep_ce_fwd_bwd_ops ->
 *                   dlogits for ep_linear_bwd_ops; ep_bn_bwd_ops -> g_out copies and delta for ep_bwd_proj_ops;
 *   EP_OPS_FP32     this call runs its contractions as fp32 FMAs on the CUDA cores (what ep_set_gemm_mode(1) selects
 *                   process-wide, here per call and per thread: evaluation next to training on another stream).
 * A flag is a promise by the caller; with ops == 0 a *_ops call behaves exactly like the plain one.  Where a shape is
 * outside what the copies cover (bf16 rows need F % 8 == 0, K % 8 == 0, c % 4 == 0) producers and consumers fall
 * back to self-made copies by the same rule, so the flags are always safe to pass.  Results are identical either
 * way (same kernels, same operand bits); ep_linear_bwd_ops with EP_OPS_INPUT also moves the classifier weight
 * gradient from the mma.sync TF32 kernel to the tcgen05 3-term GEMM (4e-6 instead of 3e-4 relative error). */
#define EP_OPS_WEIGHTS 1
#define EP_OPS_INPUT   2
#define EP_OPS_FP32    4
/* ep_bwd_proj_ops only: run one half of the call -- everything but d_v_w / d_v_b (NO_DW: delta and dP, what
 * ep_bwd_pool needs), or d_v_w / d_v_b alone (ONLY_DW).  The halves are independent, so a caller can issue them on two
 * streams (the weight gradient is not needed before the optimizer / the gradient exchange). */
#define EP_OPS_NO_DW   8
#define EP_OPS_ONLY_DW 16
/* Writes the weight-derived operand copies: scale * cls_token as bf16 hi/lo rows (EP_REFRESH_QUERIES), v.weight as
 * [hi|hi|lo] rows and as per-query transposed [hi|hi|lo] rows (EP_REFRESH_VALUE), fc.weight (K, F = D / d_out) as
 * [hi|hi|lo] rows and transposed (EP_REFRESH_FC).  One launch.  `which` selects the groups (0 = all): a training loop
 * whose optimizer updates v.weight / fc.weight before the query gradient exists refreshes them early, on another
 * stream.  fc_w / lin_workspace may be NULL (pooling head only). */
#define EP_REFRESH_QUERIES 1
#define EP_REFRESH_VALUE   2
#define EP_REFRESH_FC      4
#define EP_REFRESH_ALL     7
int ep_refresh_operands(const float* cls_token, const float* v_w, float scale, int x_dtype, int B, int N, int D, int M,
                        int d_out, void* workspace, size_t workspace_bytes, const float* fc_w, int K,
                        void* lin_workspace, size_t lin_workspace_bytes, int which, void* stream);
int ep_fwd_ops(const void* x, int x_dtype, const float* cls_token, const float* v_w, const float* v_b,
               float scale, int B, int N, int D, int M, int d_out,
               float* out, float* S, float* rowmax, float* rowsum, float* P, float* attn,
               void* workspace, size_t workspace_bytes, int ops, void* stream);
int ep_bwd_proj_ops(const float* g_out, const float* P, const float* out, const float* v_w, const float* v_b, int x_dtype,
                    int B, int N, int D, int M, int d_out, float* d_v_w, float* d_v_b,
                    void* workspace, size_t workspace_bytes, int ops, void* stream);
/* ep_bn_fwd + the operand copy of y for a Linear(F, K) that follows (lin_workspace of
 * ep_linear_workspace_bytes(B, F, K); NULL = plain ep_bn_fwd). */
int ep_bn_fwd_ops(const float* h, int B, int F, float eps, float momentum, int training,
                  float* running_mean, float* running_var, long long* num_batches_tracked,
                  float* y, float* save_mean, float* save_invstd, int K, void* lin_workspace, size_t lin_workspace_bytes,
                  int ops, void* stream);
/* ep_bn_bwd where dh is the g_out of ep_bwd_proj_ops: also leaves in `workspace` the operand copies of dh and
 * delta = dh . (out - v_b) per (sample, query) -- `out`, `v_b` as in ep_bwd. */
int ep_bn_bwd_ops(const float* dy, const float* y, const float* save_invstd, int B, int F, float* dh,
                  const float* out, const float* v_b, int x_dtype, int N, int D, int M, int d_out,
                  void* workspace, size_t workspace_bytes, int ops, void* stream);
int ep_linear_fwd_ops(const float* y, const float* W, const float* b, int B, int F, int K, float* logits,
                      void* lin_workspace, size_t lin_workspace_bytes, int ops, void* stream);
int ep_linear_bwd_ops(const float* dlogits, const float* y, const float* W, int B, int F, int K,
                      float* dW, float* db, float* dy, void* lin_workspace, size_t lin_workspace_bytes, int ops,
                      void* stream);
/* ep_ce_fwd_bwd with a deterministic loss and no zeroing by the caller: step_loss[0] = sum_b nll_b * loss_scale is
 * OVERWRITTEN (rows summed in a fixed order by the last CTA to finish), loss_acc[0] += that value when non-NULL (a
 * running meter).  scratch: B + 1 floats; word B is a counter that must be 0 before the first call and is left 0.
 * With lin_workspace (and F of the Linear(F, K) whose backward follows) the operand copy of dlogits is written too. */
int ep_ce_fwd_bwd_ops(const float* logits, const long long* targets, int B, int K, float loss_scale, float grad_scale,
                      float* step_loss, float* loss_acc, float* dlogits, int* correct, float* scratch,
                      int F, void* lin_workspace, size_t lin_workspace_bytes, int ops, void* stream);

/* The other two optimizers main_linprobe.py:403-408 can build ({"lars": LARS, "adamw": AdamW}, else SGD), one
 * launch for all tensors; pointer tables are HOST arrays of device pointers, hyper is a DEVICE array.
 * ep_adamw_step -- torch.optim.AdamW: hyper = {lr, beta1, beta2, eps, weight_decay, grad_scale,
 *                  1 - beta1^t, 1 - beta2^t} (the two bias corrections of step t, computed by the host).
 * ep_sgd_step   -- torch.optim.SGD (dampening 0, no nesterov): hyper = {lr, weight_decay, momentum, grad_scale,
 *                  first_step != 0}; momentum_buf_host may be NULL when momentum == 0. */
int ep_adamw_step(int n, float* const* params_host, const float* const* grads_host, float* const* exp_avg_host,
                  float* const* exp_avg_sq_host, const long long* numels_host, const float* hyper, void* stream);
int ep_sgd_step(int n, float* const* params_host, const float* const* grads_host, float* const* momentum_buf_host,
                const long long* numels_host, const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EP_B200_H_ */
