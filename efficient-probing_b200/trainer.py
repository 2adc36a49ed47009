"""EPHeadTrainer -- the reference's per-batch hot loop (engine_finetune.py:40-91) for an EP probe head
trained on cached tokens, restructured for a 0.2 ms step:

    tokens -> ep_fwd -> BatchNorm1d -> Linear -> CrossEntropy(+top-1) -> backward of all of it
           -> gradient all-reduce (NCCL, flat buffer) -> LARS

Every stage is one or two launches of libep_b200 on one stream; nothing synchronises with the host
(the reference's per-step ``loss.item()``, ``acc.item()``, ``cuda.synchronize()`` and scalar all-reduce,
engine_finetune.py:64,66,79-80,91, become device-side accumulators read when the caller asks), and the
whole step is captured into a CUDA graph.  The parameters updated are the tensors of the
``nn.Sequential(EfficientProbing, BatchNorm1d, Linear)`` passed in, in place, so ``head.state_dict()``
stays the checkpoint format of util/misc.py:304-332.

Multi-GPU (main_linprobe.py:581-583 DDP semantics): one process per GPU, batch sharded, BatchNorm
statistics per GPU (no SyncBN), gradients summed with one NCCL all-reduce over a flat fp32 buffer
``[fc.weight, fc.bias, v.weight, (v.bias), cls_token]`` and averaged inside the LARS kernel.  The first
four become ready before the token-streaming half of the backward (ep_bwd_pool), so their all-reduce is
issued on a communication stream underneath it; cls_token's 0.1-0.5 MB follows."""
from typing import Optional

import os

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from .ep import EfficientProbing
from .flatgrad import FlatGradLayout, allreduce_sum_
from .optim import lars_launch, adamw_launch, sgd_launch

BN_EPS_DEFAULT = 1e-6


def default_comm_sms(world: int) -> int:
    """SMs the streaming backward kernels leave to the overlapped all-reduce (see EPHeadTrainer, comm_sms)."""
    return 16 if world >= 2 else 0       # measured at c2: 2 GPUs 0.696 (0) / 0.678 (8) / 0.673 ms (16); 8 GPUs 0.704 ms (16)


class EPHeadTrainer:
    def __init__(self, head: nn.Sequential, batch_size: int, num_tokens: int, *, lr: float = 0.1,
                 weight_decay: float = 0.0, momentum: float = 0.9, trust_coefficient: float = 0.001,
                 x_dtype: torch.dtype = torch.bfloat16, process_group=None, use_graph: bool = True,
                 overlap_comm: bool = True, optimizer: str = "lars", betas=(0.9, 0.999), eps: float = 1e-8,
                 accum_iter: int = 1, comm_sms: Optional[int] = None, broadcast_buffers: str = "eval",
                 fuse_operands: Optional[bool] = None):
        """comm_sms: SMs left free for the overlapped gradient all-reduce while the token-streaming half of the
        backward pass runs (multi-GPU only; pair it with NCCL_MAX_CTAS <= comm_sms set before the process group is
        created -- bench.py does -- so the collective's CTAs fit there).  None = default_comm_sms(world): measured
        on B200/NVSwitch, c2: 8 GPUs 0.827 ms/step with 0, 0.789 with 8, 0.767 with 16; 2 GPUs are fastest with 0.
        optimizer: "lars" (util/lars.py, the published protocol), "adamw" (torch.optim.AdamW defaults) or "sgd"
        (torch.optim.SGD, momentum as given) -- the three main_linprobe.py:403-408 can build.
        accum_iter: gradient accumulation as engine_finetune.py:72-77 (loss / accum_iter, optimizer every k-th call).
        Multi-GPU replicas are made identical here, as DistributedDataParallel's constructor does
        (main_linprobe.py:581-583): rank 0's parameters and BatchNorm buffers are broadcast to every rank (the
        reference seeds each rank with seed + rank, main_linprobe.py:517, and relies on exactly that).
        broadcast_buffers: DDP(broadcast_buffers=True) replaces every rank's BatchNorm running statistics by rank 0's
        before each forward (batch statistics stay per GPU: no SyncBN).  In training mode nothing reads them but
        their own update, so the only observable effect is which statistics evaluation and checkpoints see:
        "eval" (default) broadcasts rank 0's before eval_logits / state export -- same predictions, no per-step
        message; "step" also broadcasts before every training step, as DDP literally does; "off" never."""
        # fuse_operands (default on; EP_FUSE_OPERANDS=0 turns it off): the bf16 hi/lo operand copies the tcgen05 GEMMs
        # read are written by the kernels that produce the data -- weights once per step by ep_refresh_operands right
        # after the optimizer, activations by BatchNorm / cross-entropy / BatchNorm-backward -- instead of by eight
        # extra launches (the *_ops entry points of include/ep_b200.h).  Same kernels, same operand bits, same results.
        if fuse_operands is None:
            fuse_operands = os.environ.get("EP_FUSE_OPERANDS", "1") != "0"
        self.fuse_ops = bool(fuse_operands)
        self._ops_key = None                               # parameter versions the weight copies were made from
        self.parallel_dw = os.environ.get("EP_PARALLEL_DW", "1") != "0"
        # single GPU: v.weight / fc gradients are final after part 1, so their optimizer update and operand refresh run on
        # the side branch underneath the token-streaming backward; only cls_token's follow it
        # (measured at c2: 0.598 ms against 0.592 ms without -- nothing can share an SM with the one-pass kernels, so the
        # early update only delays them; mode 2 forks after the streaming backward instead.  Off by default.)
        self.split_update = int(os.environ.get("EP_SPLIT_UPDATE", "0"))
        if broadcast_buffers not in ("eval", "step", "off"):
            raise ValueError("broadcast_buffers must be 'eval', 'step' or 'off'")
        if optimizer not in ("lars", "adamw", "sgd"):
            raise ValueError("optimizer must be 'lars', 'adamw' or 'sgd'")
        self.optimizer, self.betas, self.eps_opt, self.accum_iter = optimizer, betas, eps, max(1, int(accum_iter))
        pool, bn, fc = head[0], head[1], head[2]
        if not isinstance(pool, EfficientProbing) or not isinstance(bn, nn.BatchNorm1d) or not isinstance(fc, nn.Linear):
            raise TypeError("head must be Sequential(EfficientProbing, BatchNorm1d, Linear) (probe_heads.py:106)")
        if bn.affine:
            raise ValueError("the probe's BatchNorm1d is affine=False (probe_heads.py:110)")
        self.head, self.pool, self.bn, self.fc = head, pool, bn, fc
        self.lib = _lib.load()
        dev = pool.cls_token.device
        _lib.require_cuda(pool.cls_token, "head parameters")
        _lib.check(self.lib.ep_device_check(), "ep_device_check")
        self.dev = dev
        self.B, self.N = int(batch_size), int(num_tokens)
        self.D = pool.cls_token.shape[2]
        self.M = pool.num_queries
        self.d_out = pool.d_out
        self.Dp = self.D // self.d_out
        self.K = fc.out_features
        self.x_dtype = x_dtype
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.use_graph = use_graph
        self.overlap_comm = overlap_comm and self.world > 1
        self.comm_sms = (default_comm_sms(self.world) if comm_sms is None else int(comm_sms)) if self.overlap_comm else 0
        self.broadcast_buffers = broadcast_buffers if self.world > 1 else "off"

        f32 = dict(dtype=torch.float32, device=dev)
        B, N, D, M, Dp, K = self.B, self.N, self.D, self.M, self.Dp, self.K
        # parameters, in the reference's parameters() order: cls_token, v.weight, [v.bias], fc.weight, fc.bias
        self.params = [pool.cls_token, pool.v.weight] + ([pool.v.bias] if pool.v.bias is not None else []) + \
                      [fc.weight, fc.bias]
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("head parameters must be contiguous fp32")
        self.trust = [p.ndim > 1 for p in self.params]                                   # util/lars.py:21
        # flat gradient buffer: large, early-ready tensors first, cls_token last (flatgrad.py)
        self.layout = FlatGradLayout(K, Dp, D, M, pool.v.bias is not None)
        self.flat_grad = self.layout.allocate(dev)
        self.g = self.layout.views(self.flat_grad)
        self.n_early = self.layout.early
        self.grads = [self.g["cls"], self.g["v_w"]] + ([self.g["v_b"]] if pool.v.bias is not None else []) + \
                     [self.g["fc_w"], self.g["fc_b"]]
        self.mus = [torch.zeros_like(p) for p in self.params]           # LARS mu / SGD momentum_buffer / AdamW exp_avg
        self.sq = [torch.zeros_like(p) for p in self.params] if optimizer == "adamw" else None   # AdamW exp_avg_sq
        self.opt_steps = 0
        self.accum = torch.zeros_like(self.flat_grad) if self.accum_iter > 1 else None
        if self.accum is not None:
            av = self.layout.views(self.accum)
            self._accum_grads = [av["cls"], av["v_w"]] + ([av["v_b"]] if pool.v.bias is not None else []) + \
                                [av["fc_w"], av["fc_b"]]
        self.micro = 0
        # activations
        self.x = torch.empty(B, N, D, dtype=x_dtype, device=dev)
        self.targets = torch.empty(B, dtype=torch.int64, device=dev)
        self.out = torch.empty(B, Dp, **f32)
        self.rowmax = torch.empty(B, M, **f32)
        self.rowsum = torch.empty(B, M, **f32)
        self.P = torch.empty(B, M, D, **f32)
        self.S = torch.empty(B, M, N, **f32)               # logits, saved for the backward pass
        self.y = torch.empty(B, Dp, **f32)
        self.save_mean = torch.empty(Dp, **f32)
        self.save_invstd = torch.empty(Dp, **f32)
        self.logits = torch.empty(B, K, **f32)
        self.dlogits = torch.empty(B, K, **f32)
        self.dy = torch.empty(B, Dp, **f32)
        self.dout = torch.empty(B, Dp, **f32)
        self.loss_sum = torch.zeros(1, **f32)            # running sum of per-step mean losses
        self.step_loss = torch.zeros(1, **f32)           # mean loss of the last step
        self.correct = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lr, self._wd, self._mom, self._trust = float(lr), float(weight_decay), float(momentum), float(trust_coefficient)
        self.hyper_host = torch.zeros(8)                   # pageable on purpose: the H2D copy stages it at call time,
        self.hyper = torch.zeros(8, **f32)                 # so rewriting it for the next step cannot race the GPU
        self._write_hyper()
        self.lars_scratch = torch.empty(8192, **f32)       # EP_LARS_SCRATCH_FLOATS
        self.lars_scratch2 = torch.empty(8192, **f32)      # (the early group's, when the update runs as two groups)
        self.ce_scratch = torch.zeros(B + 8, **f32)        # per-sample losses + the counter word of ep_ce_fwd_bwd_ops
        self.ws = torch.empty(max(16, self.lib.ep_workspace_bytes(B, N, D, M, self.d_out)), dtype=torch.uint8, device=dev)
        self.lin_ws = torch.empty(max(16, self.lib.ep_linear_workspace_bytes(B, Dp, K)), dtype=torch.uint8, device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev) if self.overlap_comm else None
        self._ev_early = torch.cuda.Event() if self.overlap_comm else None
        self.two_group = os.environ.get("EP_TWO_GROUP_UPDATE", "0") != "0"   # measured at 2 GPUs, c2: 0.645 ms against 0.631 ms as one group (a fourth graph launch costs more than the hidden 131 KB all-reduce)
        # BatchNorm running statistics as two views of one flat tensor, so that DDP's buffer broadcast is one message
        self.bn_flat = torch.cat([bn.running_mean.detach().reshape(-1), bn.running_var.detach().reshape(-1)]).contiguous()
        bn.running_mean.data = self.bn_flat[:Dp]
        bn.running_var.data = self.bn_flat[Dp:]
        if self.world > 1:
            self.sync_replicas()
        self.side_stream = torch.cuda.Stream(device=dev)
        self._cx, self._ct = self.x, self.targets           # the batch the next launch sequence reads
        self._registered = {}                               # (x ptr, targets ptr) -> (x, targets) kept alive
        self.graphs = {}                                    # (x ptr, targets ptr) -> captured step
        self.steps = 0
        self.launches_per_step = None
        # pinned staging for the host-buffer API
        self._hx = self._ht = self._hloss = None

    @torch.no_grad()
    def sync_replicas(self):
        """Rank 0's parameters, BatchNorm buffers and optimizer state -> every rank (DDP's constructor broadcast;
        call it again after load_state_dict / load_optimizer_state_dict on rank 0 alone)."""
        if self.world == 1:
            return
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        for t in self.params + self.mus + (self.sq or []) + [self.bn_flat, self.bn.num_batches_tracked]:
            dist.broadcast(t.data, src=src, group=self.group)
        self.invalidate_operands()
        steps = torch.tensor([self.opt_steps], dtype=torch.int64, device=self.dev)
        dist.broadcast(steps, src=src, group=self.group)
        self.opt_steps = int(steps.item())
        if self.optimizer != "lars":
            self._write_hyper()

    def _broadcast_buffers(self):
        """Rank 0's BatchNorm running statistics -> every rank (main_linprobe.py:582, DDP's buffer broadcast)."""
        if self.world == 1 or self.broadcast_buffers == "off":
            return
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast(self.bn_flat, src=src, group=self.group)

    # ------------------------------------------------------------------ one step, stream-ordered
    # ------------------------------------------------------------------ operand copies (ABI 2)
    def invalidate_operands(self):
        """The weight-derived operand copies are stale: call this after changing a parameter through ``.data`` or a raw
        pointer (in-place tensor ops, ``load_state_dict`` included, are noticed through the tensors' version counters)."""
        self._ops_key = None

    def _refresh_launch(self, which: int = 0):
        """ep_refresh_operands on the current stream (captured into the step's graph right after the optimizer);
        which: 0 = every copy, 1 = the queries', 6 = v.weight's and fc.weight's."""
        pool, fc = self.pool, self.fc
        _lib.check(self.lib.ep_refresh_operands(pool.cls_token.data_ptr(), pool.v.weight.data_ptr(), float(pool.scale),
                                                _lib.x_dtype_code(self.x), self.B, self.N, self.D, self.M, self.d_out,
                                                self.ws.data_ptr(), self.ws.numel(), fc.weight.data_ptr(), self.K,
                                                self.lin_ws.data_ptr(), self.lin_ws.numel(), which,
                                                _lib.stream_ptr(self.dev)), "ep_refresh_operands")

    def _ops_check(self):
        """Make the weight copies current if a parameter changed behind the trainer's back since they were written."""
        if not self.fuse_ops:
            return
        key = tuple((p.data_ptr(), p._version) for p in self.params)
        if key != self._ops_key:
            self._refresh_launch()
            self._ops_key = key

    def _forward(self, training: bool, fp32: bool = False):
        lib, s = self.lib, _lib.stream_ptr(self.dev)
        B, N, D, M, Dp, K = self.B, self.N, self.D, self.M, self.Dp, self.K
        pool, bn, fc = self.pool, self.bn, self.fc
        Wf, If = (_lib.EP_OPS_WEIGHTS, _lib.EP_OPS_INPUT) if self.fuse_ops else (0, 0)
        mode = _lib.EP_OPS_FP32 if fp32 else 0                 # evaluation: fp32 contractions, for this call only
        lin_ws, lin_n = (self.lin_ws.data_ptr(), self.lin_ws.numel())
        _lib.check(lib.ep_fwd_ops(self._cx.data_ptr(), _lib.x_dtype_code(self._cx), pool.cls_token.data_ptr(),
                                  pool.v.weight.data_ptr(), _lib.ptr(pool.v.bias), float(pool.scale), B, N, D, M, self.d_out,
                                  self.out.data_ptr(), self.S.data_ptr(), self.rowmax.data_ptr(), self.rowsum.data_ptr(),
                                  self.P.data_ptr(), None, self.ws.data_ptr(), self.ws.numel(), Wf | mode, s), "ep_fwd")
        _lib.check(lib.ep_bn_fwd_ops(self.out.data_ptr(), B, Dp, float(bn.eps), float(bn.momentum), int(training),
                                     bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                     bn.num_batches_tracked.data_ptr(), self.y.data_ptr(), self.save_mean.data_ptr(),
                                     self.save_invstd.data_ptr(), K, lin_ws if self.fuse_ops else None, lin_n, mode, s),
                   "ep_bn_fwd")
        _lib.check(lib.ep_linear_fwd_ops(self.y.data_ptr(), fc.weight.data_ptr(), fc.bias.data_ptr(), B, Dp, K,
                                         self.logits.data_ptr(), lin_ws, lin_n, Wf | If | mode, s), "ep_linear_fwd")

    # The step in three stream-ordered parts; the gradient exchange sits between them.
    def _part1(self, join: bool = True):
        """forward, loss, and every gradient that does not need the tokens again"""
        lib, s = self.lib, _lib.stream_ptr(self.dev)
        B, N, D, M, Dp, K = self.B, self.N, self.D, self.M, self.Dp, self.K
        pool, fc = self.pool, self.fc
        self._forward(training=True)
        Wf, If = (_lib.EP_OPS_WEIGHTS, _lib.EP_OPS_INPUT) if self.fuse_ops else (0, 0)
        lin_ws, lin_n = self.lin_ws.data_ptr(), self.lin_ws.numel()
        xdt = _lib.x_dtype_code(self._cx)
        # loss (overwrites step_loss, adds it to the running meter, deterministic), dlogits and -- fused -- their operand copy
        _lib.check(lib.ep_ce_fwd_bwd_ops(self.logits.data_ptr(), self._ct.data_ptr(), B, K, 1.0 / B, 1.0 / B,
                                         self.step_loss.data_ptr(), self.loss_sum.data_ptr(), self.dlogits.data_ptr(),
                                         self.correct.data_ptr(), self.ce_scratch.data_ptr(), Dp,
                                         lin_ws if self.fuse_ops else None, lin_n, 0, s), "ep_ce_fwd_bwd")
        # classifier gradients: dW/db need only (dlogits, y) and nothing downstream needs them before the
        # exchange, so they run on a side stream (a parallel branch of the captured graph) next to the dy chain
        cur = torch.cuda.current_stream(self.dev)
        self.side_stream.wait_stream(cur)
        with torch.cuda.stream(self.side_stream):
            _lib.check(lib.ep_linear_bwd_ops(self.dlogits.data_ptr(), self.y.data_ptr(), fc.weight.data_ptr(), B, Dp, K,
                                             self.g["fc_w"].data_ptr(), self.g["fc_b"].data_ptr(), None,
                                             lin_ws if self.fuse_ops else None, lin_n if self.fuse_ops else 0, If,
                                             _lib.stream_ptr(self.dev)), "ep_linear_bwd (dW, db)")
        _lib.check(lib.ep_linear_bwd_ops(self.dlogits.data_ptr(), self.y.data_ptr(), fc.weight.data_ptr(), B, Dp, K,
                                         None, None, self.dy.data_ptr(), lin_ws, lin_n, Wf | If, s), "ep_linear_bwd (dy)")
        d_vb = self.g["v_b"].data_ptr() if pool.v.bias is not None else None
        if self.fuse_ops:        # BatchNorm backward + delta and the operand copies of its result for ep_bwd_proj
            _lib.check(lib.ep_bn_bwd_ops(self.dy.data_ptr(), self.y.data_ptr(), self.save_invstd.data_ptr(), B, Dp,
                                         self.dout.data_ptr(), self.out.data_ptr(), _lib.ptr(pool.v.bias), xdt, N, D, M,
                                         self.d_out, self.ws.data_ptr(), self.ws.numel(), 0, s), "ep_bn_bwd")
        else:
            _lib.check(lib.ep_bn_bwd(self.dy.data_ptr(), self.y.data_ptr(), self.save_invstd.data_ptr(), B, Dp,
                                     self.dout.data_ptr(), s), "ep_bn_bwd")
        def bwd_proj(ops, stream):
            _lib.check(lib.ep_bwd_proj_ops(self.dout.data_ptr(), self.P.data_ptr(), self.out.data_ptr(),
                                           pool.v.weight.data_ptr(), _lib.ptr(pool.v.bias), xdt, B, N, D, M, self.d_out,
                                           self.g["v_w"].data_ptr(), d_vb, self.ws.data_ptr(), self.ws.numel(), ops, stream),
                       "ep_bwd_proj")
        if self.fuse_ops and self.parallel_dw:
            # d v.weight (reads the 134 MB of P) is not needed before the exchange / the optimizer: it joins the classifier
            # gradients on the side branch and runs next to dP (134 MB written) instead of in front of it
            self.side_stream.wait_stream(cur)
            with torch.cuda.stream(self.side_stream):
                bwd_proj(Wf | If | _lib.EP_OPS_ONLY_DW, _lib.stream_ptr(self.dev))
            bwd_proj(Wf | If | _lib.EP_OPS_NO_DW, s)
        else:
            bwd_proj(Wf | If, s)
        if join:
            cur.wait_stream(self.side_stream)

    def _part2(self):
        """token-streaming half of the backward pass: d cls_token"""
        lib, s = self.lib, _lib.stream_ptr(self.dev)
        pool = self.pool
        if self.comm_sms > 0:          # the early all-reduce runs underneath: leave it SMs (statically scheduled CTAs
            lib.ep_set_sm_limit(max(1, torch.cuda.get_device_properties(self.dev).multi_processor_count - self.comm_sms))
        try:
            self._bwd_pool(lib, s, pool)
        finally:
            if self.comm_sms > 0:      # that land behind a collective's CTA would stretch the kernel by its duration)
                lib.ep_set_sm_limit(0)

    def _bwd_pool(self, lib, s, pool):
        _lib.check(lib.ep_bwd_pool(self._cx.data_ptr(), _lib.x_dtype_code(self._cx), pool.cls_token.data_ptr(),
                                   float(pool.scale), self.B, self.N, self.D, self.M, self.d_out, self.S.data_ptr(),
                                   self.rowmax.data_ptr(), self.rowsum.data_ptr(), self.g["cls"].data_ptr(),
                                   self.ws.data_ptr(), self.ws.numel(), s), "ep_bwd_pool")

    def _part3(self):
        if self.accum is not None:                         # engine_finetune.py:72-77
            self.accum.add_(self.flat_grad)
            return
        self._apply_optimizer(self.grads)
        if self.fuse_ops:                                  # operand copies of the updated weights for the next step
            self._refresh_launch()

    def _update_group(self, early: bool):
        """Optimizer + operand refresh for one group of parameters: early = all but cls_token (gradients final after
        part 1), late = cls_token (its gradient comes out of the token-streaming backward)."""
        idx = list(range(1, len(self.params))) if early else [0]
        self._apply_optimizer(self.grads, idx, self.lars_scratch2 if early else self.lars_scratch)
        if self.fuse_ops:
            self._refresh_launch(6 if early else 1)

    def _apply_optimizer(self, grads, idx=None, scratch=None):
        pick = (lambda l: l) if idx is None else (lambda l: [l[i] for i in idx])
        if self.optimizer == "lars":
            lars_launch(pick(self.params), pick(grads), pick(self.mus), pick(self.trust), self.hyper,
                        self.lars_scratch if scratch is None else scratch)
        elif self.optimizer == "adamw":
            adamw_launch(pick(self.params), pick(grads), pick(self.mus), pick(self.sq), self.hyper)
        else:
            sgd_launch(pick(self.params), pick(grads), pick(self.mus) if self._mom != 0.0 else None, self.hyper)

    def _write_hyper(self):
        """Host scalars of the next optimizer step -> pinned -> device (stream-ordered, no sync)."""
        gs = 1.0 / (self.world * self.accum_iter)
        h = self.hyper_host
        if self.optimizer == "lars":
            h[:5] = torch.tensor([self._lr, self._wd, self._mom, self._trust, gs])
        elif self.optimizer == "adamw":
            t = self.opt_steps + 1
            h[:8] = torch.tensor([self._lr, self.betas[0], self.betas[1], self.eps_opt, self._wd, gs,
                                  1.0 - self.betas[0] ** t, 1.0 - self.betas[1] ** t], dtype=torch.float64).float()
        else:
            h[:5] = torch.tensor([self._lr, self._wd, self._mom, gs, 1.0 if self.opt_steps == 0 else 0.0])
        self.hyper.copy_(h)

    def _after_run(self):
        """Host-side bookkeeping after one (micro-)step has been enqueued."""
        self.micro += 1
        if self.accum is not None:
            if self.micro % self.accum_iter == 0:          # engine_finetune.py:73-77: step + zero_grad every k-th call
                self._apply_optimizer(self._accum_grads)
                if self.fuse_ops:
                    self._refresh_launch()
                self.accum.zero_()
                self.opt_steps += 1
                if self.optimizer != "lars":
                    self._write_hyper()
        else:
            self.opt_steps += 1
            if self.optimizer != "lars":                   # bias corrections / first-step flag change every step
                self._write_hyper()

    def _exchange_early(self):
        """[fc, v] gradients are final after part 1: reduce them underneath part 2 on the comm stream"""
        if self.world == 1:
            return
        if self.overlap_comm:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.comm_stream):
                allreduce_sum_(self.flat_grad, self.group, 0, self.n_early)

    def _exchange_rest(self):
        if self.world == 1:
            return
        if self.overlap_comm:
            allreduce_sum_(self.flat_grad, self.group, self.n_early)
            torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)
        else:
            allreduce_sum_(self.flat_grad, self.group)

    # Multi-GPU with overlap: the update runs as two groups so that cls_token's all-reduce (131 KB: pure latency, and
    # the only exposed collective of the step) hides behind the optimizer update and operand refresh of the early group:
    #   comm stream:  all-reduce [fc, v]  (under part 2) ............ all-reduce cls_token
    #   main stream:  part 2 (token-streaming backward) | update early group | (join) update cls_token
    def _two_group_update(self):
        return self.world > 1 and self.overlap_comm and self.accum is None and self.two_group

    def _exchange_rest_async(self):
        """cls_token's all-reduce on the comm stream (behind the early one), main does not wait yet"""
        cur = torch.cuda.current_stream(self.dev)
        self._ev_early.record(self.comm_stream)               # the early all-reduce is what main needs first
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            allreduce_sum_(self.flat_grad, self.group, self.n_early)
        cur.wait_event(self._ev_early)

    def _join_comm(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)

    def _upd_early(self):
        self._update_group(early=True)

    def _upd_late(self):
        self._update_group(early=False)

    def _step_body(self):
        if self.world == 1 and self.accum is None and self.split_update == 2:
            self._part1()
            self._part2()
            cur = torch.cuda.current_stream(self.dev)
            self.side_stream.wait_stream(cur)
            with torch.cuda.stream(self.side_stream):
                self._update_group(early=True)
            self._update_group(early=False)
            cur.wait_stream(self.side_stream)
            return
        if self.world == 1 and self.accum is None and self.split_update == 1:
            # one GPU: the early group's update rides on the side branch (behind the weight-gradient GEMMs it depends on),
            # next to dP and the token-streaming backward; main joins it at the end of the step
            self._part1(join=False)
            # (the side branch first waits for main's dP GEMM: it reads the weight copies this update rewrites)
            self.side_stream.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.side_stream):
                self._update_group(early=True)
            self._part2()
            self._update_group(early=False)
            torch.cuda.current_stream(self.dev).wait_stream(self.side_stream)
            return
        self._part1()
        self._exchange_early()
        self._part2()
        if self._two_group_update():
            self._exchange_rest_async()
            self._upd_early()
            self._join_comm()
            self._upd_late()
            return
        self._exchange_rest()
        self._part3()

    def _ensure_graphs(self):
        """Captured graphs of the step for the current batch slot (built on first use)."""
        key = (self._cx.data_ptr(), self._ct.data_ptr())
        graphs = self.graphs.get(key)
        if graphs is None:
            if not self.graphs:
                # one eager step on a side stream first (lazy module loading, NCCL communicator set-up)
                snap = self._snapshot()
                self._ops_check()
                side = torch.cuda.Stream(device=self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(side):
                    n0 = self.lib.ep_launch_count()
                    self._step_body()
                    self.launches_per_step = int(self.lib.ep_launch_count() - n0)
                torch.cuda.current_stream(self.dev).wait_stream(side)
                torch.cuda.synchronize(self.dev)
                self._restore(snap)
            # single GPU: the whole step is one graph.  Multi-GPU: the NCCL all-reduces stay outside the
            # graphs (three graphs per step with the two exchanges between them) -- capturing the
            # collectives inside a graph deadlocked on this stack, and the eager calls cost microseconds.
            one_graph = self.world == 1 or os.environ.get("EP_GRAPH_NCCL", "0") == "1"   # experimental: NCCL captured too
            parts = [self._step_body] if one_graph else \
                    ([self._part1, self._part2, self._upd_early, self._upd_late] if self._two_group_update() else
                     [self._part1, self._part2, self._part3])
            graphs = []
            for part in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    part()
                graphs.append(g)
            self.graphs[key] = graphs
        return graphs

    def _run(self):
        if not self.use_graph:
            self._ops_check()
            if self.launches_per_step is None:
                n0 = self.lib.ep_launch_count()
                self._step_body()
                self.launches_per_step = int(self.lib.ep_launch_count() - n0)
            else:
                self._step_body()
            return
        graphs = self._ensure_graphs()
        self._ops_check()                                   # (after the capture's warm-up step has been rolled back)
        if len(graphs) == 1:
            graphs[0].replay()
        elif len(graphs) == 4:
            graphs[0].replay()
            self._exchange_early()
            graphs[1].replay()
            self._exchange_rest_async()
            graphs[2].replay()
            self._join_comm()
            graphs[3].replay()
        else:
            graphs[0].replay()
            self._exchange_early()
            graphs[1].replay()
            self._exchange_rest()
            graphs[2].replay()

    @torch.no_grad()
    def prepare(self, x: torch.Tensor, targets: torch.Tensor):
        """Set-up for a registered batch slot without training on it: build (capture) the step's CUDA graph now, so the
        first train_step on it is already a replay.  Parameters, optimizer state and meters are left untouched."""
        key = (x.data_ptr(), targets.data_ptr())
        if key not in self._registered:
            self.register_batch(x, targets)
        if not self.use_graph:
            return
        self._cx, self._ct = self._registered[key]
        self._ensure_graphs()
        self._ops_check()

    def _snapshot(self):
        bn = self.bn
        return ([p.detach().clone() for p in self.params], [m.clone() for m in self.mus],
                [q.clone() for q in self.sq] if self.sq is not None else None,
                self.accum.clone() if self.accum is not None else None, bn.running_mean.clone(),
                bn.running_var.clone(), bn.num_batches_tracked.clone(), self.loss_sum.clone(), self.correct.clone())

    def _restore(self, snap):
        ps, ms, sq, acc, rm, rv, nbt, ls, cr = snap
        with torch.no_grad():
            for p, q in zip(self.params, ps):
                p.copy_(q)
            for m, q in zip(self.mus, ms):
                m.copy_(q)
            if sq is not None:
                for m, q in zip(self.sq, sq):
                    m.copy_(q)
            if acc is not None:
                self.accum.copy_(acc)
            self.bn.running_mean.copy_(rm)
            self.bn.running_var.copy_(rv)
            self.bn.num_batches_tracked.copy_(nbt)
            self.loss_sum.copy_(ls)
            self.correct.copy_(cr)

    # ------------------------------------------------------------------ public API
    def set_lr(self, lr: float, weight_decay: Optional[float] = None):
        """Host scalar from the schedule (util/lr_sched.py) -> device, without a sync."""
        self._lr = float(lr)
        if weight_decay is not None:
            self._wd = float(weight_decay)
        self._write_hyper()

    @torch.no_grad()
    def train_step(self, x: torch.Tensor, targets: torch.Tensor, lr: Optional[float] = None):
        """One optimisation step on a device-resident batch: x (B, N, D), targets (B,) int64.
        Returns nothing; read ``mean_loss()`` / ``top1()`` when needed."""
        if x.shape != self.x.shape or x.dtype != self.x_dtype:
            raise ValueError(f"x must be {tuple(self.x.shape)} {self.x_dtype}, got {tuple(x.shape)} {x.dtype}")
        if lr is not None:
            self.set_lr(lr)
        key = (x.data_ptr(), targets.data_ptr())
        if key in self._registered:
            self._cx, self._ct = self._registered[key]        # resident batch: read it where it lies
        else:
            self._cx, self._ct = self.x, self.targets
            if x.data_ptr() != self.x.data_ptr():
                self.x.copy_(x, non_blocking=True)
            if targets.data_ptr() != self.targets.data_ptr():
                self.targets.copy_(targets, non_blocking=True)
        if self.broadcast_buffers == "step":
            self._broadcast_buffers()
        self._run()
        self._after_run()
        self.steps += 1

    def register_batch(self, x: torch.Tensor, targets: torch.Tensor):
        """Declare a device-resident batch (e.g. one slot of a token-cache pool) that train_step may read in
        place, without the staging copy; each registered batch gets its own captured graph."""
        if x.shape != self.x.shape or x.dtype != self.x_dtype or not x.is_contiguous() or x.device != self.dev:
            raise ValueError("registered batch must match the trainer's (B, N, D), dtype and device")
        if targets.dtype != torch.int64 or targets.shape != (self.B,):
            raise ValueError("targets must be (B,) int64")
        self._registered[(x.data_ptr(), targets.data_ptr())] = (x, targets)

    @torch.no_grad()
    def train_step_host(self, x_host: torch.Tensor, targets_host: torch.Tensor, lr: Optional[float] = None,
                        next_x_host: Optional[torch.Tensor] = None,
                        next_targets_host: Optional[torch.Tensor] = None) -> float:
        """End-to-end step from HOST buffers: pinned-memory H2D copy of tokens and labels, the step, and a
        D2H read of the step's mean loss (the reference loop's samples.to(device) ... loss.item()).

        ``next_*``: the following step's host batch, if known -- its H2D copy is issued on a copy stream into
        the other device slot right after this step is launched, so the 0.5 GB transfer of step i+1 overlaps
        the kernels of step i (every batch still crosses PCIe exactly once, inside the caller's loop)."""
        if self._hloss is None:
            self._hloss = torch.empty(1, dtype=torch.float32).pin_memory()
            self._h2d_stream = torch.cuda.Stream(device=self.dev)
            self._slots = [(self.x, self.targets),
                           (torch.empty_like(self.x), torch.empty_like(self.targets))]
            for sx, st in self._slots:
                self.register_batch(sx, st)
            self._slot_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._prefetched = None           # (slot index, x_host data_ptr)
            self._cur_slot = 0

        def pin(xh, th):
            if xh.is_pinned() and th.is_pinned():
                return xh, th
            if self._hx is None:
                self._hx = torch.empty(self.x.shape, dtype=self.x_dtype).pin_memory()
                self._ht = torch.empty(self.B, dtype=torch.int64).pin_memory()
            self._hx.copy_(xh)
            self._ht.copy_(th)
            return self._hx, self._ht

        cur = torch.cuda.current_stream(self.dev)
        if lr is not None:
            self.set_lr(lr)
        if self._prefetched is not None and self._prefetched[1] == x_host.data_ptr():
            slot = self._prefetched[0]                       # already on its way: wait for the copy only
            cur.wait_event(self._slot_ready[slot])
        else:
            slot = self._cur_slot
            xh, th = pin(x_host, targets_host)
            self._slots[slot][0].copy_(xh, non_blocking=True)
            self._slots[slot][1].copy_(th, non_blocking=True)
        self._prefetched = None
        self._cx, self._ct = self._slots[slot]
        if self.broadcast_buffers == "step":
            self._broadcast_buffers()
        self._run()
        self._after_run()
        self.steps += 1
        if next_x_host is not None and next_x_host.is_pinned() and next_targets_host.is_pinned():
            other = slot ^ 1                                  # last read by step i-1, which has completed (sync below)
            with torch.cuda.stream(self._h2d_stream):
                self._slots[other][0].copy_(next_x_host, non_blocking=True)
                self._slots[other][1].copy_(next_targets_host, non_blocking=True)
                self._slot_ready[other].record(self._h2d_stream)
            self._prefetched = (other, next_x_host.data_ptr())
            self._cur_slot = other
        self._hloss.copy_(self.step_loss, non_blocking=True)
        cur.synchronize()
        return float(self._hloss[0])

    @torch.no_grad()
    def eval_logits(self, x: torch.Tensor) -> torch.Tensor:
        """model.eval() forward (BatchNorm on running statistics), engine_finetune.py:106-166.  Accepts any batch of
        1..B samples (the reference's validation loader keeps its last partial batch); returns (b, K) logits."""
        if x.dim() != 3 or tuple(x.shape[1:]) != (self.N, self.D) or not 1 <= x.shape[0] <= self.B or x.dtype != self.x_dtype:
            raise ValueError(f"x must be (b <= {self.B}, {self.N}, {self.D}) {self.x_dtype}, got {tuple(x.shape)} {x.dtype}")
        b = x.shape[0]
        self._broadcast_buffers()             # every rank evaluates on rank 0's running statistics, as under DDP
        self._cx = self.x
        self.x[:b].copy_(x, non_blocking=True)
        if b < self.B:
            self.x[b:].zero_()                # samples are independent in eval mode: the padding rows are dropped below
        self._ops_check()
        self._forward(training=False, fp32=True)   # evaluation: fp32 contractions (per call), predictions must not move
        return self.logits[:b].clone()

    def mean_loss(self) -> float:
        """Mean of the per-step losses since the last reset (one D2H sync, on demand)."""
        v = float(self.loss_sum.item()) / max(self.steps, 1)
        return v

    def top1(self) -> float:
        return float(self.correct.item()) / max(self.steps * self.B, 1)

    def reset_meters(self):
        self.loss_sum.zero_()
        self.correct.zero_()
        self.steps = 0

    def optimizer_state_dict(self):
        """torch.optim-style state for util/misc.py:322 checkpoints, indexed in parameters() order:
        LARS {'mu'}, SGD {'momentum_buffer'}, AdamW {'step', 'exp_avg', 'exp_avg_sq'}."""
        state = {}
        for i, m in enumerate(self.mus):
            if self.optimizer == "lars":
                state[i] = {"mu": m.clone()}
            elif self.optimizer == "sgd":
                state[i] = {"momentum_buffer": m.clone()} if self._mom != 0.0 else {}
            else:
                state[i] = {"step": torch.tensor(float(self.opt_steps)), "exp_avg": m.clone(), "exp_avg_sq": self.sq[i].clone()}
        group = {"lr": self._lr, "weight_decay": self._wd, "params": list(range(len(self.params)))}
        if self.optimizer == "lars":
            group.update(momentum=self._mom, trust_coefficient=self._trust)
        elif self.optimizer == "sgd":
            group.update(momentum=self._mom, dampening=0, nesterov=False)
        else:
            group.update(betas=tuple(self.betas), eps=self.eps_opt)
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        key = {"lars": "mu", "sgd": "momentum_buffer", "adamw": "exp_avg"}[self.optimizer]
        for i, m in enumerate(self.mus):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            if key in st:
                m.copy_(st[key].to(m.device))
            if self.optimizer == "adamw":
                if "exp_avg_sq" in st:
                    self.sq[i].copy_(st["exp_avg_sq"].to(m.device))
                if "step" in st:
                    self.opt_steps = int(float(st["step"]))
            elif key in st:
                self.opt_steps = max(self.opt_steps, 1)
        g = sd["param_groups"][0]
        self._lr, self._wd = float(g["lr"]), float(g["weight_decay"])
        if self.optimizer != "adamw":
            self._mom = float(g.get("momentum", self._mom))
        if self.optimizer == "lars":
            self._trust = float(g.get("trust_coefficient", self._trust))
        self._write_hyper()
