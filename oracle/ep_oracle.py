"""CPU oracle for the EP probe-head hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch, fp32 or fp64, no autocast) of the reference algorithm
for the one path this repository accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``efficient-probing_b200/``) never does and fails loudly when its CUDA library is missing.

Every function cites the reference lines it restates (paths relative to the reference repo
``billpsomas/efficient-probing``):

  ep_forward            poolings/ep.py:28-47      (q/k/v wiring, softmax, weighted sum)
  ep_init               poolings/ep.py:17-26      (creation order: v Linear, then cls_token)
  ep_attention          tools/ep_attention_maps.py:51-58
  build_head            probe_heads.py:75-76,104-106,109-110
  cross_entropy         main_linprobe.py:589 (nn.CrossEntropyLoss default = mean NLL of log-softmax)
  batchnorm_train/eval  probe_heads.py:109-110 (BatchNorm1d(affine=False, eps=1e-6), momentum 0.1)
  lars_step             util/lars.py:13-37
  cosine_lr             util/lr_sched.py:3-15
  ep_forward_pooled     the pool-then-project re-association used by the CUDA kernels
                        (SURVEY.md section 0); algebraically identical to ep_forward.
  ep_backward_pooled    closed-form gradients of the same, checked
                        against autograd of ep_forward in tests/test_oracle.py.

Pinning: the reference ships no golden vectors for forward/backward values, so
the oracle is pinned against outputs of the reference module itself, imported from
/root/reference in the build container by ``tests/golden/make_golden.py``; the resulting small
fixtures are committed under ``tests/golden/`` and ``tests/test_oracle.py`` checks this file
against them, plus the reference's parameter-count known answers (logs/*/ep.txt:9).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
@dataclass
class EPParams:
    """Parameters of Sequential(EfficientProbing, BatchNorm1d(affine=False), Linear)."""
    cls_token: torch.Tensor            # (1, M, D)      state_dict key 0.cls_token
    v_weight: torch.Tensor             # (D/d_out, D)   0.v.weight
    v_bias: Optional[torch.Tensor]     # (D/d_out,)     0.v.bias (only if qkv_bias)
    running_mean: torch.Tensor         # (D/d_out,)     1.running_mean
    running_var: torch.Tensor          # (D/d_out,)     1.running_var
    num_batches_tracked: int           #                1.num_batches_tracked
    fc_weight: torch.Tensor            # (K, D/d_out)   2.weight
    fc_bias: torch.Tensor              # (K,)           2.bias
    num_queries: int
    d_out: int
    scale: float

    def trainable(self):
        """parameters() order of the reference Sequential: cls_token, v.weight, [v.bias], fc.weight, fc.bias."""
        out = [("0.cls_token", self.cls_token), ("0.v.weight", self.v_weight)]
        if self.v_bias is not None:
            out.append(("0.v.bias", self.v_bias))
        out += [("2.weight", self.fc_weight), ("2.bias", self.fc_bias)]
        return out

    def clone(self, dtype=None):
        c = lambda t: None if t is None else t.detach().clone().to(dtype or t.dtype)
        return EPParams(c(self.cls_token), c(self.v_weight), c(self.v_bias), c(self.running_mean),
                        c(self.running_var), self.num_batches_tracked, c(self.fc_weight), c(self.fc_bias),
                        self.num_queries, self.d_out, self.scale)


def ep_init(dim: int, num_queries: int = 32, d_out: int = 1, qkv_bias: bool = False,
            qk_scale: Optional[float] = None, num_heads: int = 1):
    """Draw EP parameters in the reference's creation order (ep.py:25-26): the value Linear
    first (kaiming-uniform weight, then bias), then ``cls_token = randn(1, M, D) * 0.02``."""
    if num_heads != 1:
        raise ValueError("the reference forward only works for num_heads == 1 (ep.py:45)")
    v = torch.nn.Linear(dim, dim // d_out, bias=qkv_bias)
    cls_token = torch.randn(1, num_queries, dim) * 0.02
    scale = qk_scale or (dim // num_heads) ** -0.5          # ep.py:19-20
    return cls_token, v.weight.detach(), (v.bias.detach() if qkv_bias else None), scale


def build_head(dim: int, num_queries: int, nb_classes: int, d_out: int = 1,
               qkv_bias: bool = False, seed: Optional[int] = 0) -> EPParams:
    """probe_heads.py:104-106: pooling is built first, then the fresh classifier Linear
    (probe_heads.py:75-76), BatchNorm1d(affine=False, eps=1e-6) in between (no RNG draws)."""
    if seed is not None:
        torch.manual_seed(seed)
    cls_token, v_w, v_b, scale = ep_init(dim, num_queries, d_out, qkv_bias)
    fc = torch.nn.Linear(dim // d_out, nb_classes, bias=True)
    w = dim // d_out
    return EPParams(cls_token, v_w, v_b, torch.zeros(w), torch.ones(w), 0,
                    fc.weight.detach(), fc.bias.detach(), num_queries, d_out, scale)


def param_count(dim: int, num_queries: int, nb_classes: int = 1000) -> int:
    """tools/gen_leaderboard.py:493-505: params = C^2 + C*q + 1000*C + 1000 (d_out=1, no bias)."""
    return dim * dim + dim * num_queries + nb_classes * dim + nb_classes


# --------------------------------------------------------------------------------------------
# the pooling, as the reference computes it (value projection on every token)
# --------------------------------------------------------------------------------------------
def ep_forward(x, cls_token, v_weight, v_bias, scale: float, num_queries: int, d_out: int,
               return_attn: bool = False):
    """poolings/ep.py:28-47 with num_heads == 1.

    x (B, N, C) -> (B, C // d_out); optionally also the attention map (B, M, N)."""
    B, N, C = x.shape
    M = num_queries
    c = C // (d_out * M)                                    # channels owned by one query
    q = cls_token.expand(B, -1, -1) * scale                 # ep.py:35,39   (B, M, C)
    logits = torch.einsum("bmc,bnc->bmn", q, x)             # ep.py:42
    attn = logits.softmax(dim=-1)                           # ep.py:43
    v = F.linear(x, v_weight, v_bias)                       # ep.py:25,40   (B, N, C')
    v = v.reshape(B, N, M, c)                               # query m owns channels [m*c, (m+1)*c)
    out = torch.einsum("bmn,bnmc->bmc", attn, v)            # ep.py:44
    out = out.reshape(B, C // d_out)                        # ep.py:45
    return (out, attn) if return_attn else out


def ep_attention(tokens, cls_token):
    """tools/ep_attention_maps.py:51-58: per-image attention (Q, N) = softmax(cls*C^-0.5 @ tokens^T)."""
    C = tokens.shape[-1]
    q = cls_token * (C ** -0.5)
    return (q @ tokens.transpose(0, 1)).softmax(dim=-1)


# --------------------------------------------------------------------------------------------
# the same pooling, re-associated the way the CUDA path computes it
# --------------------------------------------------------------------------------------------
def ep_forward_pooled(x, cls_token, v_weight, v_bias, scale, num_queries, d_out):
    """Pool raw tokens first, project after:  P[b,m] = sum_n A[b,m,n] x[b,n];
    out[b, m*c:(m+1)*c] = W[m*c:(m+1)*c] @ P[b,m] + bias.   Returns (out, attn, P, rowmax, rowsum)."""
    B, N, C = x.shape
    M = num_queries
    c = C // (d_out * M)
    q = cls_token[0] * scale
    logits = torch.einsum("mc,bnc->bmn", q, x)
    rowmax = logits.max(dim=-1).values
    e = torch.exp(logits - rowmax[..., None])
    rowsum = e.sum(dim=-1)
    attn = e / rowsum[..., None]
    P = torch.einsum("bmn,bnc->bmc", attn, x)               # (B, M, C)
    Wm = v_weight.reshape(M, c, C)
    out = torch.einsum("mjc,bmc->bmj", Wm, P).reshape(B, M * c)
    if v_bias is not None:
        out = out + v_bias
    return out, attn, P, rowmax, rowsum


def ep_backward_pooled(x, cls_token, v_weight, scale, num_queries, d_out, attn, P, g, want_dx=False):
    """Closed-form gradients for ep_forward_pooled given g = dL/d out (B, C').

    Returns dict(d_cls_token (1,M,C), d_v_weight, d_v_bias, [d_x])."""
    B, N, C = x.shape
    M = num_queries
    c = C // (d_out * M)
    dU = g.reshape(B, M, c)
    Wm = v_weight.reshape(M, c, C)
    d_v_weight = torch.einsum("bmj,bmc->mjc", dU, P).reshape(M * c, C)
    d_v_bias = g.sum(dim=0)
    dP = torch.einsum("bmj,mjc->bmc", dU, Wm)               # (B, M, C)
    delta = (dP * P).sum(dim=-1)                            # (B, M)  = sum_n A dA
    dA = torch.einsum("bmc,bnc->bmn", dP, x)
    dS = attn * (dA - delta[..., None])
    dq = torch.einsum("bmn,bnc->mc", dS, x)
    out = {"d_cls_token": (dq * scale)[None], "d_v_weight": d_v_weight, "d_v_bias": d_v_bias}
    if want_dx:
        q = cls_token[0] * scale
        out["d_x"] = torch.einsum("bmn,bmc->bnc", attn, dP) + torch.einsum("bmn,mc->bnc", dS, q)
    return out


# --------------------------------------------------------------------------------------------
# BatchNorm1d(affine=False, eps=1e-6), classifier, loss
# --------------------------------------------------------------------------------------------
BN_EPS = 1e-6          # probe_heads.py:110
BN_MOMENTUM = 0.1      # torch default, untouched by the reference


def batchnorm_train(h, running_mean, running_var, num_batches_tracked, eps=BN_EPS, momentum=BN_MOMENTUM):
    """Train-mode BatchNorm1d: normalise with biased batch variance, update running stats with
    the unbiased one.  Returns (y, new_running_mean, new_running_var, new_num_batches_tracked)."""
    B = h.shape[0]
    mean = h.mean(dim=0)
    var = h.var(dim=0, unbiased=False)
    y = (h - mean) / torch.sqrt(var + eps)
    unbiased = var * (B / (B - 1)) if B > 1 else var
    new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
    new_rv = (1 - momentum) * running_var + momentum * unbiased.detach()
    return y, new_rm, new_rv, num_batches_tracked + 1


def batchnorm_eval(h, running_mean, running_var, eps=BN_EPS):
    return (h - running_mean) / torch.sqrt(running_var + eps)


def cross_entropy(logits, targets):
    """nn.CrossEntropyLoss() defaults: mean over the batch of -log_softmax(logits)[target]."""
    lse = torch.logsumexp(logits, dim=-1)
    picked = logits.gather(1, targets[:, None].long())[:, 0]
    return (lse - picked).mean()


def head_forward(p: EPParams, x, train: bool = True, pooled: bool = False):
    """Sequential(EP, BN, Linear) forward.  Returns dict(out, attn, y, logits, bn=(rm, rv, nbt))."""
    if pooled:
        out, attn, P, rowmax, rowsum = ep_forward_pooled(x, p.cls_token, p.v_weight, p.v_bias, p.scale,
                                                          p.num_queries, p.d_out)
    else:
        out, attn = ep_forward(x, p.cls_token, p.v_weight, p.v_bias, p.scale, p.num_queries, p.d_out,
                               return_attn=True)
    if train:
        y, rm, rv, nbt = batchnorm_train(out, p.running_mean, p.running_var, p.num_batches_tracked)
    else:
        y = batchnorm_eval(out, p.running_mean, p.running_var)
        rm, rv, nbt = p.running_mean, p.running_var, p.num_batches_tracked
    logits = F.linear(y, p.fc_weight, p.fc_bias)            # probe_heads.py:76
    return {"out": out, "attn": attn, "y": y, "logits": logits, "bn": (rm, rv, nbt)}


def head_loss_and_grads(p: EPParams, x, targets, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """One fwd+bwd of the reference head through autograd, in ``dtype``, train mode.
    Returns loss, logits, attn, out and the gradient of every trainable tensor by state_dict key."""
    q = p.clone(dtype)
    leaves = [t.requires_grad_(True) for _, t in q.trainable()]
    r = head_forward(q, x.to(dtype), train=True, pooled=False)
    loss = cross_entropy(r["logits"], targets)
    grads = torch.autograd.grad(loss, leaves)
    res = {"loss": loss.detach(), "logits": r["logits"].detach(), "attn": r["attn"].detach(),
           "out": r["out"].detach(), "y": r["y"].detach(),
           "running_mean": r["bn"][0].detach(), "running_var": r["bn"][1].detach()}
    for (name, _), g in zip(q.trainable(), grads):
        res["grad." + name] = g.detach()
    return res


def head_loss_and_grads_pooled(p: EPParams, x, targets, dtype=torch.float64) -> Dict[str, torch.Tensor]:
    """Same result as head_loss_and_grads, computed through the pool-then-project closed form
    (ep_forward_pooled + autograd of BN/Linear/CE only + ep_backward_pooled): no (B, N, D') value tensor,
    so BASELINE-sized batches finish in seconds.  tests/test_oracle.py holds it to head_loss_and_grads."""
    q = p.clone(dtype)
    xd = x.to(dtype)
    out, attn, P, rowmax, rowsum = ep_forward_pooled(xd, q.cls_token, q.v_weight, q.v_bias, q.scale,
                                                     q.num_queries, q.d_out)
    out_leaf = out.detach().requires_grad_(True)
    fcw, fcb = q.fc_weight.requires_grad_(True), q.fc_bias.requires_grad_(True)
    y, rm, rv, _ = batchnorm_train(out_leaf, q.running_mean, q.running_var, q.num_batches_tracked)
    logits = F.linear(y, fcw, fcb)
    loss = cross_entropy(logits, targets)
    g_out, g_fcw, g_fcb = torch.autograd.grad(loss, [out_leaf, fcw, fcb])
    cf = ep_backward_pooled(xd, q.cls_token, q.v_weight, q.scale, q.num_queries, q.d_out, attn, P, g_out)
    res = {"loss": loss.detach(), "logits": logits.detach(), "attn": attn, "out": out, "y": y.detach(),
           "running_mean": rm.detach(), "running_var": rv.detach(), "P": P, "rowmax": rowmax, "rowsum": rowsum,
           "g_out": g_out, "grad.0.cls_token": cf["d_cls_token"], "grad.0.v.weight": cf["d_v_weight"],
           "grad.2.weight": g_fcw, "grad.2.bias": g_fcb}
    if q.v_bias is not None:
        res["grad.0.v.bias"] = cf["d_v_bias"]
    return res


# --------------------------------------------------------------------------------------------
# optimizer + schedule
# --------------------------------------------------------------------------------------------
def lars_step(params, grads, mus, lr: float, weight_decay: float = 0.0, momentum: float = 0.9,
              trust_coefficient: float = 0.001):
    """util/lars.py:13-37.  ``params``/``grads``/``mus`` are parallel lists; returns new (params, mus).
    Trust-ratio scaling and weight decay only for tensors with ndim > 1 (lars.py:21)."""
    new_p, new_mu = [], []
    for p, g, mu in zip(params, grads, mus):
        dp = g
        if p.ndim > 1:
            dp = dp + weight_decay * p                      # lars.py:22
            pn = torch.linalg.vector_norm(p)                # lars.py:23
            un = torch.linalg.vector_norm(dp)               # lars.py:24
            q = torch.where(pn > 0, torch.where(un > 0, trust_coefficient * pn / un, torch.ones_like(pn)),
                            torch.ones_like(pn))            # lars.py:26-29
            dp = dp * q
        mu = mu * momentum + dp                             # lars.py:36
        new_mu.append(mu)
        new_p.append(p - lr * mu)                           # lars.py:37
    return new_p, new_mu


def cosine_lr(epoch: float, lr: float, min_lr: float, warmup_epochs: float, epochs: float) -> float:
    """util/lr_sched.py:3-15: linear warm-up then half-cycle cosine; ``epoch`` is fractional
    (engine_finetune.py:43-44 passes data_iter_step / len(loader) + epoch)."""
    if epoch < warmup_epochs:
        return lr * epoch / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) /
                                                           (epochs - warmup_epochs)))


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def synthetic_tokens(B, N, D, seed=1234, class_shift: Optional[torch.Tensor] = None, spread: float = 1.0):
    """N(0,1) tokens rounded to bf16 (returned as bf16); optional per-sample class-dependent
    mean shift +0.5 on channel (y mod D) so that top-1 predictions are not degenerate."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, D, generator=g) * spread
    if class_shift is not None:
        idx = (class_shift % D).long()
        x[torch.arange(B), :, idx] += 0.5
    return x.to(torch.bfloat16)


def synthetic_labels(B, K, seed=4321):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, K, (B,), generator=g)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative L2 error ||a-b|| / ||b|| in fp64 (the metric every parity test states)."""
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| (a scale-relative max-norm error)."""
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
