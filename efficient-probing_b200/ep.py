"""EfficientProbing -- drop-in for the reference's ``poolings/ep.py`` backed by libep_b200 (sm_100a).

Same constructor arguments, attributes, ``forward(x, cls=None, **_)`` signature and state_dict layout
(``v.weight``, ``[v.bias]``, ``cls_token``) as poolings/ep.py:7-47, so ``probe_heads.build_probe_head``,
``main_linprobe.py`` and ``engine_finetune.py`` can use it unchanged.  Inside, one
``torch.autograd.Function`` calls ``ep_fwd`` / ``ep_bwd`` of the C ABI (include/ep_b200.h)."""
from typing import Any, Optional

import torch
from torch import nn

from . import _lib


def _workspace(x, nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=x.device)


class EPPoolFunction(torch.autograd.Function):
    """out = EP(x; cls_token, v.weight, v.bias).  Saves x (no copy), the logits, the softmax row statistics and
    the pooled tokens P; backward returns gradients for cls_token, v.weight and v.bias through ``ep_bwd``.
    The probe trains on a frozen backbone (main_linprobe.py:393-400), so dL/dx is only computed when x
    requires grad (``--finetuning``), and per-sample queries (``cls_batched``: cls_token is (B, M, D), the
    ``cls=`` argument of ep.py:32-33) are supported the same way: both go through ``ep_fwd_ex`` /
    ``ep_bwd_ex`` on the general kernel family."""

    @staticmethod
    def forward(ctx, x, cls_token, v_weight, v_bias, scale, num_queries, d_out, return_attn, cls_batched=False):
        lib = _lib.load()
        _lib.require_cuda(x, "x")
        if x.dim() != 3:
            raise ValueError(f"x must be (B, N, C), got {tuple(x.shape)}")
        x = x.contiguous()
        B, N, D = x.shape
        M = int(num_queries)
        want = (B, M, D) if cls_batched else (1, M, D)
        if tuple(cls_token.shape) != want:
            raise ValueError(f"cls_token must be {want}, got {tuple(cls_token.shape)}")
        cls32 = cls_token.detach().float().contiguous()
        w32 = v_weight.detach().float().contiguous()
        b32 = None if v_bias is None else v_bias.detach().float().contiguous()
        dev = x.device
        Dp = D // d_out
        out = torch.empty(B, Dp, dtype=torch.float32, device=dev)
        rowmax = torch.empty(B, M, dtype=torch.float32, device=dev)
        rowsum = torch.empty(B, M, dtype=torch.float32, device=dev)
        P = torch.empty(B, M, D, dtype=torch.float32, device=dev)
        S = torch.empty(B, M, N, dtype=torch.float32, device=dev)
        attn = torch.empty(B, M, N, dtype=torch.float32, device=dev) if return_attn else None
        nbytes = lib.ep_workspace_bytes(B, N, D, M, d_out)
        ws = _workspace(x, nbytes)
        xt = _lib.x_dtype_code(x)
        with torch.cuda.device(dev):
            if cls_batched:
                rc = lib.ep_fwd_ex(x.data_ptr(), xt, cls32.data_ptr(), 1, w32.data_ptr(), _lib.ptr(b32),
                                   float(scale), B, N, D, M, int(d_out), out.data_ptr(), S.data_ptr(),
                                   rowmax.data_ptr(), rowsum.data_ptr(), P.data_ptr(), _lib.ptr(attn), ws.data_ptr(),
                                   ws.numel(), _lib.stream_ptr(dev))
                p_layout = 0
            else:
                p_layout = lib.ep_pooled_layout(xt, B, N, D, M, int(d_out))      # as ep_fwd will write it
                rc = lib.ep_fwd(x.data_ptr(), xt, cls32.data_ptr(), w32.data_ptr(), _lib.ptr(b32),
                                float(scale), B, N, D, M, int(d_out), out.data_ptr(), S.data_ptr(), rowmax.data_ptr(),
                                rowsum.data_ptr(), P.data_ptr(), _lib.ptr(attn), ws.data_ptr(), ws.numel(),
                                _lib.stream_ptr(dev))
        _lib.check(rc, "ep_fwd_ex" if cls_batched else "ep_fwd")
        ctx.save_for_backward(x, cls32, w32, S, rowmax, rowsum, P, out, b32)
        ctx.meta = (float(scale), M, int(d_out), v_bias is not None, cls_token.dtype, v_weight.dtype)
        ctx.ext = (bool(cls_batched), int(p_layout))
        if return_attn:
            ctx.mark_non_differentiable(attn)
            return out, attn
        return out

    @staticmethod
    def backward(ctx, g, *unused):
        lib = _lib.load()
        x, cls32, w32, S, rowmax, rowsum, P, out, b32 = ctx.saved_tensors
        scale, M, d_out, has_bias, cls_dtype, w_dtype = ctx.meta
        B, N, D = x.shape
        dev = x.device
        g = g.detach().float().contiguous()
        cls_batched, p_layout = ctx.ext
        want_dx = ctx.needs_input_grad[0]
        d_cls = torch.empty(B if cls_batched else 1, M, D, dtype=torch.float32, device=dev)
        d_w = torch.empty_like(w32)
        d_b = torch.empty(D // d_out, dtype=torch.float32, device=dev) if has_bias else None
        ws = _workspace(x, lib.ep_workspace_bytes(B, N, D, M, d_out))
        dx = torch.empty_like(x) if want_dx else None
        with torch.cuda.device(dev):
            if cls_batched or want_dx:
                rc = lib.ep_bwd_ex(x.data_ptr(), _lib.x_dtype_code(x), cls32.data_ptr(), int(cls_batched), w32.data_ptr(),
                                   scale, B, N, D, M, d_out, S.data_ptr(), rowmax.data_ptr(), rowsum.data_ptr(),
                                   P.data_ptr(), p_layout, out.data_ptr(), _lib.ptr(b32), g.data_ptr(),
                                   d_cls.data_ptr(), d_w.data_ptr(), _lib.ptr(d_b), _lib.ptr(dx), ws.data_ptr(),
                                   ws.numel(), _lib.stream_ptr(dev))
            else:
                # ep_bwd re-derives the layout of the saved P from the process-wide kernel / GEMM mode: refuse to
                # read it back under another mode than the one the forward wrote it in
                if lib.ep_pooled_layout(_lib.x_dtype_code(x), B, N, D, M, d_out) != p_layout:
                    raise RuntimeError("ep_set_kernel_mode / ep_set_gemm_mode changed between forward and backward: "
                                       "the saved pooled tokens are in the other layout")
                rc = lib.ep_bwd(x.data_ptr(), _lib.x_dtype_code(x), cls32.data_ptr(), w32.data_ptr(), scale,
                                B, N, D, M, d_out, S.data_ptr(), rowmax.data_ptr(), rowsum.data_ptr(), P.data_ptr(),
                                out.data_ptr(), _lib.ptr(b32), g.data_ptr(),
                                d_cls.data_ptr(), d_w.data_ptr(), _lib.ptr(d_b), ws.data_ptr(), ws.numel(),
                                _lib.stream_ptr(dev))
        _lib.check(rc, "ep_bwd")
        return dx, d_cls.to(cls_dtype), d_w.to(w_dtype), d_b, None, None, None, None, None


class EfficientProbing(nn.Module):
    """Multi-query cross-attention pooling (poolings/ep.py:7-47)."""

    def __init__(self, dim: int, num_heads: int = 1, qkv_bias: bool = False, qk_scale: Optional[float] = None,
                 num_queries: int = 32, d_out: int = 1):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.d_out = d_out
        self.num_queries = num_queries
        # creation order fixes the RNG draws (probe_heads.py:14-16): value Linear first, then the queries
        self.v = nn.Linear(dim, dim // d_out, bias=qkv_bias)
        self.cls_token = nn.Parameter(torch.randn(1, num_queries, dim) * 0.02)

    def _check(self, x, cls):
        if self.num_heads != 1:
            # the reference forward itself fails for num_heads > 1 (view at ep.py:45)
            raise RuntimeError("EfficientProbing supports num_heads == 1 only (as the reference, ep.py:45)")
        C = x.shape[-1]
        if C % (self.d_out * self.num_queries) != 0:
            raise RuntimeError(f"shape '[{x.shape[0]}, {x.shape[1]}, {self.num_queries}, "
                               f"{C // (self.d_out * self.num_queries)}]' is invalid for input of size "
                               f"{x.shape[0] * x.shape[1] * (C // self.d_out)}")     # ep.py:40 reshape

    def forward(self, x: torch.Tensor, cls=None, **_: Any) -> torch.Tensor:
        self._check(x, cls)
        if cls is not None:
            # external per-sample queries replace the learned ones (ep.py:32-33,35: reshaped to (B, M, C))
            q = cls.reshape(x.shape[0], self.num_queries, x.shape[-1])
            return EPPoolFunction.apply(x, q, self.v.weight, self.v.bias, self.scale, self.num_queries, self.d_out,
                                        False, True)
        return EPPoolFunction.apply(x, self.cls_token, self.v.weight, self.v.bias, self.scale,
                                    self.num_queries, self.d_out, False, False)

    @torch.no_grad()
    def attention_maps(self, x: torch.Tensor) -> torch.Tensor:
        """(B, M, N) attention of every query over the tokens -- tools/ep_attention_maps.py:51-58."""
        return ep_attention(x, self.cls_token[0], self.scale)


@torch.no_grad()
def ep_attention(tokens: torch.Tensor, cls_token: torch.Tensor, scale: Optional[float] = None) -> torch.Tensor:
    """attn[(b,) q, n] = softmax(cls_token * C^-0.5 @ tokens^T) -- tools/ep_attention_maps.py:51-58.
    tokens (N, C) or (B, N, C) on the GPU; cls_token (Q, C)."""
    lib = _lib.load()
    _lib.require_cuda(tokens, "tokens")
    squeeze = tokens.dim() == 2
    x = (tokens[None] if squeeze else tokens).contiguous()
    B, N, D = x.shape
    q = cls_token.detach().reshape(-1, D).float().contiguous().to(x.device)
    M = q.shape[0]
    attn = torch.empty(B, M, N, dtype=torch.float32, device=x.device)
    ws = _workspace(x, lib.ep_workspace_bytes(B, N, D, M, 1))
    with torch.cuda.device(x.device):
        rc = lib.ep_attention_maps(x.data_ptr(), _lib.x_dtype_code(x), q.data_ptr(),
                                   float(D ** -0.5 if scale is None else scale), B, N, D, M, attn.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
    _lib.check(rc, "ep_attention_maps")
    return attn[0] if squeeze else attn
