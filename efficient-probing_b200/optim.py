"""LARS and the LR schedule of the reference (util/lars.py, util/lr_sched.py), with the optimizer step
fused into two kernel launches for all tensors (ep_lars_step of the C ABI)."""
import ctypes
import math

import torch

from . import _lib

_HYPER = 5     # lr, weight_decay, momentum, trust_coefficient, grad_scale


def lars_launch(params, grads, mus, trust_flags, hyper_dev, scratch):
    """One fused LARS update for up to 8 fp32 CUDA tensors (util/lars.py:13-37)."""
    lib = _lib.load()
    n = len(params)
    P = (ctypes.c_void_p * n)(*[p.data_ptr() for p in params])
    G = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
    Mu = (ctypes.c_void_p * n)(*[m.data_ptr() for m in mus])
    Nn = (ctypes.c_longlong * n)(*[p.numel() for p in params])
    T = (ctypes.c_int * n)(*[int(t) for t in trust_flags])
    dev = params[0].device
    with torch.cuda.device(dev):
        rc = lib.ep_lars_step(n, P, G, Mu, Nn, T, hyper_dev.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "ep_lars_step")


def _ptr_table(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def adamw_launch(params, grads, exp_avg, exp_avg_sq, hyper_dev):
    """One fused torch.optim.AdamW update for up to 8 fp32 CUDA tensors (hyper: see ep_adamw_step)."""
    lib = _lib.load()
    n = len(params)
    Nn = (ctypes.c_longlong * n)(*[p.numel() for p in params])
    dev = params[0].device
    with torch.cuda.device(dev):
        rc = lib.ep_adamw_step(n, _ptr_table(params), _ptr_table(grads), _ptr_table(exp_avg), _ptr_table(exp_avg_sq), Nn,
                               hyper_dev.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "ep_adamw_step")


def sgd_launch(params, grads, momentum_bufs, hyper_dev):
    """One fused torch.optim.SGD update (dampening 0, no nesterov); momentum_bufs may be None."""
    lib = _lib.load()
    n = len(params)
    Nn = (ctypes.c_longlong * n)(*[p.numel() for p in params])
    dev = params[0].device
    with torch.cuda.device(dev):
        rc = lib.ep_sgd_step(n, _ptr_table(params), _ptr_table(grads),
                             _ptr_table(momentum_bufs) if momentum_bufs is not None else None, Nn,
                             hyper_dev.data_ptr(), _lib.stream_ptr(dev))
    _lib.check(rc, "ep_sgd_step")


class LARS(torch.optim.Optimizer):
    """LARS optimizer, no rate scaling or weight decay for parameters <= 1D (util/lars.py:4-37).

    Same constructor, param_groups and per-parameter state (``'mu'``) as the reference class, so
    optimizer checkpoints written by util/misc.py:304-332 load into it.  Parameters must be fp32
    CUDA tensors; there is no CPU step."""

    def __init__(self, params, lr=0, weight_decay=0, momentum=0.9, trust_coefficient=0.001):
        defaults = dict(lr=lr, weight_decay=weight_decay, momentum=momentum, trust_coefficient=trust_coefficient)
        super().__init__(params, defaults)
        self._hyper = {}

    @torch.no_grad()
    def step(self):
        for gi, g in enumerate(self.param_groups):
            todo = [p for p in g["params"] if p.grad is not None]
            if not todo:
                continue
            for p in todo:
                _lib.require_cuda(p, "LARS parameter")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise TypeError("LARS (ep_lars_step) handles fp32 parameters and gradients only")
                if not p.is_contiguous():     # the kernel updates in place: a contiguous copy would swallow the step
                    raise ValueError("LARS (ep_lars_step) updates parameters in place and needs them contiguous")
                if "mu" not in self.state[p]:
                    self.state[p]["mu"] = torch.zeros_like(p)
            dev = todo[0].device
            key = (gi, dev)
            if key not in self._hyper:
                self._hyper[key] = (torch.empty(_HYPER, dtype=torch.float32, device=dev),
                                    torch.empty(8192, dtype=torch.float32, device=dev))
            hyper, scratch = self._hyper[key]
            hyper.copy_(torch.tensor([g["lr"], g["weight_decay"], g["momentum"], g["trust_coefficient"], 1.0],
                                     dtype=torch.float32), non_blocking=True)
            for i in range(0, len(todo), 8):
                chunk = todo[i:i + 8]
                lars_launch([p.data for p in chunk],
                            [p.grad.contiguous() for p in chunk], [self.state[p]["mu"] for p in chunk],
                            [p.ndim > 1 for p in chunk], hyper, scratch)


def adjust_learning_rate(optimizer, epoch, args):
    """Decay the learning rate with half-cycle cosine after warmup (util/lr_sched.py:3-15)."""
    lr = cosine_lr(epoch, args.lr, args.min_lr, args.warmup_epochs, args.epochs)
    for param_group in optimizer.param_groups:
        if "lr_scale" in param_group:
            param_group["lr"] = lr * param_group["lr_scale"]
        else:
            param_group["lr"] = lr
    return lr


def cosine_lr(epoch, lr, min_lr, warmup_epochs, epochs):
    if epoch < warmup_epochs:
        return lr * epoch / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))
