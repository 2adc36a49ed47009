"""Flat gradient buffer of the EP head and the data-parallel exchange over it (host-side logic of the
multi-GPU path; main_linprobe.py:581-583 wraps the head in DDP, whose bucketed all-reduce-mean this is).

Layout: ``[fc.weight, fc.bias, v.weight, (v.bias), cls_token]``, every slice starting on a 256-byte
boundary (vector stores / TMA in the kernels), zero padding in between.  The first four are final before
the token-streaming half of the backward pass, so ``early`` marks the prefix that can be reduced
underneath it; ``cls_token`` (M*D floats) follows.  Ranks sum; the 1/world factor is applied inside the
LARS kernel (``hyper[4]``)."""
from collections import OrderedDict

import torch
import torch.distributed as dist

ALIGN = 64            # floats


def _pad(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class FlatGradLayout:
    def __init__(self, K, Dp, D, M, has_v_bias):
        self.sizes = OrderedDict([("fc_w", K * Dp), ("fc_b", K), ("v_w", Dp * D), ("v_b", Dp if has_v_bias else 0),
                                  ("cls", M * D)])
        self.offsets, off = {}, 0
        for k, n in self.sizes.items():
            self.offsets[k] = off
            off += _pad(n)
        self.total = off
        self.early = self.offsets["cls"]            # [0, early) is ready before ep_bwd_pool runs

    def allocate(self, device):
        return torch.zeros(self.total, dtype=torch.float32, device=device)

    def views(self, flat):
        return {k: flat[o:o + self.sizes[k]] for k, o in self.offsets.items()}


def shard_range(rank, world, global_batch):
    """Samples [lo, hi) of a global batch owned by ``rank`` (equal shards; drop_last semantics of
    main_linprobe.py:313-314 make the local mean-loss gradients average to the global mean)."""
    if global_batch % world:
        raise ValueError("global batch must divide evenly over the ranks (drop_last=True in the reference)")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def allreduce_sum_(flat, group=None, lo=0, hi=None):
    """In-place sum over ranks of flat[lo:hi] (NCCL on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat[lo:hi if hi is not None else flat.numel()], group=group)
    return flat
