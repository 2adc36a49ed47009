"""efficient_probing_b200 -- B200-native EP probe head (one hot path of billpsomas/efficient-probing).

Public surface mirrors the reference for this path:
  EfficientProbing        poolings/ep.py            (module, drop-in)
  build_probe_head        probe_heads.py:87-106     (Sequential(EP, BatchNorm1d(affine=False), Linear))
  EPHeadTrainer           engine_finetune.py:22-103 hot loop for cached tokens, fused + graph-captured
  LARS, adjust_learning_rate   util/lars.py, util/lr_sched.py
  ep_attention            tools/ep_attention_maps.py:51-58
"""
from .ep import EfficientProbing, EPPoolFunction, ep_attention                      # noqa: F401
from .probe_heads import build_probe_head, make_ep_head, POOLINGS                    # noqa: F401
from .optim import LARS, adjust_learning_rate                                        # noqa: F401
from .trainer import EPHeadTrainer                                                   # noqa: F401
from .flatgrad import FlatGradLayout, shard_range, allreduce_sum_                    # noqa: F401
from .token_cache import TokenShard, TokenStream, write_shard, load_reference_npz, epoch_order   # noqa: F401
from .head_io import load_head, save_checkpoint, export_head, evaluate                # noqa: F401
from . import _lib                                                                   # noqa: F401

__all__ = ["EfficientProbing", "EPPoolFunction", "ep_attention", "build_probe_head", "make_ep_head", "POOLINGS",
           "LARS", "adjust_learning_rate", "EPHeadTrainer", "FlatGradLayout", "shard_range", "TokenShard", "TokenStream",
           "write_shard", "load_reference_npz", "load_head", "save_checkpoint", "export_head", "evaluate"]
