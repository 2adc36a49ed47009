"""Head files: the checkpoint written by the reference's training loop and the public export format
(SURVEY.md 8f rows 2 and 4), plus the evaluation loop over cached tokens.

* training checkpoint -- ``util/misc.py:304-332``: ``{'saved_module': 'head', 'model': head.state_dict(),
  'optimizer': ..., 'epoch': ..., 'scaler': ..., 'args': ...}``; keys ``0.cls_token, 0.v.weight, [0.v.bias],
  1.running_mean, 1.running_var, 1.num_batches_tracked, 2.weight, 2.bias``.
* exported head -- ``tools/export_ep_heads.py:125``: ``{'state_dict': sd, 'meta': {...}}``.
* ``tools/ep_attention_maps.py:39-48`` finds the queries by the key suffix ``cls_token``; so does ``load_head``.
"""
from typing import Iterable, Optional, Tuple

import torch

from .probe_heads import make_ep_head


def head_state_dict_from_file(obj) -> Tuple[dict, dict]:
    """Accepts a loaded checkpoint / export (or a path) and returns (state_dict of the Sequential, meta)."""
    if isinstance(obj, str):
        obj = torch.load(obj, map_location="cpu", weights_only=False)
    if "state_dict" in obj and isinstance(obj["state_dict"], dict):          # tools/export_ep_heads.py:125
        return obj["state_dict"], dict(obj.get("meta", {}))
    if "model" in obj and isinstance(obj["model"], dict):                    # util/misc.py:318-326
        meta = {k: obj[k] for k in ("epoch", "saved_module", "test_stats") if k in obj}
        return obj["model"], meta
    if any(k.endswith("cls_token") for k in obj):                            # a bare state_dict
        return obj, {}
    raise ValueError("not an EP head file: no 'state_dict', 'model' or cls_token key")


def load_head(obj, device=None):
    """Build ``Sequential(EfficientProbing, BatchNorm1d, Linear)`` with the shapes found in the file and load it.
    Returns (head, meta).  Keys may carry a ``head.`` / ``module.head.`` prefix (full-model checkpoints)."""
    sd, meta = head_state_dict_from_file(obj)
    key = next((k for k in sd if k.endswith("cls_token")), None)
    if key is None:
        raise ValueError(f"no cls_token in the checkpoint -- not an EP head? keys: {list(sd)[:8]}")
    prefix = key[: -len("0.cls_token")]
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    _, M, D = sd["0.cls_token"].shape
    Dp = sd["0.v.weight"].shape[0]
    K = sd["2.weight"].shape[0]
    if sd["0.v.weight"].shape[1] != D or sd["2.weight"].shape[1] != Dp or D % Dp:
        raise ValueError("inconsistent EP head shapes")
    head = make_ep_head(D, num_queries=M, nb_classes=K, d_out=D // Dp, qkv_bias="0.v.bias" in sd)
    head.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    meta.update(dim=D, num_queries=M, d_out=D // Dp, nb_classes=K)
    return (head.to(device) if device is not None else head), meta


def grad_scaler_state(scale: float = 65536.0) -> dict:
    """State dict of a fresh ``torch.cuda.amp.GradScaler`` -- what ``NativeScalerWithGradNormCount.state_dict()``
    (util/misc.py:283-286) returns and ``load_model`` (util/misc.py:385) feeds to ``loss_scaler.load_state_dict``.
    The bf16 / fp32 step of this package needs no loss scaling; the entry exists so the reference can resume."""
    return {"scale": float(scale), "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}


def save_checkpoint(path, head, optimizer_state, epoch=0, args=None, test_stats=None, scaler_state=None):
    """Write the head-only checkpoint of util/misc.py:304-332 (what ``--resume`` / ``--auto_resume`` read).
    ``optimizer_state``: ``EPHeadTrainer.optimizer_state_dict()`` or a torch optimizer's ``state_dict()`` -- required,
    because the reference's ``load_model`` (util/misc.py:381-385) calls ``optimizer.load_state_dict`` on it."""
    if not isinstance(optimizer_state, dict) or "state" not in optimizer_state or "param_groups" not in optimizer_state:
        raise ValueError("optimizer_state must be an optimizer state_dict ({'state': ..., 'param_groups': ...})")
    torch.save({"saved_module": "head", "model": {k: v.detach().cpu() for k, v in head.state_dict().items()},
                "optimizer": optimizer_state, "epoch": epoch,
                "scaler": dict(scaler_state) if scaler_state is not None else grad_scaler_state(), "args": args,
                "test_stats": test_stats}, path)


def export_head(path, head, meta: Optional[dict] = None):
    """Write the public export format of tools/export_ep_heads.py:125."""
    torch.save({"state_dict": {k: v.detach().cpu() for k, v in head.state_dict().items()}, "meta": dict(meta or {})},
               path)


@torch.no_grad()
def evaluate(trainer, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]):
    """engine_finetune.py:106-166 on cached tokens: eval-mode forward (BatchNorm on running statistics,
    fp32 contractions), mean cross-entropy, top-1 and top-5 accuracy in percent."""
    n = 0
    loss_sum = torch.zeros((), dtype=torch.float64, device=trainer.dev)
    c1 = torch.zeros((), dtype=torch.int64, device=trainer.dev)
    c5 = torch.zeros((), dtype=torch.int64, device=trainer.dev)
    for x, y in batches:
        logits = trainer.eval_logits(x)
        y = y.to(trainer.dev)
        loss_sum += torch.nn.functional.cross_entropy(logits, y, reduction="sum").double()
        top5 = logits.topk(min(5, logits.shape[1]), dim=1).indices
        c1 += (top5[:, 0] == y).sum()
        c5 += (top5 == y[:, None]).any(dim=1).sum()
        n += y.numel()
    return {"loss": float(loss_sum) / max(n, 1), "acc1": 100.0 * float(c1) / max(n, 1), "acc5": 100.0 * float(c5) / max(n, 1),
            "n": n}
