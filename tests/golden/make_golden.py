"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

What is imported from the reference, unmodified:
  poolings/ep.py            EfficientProbing           (the module under test)
  util/lars.py              LARS                       (optimizer step)
  util/lr_sched.py          adjust_learning_rate       (schedule)
  tools/ep_attention_maps.py  ep_attention             (attention-map definition; its module-level
                                                        matplotlib/PIL imports are stubbed, the
                                                        function itself is pure torch)
What is restated here because probe_heads.py / main_linprobe.py need timm / open_clip (absent):
  Sequential(EP, BatchNorm1d(affine=False, eps=1e-6), Linear)   probe_heads.py:75-76,104-106,109-110
  CrossEntropyLoss()                                            main_linprobe.py:589

Outputs: tests/golden/case_*.npz (inputs, parameters, forward values, gradients, post-LARS
parameters, all fp32 computed by the reference in fp32 and again in fp64) and
tests/golden/fingerprints.json (init hashes in the style of tools/inv_heads.py:102-120,
parameter-count known answers from logs/*/ep.txt:9, lr-schedule samples).
"""
import hashlib
import json
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch
from torch import nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from poolings.ep import EfficientProbing            # noqa: E402
from util.lars import LARS                          # noqa: E402
from util.lr_sched import adjust_learning_rate      # noqa: E402


class _Stub(types.ModuleType):
    """Placeholder for plotting-only imports of tools/ep_attention_maps.py (never called)."""
    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Stub(self.__name__ + "." + attr)

    def __call__(self, *a, **k):
        return None


def _import_ep_attention():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "PIL", "PIL.Image",
                 "sklearn", "sklearn.cluster", "torchvision", "torchvision.transforms"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ep_attention_maps", REF + "/tools/ep_attention_maps.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ep_attention


def ref_head(dim, M, K, d_out, qkv_bias, seed=0):
    torch.manual_seed(seed)
    pool = EfficientProbing(dim=dim, num_queries=M, d_out=d_out, qkv_bias=qkv_bias)   # probe_heads.py:75,104
    fc = nn.Linear(dim // d_out, K, bias=True)                                          # probe_heads.py:76,105
    return nn.Sequential(pool, nn.BatchNorm1d(fc.in_features, affine=False, eps=1e-6), fc)


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def run_case(name, B, N, D, M, K, d_out, qkv_bias, spread=1.0, q_gain=1.0, lr=0.05, wd=0.0):
    ep_attention = _import_ep_attention()
    head = ref_head(D, M, K, d_out, qkv_bias)
    with torch.no_grad():
        head[0].cls_token.mul_(q_gain)              # optional: sharper attention than the 0.02 init gives
    g = torch.Generator().manual_seed(1234)
    x = bf16_round(torch.randn(B, N, D, generator=g) * spread)
    y = torch.randint(0, K, (B,), generator=torch.Generator().manual_seed(4321))
    out = {"x": x.numpy(), "targets": y.numpy().astype(np.int64),
           "meta": np.array(json.dumps(dict(B=B, N=N, D=D, M=M, K=K, d_out=d_out, qkv_bias=qkv_bias,
                                            lr=lr, weight_decay=wd, scale=float(head[0].scale))))}
    for k, v in head.state_dict().items():
        out["param." + k] = v.detach().numpy().copy()

    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        h = ref_head(D, M, K, d_out, qkv_bias).to(dt)
        h.load_state_dict({k: v.to(dt) if v.is_floating_point() else v for k, v in head.state_dict().items()})
        h.train()
        pooled = h[0](x.to(dt))
        logits = h[2](h[1](pooled))
        loss = nn.CrossEntropyLoss()(logits, y)                                    # main_linprobe.py:589
        loss.backward()
        out[f"{tag}.out"] = pooled.detach().numpy().copy()
        out[f"{tag}.logits"] = logits.detach().numpy().copy()
        out[f"{tag}.loss"] = loss.detach().numpy().copy()
        out[f"{tag}.attn"] = torch.stack([ep_attention(x[b].to(dt), h[0].cls_token[0].detach())
                                          for b in range(B)]).numpy()
        for k, p in h.named_parameters():
            out[f"{tag}.grad.{k}"] = p.grad.detach().numpy().copy()
        out[f"{tag}.running_mean"] = h[1].running_mean.detach().numpy().copy()
        out[f"{tag}.running_var"] = h[1].running_var.detach().numpy().copy()
        # eval-mode logits with the just-updated running stats (engine_finetune.py:106-166 semantics)
        h.eval()
        with torch.no_grad():
            out[f"{tag}.eval_logits"] = h(x.to(dt)).numpy().copy()
        h.train()
        # two LARS steps (momentum state matters on the second)                 util/lars.py
        opt = LARS(h.parameters(), lr=lr, weight_decay=wd)                     # main_linprobe.py:403-408
        opt.step()
        opt.zero_grad()
        loss2 = nn.CrossEntropyLoss()(h(x.to(dt)), y)
        loss2.backward()
        opt.step()
        out[f"{tag}.loss_step2"] = loss2.detach().numpy().copy()
        for k, p in h.named_parameters():
            out[f"{tag}.after2.{k}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **out)
    print("wrote case", name, {k: v.shape for k, v in out.items() if k.startswith("f32.")})


def fingerprints():
    fp = {"init_sha256": {}, "param_count": {}, "lr_sched": [], "state_dict_keys": {}}
    for D, M, d_out, bias in [(768, 32, 1, False), (768, 32, 2, False), (768, 32, 4, False),
                              (1024, 32, 1, False), (768, 8, 1, False), (64, 8, 1, True)]:
        head = ref_head(D, M, 1000, d_out, bias, seed=0)
        h = hashlib.sha256()
        for n, p in sorted(head.named_parameters()):                          # tools/inv_heads.py:113-116
            h.update(n.encode())
            h.update(p.detach().float().cpu().numpy().tobytes())
        key = f"D{D}_M{M}_dout{d_out}_bias{int(bias)}"
        fp["init_sha256"][key] = h.hexdigest()
        fp["state_dict_keys"][key] = list(head.state_dict().keys())
        fp["param_count"][key] = sum(p.numel() for p in head.parameters())
    # known answers printed by the reference's own training logs (logs/*/ep.txt:9)
    fp["param_count_logs"] = {"768": 1383400, "1024": 2106344, "1152": 2516968, "1664": 4487144, "4096": 21005288}
    args = Namespace(lr=0.4, min_lr=1e-6, warmup_epochs=10, epochs=90)
    opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=0.0)
    for e in [0.0, 0.5, 3.25, 9.999, 10.0, 10.5, 45.0, 89.99]:
        fp["lr_sched"].append([e, adjust_learning_rate(opt, e, args)])
    fp["lr_sched_args"] = vars(args)
    json.dump(fp, open(os.path.join(HERE, "fingerprints.json"), "w"), indent=1)
    print("wrote fingerprints")


if __name__ == "__main__":
    torch.set_num_threads(1)                       # bit-stable reductions
    run_case("small", B=4, N=19, D=64, M=8, K=10, d_out=1, qkv_bias=False)
    run_case("dout2_bias", B=3, N=7, D=64, M=4, K=5, d_out=2, qkv_bias=True, wd=1e-3)
    run_case("m32_sharp", B=5, N=33, D=128, M=32, K=12, d_out=1, qkv_bias=False, spread=2.0, q_gain=40.0)
    run_case("cls197", B=4, N=197, D=128, M=8, K=16, d_out=1, qkv_bias=False, q_gain=10.0)
    fingerprints()
