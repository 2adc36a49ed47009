"""Developer tool: CUDA-event time of every C-ABI stage of one EP-head training step (no graph)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E
from efficient_probing_b200 import _lib
from efficient_probing_b200.optim import lars_launch

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1024); ap.add_argument("--N", type=int, default=257)
ap.add_argument("--D", type=int, default=1024); ap.add_argument("--M", type=int, default=32)
ap.add_argument("--K", type=int, default=1000); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--only", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load(); lib.ep_set_kernel_mode(a.mode); lib.ep_set_debug(a.debug)
torch.manual_seed(0)
head = E.make_ep_head(a.D, a.M, a.K).to(dev)
tr = E.EPHeadTrainer(head, a.B, a.N, use_graph=False)
B, N, D, M, K, Dp = a.B, a.N, a.D, a.M, a.K, a.D
pool = [torch.randn(B, N, D, device=dev).to(torch.bfloat16) for _ in range(6)]
tr.targets.copy_(torch.randint(0, K, (B,), device=dev))
s = lambda: _lib.stream_ptr(dev)
pl, fc, bn = head[0], head[2], head[1]
xt = 0
stages = {
 "ep_fwd": lambda x: lib.ep_fwd(x.data_ptr(), xt, pl.cls_token.data_ptr(), pl.v.weight.data_ptr(), None, float(pl.scale), B, N, D, M, 1, tr.out.data_ptr(), tr.S.data_ptr(), tr.rowmax.data_ptr(), tr.rowsum.data_ptr(), tr.P.data_ptr(), None, tr.ws.data_ptr(), tr.ws.numel(), s()),
 "bn_fwd": lambda x: lib.ep_bn_fwd(tr.out.data_ptr(), B, Dp, 1e-6, 0.1, 1, bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), tr.y.data_ptr(), tr.save_mean.data_ptr(), tr.save_invstd.data_ptr(), s()),
 "linear_fwd": lambda x: lib.ep_linear_fwd(tr.y.data_ptr(), fc.weight.data_ptr(), fc.bias.data_ptr(), B, Dp, K, tr.logits.data_ptr(), tr.lin_ws.data_ptr(), tr.lin_ws.numel(), s()),
 "ce": lambda x: lib.ep_ce_fwd_bwd(tr.logits.data_ptr(), tr.targets.data_ptr(), B, K, 1.0 / B, 1.0 / B, tr.step_loss.data_ptr(), tr.dlogits.data_ptr(), tr.correct.data_ptr(), s()),
 "linear_bwd": lambda x: lib.ep_linear_bwd(tr.dlogits.data_ptr(), tr.y.data_ptr(), fc.weight.data_ptr(), B, Dp, K, tr.g["fc_w"].data_ptr(), tr.g["fc_b"].data_ptr(), tr.dy.data_ptr(), tr.lin_ws.data_ptr(), tr.lin_ws.numel(), s()),
 "bn_bwd": lambda x: lib.ep_bn_bwd(tr.dy.data_ptr(), tr.y.data_ptr(), tr.save_invstd.data_ptr(), B, Dp, tr.dout.data_ptr(), s()),
 "bwd_proj": lambda x: lib.ep_bwd_proj(tr.dout.data_ptr(), tr.P.data_ptr(), tr.out.data_ptr(), pl.v.weight.data_ptr(), None, xt, B, N, D, M, 1, tr.g["v_w"].data_ptr(), None, tr.ws.data_ptr(), tr.ws.numel(), s()),
 "bwd_pool": lambda x: lib.ep_bwd_pool(x.data_ptr(), xt, pl.cls_token.data_ptr(), float(pl.scale), B, N, D, M, 1, tr.S.data_ptr(), tr.rowmax.data_ptr(), tr.rowsum.data_ptr(), tr.g["cls"].data_ptr(), tr.ws.data_ptr(), tr.ws.numel(), s()),
 "lars": lambda x: (lars_launch(tr.params, tr.grads, tr.mus, tr.trust, tr.hyper, tr.lars_scratch), 0)[1],
}
res = {}
for name, fn in stages.items():
    if a.only and name not in a.only.split(","):
        continue
    for i in range(3):
        _lib.check(fn(pool[i % 6]), name)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for i, (e0, e1) in enumerate(ev):
        e0.record(); _lib.check(fn(pool[i % 6]), name); e1.record()
    torch.cuda.synchronize()
    res[name] = round(sum(e0.elapsed_time(e1) for e0, e1 in ev) / a.iters * 1e3, 1)
res["total_us"] = round(sum(res.values()), 1)
if a.debug & 32:
    agg = {}
    for nm, us in _lib.kernel_timings():
        agg.setdefault(nm, []).append(us)
    res["kernels_us"] = {k: round(sum(v[-a.iters:]) / len(v[-a.iters:]), 1) for k, v in agg.items()}
res["family"] = lib.ep_last_kernel_family()
res["shape"] = [B, N, D, M, K]
print(json.dumps(res))
