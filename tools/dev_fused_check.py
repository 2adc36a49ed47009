"""Developer check of the one-pass fused kernels: against the four-kernel path (ep_set_debug(1024)) and the oracle,
then per-kernel timings at full size.  Run on the GPU box: python tools/dev_fused_check.py [quick]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import efficient_probing_b200 as E
from oracle import ep_oracle as O

DEV = "cuda:0"
lib = E._lib.load()


def step(B, N, D, M, K, flags, x, y, p):
    from test_parity_gpu import head_from_params
    lib.ep_set_debug(flags)
    head = head_from_params(p, K)
    tr = E.EPHeadTrainer(head, B, N, lr=0.0, use_graph=False)
    tr.train_step(x, y)
    torch.cuda.synchronize()
    lib.ep_set_debug(0)
    return tr


def compare(B, N, D, M, K=100, gain=20.0, oracle=False):
    p = O.build_head(D, M, K, seed=0)
    p.cls_token = p.cls_token * gain
    x = O.synthetic_tokens(B, N, D, seed=5)
    y = O.synthetic_labels(B, K)
    xg, yg = x.to(DEV), y.to(DEV)
    a = step(B, N, D, M, K, 0, xg, yg, p)
    b = step(B, N, D, M, K, 1024, xg, yg, p)
    errs = {k: O.rel_err(getattr(a, k).cpu(), getattr(b, k).cpu()) for k in ("out", "S", "rowmax", "rowsum", "logits")}
    errs["d_cls"] = O.rel_err(a.g["cls"].cpu(), b.g["cls"].cpu())
    errs["d_v_w"] = O.rel_err(a.g["v_w"].cpu(), b.g["v_w"].cpu())
    msg = f"B{B} N{N} D{D} M{M}: fused vs 4-kernel " + " ".join(f"{k} {v:.1e}" for k, v in errs.items())
    if oracle:
        ref = O.head_loss_and_grads_pooled(p, x, y, dtype=torch.float64)
        msg += (f" | vs oracle out {O.rel_err(a.out.cpu(), ref['out']):.1e} d_cls {O.rel_err(a.g['cls'].cpu().reshape(-1), ref['grad.0.cls_token'].reshape(-1)):.1e}"
                f" d_v_w {O.rel_err(a.g['v_w'].cpu().reshape(-1), ref['grad.0.v.weight'].reshape(-1)):.1e}")
    print(msg, flush=True)
    return max(errs.values())


def timings(B, N, D, M, K=1000, flags=0, reps=5):
    torch.manual_seed(0)
    head = E.make_ep_head(D, M, K).to(DEV)
    tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=False)
    xs = [torch.randn(B, N, D, device=DEV).to(torch.bfloat16) for _ in range(3)]
    y = torch.randint(0, K, (B,), device=DEV)
    lib.ep_set_debug(flags)
    for i in range(2):
        tr.train_step(xs[i], y)
    torch.cuda.synchronize()
    lib.ep_set_debug(32 | flags)
    E._lib.kernel_timings()
    for i in range(reps):
        tr.train_step(xs[i % 3], y)
    torch.cuda.synchronize()
    lib.ep_set_debug(0)
    agg = {}
    for nm, us in E._lib.kernel_timings():
        agg.setdefault(nm, []).append(us)
    print(f"timings B{B} N{N} D{D} M{M} flags={flags}: " + ", ".join(f"{k} {sum(v)/len(v):.1f}" for k, v in agg.items()), flush=True)


if __name__ == "__main__":
    worst = 0.0
    worst = max(worst, compare(64, 257, 1024, 32, oracle=True))
    worst = max(worst, compare(64, 197, 768, 8, oracle=True))
    worst = max(worst, compare(64, 256, 1152, 32))
    worst = max(worst, compare(320, 70, 256, 8))
    worst = max(worst, compare(192, 257, 1024, 32))
    worst = max(worst, compare(20, 129, 256, 16))
    worst = max(worst, compare(7, 128, 128, 64))
    worst = max(worst, compare(5, 1, 128, 8))
    print("worst fused-vs-4-kernel error", worst)
    if len(sys.argv) < 2:
        for M in (32, 8):
            timings(1024, 257, 1024, M)
            timings(1024, 257, 1024, M, flags=1024)
        for lead in (0, 1, 3, 4):
            timings(1024, 257, 1024, 32, flags=(lead + 1) << 16)
        timings(1024, 256, 1152, 32)
        timings(1024, 256, 1152, 32, flags=1024)
