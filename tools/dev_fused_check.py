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


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("trace", "sweep")):
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
        for lead in (0, 2, 3, 6):
            timings(1024, 257, 1024, 32, flags=(lead + 1) << 16)
        timings(1024, 256, 1152, 32)
        timings(1024, 256, 1152, 32, flags=1024)


def trace(B, N, D, M, bwd=False, flags=0):
    """Pipeline hand-off stamps of CTA 0 (ep_set_debug bit 11), printed in microseconds at 1.965 GHz."""
    import ctypes
    torch.manual_seed(0)
    head = E.make_ep_head(D, M, 1000).to(DEV)
    tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=False)
    x = torch.randn(B, N, D, device=DEV).to(torch.bfloat16)
    y = torch.randint(0, 1000, (B,), device=DEV)
    tr.train_step(x, y)
    torch.cuda.synchronize()
    lib.ep_set_debug(2048 | flags)
    lib.ep_set_sm_limit(int(os.environ.get("EP_SM_LIMIT", "0")))
    if bwd:
        tr._cx, tr._ct = x, y
        tr._part2()
    else:
        tr._cx, tr._ct = x, y
        tr._forward(True)
    torch.cuda.synchronize()
    lib.ep_set_debug(0)
    lib.ep_set_sm_limit(0)
    buf = (ctypes.c_longlong * 128)()
    lib.ep_debug_trace(ctypes.cast(buf, ctypes.c_void_p), 128)
    t = [buf[i] for i in range(128)]
    t0 = min(v for v in t[:112] if v > 0)
    names = {0: "P:wait_blocks", 1: "P:start", 3: "P:issued", 2: "L:issued", 4: "epi:pass1", 5: "epi:pass2_begin", 8: "epi:begin", 9: "epi:logits_ready",
             10: "epi:P_prev_done", 12: "epi:blocks_written", 13: "epi:drains_done"}
    print(f"trace {'bwd' if bwd else 'fwd'} B{B} N{N} D{D} M{M}")
    for i in range(2, 5):
        row = sorted((t[i * 16 + k], names[k]) for k in names if t[i * 16 + k] > 0)
        print(f"  sample {i}: " + "  ".join(f"{nm}@{(v - t0) / 1965.0:.1f}" for v, nm in row))
    print(f"  L producer: ring-full wait {t[120] / 1965.0:.1f} us of {t[121] / 1965.0:.1f};  P producer: {t[125] / 1965.0:.1f} of {t[126] / 1965.0:.1f};  "
          f"L warp: load wait {t[122] / 1965.0:.1f} us, order wait {t[123] / 1965.0:.1f} us of {t[124] / 1965.0:.1f};  "
          f"P warp: load wait {t[117] / 1965.0:.1f} us, block wait {t[118] / 1965.0:.1f} us of {t[119] / 1965.0:.1f}")
    print(f"  L warp issue {t[112] / 1965.0:.1f} us;  P warp: drain (pfree) wait {t[113] / 1965.0:.1f} us, issue {t[114] / 1965.0:.1f} us")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sweep":
    for lead in (2, 4, 6, 8):
        for pf in (0, 3, 6):
            print(f"lead {lead} pf {pf}", end=": ")
            timings(1024, 257, 1024, 32, flags=((lead + 1) << 16) | ((pf + 1) << 22) | (4 << 25))

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "trace":
    trace(1024, 257, 1024, 32)
    trace(1024, 257, 1024, 32, bwd=True)
