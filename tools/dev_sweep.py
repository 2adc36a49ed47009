"""Developer sweep: per-kernel time of the one-pass fused kernels against the CTA count (ep_set_sm_limit) and a few
plan knobs.  Run on the GPU box: python tools/dev_sweep.py [config] [queries]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E

CFG = {"c1": (64, 197, 768), "c2": (1024, 257, 1024), "c3": (1024, 256, 1152), "c4": (1024, 730, 1664), "c5": (1024, 201, 4096),
       "t128": (2048, 128, 1024), "t256": (1024, 256, 1024), "t64": (4096, 64, 1024)}   # L2 experiments: same bytes, shorter samples
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
limits = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 132, 120, 112, 104, 96, 88, 80, 74]
flag_list = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
B, N, D = CFG[cfg]
dev = "cuda:0"
lib = E._lib.load()
torch.manual_seed(0)
head = E.make_ep_head(D, M, 1000).to(dev)
tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=False)
xs = [torch.randn(B, N, D, device=dev).to(torch.bfloat16) for _ in range(3)]
y = torch.randint(0, 1000, (B,), device=dev)
for flags in flag_list:
    for lim in limits:
        lib.ep_set_sm_limit(lim)
        lib.ep_set_debug(flags)
        for i in range(2):
            tr.train_step(xs[i], y)
        torch.cuda.synchronize()
        lib.ep_set_debug(32 | flags)
        E._lib.kernel_timings()
        for i in range(6):
            tr.train_step(xs[i % 3], y)
        torch.cuda.synchronize()
        lib.ep_set_debug(0)
        agg = {}
        for nm, us in E._lib.kernel_timings():
            agg.setdefault(nm, []).append(us)
        keep = {k: sum(v) / len(v) for k, v in agg.items()}
        tot = sum(keep.values())
        print(f"{cfg} M{M} flags={flags} sm_limit={lim}: " + ", ".join(f"{k} {v:.1f}" for k, v in keep.items() if "fused" in k or "ks" in k or "kp" in k or "pool" in k)
              + f" | all stages {tot:.1f}", flush=True)
lib.ep_set_sm_limit(0)
