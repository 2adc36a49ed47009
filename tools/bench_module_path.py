"""The drop-in MODULE path, timed: what main_linprobe.py / engine_finetune.py run when only poolings.ep.EfficientProbing
is swapped for efficient_probing_b200.EfficientProbing -- Sequential(EP, BatchNorm1d, Linear) under torch autograd,
nn.CrossEntropyLoss, the package's LARS (util/lars.py semantics) -- against EPHeadTrainer's captured step on the same
batch.  python tools/bench_module_path.py [config] [queries]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E

CFG = {"c1": (64, 197, 768), "c2": (1024, 257, 1024), "c3": (1024, 256, 1152), "c4": (1024, 730, 1664), "c5": (1024, 201, 4096)}
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B, N, D = CFG[cfg]
K, dev = 1000, "cuda:0"
torch.manual_seed(0)
head = E.make_ep_head(D, M, K).to(dev)
opt = E.optim.LARS(head.parameters(), lr=0.1, weight_decay=0.0)
crit = torch.nn.CrossEntropyLoss()
xs = [torch.randn(B, N, D, device=dev).to(torch.bfloat16) for _ in range(4)]
y = torch.randint(0, K, (B,), device=dev)


def step(i):
    out = head(xs[i % 4])
    loss = crit(out, y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for i in range(5):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 30
e0.record()
for i in range(steps):
    step(i)
e1.record()
torch.cuda.synchronize()
ms_module = e0.elapsed_time(e1) / steps

head2 = E.make_ep_head(D, M, K).to(dev)
tr = E.EPHeadTrainer(head2, B, N, lr=0.1)
for i in range(4):
    tr.register_batch(xs[i], y)
    tr.prepare(xs[i], y)
for i in range(5):
    tr.train_step(xs[i % 4], y)
torch.cuda.synchronize()
e0.record()
for i in range(steps):
    tr.train_step(xs[i % 4], y)
e1.record()
torch.cuda.synchronize()
ms_trainer = e0.elapsed_time(e1) / steps
print(json.dumps({"config": cfg, "queries": M, "per_gpu_batch": B, "module_path_ms_per_step": ms_module,
                  "module_path_tokens_per_s": B * N / (ms_module * 1e-3), "trainer_ms_per_step": ms_trainer,
                  "note": "module path = nn.Sequential(EfficientProbing[libep_b200 autograd Function], BatchNorm1d, Linear) + "
                          "CrossEntropyLoss + LARS under eager torch autograd (cuBLAS / ATen for everything but the pooling); "
                          "trainer = EPHeadTrainer's captured graph"}))
