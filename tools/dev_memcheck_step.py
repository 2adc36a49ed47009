"""Developer tool: one training step (no graph) on a small tcgen05-eligible shape -- run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, N, D, M, K) in [(64, 70, 256, 8, 40), (8, 257, 1024, 32, 16)]:
    head = E.make_ep_head(D, M, K).to(dev)
    tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=False)
    x = torch.randn(B, N, D, device=dev).to(torch.bfloat16)
    y = torch.randint(0, K, (B,), device=dev)
    tr.train_step(x, y)
    torch.cuda.synchronize()
    print((B, N, D, M, K), "family", E._lib.load().ep_last_kernel_family(), "loss", float(tr.step_loss), flush=True)
