"""Developer check: tcgen05 kernel family (2) against the general family (1) and the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E
from oracle import ep_oracle as O

lib = E._lib.load()
dev = "cuda:0"
shapes = [(4, 257, 256, 8), (3, 130, 128, 8), (8, 197, 768, 8), (4, 257, 1024, 32), (2, 730, 1664, 32), (2, 201, 4096, 32),
          (5, 64, 384, 12), (2, 1, 128, 8), (150, 257, 1024, 32)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (B, N, D, M) in shapes:
    torch.manual_seed(0)
    pool = E.EfficientProbing(D, num_queries=M).to(dev)
    with torch.no_grad():
        pool.cls_token.mul_(15.0)
    x = O.synthetic_tokens(B, N, D, seed=1).to(dev)
    res = {}
    for fam in (1, 2):
        lib.ep_set_kernel_mode(fam)
        if lib.ep_kernel_family_for(0, B, N, D, M) != fam:
            print((B, N, D, M), "family", fam, "unsupported"); continue
        pool.zero_grad()
        out, attn = E.EPPoolFunction.apply(x, pool.cls_token, pool.v.weight, None, pool.scale, M, 1, True)
        g = torch.randn(out.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        (out * g).sum().backward()
        torch.cuda.synchronize()
        res[fam] = (out.detach().clone(), attn.clone(), pool.cls_token.grad.clone(), pool.v.weight.grad.clone())
    if 2 in res:
        xs = x[:4].cpu().double()
        o, a = O.ep_forward(xs, pool.cls_token.detach().cpu().double(), pool.v.weight.detach().cpu().double(), None,
                            pool.scale, M, 1, True)
        print((B, N, D, M), "fam2 vs fam1: out %.2e attn %.2e dcls %.2e dvw %.2e | fam2 vs oracle: out %.2e attn %.2e" % (
            O.rel_err(res[2][0].cpu(), res[1][0].cpu()), O.rel_err(res[2][1].cpu(), res[1][1].cpu()),
            O.rel_err(res[2][2].cpu(), res[1][2].cpu()), O.rel_err(res[2][3].cpu(), res[1][3].cpu()),
            O.rel_err(res[2][0][:4].cpu(), o), O.rel_err(res[2][1][:4].cpu(), a)), flush=True)
lib.ep_set_kernel_mode(0)
