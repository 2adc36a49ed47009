"""A few eager training steps of the EP head at a BASELINE config, for ncu (tools/profile_round*.sh):
python tools/profile_step.py [config] [queries] [steps] [debug flags] [sm limit]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E

CFG = {"c1": (64, 197, 768), "c2": (1024, 257, 1024), "c3": (1024, 256, 1152), "c4": (1024, 730, 1664), "c5": (1024, 201, 4096)}
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sm_limit = int(sys.argv[5]) if len(sys.argv) > 5 else 0
persist_mb = int(sys.argv[6]) if len(sys.argv) > 6 else -1
B, N, D = CFG[cfg]
dev = "cuda:0"
if persist_mb >= 0:   # experiment: L2 set-aside for evict_last (persisting) lines
    import ctypes
    rt = ctypes.CDLL('libcudart.so')
    torch.cuda.init(); torch.zeros(1, device=dev)
    v = ctypes.c_int(0)
    rt.cudaDeviceGetAttribute(ctypes.byref(v), 108, 0)   # cudaDevAttrMaxPersistingL2CacheSize
    want = min(persist_mb << 20, v.value)
    rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(want))   # cudaLimitPersistingL2CacheSize
    print('max persisting L2', v.value >> 20, 'MB; set', want >> 20, 'MB rc', rc)
torch.manual_seed(0)
head = E.make_ep_head(D, M, 1000).to(dev)
tr = E.EPHeadTrainer(head, B, N, lr=0.1, use_graph=False)
xs = [torch.randn(B, N, D, device=dev).to(torch.bfloat16) for _ in range(3)]
y = torch.randint(0, 1000, (B,), device=dev)
E._lib.load().ep_set_debug(flags)
E._lib.load().ep_set_sm_limit(sm_limit)
for i in range(steps):
    tr.train_step(xs[i % 3], y)
torch.cuda.synchronize()
print("loss", float(tr.step_loss))
