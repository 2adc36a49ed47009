"""Developer check of the TF32 tensor-core GEMM wrappers against torch (fp64) on random data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_probing_b200 as E
from efficient_probing_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
s = lambda: _lib.stream_ptr(dev)
for (B, F, K) in [(1024, 1024, 1000), (64, 128, 40), (200, 256, 16)]:
    y = torch.randn(B, F, device=dev); W = torch.randn(K, F, device=dev) * 0.1; b = torch.randn(K, device=dev)
    dl = torch.randn(B, K, device=dev)
    ws = torch.empty(lib.ep_linear_workspace_bytes(B, F, K), dtype=torch.uint8, device=dev)
    logits = torch.empty(B, K, device=dev); dW = torch.empty(K, F, device=dev); db = torch.empty(K, device=dev); dy = torch.empty(B, F, device=dev)
    _lib.check(lib.ep_linear_fwd(y.data_ptr(), W.data_ptr(), b.data_ptr(), B, F, K, logits.data_ptr(), ws.data_ptr(), ws.numel(), s()), "fwd")
    _lib.check(lib.ep_linear_bwd(dl.data_ptr(), y.data_ptr(), W.data_ptr(), B, F, K, dW.data_ptr(), db.data_ptr(), dy.data_ptr(), ws.data_ptr(), ws.numel(), s()), "bwd")
    torch.cuda.synchronize()
    print((B, F, K), "logits %.2e dW %.2e db %.2e dy %.2e" % (rel(logits, y.double() @ W.double().T + b.double()),
          rel(dW, dl.double().T @ y.double()), rel(db, dl.double().sum(0)), rel(dy, dl.double() @ W.double())), flush=True)
def pooled(P, xt, B, N, D, M, d_out):
    if lib.ep_pooled_layout(xt, B, N, D, M, d_out) == 1:
        hl = P.view(torch.bfloat16).reshape(B, M, 2, D).double()
        return hl[:, :, 0] + hl[:, :, 1]
    return P.double()
for (B, N, D, M, d_out, xt) in [(256, 5, 256, 8, 1, 1), (130, 3, 128, 32, 1, 1), (64, 4, 256, 8, 2, 1), (256, 70, 256, 8, 1, 0),
                                (128, 257, 1024, 32, 1, 0), (64, 197, 768, 12, 1, 0), (192, 100, 512, 8, 2, 0), (64, 40, 1152, 32, 1, 0)]:
    Dp = D // d_out; c = Dp // M
    x = torch.randn(B, N, D, device=dev).to(torch.bfloat16 if xt == 0 else torch.float32)
    cls = torch.randn(M, D, device=dev) * 0.5; W = torch.randn(Dp, D, device=dev) * 0.1
    out = torch.empty(B, Dp, device=dev); S = torch.empty(B, M, N, device=dev); rm = torch.empty(B, M, device=dev); rs = torch.empty(B, M, device=dev)
    P = torch.empty(B, M, D, device=dev)
    ws = torch.empty(lib.ep_workspace_bytes(B, N, D, M, d_out), dtype=torch.uint8, device=dev)
    _lib.check(lib.ep_fwd(x.data_ptr(), xt, cls.data_ptr(), W.data_ptr(), None, D ** -0.5, B, N, D, M, d_out, out.data_ptr(), S.data_ptr(), rm.data_ptr(), rs.data_ptr(), P.data_ptr(), None, ws.data_ptr(), ws.numel(), s()), "ep_fwd")
    g = torch.randn(B, Dp, device=dev); dvw = torch.empty(Dp, D, device=dev)
    _lib.check(lib.ep_bwd_proj(g.data_ptr(), P.data_ptr(), out.data_ptr(), W.data_ptr(), None, xt, B, N, D, M, d_out, dvw.data_ptr(), None, ws.data_ptr(), ws.numel(), s()), "bwd_proj")
    torch.cuda.synchronize()
    Wm = W.double().reshape(M, c, D)
    Pd = pooled(P, xt, B, N, D, M, d_out)
    A = torch.softmax(torch.einsum("md,bnd->bmn", cls.double(), x.double()) * D ** -0.5, -1)
    ref_P = torch.einsum("bmn,bnd->bmd", A, x.double())
    ref_out = torch.einsum("mjc,bmc->bmj", Wm, Pd).reshape(B, Dp)
    ref_dvw = torch.einsum("bmj,bmc->mjc", g.double().reshape(B, M, c), Pd).reshape(Dp, D)
    # dP sits at the start of the pooling part of the workspace: find it via the known layout (w_r, g_r first)
    au = lambda v: (v + 255) // 256 * 256
    off = au(3 * D * D * 4) + au(3 * B * D * 4) + au(3 * B * D * 2)  # dP (fp32 path, x_dtype=1 -> general family)
    lay = lib.ep_pooled_layout(xt, B, N, D, M, d_out)
    msg = "layout %d P %.2e out %.2e d_v_w %.2e" % (lay, rel(Pd, ref_P), rel(out, ref_out), rel(dvw, ref_dvw))
    if xt == 1:
        dP = ws[off: off + B * M * D * 4].view(torch.float32).reshape(B, M, D)
        ref_dP = torch.einsum("bmj,mjc->bmc", g.double().reshape(B, M, c), Wm)
        msg += " dP %.2e" % rel(dP, ref_dP)
    print((B, N, D, M, d_out, xt), msg, flush=True)
