"""GPU parity tests: the CUDA path (through the C ABI, via the drop-in module and the fused trainer)
against the committed outputs of the reference (tests/golden) and against the CPU oracle on seeded
inputs.  Tolerances are BASELINE.json's: attention maps and logits 1e-3, parameter gradients 2e-3
(relative L2 error, fp32 accumulate)."""
import pytest
import torch
from torch import nn

import efficient_probing_b200 as E
from oracle import ep_oracle as O
from conftest import Golden, GOLDEN_CASES

pytestmark = pytest.mark.gpu
TOL_FWD, TOL_GRAD = 1e-3, 2e-3
DEV = "cuda:0"


@pytest.fixture(params=[1, 2], ids=["general", "tcgen05"])
def family(request):
    """Force one kernel family of the pooling (ep_set_kernel_mode) for the duration of a test."""
    lib = E._lib.load()
    lib.ep_set_kernel_mode(request.param)
    yield request.param
    lib.ep_set_kernel_mode(0)


def require_family(family, B, N, D, M, dtype=torch.bfloat16):
    if E._lib.load().ep_kernel_family_for(0 if dtype == torch.bfloat16 else 1, B, N, D, M) != family:
        pytest.skip("shape not covered by the tcgen05 kernels (D % 128 != 0 or fp32 tokens)")


def head_from_params(p: O.EPParams, K):
    D = p.cls_token.shape[2]
    h = E.make_ep_head(D, p.num_queries, K, d_out=p.d_out, qkv_bias=p.v_bias is not None)
    sd = {"0.cls_token": p.cls_token, "0.v.weight": p.v_weight, "1.running_mean": p.running_mean,
          "1.running_var": p.running_var, "1.num_batches_tracked": torch.tensor(p.num_batches_tracked),
          "2.weight": p.fc_weight, "2.bias": p.fc_bias}
    if p.v_bias is not None:
        sd["0.v.bias"] = p.v_bias
    h.load_state_dict(sd)                              # reference checkpoint keys load unchanged
    return h.to(DEV)


def close(got, ref, tol, what):
    got, ref = got.detach(), ref.detach().to(got.device)
    err = (got.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)
    if ref.double().norm() < 1e-12:                    # identically-zero reference (bias under BatchNorm)
        assert got.double().norm() < 1e-6, what
        return
    assert float(err) < tol, f"{what}: rel err {float(err):.3e} > {tol}"


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float32])
def test_module_forward_backward_vs_reference_golden(name, xdt):
    g = Golden(name)
    head = head_from_params(g.params(), g.meta["K"])
    head.train()
    x = g.t("x").to(DEV).to(xdt)                       # fixture tokens are bf16-representable
    y = g.t("targets").to(DEV)
    pooled = head[0](x)
    logits = head[2](head[1](pooled))
    loss = nn.CrossEntropyLoss()(logits, y)
    loss.backward()
    close(pooled, g.t("f64.out"), TOL_FWD, "out")
    close(logits, g.t("f64.logits"), TOL_FWD, "logits")
    close(head[0].attention_maps(x), g.t("f64.attn"), TOL_FWD, "attn")
    close(E.ep_attention(x[0], head[0].cls_token[0]), g.t("f64.attn")[0], TOL_FWD, "ep_attention")
    assert abs(float(loss) - float(g.z["f64.loss"])) < 1e-3 * abs(float(g.z["f64.loss"]))
    for k, p in head.named_parameters():
        close(p.grad, g.t("f64.grad." + k), TOL_GRAD, "grad " + k)
    close(head[1].running_mean, g.t("f64.running_mean"), TOL_FWD, "running_mean")
    close(head[1].running_var, g.t("f64.running_var"), TOL_FWD, "running_var")


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("graph", [False, True])
def test_trainer_two_lars_steps_vs_reference_golden(name, graph):
    g = Golden(name)
    m = g.meta
    head = head_from_params(g.params(), m["K"])
    x = g.t("x").to(DEV).to(torch.bfloat16)
    y = g.t("targets").to(DEV)
    tr = E.EPHeadTrainer(head, m["B"], m["N"], lr=m["lr"], weight_decay=m["weight_decay"], use_graph=graph)
    tr.train_step(x, y)
    torch.cuda.synchronize()
    assert abs(float(tr.step_loss) - float(g.z["f64.loss"])) < 1e-3 * abs(float(g.z["f64.loss"]))
    close(tr.logits, g.t("f64.logits"), TOL_FWD, "logits")
    for k, grad in zip([n for n, _ in head.named_parameters()], tr.grads):
        close(grad.reshape(-1), g.t("f64.grad." + k).reshape(-1), TOL_GRAD, "grad " + k)
    tr.train_step(x, y)
    torch.cuda.synchronize()
    assert abs(float(tr.step_loss) - float(g.z["f64.loss_step2"])) < 1e-3 * abs(float(g.z["f64.loss_step2"]))
    for k, p in head.named_parameters():
        close(p, g.t("f64.after2." + k), 1e-3, "param after 2 LARS steps " + k)
    assert int(head[1].num_batches_tracked) == 2
    # eval-mode logits on running statistics (engine_finetune.py:106-166)
    head2 = head_from_params(g.params(), m["K"])
    head2[1].running_mean.copy_(g.t("f64.running_mean").float())
    head2[1].running_var.copy_(g.t("f64.running_var").float())
    tr2 = E.EPHeadTrainer(head2, m["B"], m["N"], use_graph=False)
    close(tr2.eval_logits(x), g.t("f64.eval_logits"), TOL_FWD, "eval logits")


def test_auto_mode_prefers_tcgen05_for_baseline_shapes():
    lib = E._lib.load()
    lib.ep_set_kernel_mode(0)
    for (N, D) in [(197, 768), (257, 1024), (256, 1152), (730, 1664), (201, 4096)]:
        for M in (8, 32):
            assert lib.ep_kernel_family_for(0, 1024, N, D, M) == 2, (N, D, M)
    assert lib.ep_kernel_family_for(1, 8, 197, 768, 8) == 1       # fp32 tokens: general kernels
    assert lib.ep_kernel_family_for(0, 8, 50, 72, 8) == 1         # D % 128 != 0: general kernels


CASES = [  # B, N, D, M, K, d_out, bias, spread, q_gain
    (8, 197, 768, 8, 1000, 1, False, 1.0, 1.0),        # BASELINE config 1 shape (smaller batch)
    (4, 257, 1024, 32, 1000, 1, False, 1.0, 25.0),     # config 2 shape, sharpened attention
    (3, 256, 1152, 32, 100, 1, False, 1.0, 1.0),       # config 3 shape
    (2, 730, 1664, 32, 64, 1, False, 1.0, 10.0),       # config 4 shape (long N)
    (2, 201, 4096, 32, 32, 1, False, 1.0, 10.0),       # config 5 shape (wide D)
    (5, 50, 384, 12, 10, 2, True, 8.0, 20.0),          # d_out=2, bias, M not a power of two, logits ~ +-30
    (1, 1, 64, 8, 4, 1, False, 1.0, 1.0),              # single token: attention == 1
    (3, 1370, 768, 8, 10, 1, False, 1.0, 5.0),         # Franca@518 token count
    (150, 129, 256, 16, 10, 1, False, 1.0, 10.0),      # more samples than SMs, one token past a tile edge
    (7, 128, 128, 64, 10, 1, False, 1.0, 10.0),        # M = 64 (widest operand), N exactly one tile
    (64, 40, 1152, 32, 10, 1, False, 1.0, 10.0),       # B % 64 == 0 with c = 36 (c % 8 != 0): hi/lo P, ragged column tiles
    (128, 20, 256, 8, 16, 2, True, 1.0, 10.0),         # hi/lo P with d_out = 2 and bias
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_N%d_D%d_M%d_K%d_do%d_b%d" % c[:7])
def test_seeded_shapes_vs_oracle(case, family):
    B, N, D, M, K, d_out, bias, spread, q_gain = case
    require_family(family, B, N, D, M)
    p = O.build_head(D, M, K, d_out=d_out, qkv_bias=bias, seed=0)
    p.cls_token = p.cls_token * q_gain
    x = O.synthetic_tokens(B, N, D, seed=1234, spread=spread)
    y = O.synthetic_labels(B, K)
    if B > 1:
        ref = O.head_loss_and_grads(p, x.float(), y, dtype=torch.float64)
    else:
        o, a = O.ep_forward(x.double(), p.cls_token.double(), p.v_weight.double(), None, p.scale, M, d_out, True)
        ref = {"out": o, "attn": a}
    head = head_from_params(p, K)
    head.train()
    xg, yg = x.to(DEV), y.to(DEV)
    pooled = head[0](xg)
    close(pooled, ref["out"], TOL_FWD, "out")
    close(head[0].attention_maps(xg), ref["attn"], TOL_FWD, "attn")
    if B == 1:                                         # BatchNorm1d cannot train on one sample (torch raises too)
        return
    logits = head[2](head[1](pooled))
    nn.CrossEntropyLoss()(logits, yg).backward()
    if B > 1:
        close(logits, ref["logits"], TOL_FWD, "logits")
        for k, prm in head.named_parameters():
            close(prm.grad, ref["grad." + k], TOL_GRAD, "grad " + k)
    # the fused trainer computes the same step
    head_t = head_from_params(p, K)
    tr = E.EPHeadTrainer(head_t, B, N, lr=0.0, use_graph=False)
    tr.train_step(xg, yg)
    assert E._lib.load().ep_last_kernel_family() == family
    if B > 1:
        close(tr.logits, ref["logits"], TOL_FWD, "trainer logits")
        for k, grad in zip([n for n, _ in head_t.named_parameters()], tr.grads):
            close(grad.reshape(-1), ref["grad." + k].reshape(-1), TOL_GRAD, "trainer grad " + k)


def test_lars_optimizer_matches_oracle():
    torch.manual_seed(3)
    shapes = [(1, 8, 64), (64, 64), (64,), (10, 64), (10,)]
    ps = [torch.randn(s) for s in shapes]
    gs = [torch.randn(s) * 0.1 for s in shapes]
    params = [nn.Parameter(p.clone().to(DEV)) for p in ps]
    opt = E.LARS(params, lr=0.3, weight_decay=1e-2)
    mus = [torch.zeros_like(p) for p in ps]
    cur = [p.double() for p in ps]
    mus = [m.double() for m in mus]
    for it in range(3):
        for p, g in zip(params, gs):
            p.grad = (g * (it + 1)).to(DEV)
        opt.step()
        cur, mus = O.lars_step(cur, [g.double() * (it + 1) for g in gs], mus, lr=0.3, weight_decay=1e-2)
    for p, c in zip(params, cur):
        close(p.detach(), c, 1e-5, "lars")
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"mu"}           # same optimizer-state layout as util/lars.py


def test_errors_are_loud():
    head = E.make_ep_head(64, 8, 10).to(DEV)
    with pytest.raises(TypeError):
        head[0](torch.randn(2, 5, 64, device=DEV, dtype=torch.float16))
    with pytest.raises(ValueError):
        head[0](torch.zeros(2, 5, 72, device=DEV))        # cls_token is (1, 8, 64)
    with pytest.raises(RuntimeError):
        head[0](torch.zeros(2, 5, 64, device=DEV), cls=torch.zeros(2, 4, 64, device=DEV))   # not (B, M, C): ep.py:35
    assert E._lib.load().ep_device_check() == 0


@pytest.mark.parametrize("B", [64, 20], ids=["hilo_rows", "fp32"])
def test_saved_pooled_tokens_vs_oracle(B):
    """The saved tensor P of ep_fwd, in whichever layout ep_pooled_layout announces (bf16 hi/lo rows when the
    tcgen05 GEMMs consume it, else fp32), against the oracle's pooled tokens; and the saved logits / statistics."""
    lib = E._lib.load()
    N, D, M = 70, 256, 8
    p = O.build_head(D, M, 10, seed=0)
    p.cls_token = p.cls_token * 10.0
    x = O.synthetic_tokens(B, N, D, seed=21)
    _, attn, P_ref, rmax_ref, rsum_ref = O.ep_forward_pooled(x.double(), p.cls_token.double(), p.v_weight.double(), None,
                                                            p.scale, M, 1)
    xg = x.to(DEV)
    out, S, rmax, rsum = (torch.empty(s_, device=DEV) for s_ in ((B, D), (B, M, N), (B, M), (B, M)))
    P = torch.empty(B, M, D, device=DEV)
    ws = torch.empty(lib.ep_workspace_bytes(B, N, D, M, 1), dtype=torch.uint8, device=DEV)
    cls, w = p.cls_token.to(DEV).contiguous(), p.v_weight.to(DEV).contiguous()
    E._lib.check(lib.ep_fwd(xg.data_ptr(), 0, cls.data_ptr(), w.data_ptr(), None, float(p.scale), B, N, D, M, 1,
                            out.data_ptr(), S.data_ptr(), rmax.data_ptr(), rsum.data_ptr(), P.data_ptr(), None,
                            ws.data_ptr(), ws.numel(), E._lib.stream_ptr(torch.device(DEV))), "ep_fwd")
    layout = lib.ep_pooled_layout(0, B, N, D, M, 1)
    assert layout == (1 if B % 64 == 0 else 0)
    if layout == 1:
        hl = P.view(torch.bfloat16).reshape(B, M, 2, D).double()
        assert float(hl[:, :, 1].abs().max()) <= float(hl[:, :, 0].abs().max()) * 2.0 ** -7     # lo is a correction
        P_got = hl[:, :, 0] + hl[:, :, 1]
    else:
        P_got = P.double()
    close(P_got, P_ref, 2e-5, "P")
    close(rmax, rmax_ref, 1e-5, "rowmax")
    close(rsum, rsum_ref, 1e-5, "rowsum")
    close(torch.softmax(S.double(), -1), attn, 1e-5, "softmax(S)")


def test_fused_and_separate_softmax_agree():
    """ep_set_debug(512) splits the forward softmax out of the logit kernel: both paths must give the same
    pooled output, saved statistics and attention-derived gradients."""
    lib = E._lib.load()
    B, N, D, M = 20, 257, 1024, 32
    p = O.build_head(D, M, 10, seed=0)
    p.cls_token = p.cls_token * 25.0
    x = O.synthetic_tokens(B, N, D, seed=3).to(DEV)
    g = torch.randn(B, D, device=DEV)
    res = []
    for flags in (0, 512):
        lib.ep_set_debug(flags)
        try:
            pool = E.EfficientProbing(D, num_queries=M).to(DEV)
            with torch.no_grad():
                pool.cls_token.copy_(p.cls_token); pool.v.weight.copy_(p.v_weight)
            out = pool(x)
            out.backward(g)
            res.append((out.detach(), pool.cls_token.grad.clone(), pool.v.weight.grad.clone()))
        finally:
            lib.ep_set_debug(0)
    assert lib.ep_last_kernel_family() == 2
    for a, b_, what in zip(res[0], res[1], ("out", "d cls_token", "d v.weight")):
        close(a, b_, 1e-5, what)


DX_CASES = [  # B, N, D, M, d_out, bias, q_gain
    (4, 257, 1024, 32, 1, False, 25.0),                # config 2 shape (tcgen05 forward, hi/lo P when B % 64 == 0)
    (64, 70, 256, 8, 1, False, 10.0),                  # B % 64 == 0: P saved as bf16 hi/lo rows by the forward
    (5, 50, 384, 12, 2, True, 20.0),                   # d_out = 2, bias, M not a power of two
    (3, 33, 128, 64, 1, False, 10.0),                  # M = 64
    (2, 19, 72, 6, 1, False, 5.0),                     # M % 4 != 0, D % 128 != 0
]


@pytest.mark.parametrize("case", DX_CASES, ids=lambda c: "B%d_N%d_D%d_M%d_do%d_b%d" % c[:6])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_input_gradient_vs_oracle(case, xdt):
    """--finetuning (main_linprobe.py:152-154): dL/dx through the pooling, against the closed form of the
    oracle and (fp32) against autograd through the reference formulation."""
    B, N, D, M, d_out, bias, q_gain = case
    p = O.build_head(D, M, 10, d_out=d_out, qkv_bias=bias, seed=0)
    p.cls_token = p.cls_token * q_gain
    x = O.synthetic_tokens(B, N, D, seed=99)
    g = torch.randn(B, D // d_out, generator=torch.Generator().manual_seed(5))
    xr = x.double().requires_grad_(True)
    prm = [t.double().requires_grad_(True) for t in (p.cls_token, p.v_weight)]
    bref = p.v_bias.double().requires_grad_(True) if bias else None
    O.ep_forward(xr, prm[0], prm[1], bref, p.scale, M, d_out).backward(g.double())
    out_p, attn, P, _, _ = O.ep_forward_pooled(x.double(), p.cls_token.double(), p.v_weight.double(),
                                              p.v_bias.double() if bias else None, p.scale, M, d_out)
    cf = O.ep_backward_pooled(x.double(), p.cls_token.double(), p.v_weight.double(), p.scale, M, d_out, attn, P,
                              g.double(), want_dx=True)
    assert O.rel_err(cf["d_x"], xr.grad) < 1e-10         # the oracle's closed form == autograd of ep.py:28-47
    pool = E.EfficientProbing(D, num_queries=M, d_out=d_out, qkv_bias=bias).to(DEV)
    with torch.no_grad():
        pool.cls_token.copy_(p.cls_token); pool.v.weight.copy_(p.v_weight)
        if bias:
            pool.v.bias.copy_(p.v_bias)
    xg = x.to(DEV, xdt).requires_grad_(True)
    pool(xg).backward(g.to(DEV))
    assert xg.grad.dtype == xdt and xg.grad.shape == xg.shape
    # bf16 tokens get a bf16 gradient: 2^-9 rounding per element on top of the fp32-accumulated value
    close(xg.grad.float(), xr.grad, 2e-4 if xdt == torch.float32 else 3e-3, "dx")
    close(pool.cls_token.grad, prm[0].grad, TOL_GRAD, "d cls_token")
    close(pool.v.weight.grad, prm[1].grad, TOL_GRAD, "d v.weight")
    if bias:
        close(pool.v.bias.grad, bref.grad, TOL_GRAD, "d v.bias")


@pytest.mark.parametrize("case", DX_CASES[1:], ids=lambda c: "B%d_N%d_D%d_M%d_do%d_b%d" % c[:6])
def test_external_queries_vs_reference_formulation(case):
    """forward(x, cls=...) (ep.py:32-33): per-sample queries replace cls_token; gradients flow to them."""
    B, N, D, M, d_out, bias, q_gain = case
    p = O.build_head(D, M, 10, d_out=d_out, qkv_bias=bias, seed=0)
    x = O.synthetic_tokens(B, N, D, seed=7)
    gen = torch.Generator().manual_seed(11)
    q = torch.randn(B, M, D, generator=gen) * 0.02 * q_gain
    g = torch.randn(B, D // d_out, generator=gen)
    xr, qr = x.double().requires_grad_(True), q.double().requires_grad_(True)
    wr = p.v_weight.double().requires_grad_(True)
    bref = p.v_bias.double().requires_grad_(True) if bias else None
    ref_out = O.ep_forward(xr, qr, wr, bref, p.scale, M, d_out)
    ref_out.backward(g.double())
    pool = E.EfficientProbing(D, num_queries=M, d_out=d_out, qkv_bias=bias).to(DEV)
    with torch.no_grad():
        pool.v.weight.copy_(p.v_weight)
        if bias:
            pool.v.bias.copy_(p.v_bias)
    for want_dx in (False, True):
        pool.zero_grad()
        xg = x.to(DEV).requires_grad_(want_dx)               # bf16 tokens
        qg = q.to(DEV).requires_grad_(True)
        out = pool(xg, cls=qg)
        close(out, ref_out.detach(), TOL_FWD, "out")
        out.backward(g.to(DEV))
        close(qg.grad, qr.grad, TOL_GRAD, "d cls (per sample)")
        close(pool.v.weight.grad, wr.grad, TOL_GRAD, "d v.weight")
        assert pool.cls_token.grad is None                   # the learned queries are unused, as in the reference
        if want_dx:
            close(xg.grad.float(), xr.grad, 3e-3, "dx")


def test_host_buffer_step_matches_device_step():
    """train_step_host (pinned host buffers, prefetch of the next batch) == train_step on resident tensors."""
    B, N, D, M, K = 16, 65, 128, 8, 24
    xs = [O.synthetic_tokens(B, N, D, seed=40 + i) for i in range(3)]
    ys = [O.synthetic_labels(B, K, seed=50 + i) for i in range(3)]
    torch.manual_seed(0)
    h1 = E.make_ep_head(D, M, K).to(DEV)
    torch.manual_seed(0)
    h2 = E.make_ep_head(D, M, K).to(DEV)
    t1 = E.EPHeadTrainer(h1, B, N, lr=0.5, use_graph=True)
    t2 = E.EPHeadTrainer(h2, B, N, lr=0.5, use_graph=True)
    hx = [x.pin_memory() for x in xs]
    hy = [y.pin_memory() for y in ys]
    for i in range(3):
        t1.train_step(xs[i].to(DEV), ys[i].to(DEV))
        l1 = float(t1.step_loss)
        nxt = (i + 1) % 3
        l2 = t2.train_step_host(hx[i], hy[i], next_x_host=hx[nxt], next_targets_host=hy[nxt])
        assert abs(l1 - l2) <= 1e-6 * abs(l1), (i, l1, l2)
    l3 = t2.train_step_host(xs[0], ys[0])                 # pageable host tensors are staged through pinned memory
    t1.train_step(xs[0].to(DEV), ys[0].to(DEV))
    assert abs(float(t1.step_loss) - l3) <= 1e-6 * abs(l3)
    for p1, p2 in zip(h1.parameters(), h2.parameters()):
        assert torch.equal(p1, p2)


@pytest.mark.parametrize("opt", ["adamw", "sgd", "sgd_nomom"])
def test_other_optimizers_match_torch(opt):
    """main_linprobe.py:403-408 builds AdamW or SGD besides LARS: the fused kernels follow torch.optim."""
    B, N, D, M, K = 16, 33, 128, 8, 10
    p = O.build_head(D, M, K, seed=0)
    p.cls_token = p.cls_token * 10.0
    head = head_from_params(p, K)
    ref_head = head_from_params(p, K)
    mom = 0.0 if opt == "sgd_nomom" else 0.9
    kw = dict(lr=0.05, weight_decay=0.01, use_graph=True)
    tr = E.EPHeadTrainer(head, B, N, optimizer="adamw" if opt == "adamw" else "sgd", momentum=mom, **kw)
    ropt = (torch.optim.AdamW(ref_head.parameters(), lr=0.05, weight_decay=0.01) if opt == "adamw" else
            torch.optim.SGD(ref_head.parameters(), lr=0.05, weight_decay=0.01, momentum=mom))
    ref_head.train()
    for it in range(4):
        x = O.synthetic_tokens(B, N, D, seed=70 + it).to(DEV)
        y = O.synthetic_labels(B, K, seed=80 + it).to(DEV)
        tr.train_step(x, y)
        ropt.zero_grad()
        nn.CrossEntropyLoss()(ref_head(x), y).backward()      # the drop-in module + torch BN/Linear/CE/optimizer
        ropt.step()
    for (k, a), (_, b) in zip(head.named_parameters(), ref_head.named_parameters()):
        close(a.detach(), b.detach(), 2e-3, f"{opt} {k}")
    sd = tr.optimizer_state_dict()
    assert set(sd["state"][0]) == ({"step", "exp_avg", "exp_avg_sq"} if opt == "adamw" else
                                   ({"momentum_buffer"} if mom else set()))


def test_accum_iter_matches_large_batch_gradient():
    """engine_finetune.py:72-77: k micro-steps with loss / k, then one optimizer step."""
    B, N, D, M, K = 8, 20, 128, 8, 10
    p = O.build_head(D, M, K, seed=0)
    xs = [O.synthetic_tokens(B, N, D, seed=90 + i).to(DEV) for i in range(2)]
    ys = [O.synthetic_labels(B, K, seed=95 + i).to(DEV) for i in range(2)]
    h_acc, h_ref = head_from_params(p, K), head_from_params(p, K)
    tr = E.EPHeadTrainer(h_acc, B, N, lr=0.1, accum_iter=2, use_graph=True)
    before = [q.detach().clone() for q in h_acc.parameters()]
    tr.train_step(xs[0], ys[0])
    for q, b in zip(h_acc.parameters(), before):
        assert torch.equal(q, b)                               # no update after the first micro-step
    tr.train_step(xs[1], ys[1])
    opt = E.LARS(h_ref.parameters(), lr=0.1)
    h_ref.train()
    for x, y in zip(xs, ys):
        (nn.CrossEntropyLoss()(h_ref(x), y) / 2).backward()
    opt.step()
    for (k, a), (_, b) in zip(h_acc.named_parameters(), h_ref.named_parameters()):
        close(a.detach(), b.detach(), 1e-3, "accum " + k)


FUSE_CASES = [
    # B, N, D, M, K, d_out, bias   (B % 64 == 0: hi/lo P and every tcgen05 GEMM; the others exercise the fall-backs)
    (64, 70, 256, 8, 40, 1, 0),          # c = 32: bf16 copies everywhere
    (128, 33, 1152, 32, 1000, 1, 0),     # c = 36: tf32 copies of g / W_m^T, bf16 for the rest (config 3's head)
    (64, 40, 256, 8, 24, 2, 1),          # d_out = 2 with bias: delta subtracts the bias
    (20, 17, 128, 8, 10, 1, 0),          # B % 64 != 0: fp32 P, mma.sync weight gradient, g^T copy unused
    (64, 50, 256, 4, 36, 1, 0),          # K % 8 != 0: classifier copies fall back to the self-made ones
]


@pytest.mark.parametrize("case", FUSE_CASES, ids=lambda c: "B%d_N%d_D%d_M%d_K%d_do%d_b%d" % c)
@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_operand_copies_by_producers_match_self_contained_calls(case, graph):
    """ABI 2 (*_ops entry points): operand copies written by the producing kernels + ep_refresh_operands give the
    step the self-contained calls give -- the same operand bits reach the same GEMMs.  Only the classifier weight
    gradient changes kernel (tcgen05 3-term instead of mma.sync TF32: closer to the oracle, compared at 1e-3)."""
    B, N, D, M, K, d_out, bias = case
    p = O.build_head(D, M, K, seed=3, d_out=d_out, qkv_bias=bool(bias))
    p.cls_token = p.cls_token * 8.0
    h_f, h_s = head_from_params(p, K), head_from_params(p, K)
    t_f = E.EPHeadTrainer(h_f, B, N, lr=0.3, weight_decay=1e-4, use_graph=graph, fuse_operands=True)
    t_s = E.EPHeadTrainer(h_s, B, N, lr=0.3, weight_decay=1e-4, use_graph=graph, fuse_operands=False)
    for it in range(3):
        x = O.synthetic_tokens(B, N, D, seed=300 + it).to(DEV)
        y = O.synthetic_labels(B, K, seed=310 + it).to(DEV)
        if it == 2:                       # a parameter changed behind the trainers' back: the copies must follow
            with torch.no_grad():
                for h in (h_f, h_s):
                    h[0].cls_token.mul_(1.25)
                    h[2].weight.add_(0.01)
        t_f.train_step(x, y)
        t_s.train_step(x, y)
        torch.cuda.synchronize()
        assert abs(float(t_f.step_loss) - float(t_s.step_loss)) <= (2e-6 if it == 0 else 1e-4) * abs(float(t_s.step_loss)), it
        for k in ("cls", "v_w", "fc_b"):
            close(t_f.g[k], t_s.g[k], 2e-5 if it == 0 else 1e-3, f"step {it} grad {k}")
        close(t_f.g["fc_w"], t_s.g["fc_w"], 1e-3, f"step {it} grad fc_w")
        # (from the second step on the parameters differ by what the two classifier weight-gradient kernels differ)
        close(t_f.logits, t_s.logits, 2e-6 if it == 0 else 1e-4, f"step {it} logits")
    assert t_f.launches_per_step < t_s.launches_per_step
    assert abs(t_f.mean_loss() - t_s.mean_loss()) <= 1e-4 * abs(t_s.mean_loss())
    xe = O.synthetic_tokens(B - 3 if B > 8 else B, N, D, seed=399).to(DEV)
    assert torch.equal(t_f.eval_logits(xe).argmax(1), t_s.eval_logits(xe).argmax(1))
    close(t_f.eval_logits(xe), t_s.eval_logits(xe), 1e-3, "eval logits")
