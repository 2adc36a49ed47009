"""CPU tests: the oracle (oracle/ep_oracle.py) against fixtures produced by the reference itself."""
import json
import os

import pytest
import torch

from oracle import ep_oracle as O
from conftest import GOLDEN


def test_forward_matches_reference_f64(golden):
    p = golden.params(torch.float64)
    x = golden.t("x", torch.float64)
    r = O.head_forward(p, x, train=True)
    assert O.rel_err(r["out"], golden.t("f64.out")) < 1e-12
    assert O.rel_err(r["attn"], golden.t("f64.attn")) < 1e-12
    assert O.rel_err(r["logits"], golden.t("f64.logits")) < 1e-11
    assert O.rel_err(r["bn"][0], golden.t("f64.running_mean")) < 1e-12
    assert O.rel_err(r["bn"][1], golden.t("f64.running_var")) < 1e-12


def test_forward_matches_reference_f32(golden):
    p = golden.params(torch.float32)
    x = golden.t("x")
    r = O.head_forward(p, x, train=True)
    assert O.rel_err(r["out"], golden.t("f32.out")) < 2e-6
    assert O.rel_err(r["attn"], golden.t("f32.attn")) < 2e-6
    assert O.rel_err(r["logits"], golden.t("f32.logits")) < 2e-5


def test_pooled_reassociation_equals_reference(golden):
    """The pool-then-project form used by the CUDA kernels is the same function (fp64: rounding only)."""
    p = golden.params(torch.float64)
    x = golden.t("x", torch.float64)
    out, attn, P, rowmax, rowsum = O.ep_forward_pooled(x, p.cls_token, p.v_weight, p.v_bias, p.scale,
                                                       p.num_queries, p.d_out)
    assert O.rel_err(out, golden.t("f64.out")) < 1e-12
    assert O.rel_err(attn, golden.t("f64.attn")) < 1e-12


def test_grads_match_reference(golden):
    p = golden.params(torch.float64)
    r = O.head_loss_and_grads(p, golden.t("x"), golden.t("targets"), dtype=torch.float64)
    assert abs(float(r["loss"]) - float(golden.z["f64.loss"])) < 1e-12
    for k in golden.z:
        if k.startswith("f64.grad."):
            ref = golden.t(k)
            got = r["grad." + k[len("f64.grad."):]]
            # v.bias feeds a BatchNorm: its gradient is identically zero (1e-17 noise in the reference)
            assert (got - ref).norm() <= 1e-10 * ref.norm() + 1e-14, k


def test_closed_form_backward_matches_autograd(golden):
    p = golden.params(torch.float64)
    x = golden.t("x", torch.float64)
    out, attn, P, _, _ = O.ep_forward_pooled(x, p.cls_token, p.v_weight, p.v_bias, p.scale, p.num_queries, p.d_out)
    g = torch.randn(out.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(7))
    xr = x.clone().requires_grad_(True)
    leaves = [p.cls_token.clone().requires_grad_(True), p.v_weight.clone().requires_grad_(True)]
    vb = p.v_bias.clone().requires_grad_(True) if p.v_bias is not None else None
    o = O.ep_forward(xr, leaves[0], leaves[1], vb, p.scale, p.num_queries, p.d_out)
    grads = torch.autograd.grad((o * g).sum(), leaves + [xr] + ([vb] if vb is not None else []))
    cf = O.ep_backward_pooled(x, p.cls_token, p.v_weight, p.scale, p.num_queries, p.d_out, attn, P, g, want_dx=True)
    assert O.rel_err(cf["d_cls_token"], grads[0]) < 1e-10
    assert O.rel_err(cf["d_v_weight"], grads[1]) < 1e-10
    assert O.rel_err(cf["d_x"], grads[2]) < 1e-10
    if vb is not None:
        assert O.rel_err(cf["d_v_bias"], grads[3]) < 1e-10


def test_eval_logits(golden):
    p = golden.params(torch.float64)
    p.running_mean = golden.t("f64.running_mean")
    p.running_var = golden.t("f64.running_var")
    r = O.head_forward(p, golden.t("x", torch.float64), train=False)
    assert O.rel_err(r["logits"], golden.t("f64.eval_logits")) < 1e-11


def test_two_lars_steps_match_reference(golden):
    m = golden.meta
    p = golden.params(torch.float64)
    x, y = golden.t("x"), golden.t("targets")
    mus = None
    for step in range(2):
        r = O.head_loss_and_grads(p, x, y, dtype=torch.float64)
        names = [n for n, _ in p.trainable()]
        params = [t for _, t in p.trainable()]
        grads = [r["grad." + n] for n in names]
        mus = mus or [torch.zeros_like(t) for t in params]
        new_p, mus = O.lars_step(params, grads, mus, lr=m["lr"], weight_decay=m["weight_decay"])
        p.running_mean, p.running_var = r["running_mean"], r["running_var"]
        for n, t in zip(names, new_p):
            if n == "0.cls_token": p.cls_token = t
            elif n == "0.v.weight": p.v_weight = t
            elif n == "0.v.bias": p.v_bias = t
            elif n == "2.weight": p.fc_weight = t
            elif n == "2.bias": p.fc_bias = t
        if step == 1:
            assert abs(float(r["loss"]) - float(golden.z["f64.loss_step2"])) < 1e-10
    for n, t in p.trainable():
        assert O.rel_err(t, golden.t("f64.after2." + n)) < 1e-10, n


def test_fingerprints_init_order_counts_and_schedule():
    fp = json.load(open(os.path.join(GOLDEN, "fingerprints.json")))
    import hashlib
    for key, sha in fp["init_sha256"].items():
        D, M, d_out, bias = (int(s.lstrip("DMdoutbias")) for s in key.split("_"))
        p = O.build_head(D, M, 1000, d_out=d_out, qkv_bias=bool(bias), seed=0)
        h = hashlib.sha256()
        for n, t in sorted(p.trainable()):
            h.update(n.encode())
            h.update(t.detach().float().numpy().tobytes())
        assert h.hexdigest() == sha, key
        assert sum(t.numel() for _, t in p.trainable()) == fp["param_count"][key]
    for D, n in fp["param_count_logs"].items():          # the reference's own training logs (logs/*/ep.txt:9)
        assert O.param_count(int(D), 32) == n
    a = fp["lr_sched_args"]
    for e, lr in fp["lr_sched"]:
        assert abs(O.cosine_lr(e, a["lr"], a["min_lr"], a["warmup_epochs"], a["epochs"]) - lr) < 1e-15


def test_invalid_configs_raise_like_reference():
    with pytest.raises(ValueError):
        O.ep_init(64, num_heads=2)                        # reference: RuntimeError in forward (ep.py:45)
    with pytest.raises(RuntimeError):
        cls, w, b, s = O.ep_init(64, num_queries=5)       # 64 % 5 != 0 -> reshape fails (ep.py:40)
        O.ep_forward(torch.randn(2, 3, 64), cls, w, b, s, 5, 1)


def test_pooled_head_grads_equal_reference_formulation(golden):
    """head_loss_and_grads_pooled (closed form, used for BASELINE-sized GPU parity cases) == autograd through the
    reference formulation, every gradient and the loss."""
    p = golden.params(torch.float64)
    a = O.head_loss_and_grads(p, golden.t("x"), golden.t("targets"), dtype=torch.float64)
    b = O.head_loss_and_grads_pooled(p, golden.t("x"), golden.t("targets"), dtype=torch.float64)
    assert abs(float(a["loss"]) - float(b["loss"])) < 1e-12
    for k in a:
        if k.startswith("grad.") or k in ("logits", "out", "attn", "running_mean", "running_var"):
            assert (a[k] - b[k]).norm() <= 1e-10 * a[k].norm() + 1e-14, k
