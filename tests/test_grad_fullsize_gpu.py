"""GPU gradient parity on the code path bench.py runs (tcgen05 family, P saved as bf16 hi/lo rows: B % 64 == 0)
at the BASELINE configs' full token / channel dimensions, against the CPU oracle's pooled closed form
(oracle.head_loss_and_grads_pooled, held to autograd of the reference formulation in tests/test_oracle.py).
Tolerances are BASELINE.json's (logits 1e-3, parameter gradients 2e-3 relative L2) plus a worst-element bound
(max |err| / max |ref|)."""
import pytest
import torch

import efficient_probing_b200 as E
from oracle import ep_oracle as O
from test_parity_gpu import head_from_params, close, DEV, TOL_FWD, TOL_GRAD

pytestmark = pytest.mark.gpu
TOL_MAX = 4e-3            # worst element, relative to the largest reference element


def check_grads(tr, head, ref):
    for k, grad in zip([n for n, _ in head.named_parameters()], tr.grads):
        r = ref["grad." + k].reshape(-1)
        close(grad.reshape(-1), r, TOL_GRAD, "grad " + k)
        assert O.max_rel_err(grad.reshape(-1).cpu(), r) < TOL_MAX, ("max-norm grad " + k, O.max_rel_err(grad.reshape(-1).cpu(), r))


FULL_DIM_CASES = [  # B, N, D, M, K, q_gain   (BASELINE configs 2..5 and 1, smaller batch)
    (64, 257, 1024, 32, 1000, 25.0),
    (64, 256, 1152, 32, 1000, 20.0),
    (64, 730, 1664, 32, 100, 15.0),
    (64, 201, 4096, 32, 100, 10.0),
    (64, 197, 768, 8, 1000, 20.0),
    (128, 257, 1024, 8, 100, 25.0),
    (192, 257, 1024, 32, 100, 40.0),      # more samples than one wave of CTAs handles: >= 2 samples per CTA
]


@pytest.mark.parametrize("case", FULL_DIM_CASES, ids=lambda c: "B%d_N%d_D%d_M%d" % c[:4])
@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_bench_path_gradients_at_full_dims(case, graph):
    B, N, D, M, K, q_gain = case
    lib = E._lib.load()
    lib.ep_set_kernel_mode(2)
    try:
        assert lib.ep_pooled_layout(0, B, N, D, M, 1) == 1          # the hi/lo-P layout bench.py runs on
        p = O.build_head(D, M, K, seed=0)
        p.cls_token = p.cls_token * q_gain
        x = O.synthetic_tokens(B, N, D, seed=77)
        y = O.synthetic_labels(B, K)
        ref = O.head_loss_and_grads_pooled(p, x, y, dtype=torch.float64)
        head = head_from_params(p, K)
        tr = E.EPHeadTrainer(head, B, N, lr=0.0, use_graph=graph)
        tr.train_step(x.to(DEV), y.to(DEV))
        torch.cuda.synchronize()
        assert lib.ep_last_kernel_family() == 2
        close(tr.out, ref["out"], TOL_FWD, "out")
        assert O.max_rel_err(tr.out.cpu(), ref["out"]) < 1e-3
        close(tr.logits, ref["logits"], TOL_FWD, "logits")
        close(tr.rowmax, ref["rowmax"], 1e-4, "rowmax")
        close(tr.rowsum, ref["rowsum"], 1e-4, "rowsum")
        assert abs(float(tr.step_loss) - float(ref["loss"])) < 1e-4 * abs(float(ref["loss"]))
        check_grads(tr, head, ref)
    finally:
        lib.ep_set_kernel_mode(0)


def test_c2_full_batch_gradients_vs_oracle():
    """BASELINE config 2 exactly (B=1024, N=257, D=1024, M=32, K=1000): every gradient of the step bench.py times,
    against the oracle on the full batch (pooled closed form, fp64; about a minute of host time)."""
    B, N, D, M, K = 1024, 257, 1024, 32, 1000
    p = O.build_head(D, M, K, seed=0)
    p.cls_token = p.cls_token * 25.0
    x = O.synthetic_tokens(B, N, D, seed=1234)
    y = O.synthetic_labels(B, K)
    ref = O.head_loss_and_grads_pooled(p, x, y, dtype=torch.float64)
    head = head_from_params(p, K)
    tr = E.EPHeadTrainer(head, B, N, lr=0.0, use_graph=True)
    tr.train_step(x.to(DEV), y.to(DEV))
    torch.cuda.synchronize()
    assert E._lib.load().ep_last_kernel_family() == 2
    close(tr.logits, ref["logits"], TOL_FWD, "logits")
    close(tr.out, ref["out"], TOL_FWD, "out")
    check_grads(tr, head, ref)
