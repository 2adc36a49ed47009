"""Multi-GPU parity of the trainer on hardware (needs >= 2 GPUs; run with `gpurun --gpus 2`): two ranks, one process
each, NCCL -- main_linprobe.py:581-583 semantics: replicas made identical at construction (DDP's constructor
broadcast; the ranks are seeded seed + rank like main_linprobe.py:517), batch sharded, BatchNorm statistics per
rank, gradients averaged, every rank applies the same LARS step.  Checked: parameters bit-identical across ranks
after k steps, and equal to the oracle's mean-of-shard-gradients LARS steps on the global batch."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu
STEPS, B_LOCAL, N, D, M, K, LR = 3, 64, 70, 256, 8, 16, 0.5


def _worker(rank, world, port, graph, bcast, out_q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, ROOT)
    import efficient_probing_b200 as E
    from oracle import ep_oracle as O
    torch.manual_seed(100 + rank)                                  # different init per rank: the constructor must fix it
    head = E.make_ep_head(D, M, K).to(dev)
    with torch.no_grad():
        head[0].cls_token.mul_(10.0)
    tr = E.EPHeadTrainer(head, B_LOCAL, N, lr=LR, use_graph=graph, broadcast_buffers=bcast)
    init = {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}
    losses = []
    for step in range(STEPS):
        x = O.synthetic_tokens(world * B_LOCAL, N, D, seed=300 + step)
        y = O.synthetic_labels(world * B_LOCAL, K, seed=400 + step)
        lo, hi = E.flatgrad.shard_range(rank, world, world * B_LOCAL)
        tr.train_step(x[lo:hi].to(dev), y[lo:hi].to(dev))
        losses.append(float(tr.step_loss))
    ev = tr.eval_logits(O.synthetic_tokens(5, N, D, seed=999).to(dev)).cpu()     # partial batch, rank-0 statistics
    torch.cuda.synchronize()
    out_q.put((rank, init, {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}, losses, ev))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("graph,bcast", [(False, "eval"), (True, "step")], ids=["eager", "graph_stepbcast"])
def test_two_gpu_trainer_matches_oracle_and_ranks_agree(graph, bcast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import ep_oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200 + (1 if graph else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, graph, bcast, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, init0, sd0, loss0, ev0), (_, init1, sd1, loss1, ev1) = res
    # (a) replicas: identical after construction (rank 0's init everywhere) and after the steps, bit for bit
    for k in init0:
        assert torch.equal(init0[k], init1[k]), "constructor broadcast " + k
    for k in sd0:
        if k.startswith("1.running"):
            continue                                               # each rank folds its own batch statistics in (no SyncBN)
        assert torch.equal(sd0[k], sd1[k]), "after steps " + k
    assert torch.equal(ev0, ev1)                                   # both ranks evaluate on rank 0's running statistics
    # (b) the oracle: per-rank forward/backward on its shard (own BatchNorm statistics), mean of gradients, one LARS step
    p = O.EPParams(init0["0.cls_token"], init0["0.v.weight"], None, init0["1.running_mean"], init0["1.running_var"], 0,
                   init0["2.weight"], init0["2.bias"], M, 1, D ** -0.5).clone(torch.float64)
    mus = None
    for step in range(STEPS):
        x = O.synthetic_tokens(world * B_LOCAL, N, D, seed=300 + step)
        y = O.synthetic_labels(world * B_LOCAL, K, seed=400 + step)
        rs = [O.head_loss_and_grads_pooled(p, x[r * B_LOCAL:(r + 1) * B_LOCAL], y[r * B_LOCAL:(r + 1) * B_LOCAL]) for r in range(world)]
        names = [n for n, _ in p.trainable()]
        params = [t for _, t in p.trainable()]
        grads = [sum(r["grad." + n] for r in rs) / world for n in names]
        mus = mus or [torch.zeros_like(t) for t in params]
        new_p, mus = O.lars_step(params, grads, mus, lr=LR)
        p.cls_token, p.v_weight, p.fc_weight, p.fc_bias = new_p
        p.running_mean, p.running_var = rs[0]["running_mean"], rs[0]["running_var"]      # rank 0's chain
        assert abs(loss0[step] - float(rs[0]["loss"])) < 1e-3 * abs(float(rs[0]["loss"]))
        assert abs(loss1[step] - float(rs[1]["loss"])) < 1e-3 * abs(float(rs[1]["loss"]))
    for n, t in p.trainable():
        assert O.rel_err(sd0[n], t) < 1e-3, n
    assert O.rel_err(sd0["1.running_mean"], p.running_mean) < 1e-3
    assert O.rel_err(sd0["1.running_var"], p.running_var) < 1e-3
    # evaluation: rank 0's running statistics, 5-sample partial batch
    p.num_batches_tracked = STEPS
    ref = O.head_forward(p, O.synthetic_tokens(5, N, D, seed=999).double(), train=False)["logits"]
    assert O.rel_err(ev0, ref) < 1e-3
