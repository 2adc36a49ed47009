"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, the flat gradient buffer and
its sum all-reduce + 1/world scaling reproduce the single-process gradient of the global batch, exactly as
DDP does for the reference head (main_linprobe.py:581-583); bench.py's reference arm under a 2-rank launch."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from efficient_probing_b200.flatgrad import FlatGradLayout, shard_range, allreduce_sum_
    from oracle import ep_oracle as O
    B, N, D, M, K = 8, 11, 32, 4, 6
    p = O.build_head(D, M, K, seed=0)                       # identical init on every rank (DDP broadcast)
    x = O.synthetic_tokens(B, N, D, seed=1234).float()
    y = O.synthetic_labels(B, K)
    lo, hi = shard_range(rank, world, B)
    r = O.head_loss_and_grads(p, x[lo:hi], y[lo:hi], dtype=torch.float64)   # per-rank BN statistics, local mean loss
    lay = FlatGradLayout(K, D, D, M, False)
    flat = lay.allocate("cpu").double()
    v = lay.views(flat)
    v["fc_w"].copy_(r["grad.2.weight"].reshape(-1)); v["fc_b"].copy_(r["grad.2.bias"])
    v["v_w"].copy_(r["grad.0.v.weight"].reshape(-1)); v["cls"].copy_(r["grad.0.cls_token"].reshape(-1))
    allreduce_sum_(flat, None, 0, lay.early)                # the early part first, then the queries
    allreduce_sum_(flat, None, lay.early)
    flat *= 1.0 / world                                     # what the LARS kernel's grad_scale does
    # numpy arrays are pickled by value: a tensor would travel as a file descriptor that dies with this process
    out_q.put((rank, lay.early, lay.total, {k: t.detach().numpy().copy() for k, t in lay.views(flat).items()},
               {k: r[k].detach().numpy().copy() for k in r if k.startswith("grad.")}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_matches_ddp_average():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(a, b, c, {k: torch.from_numpy(v) for k, v in d.items()}, {k: torch.from_numpy(v) for k, v in e.items()})
           for a, b, c, d, e in res]
    (_, early, total, avg0, g0), (_, _, _, avg1, g1) = res
    assert early % 64 == 0 and total % 64 == 0
    for k, key in [("fc_w", "grad.2.weight"), ("fc_b", "grad.2.bias"), ("v_w", "grad.0.v.weight"), ("cls", "grad.0.cls_token")]:
        want = 0.5 * (g0[key] + g1[key]).reshape(-1)        # DDP: mean of the per-rank gradients
        assert torch.allclose(avg0[k], want, rtol=0, atol=1e-14) and torch.equal(avg0[k], avg1[k]), k


def test_shard_range_and_layout():
    sys.path.insert(0, ROOT)
    from efficient_probing_b200.flatgrad import FlatGradLayout, shard_range
    assert [shard_range(r, 4, 4096) for r in range(4)] == [(0, 1024), (1024, 2048), (2048, 3072), (3072, 4096)]
    with pytest.raises(ValueError):
        shard_range(0, 3, 4096)
    lay = FlatGradLayout(1000, 1024, 1024, 32, False)
    assert lay.sizes["v_b"] == 0 and lay.offsets["cls"] == lay.early
    assert all(o % 64 == 0 for o in lay.offsets.values())
    # the all-reduce payload SURVEY 2b quotes for ViT-L: (D'D + MD + D'K + K) * 4 B = 8.4 MB (+ padding)
    assert abs(lay.total * 4 - 8.425e6) < 2e4


def test_reference_arm_under_two_rank_launch():
    """bench.py --impl reference under torchrun: rank 0 alone prints the JSON line, rank 1 exits 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "bench.py"), "--impl",
           "reference", "--gpus", "2", "--steps", "1", "--warmup", "3", "--cpu-sample", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "tokens/s" and j["cpu_baseline"]["kind"] == "port"
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["value"] > 0
