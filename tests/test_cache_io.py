"""CPU tests of the rows either side of the hot path: token cache format (tools/dump_tokens.py) and head files
(util/misc.py:304-332, tools/export_ep_heads.py:125)."""
import numpy as np
import os

import pytest
import torch

import efficient_probing_b200 as E
from efficient_probing_b200.token_cache import epoch_order


def test_shard_round_trip_and_alignment(tmp_path):
    t = torch.randn(37, 19, 64)
    y = torch.randint(0, 1000, (37,))
    p = str(tmp_path / "a.eptok")
    E.write_shard(p, t, y, meta={"model": "vit_base_patch16", "cls_features": "ep"})
    s = E.TokenShard(p)
    assert (s.n, s.N, s.C) == (37, 19, 64) and s.header["meta"]["cls_features"] == "ep"
    assert torch.equal(s.tokens, t.to(torch.bfloat16)) and torch.equal(s.labels, y)
    assert s.header["tokens_offset"] % 4096 == 0 and s.header["labels_offset"] % 4096 == 0
    with pytest.raises(ValueError):
        E.write_shard(p, t[0], y)
    (tmp_path / "bad.eptok").write_bytes(b"{}" + b"\0" * 5000)
    with pytest.raises(ValueError):
        E.TokenShard(str(tmp_path / "bad.eptok"))


def test_reads_reference_dump_tokens_npz(tmp_path):
    tokens = np.random.randn(35, 196, 32).astype(np.float32)              # dump_tokens.py: (n, N, C) float32, n = 35
    p = str(tmp_path / "tokens.npz")
    np.savez_compressed(p, tokens=tokens, images=np.zeros((35, 3, 8, 8), np.float32), names=np.array(["a"] * 35))
    t, images, names = E.load_reference_npz(p)
    assert t.dtype == torch.bfloat16 and t.shape == (35, 196, 32)
    assert torch.equal(t, torch.from_numpy(tokens).to(torch.bfloat16)) and len(names) == 35
    np.savez(str(tmp_path / "x.npz"), other=tokens)
    with pytest.raises(ValueError):
        E.load_reference_npz(str(tmp_path / "x.npz"))


def test_epoch_order_partitions_like_distributed_sampler():
    total, world, batch = 1003, 4, 32
    per_rank = [epoch_order(total, epoch=3, seed=7, rank=r, world=world, batch=batch) for r in range(world)]
    assert all(len(o) == (total // world) // batch * batch for o in per_rank)          # drop_last
    allidx = torch.cat(per_rank)
    assert len(set(allidx.tolist())) == len(allidx)                                   # ranks are disjoint
    assert not torch.equal(per_rank[0], epoch_order(total, 4, 7, 0, world, batch))      # set_epoch reshuffles
    assert torch.equal(per_rank[0], epoch_order(total, 3, 7, 0, world, batch))          # and is reproducible


def test_head_files_both_formats(tmp_path):
    torch.manual_seed(0)
    head = E.make_ep_head(64, num_queries=8, nb_classes=10, d_out=2, qkv_bias=True)
    p1, p2 = str(tmp_path / "checkpoint-best.pth"), str(tmp_path / "ep_head.pth")
    E.save_checkpoint(p1, head, optimizer_state={"state": {}, "param_groups": []}, epoch=89)
    E.export_head(p2, head, meta={"arch": "ViT-B/16", "ep_queries": 8, "d_out": 2})
    ck = torch.load(p1, weights_only=False)
    assert ck["saved_module"] == "head" and ck["epoch"] == 89                         # util/misc.py:318-326
    assert list(ck["model"]) == ["0.cls_token", "0.v.weight", "0.v.bias", "1.running_mean", "1.running_var",
                                 "1.num_batches_tracked", "2.weight", "2.bias"]
    assert set(torch.load(p2, weights_only=False)) == {"state_dict", "meta"}          # export_ep_heads.py:125
    for p in (p1, p2):
        h, meta = E.load_head(p)
        assert (meta["dim"], meta["num_queries"], meta["d_out"], meta["nb_classes"]) == (64, 8, 2, 10)
        for (k, a), (_, b) in zip(head.state_dict().items(), h.state_dict().items()):
            assert torch.equal(a, b), k
    # a full-model checkpoint carries a prefix; a file without queries is refused like ep_attention_maps.py:44-46
    h, _ = E.load_head({"model": {"head." + k: v for k, v in head.state_dict().items()}})
    assert torch.equal(h[0].cls_token, head[0].cls_token)
    with pytest.raises(ValueError):
        E.load_head({"model": {"head.weight": torch.zeros(3, 3)}})


def test_batch_gather_across_shards_matches_row_lookup(tmp_path):
    """TokenStream._gather (the host side of the loader) on CPU buffers: rows from several shards, repeated rows,
    a batch from a single shard -- against a per-row lookup."""
    import numpy as np
    from efficient_probing_b200 import token_cache as T
    g = torch.Generator().manual_seed(0)
    shards = []
    for i, n in enumerate([37, 50, 13]):
        path = str(tmp_path / f"s{i}.eptok")
        T.write_shard(path, torch.randn(n, 5, 16, generator=g), torch.randint(0, 10, (n,), generator=g))
        shards.append(T.TokenShard(path))

    class HostOnly(T.TokenStream):                     # the gather needs only the shard table and one buffer pair
        def __init__(self, shards, batch):
            self.shards, self.batch = list(shards), batch
            self.offsets = np.cumsum([0] + [len(s) for s in shards])
            self._hx = [torch.empty(batch, 5, 16, dtype=torch.bfloat16)]
            self._hy = [torch.empty(batch, dtype=torch.int64)]

    for idx in (torch.tensor([0, 99, 36, 37, 40, 86, 87, 5]), torch.tensor([40, 41, 38, 86, 50, 37]),
                torch.tensor([3, 3, 0, 36])):
        st = HostOnly(shards, len(idx))
        st._gather(idx, 0)
        for j, i in enumerate(idx.tolist()):
            sh = int(np.searchsorted(st.offsets, i, side="right") - 1)
            k = i - int(st.offsets[sh])
            assert torch.equal(st._hx[0][j], shards[sh].tokens[k]) and int(st._hy[0][j]) == int(shards[sh].labels[k])


def test_checkpoint_resumes_through_reference_load_model(tmp_path):
    """The checkpoint save_checkpoint writes goes through the reference's own ``util.misc.load_model``
    (util/misc.py:335-393): head weights into ``model.head``, the optimizer state into the reference LARS, the
    scaler entry into a GradScaler.  Needs the reference checkout (build container only)."""
    import argparse
    import importlib.util
    import sys
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "util", "misc.py")):
        pytest.skip("reference checkout not present")
    sys.dont_write_bytecode = True

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    try:
        misc = load("_ref_misc", os.path.join(ref, "util", "misc.py"))
        lars = load("_ref_lars", os.path.join(ref, "util", "lars.py"))
    except Exception as e:                                    # a dependency of util/misc.py missing in this image
        pytest.skip(f"reference util modules not importable here: {e}")
    torch.manual_seed(0)
    head = E.make_ep_head(64, num_queries=8, nb_classes=10)
    mu = {i: {"mu": torch.full_like(p, 0.25 * (i + 1))} for i, p in enumerate(head.parameters())}
    opt_state = {"state": mu, "param_groups": [{"lr": 0.4, "weight_decay": 0.0, "momentum": 0.9, "trust_coefficient": 0.001,
                                                "params": list(range(len(mu)))}]}
    path = str(tmp_path / "checkpoint-3.pth")
    E.save_checkpoint(path, head, opt_state, epoch=3, test_stats={"acc1": 1.0})
    with pytest.raises(ValueError):
        E.save_checkpoint(path + ".bad", head, None)

    class Model(torch.nn.Module):                             # what main_linprobe.py hands to load_model: a model with .head
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Linear(4, 4)
            torch.manual_seed(1)
            self.head = E.make_ep_head(64, num_queries=8, nb_classes=10)
    model = Model()
    optimizer = lars.LARS(model.head.parameters(), lr=0.1)

    class Scaler:                                             # NativeScalerWithGradNormCount.load_state_dict (util/misc.py:285-286)
        def __init__(self):
            self._scaler = torch.amp.GradScaler("cpu", enabled=True)

        def load_state_dict(self, sd):
            self._scaler.load_state_dict(sd)
    args = argparse.Namespace(resume=path, start_epoch=0, eval=False, knn_eval=False)
    stats = misc.load_model(args, model, optimizer=optimizer, loss_scaler=Scaler())
    assert stats == {"acc1": 1.0} and args.start_epoch == 4
    for (k, a), (_, b) in zip(head.state_dict().items(), model.head.state_dict().items()):
        assert torch.equal(a, b), k
    for i, p in enumerate(model.head.parameters()):
        assert torch.equal(optimizer.state[p]["mu"], mu[i]["mu"])
    assert optimizer.param_groups[0]["lr"] == 0.4
