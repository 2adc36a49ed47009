"""CPU tests of the host side: module surface identical to poolings/ep.py + probe_heads.py, the C-ABI
library loads and exports every symbol include/ep_b200.h declares, argument rejection without a GPU."""
import ctypes
import hashlib
import json
import os
import re
from argparse import Namespace

import pytest
import torch
from torch import nn

import efficient_probing_b200 as E
from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ep_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ep_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(E._lib.lib_path()) if os.path.exists(E._lib.lib_path()) else None
    assert lib is not None or E._lib.load() is not None
    lib = lib or ctypes.CDLL(E._lib.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in ep_b200.h but not exported"
    assert declared == set(E._lib.EXPORTED_SYMBOLS), declared ^ set(E._lib.EXPORTED_SYMBOLS)
    L = E._lib.load()
    assert L.ep_abi_version() == 2
    assert b"NULL" in L.ep_strerror(-1)


def test_argument_rejection_needs_no_gpu():
    L = E._lib.load()
    # NULL pointers / bad shapes are rejected before any CUDA call
    assert L.ep_fwd(None, 0, None, None, None, 1.0, 1, 1, 8, 1, 1, None, None, None, None, None, None, None, 0, None) == -1
    assert L.ep_workspace_bytes(0, 1, 1, 1, 1) == 0
    assert L.ep_workspace_bytes(4, 19, 64, 8, 1) > 0
    assert L.ep_set_kernel_mode(7) == -2 and L.ep_set_kernel_mode(0) == 0
    # the extended entry points check their arguments the same way
    assert L.ep_fwd_ex(None, 0, None, 1, None, None, 1.0, 1, 1, 8, 1, 1, None, None, None, None, None, None, None, 0, None) == -1
    assert L.ep_bwd_ex(None, 0, None, 0, None, 1.0, 1, 1, 8, 1, 1, None, None, None, None, 0, None, None, None, None, None,
                       None, None, None, 0, None) == -1
    assert L.ep_pooled_layout(0, 4, 19, 64, 5, 1) == -2          # 64 % 5 != 0
    assert L.ep_pooled_layout(1, 64, 19, 128, 8, 1) == 0         # fp32 tokens: general kernels, fp32 P
    assert L.ep_set_sm_limit(-1) == -2 and L.ep_set_sm_limit(0) == 0
    assert L.ep_linear_workspace_bytes(8, 72, 10) >= 3 * 2 * (10 * 128 + 72 * 64 + 8 * 128 + 8 * 64)   # padded thirds
    # ABI 2 (*_ops variants): the same checks in front of any CUDA call
    assert L.ep_fwd_ops(None, 0, None, None, None, 1.0, 1, 1, 8, 1, 1, None, None, None, None, None, None, None, 0, 3, None) == -1
    assert L.ep_refresh_operands(None, None, 1.0, 0, 4, 19, 64, 8, 1, None, 0, None, 10, None, 0, 0, None) == -1
    one = ctypes.c_float(0.0)
    ptr = ctypes.cast(ctypes.pointer(one), ctypes.c_void_p)                       # any non-NULL address: rejected on shape / size
    assert L.ep_refresh_operands(ptr, ptr, 1.0, 0, 4, 19, 64, 5, 1, ptr, 1 << 30, None, 10, None, 0, 0, None) == -2   # 64 % 5
    assert L.ep_refresh_operands(ptr, ptr, 1.0, 0, 4, 19, 64, 8, 1, ptr, 16, None, 10, None, 0, 0, None) == -5         # workspace
    assert L.ep_bwd_proj_ops(None, None, None, None, None, 0, 4, 19, 64, 8, 1, None, None, None, 0, 3, None) == -1
    assert L.ep_bwd_proj_ops(ptr, ptr, ptr, ptr, None, 0, 4, 19, 64, 8, 1, ptr, None, ptr, 16, 3, None) == -5
    assert L.ep_bn_fwd_ops(None, 4, 8, 1e-6, 0.1, 1, None, None, None, None, None, None, 10, None, 0, 0, None) == -1
    assert L.ep_bn_bwd_ops(None, None, None, 4, 8, None, None, None, 0, 19, 64, 8, 1, None, 0, 0, None) == -1
    assert L.ep_bn_bwd_ops(ptr, ptr, ptr, 4, 9, ptr, ptr, None, 0, 19, 64, 8, 1, ptr, 1 << 30, 0, None) == -2          # F != D / d_out
    assert L.ep_linear_fwd_ops(None, None, None, 4, 8, 10, None, None, 0, 3, None) == -1
    assert L.ep_linear_bwd_ops(None, None, None, 4, 8, 10, None, None, None, None, 0, 3, None) == -1
    assert L.ep_ce_fwd_bwd_ops(None, None, 4, 10, 1.0, 1.0, None, None, None, None, None, 8, None, 0, 0, None) == -1
    assert L.ep_ce_fwd_bwd_ops(ptr, ptr, 4, 10, 1.0, 1.0, ptr, None, None, None, None, 8, None, 0, 0, None) == -1      # scratch is required


def test_module_surface_matches_reference_fingerprints():
    fp = json.load(open(os.path.join(GOLDEN, "fingerprints.json")))
    for key, sha in fp["init_sha256"].items():
        D, M, d_out, bias = (int(s.lstrip("DMdoutbias")) for s in key.split("_"))
        torch.manual_seed(0)
        head = E.make_ep_head(D, M, 1000, d_out=d_out, qkv_bias=bool(bias))
        assert list(head.state_dict().keys()) == fp["state_dict_keys"][key]
        h = hashlib.sha256()
        for n, p in sorted(head.named_parameters()):          # tools/inv_heads.py:113-116
            h.update(n.encode())
            h.update(p.detach().float().numpy().tobytes())
        assert h.hexdigest() == sha, key                      # same RNG draw order as poolings/ep.py:25-26
        assert sum(p.numel() for p in head.parameters()) == fp["param_count"][key]
    # parameters() order drives optimizer-state indices (util/misc.py:322)
    assert [n for n, _ in E.make_ep_head(64, 8, 10).named_parameters()] == \
        ["0.cls_token", "0.v.weight", "2.weight", "2.bias"]


def test_constructor_attributes_and_errors():
    m = E.EfficientProbing(64, num_queries=8, d_out=2, qkv_bias=True, qk_scale=None)
    assert (m.num_heads, m.d_out, m.num_queries) == (1, 2, 8)
    assert m.scale == 64 ** -0.5 and m.v.weight.shape == (32, 64) and m.v.bias.shape == (32,)
    assert m.cls_token.shape == (1, 8, 64)
    assert E.EfficientProbing(64, qk_scale=0.3).scale == 0.3
    with pytest.raises(RuntimeError):                         # no CPU path
        m(torch.randn(2, 5, 64))
    with pytest.raises(RuntimeError):                         # reference fails for num_heads > 1 as well (ep.py:45)
        E.EfficientProbing(64, num_heads=2, num_queries=8)(torch.randn(2, 5, 64))
    with pytest.raises(RuntimeError):                         # 64 % 5 != 0: the reference's reshape error (ep.py:40)
        E.EfficientProbing(64, num_queries=5)(torch.randn(2, 5, 64))
    with pytest.raises(RuntimeError):                         # external queries take the same (CUDA-only) path
        m(torch.randn(2, 5, 64), cls=torch.randn(2, 8, 64))


class _Stub(nn.Module):
    def __init__(self, dim, nb):
        super().__init__()
        self.head = nn.Linear(dim, nb)


def test_build_probe_head_like_reference():
    args = Namespace(cls_features="ep_all", ep_queries=16, d_out=2, nb_classes=10)
    m = _Stub(64, 10)
    old = m.head
    E.build_probe_head(m, args)
    assert isinstance(m.head[0], E.EfficientProbing) and m.head[0].num_queries == 16
    assert isinstance(m.head[1], nn.BatchNorm1d) and not m.head[1].affine and m.head[1].eps == 1e-6
    assert m.head[1].num_features == 32 and m.head[2].in_features == 32 and m.head[2] is not old
    m = _Stub(64, 10)
    old = m.head
    E.build_probe_head(m, Namespace(cls_features="cls"))      # plain linear probe keeps the encoder's head
    assert isinstance(m.head[0], nn.BatchNorm1d) and m.head[1] is old
    with pytest.raises(NotImplementedError):
        E.build_probe_head(_Stub(64, 10), Namespace(cls_features="simpool"))


def test_lr_schedule_matches_reference():
    fp = json.load(open(os.path.join(GOLDEN, "fingerprints.json")))
    args = Namespace(**fp["lr_sched_args"])
    opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=0.0)
    for e, lr in fp["lr_sched"]:
        assert abs(E.adjust_learning_rate(opt, e, args) - lr) < 1e-15
        assert opt.param_groups[0]["lr"] == E.adjust_learning_rate(opt, e, args)


def test_lars_refuses_cpu_parameters():
    p = nn.Parameter(torch.randn(4, 4))
    p.grad = torch.randn(4, 4)
    with pytest.raises(RuntimeError):
        E.LARS([p], lr=0.1).step()
