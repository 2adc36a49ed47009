import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One committed fixture written by tests/golden/make_golden.py (outputs of the reference itself)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
        self.z = {k: z[k] for k in z.files}
        self.meta = json.loads(str(self.z["meta"]))
        self.name = name

    def t(self, key, dtype=None):
        t = torch.from_numpy(self.z[key])
        return t.to(dtype) if dtype is not None else t

    def params(self, dtype=torch.float32):
        """EPParams (oracle dataclass) filled from the fixture's state_dict."""
        from oracle.ep_oracle import EPParams
        m = self.meta
        g = lambda k: self.t("param." + k, dtype)
        return EPParams(g("0.cls_token"), g("0.v.weight"), g("0.v.bias") if m["qkv_bias"] else None,
                        g("1.running_mean"), g("1.running_var"), int(self.z["param.1.num_batches_tracked"]),
                        g("2.weight"), g("2.bias"), m["M"], m["d_out"], m["scale"])


GOLDEN_CASES = ["small", "dout2_bias", "m32_sharp", "cls197"]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)
