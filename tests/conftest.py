import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def po():
    import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def b2():
    return importlib.import_module("troy-nova_b200")


# ---- parameter sets named by BASELINE.json `configs` (prime bit sizes; special prime last) ----
CONFIGS = {
    "C1": dict(scheme=1, log_n=13, bits=[40, 40, 40], t_bits=20),
    "C2": dict(scheme=1, log_n=14, bits=[60, 40, 40, 60], t_bits=20),
    "C3": dict(scheme=2, log_n=16, bits=[60] + [40] * 19 + [60], t_bits=0),
    "C4": dict(scheme=3, log_n=15, bits=[60] + [40] * 7 + [60], t_bits=20),
    "C5": dict(scheme=1, log_n=13, bits=[60, 40, 40, 60], t_bits=20),
}


def make_oracle(po, name=None, **kw):
    cfg = dict(CONFIGS[name]) if name else {}
    cfg.update(kw)
    n = 1 << cfg["log_n"]
    moduli = po.coeff_modulus_create(n, cfg["bits"])
    t = po.plain_modulus_batching(n, cfg["t_bits"]) if cfg.get("t_bits") else 0
    return po.Oracle(cfg["log_n"], moduli, t, cfg["scheme"])


def to_dev(a):
    """numpy uint64 -> torch cuda tensor (bit pattern preserved in int64)"""
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_np(t):
    return t.cpu().numpy().view(np.uint64)
