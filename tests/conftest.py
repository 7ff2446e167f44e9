import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = torch.load(GOLDEN / f"{name}.pt", map_location="cpu", weights_only=False)
        return cache[name]

    return load


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _exact_fp32():
    # parity tests compare against an fp32 oracle: keep library GEMMs/convs in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture()
def fixed_normal():
    """Replace torch.normal / torch.randn_like by one deterministic, device-independent pattern, so that the node
    hallucination of GModule (graph_matching.py:438-463: classes present in one domain only are completed with random
    draws around the seed bank) produces the SAME nodes in the CUDA engine and in the CPU oracle, and every loss
    downstream of it can be compared exactly instead of statistically."""
    real_normal, real_randn_like = torch.normal, torch.randn_like

    def pattern(shape, device):
        n = 1
        for s_ in shape:
            n *= int(s_)
        t = torch.arange(n, device=device, dtype=torch.float32) * 0.6180339887
        return ((t - t.floor()) * 2.0 - 1.0).reshape(tuple(shape))

    def normal(mean=0.0, std=1.0, size=None, generator=None, **kw):
        if torch.is_tensor(mean):
            return mean + std * pattern(mean.shape, mean.device)
        if torch.is_tensor(std):
            return mean + std * pattern(std.shape, std.device)
        return mean + std * pattern(size, "cpu")

    def randn_like(t, **kw):
        return pattern(t.shape, t.device)

    torch.normal, torch.randn_like = normal, randn_like
    try:
        yield
    finally:
        torch.normal, torch.randn_like = real_normal, real_randn_like
