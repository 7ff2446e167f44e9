"""Quick check of the tcgen05 k-NN kernel against the fp32 FFMA kernels and the CPU oracle, plus timing.
Run under `timeout` on the GPU box: a pipeline bug in a hand-written mbarrier ring shows up as a hang."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from graphecho_b200 import functional as GF, _cabi
from oracle import vig_ops as V

dev = torch.device("cuda:0")
lib = _cabi.lib()


def run(path, x, y, k, d):
    assert lib.ge_knn_graph_set_path(path) == 0
    try:
        return GF.knn_graph(x, y, k, d)
    finally:
        lib.ge_knn_graph_set_path(0)


def check(B, C, N, M, k, d, self_graph, oracle=True):
    torch.manual_seed(B * 7 + N)
    x = torch.randn(B, C, N, 1)
    y = None if self_graph else torch.randn(B, C, M, 1)
    xd, yd = x.to(dev), None if y is None else y.to(dev)
    e_tc = run(2, xd, yd, k, d)
    torch.cuda.synchronize()
    e_ff = run(1, xd, yd, k, d)
    torch.cuda.synchronize()
    mism_ff = (e_tc != e_ff).float().mean().item()
    msg = f"B{B} C{C} N{N} M{M} k{k} d{d} self={self_graph}: tc-vs-ffma mismatch {mism_ff:.5f}"
    if oracle:
        dist = V.knn_distances(x, y, None)
        ref = V.dense_dilated_knn(x, y, k, d)
        e = e_tc.cpu()
        assert torch.equal(e[1], ref[1]), "centre index"
        mism = e[0] != ref[0]
        gap = (torch.gather(dist, 2, e[0]) - torch.gather(dist, 2, ref[0])).abs()
        worst = gap[mism].max().item() if mism.any() else 0.0
        msg += f"; vs oracle mismatch {mism.float().mean().item():.5f}, worst distance gap {worst:.2e}"
        assert worst < 2e-6, msg
    print(msg, flush=True)


check(1, 32, 128, 128, 9, 1, True)
check(2, 64, 200, 200, 9, 1, True)
check(2, 256, 784, 784, 9, 1, True)
check(2, 96, 300, 140, 9, 1, False)
check(2, 64, 256, 196, 9, 2, False)
check(1, 128, 500, 500, 16, 2, True)
check(1, 256, 4096, 1024, 9, 1, False)
# all keys identical -> lowest indices
x = torch.randn(2, 64, 256, 1, device=dev)
y = torch.zeros(2, 64, 256, 1, device=dev)
e = run(2, x, y, 9, 1).cpu()
assert torch.equal(e[0], torch.arange(9).expand(2, 256, 9)), "tie rule"
print("tie rule ok")

# timing at the bench shape
x = torch.randn(256, 256, 784, 1, device=dev)
for path in (2, 1):
    for _ in range(3):
        run(path, x, None, 9, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run(path, x, None, 9, 1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"path {path}: {ms:.3f} ms  ({2 * 256 * 784 * 784 * 256 / ms / 1e9:.1f} algorithmic TFLOP/s)")
