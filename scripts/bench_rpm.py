"""Micro-benchmark of the Sinkhorn (K4) kernels: log-domain vs register-resident paths, forward and backward,
one problem and a saturating batch.  CUDA events, L2 flushed between iterations, median of 15.
Usage (GPU box): python scripts/bench_rpm.py > gpurun_out/rpm_paths.json"""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from graphecho_b200 import functional as GF, _cabi  # noqa: E402


def timed(fn, flush, iters=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for batch, n in ((1, 252), (8, 252), (74, 252), (512, 252), (512, 128), (2048, 64)):
        M = torch.randn(batch, n, n, device=dev)
        for path in (1, 2, 3):
            _cabi.lib().ge_sinkhorn_rpm_set_path(path)
            ms_f = timed(lambda: GF.sinkhorn_rpm_exp(M, 20, True), flush)
            Mr = M.clone().requires_grad_()
            P = GF.sinkhorn_rpm_exp(Mr, 20, True)
            g = torch.randn_like(P)
            ms_b = timed(lambda: torch.autograd.grad(P, Mr, g, retain_graph=True), flush)
            rows.append({"batch": batch, "n": n, "path": path, "fwd_ms": round(ms_f, 4), "bwd_ms": round(ms_b, 4),
                         "iters_per_s": round(20 * batch / (ms_f / 1e3)),
                         "fwd_GBps": round(8 * batch * n * n / ms_f / 1e6, 1),
                         "fwd_fma_GFLOPs": round(batch * n * n * 80 / ms_f / 1e6, 1)})
            print(json.dumps(rows[-1]), flush=True)
    _cabi.lib().ge_sinkhorn_rpm_set_path(0)


if __name__ == "__main__":
    main()
