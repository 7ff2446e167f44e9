"""In-graph per-call cost of the fused BatchNorm entry points at every shape of the config-2 trunk: 8 calls on distinct
tensors captured in one CUDA graph, replayed; time / 8.  bytes = compulsory traffic."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
cl = torch.channels_last
shapes = [(256, 64, 56, False), (256, 64, 28, False), (256, 256, 28, True), (256, 128, 28, False), (256, 128, 14, False), (256, 512, 14, True),
          (256, 256, 14, False), (256, 256, 7, False), (256, 1024, 7, True), (256, 512, 7, False), (256, 512, 4, False), (256, 2048, 4, True),
          (256, 512, 28, False)]
R = 8
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
print(f"{'shape':28s} {'MB':>7s} {'fwd us':>8s} {'fwd GB/s':>9s} {'bwd us':>8s} {'bwd GB/s':>9s}")
tot_f = tot_b = 0.0
for N, C, H, res in shapes:
    xs = [torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_() for _ in range(R)]
    rs = [torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl).requires_grad_() if res else None for _ in range(R)]
    bns = [torch.nn.BatchNorm2d(C).to(dev) for _ in range(R)]
    gs = [torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(R)]
    def fwd():
        with GF.domain_split(N // 2):
            return [GF.bn_act(x, bn, residual=r, relu=True) for x, r, bn in zip(xs, rs, bns)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        outs = fwd()
    torch.cuda.synchronize()
    def bwd():
        for o, g, x, r in zip(outs, gs, xs, rs):
            torch.autograd.grad(o, [x] + ([r] if r is not None else []), g, retain_graph=True)
    with torch.cuda.stream(side):
        bwd()
    torch.cuda.synchronize()
    res_t = []
    for fn in (fwd, bwd):
        g = torch.cuda.CUDAGraph()
        s = side
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res_t.append(ts[len(ts) // 2] / R * 1e3)
    mb = N * C * H * H * 2 / 1e6
    fb = mb * (3 if res else 2)
    bb = mb * (4 if res else 3) + mb / 16
    print(f"{N}x{C}x{H}x{H}{' +res' if res else '':5s} {mb:9.1f} {res_t[0]:8.1f} {fb / res_t[0] * 1e3:9.0f} {res_t[1]:8.1f} {bb / res_t[1] * 1e3:9.0f}")
