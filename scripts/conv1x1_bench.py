"""In-graph cost of the 1x1 convolutions of the config-2 trunk on cuDNN (bf16, channels_last) and as cuBLAS GEMMs:
fprop, dgrad, wgrad.  bytes = A + W + D once."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cl = torch.channels_last
shapes = [(256, 28, 64, 64), (256, 28, 64, 256), (256, 28, 256, 64), (256, 28, 256, 128), (256, 28, 256, 256), (256, 28, 512, 256),
          (256, 14, 128, 512), (256, 14, 512, 128), (256, 14, 512, 256), (256, 7, 256, 1024), (256, 7, 1024, 256), (256, 7, 1024, 512),
          (256, 4, 512, 2048), (256, 4, 2048, 512), (256, 4, 2048, 256)]
R = 6
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
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
    return ts[3] / R * 1e3
print(f"{'N x HW x Cin->Cout':26s} {'MB':>7s} | {'cudnn fprop':>11s} {'GB/s':>6s} | {'mm fprop':>9s} {'GB/s':>6s} | {'cudnn dgrad':>11s} {'mm dgrad':>9s} | {'cudnn wgrad':>11s} {'mm wgrad':>9s}")
for N, H, Ci, Co in shapes:
    xs = [torch.randn(N, Ci, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(R)]
    ws = [torch.randn(Co, Ci, 1, 1, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(R)]
    gs = [torch.randn(N, Co, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(R)]
    P = N * H * H
    x2 = [x.permute(0, 2, 3, 1).reshape(P, Ci) for x in xs]
    w2 = [w.reshape(Co, Ci) for w in ws]
    g2 = [g.permute(0, 2, 3, 1).reshape(P, Co) for g in gs]
    t_cf = graph_time(lambda: [F.conv2d(x, w) for x, w in zip(xs, ws)])
    t_mf = graph_time(lambda: [a @ w.t() for a, w in zip(x2, w2)])
    t_cd = graph_time(lambda: [torch.ops.aten.convolution_backward(g, x, w, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False]) for g, x, w in zip(gs, xs, ws)])
    t_md = graph_time(lambda: [g @ w for g, w in zip(g2, w2)])
    t_cw = graph_time(lambda: [torch.ops.aten.convolution_backward(g, x, w, None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False]) for g, x, w in zip(gs, xs, ws)])
    t_mw = graph_time(lambda: [g.t() @ a for g, a in zip(g2, x2)])
    mb = (P * (Ci + Co) + Ci * Co) * 2 / 1e6
    print(f"{N}x{H}x{H} {Ci:5d}->{Co:<5d} {mb:9.1f} | {t_cf:11.1f} {mb / t_cf * 1e3:6.0f} | {t_mf:9.1f} {mb / t_mf * 1e3:6.0f} | {t_cd:11.1f} {t_md:9.1f} | {t_cw:11.1f} {t_mw:9.1f}")
