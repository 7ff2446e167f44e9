"""Correctness + in-graph timing of the tcgen05 1x1-conv GEMM with the BatchNorm-statistics epilogue against
cuDNN conv2d (+ our three-kernel BatchNorm) at the config-2 trunk shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
cl = torch.channels_last
shapes = [(256, 28, 64, 64), (256, 28, 64, 256), (256, 28, 256, 64), (256, 28, 256, 128), (256, 28, 256, 256),
          (256, 14, 128, 512), (256, 14, 512, 128), (256, 7, 256, 1024), (256, 7, 1024, 256), (3, 7, 64, 64), (5, 9, 128, 256)]
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
print(f"{'shape':26s} {'max|err|':>9s} {'stat err':>9s} | {'cudnn':>7s} {'ours':>7s} {'GB/s':>6s} | {'cudnn+bn':>9s} {'ours+bn':>8s}")
for N, H, Ci, Co in shapes:
    torch.manual_seed(Ci + Co)
    P = N * H * H
    if not GF.conv1x1_tc_supported(P, Ci, Co):
        print(f"{N}x{H}x{H} {Ci}->{Co}: unsupported"); continue
    xs = [torch.randn(N, Ci, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(R)]
    ws = [(torch.randn(Co, Ci, 1, 1, device=dev) / Ci ** 0.5).bfloat16() for _ in range(R)]
    x2 = [x.permute(0, 2, 3, 1).reshape(P, Ci) for x in xs]
    w2 = [w.reshape(Co, Ci).contiguous() for w in ws]
    shift = torch.randn(Co, device=dev) * 0.1
    ns = N // 2
    Ps = ns * H * H
    y, part, rows, rows0 = GF.conv1x1_gemm(x2[0], w2[0], shift, True, Ps)
    ref = x2[0].float() @ w2[0].float().t()
    err = float((y.float() - ref).abs().max())
    d = ref - shift
    s0, s1 = part[:rows0].sum(0), part[rows0:].sum(0)
    st_ref = torch.stack([d[:Ps].sum(0), (d[:Ps] ** 2).sum(0)]), torch.stack([d[Ps:].sum(0), (d[Ps:] ** 2).sum(0)])
    serr = max(float(((s0 - st_ref[0]).abs() / (st_ref[0].abs() + 1)).max()), float(((s1 - st_ref[1]).abs() / (st_ref[1].abs() + 1)).max()))
    t_c = graph_time(lambda: [F.conv2d(x, w) for x, w in zip(xs, ws)])
    t_o = graph_time(lambda: [GF.conv1x1_gemm(a, w) for a, w in zip(x2, w2)])
    bns = [torch.nn.BatchNorm2d(Co).to(dev) for _ in range(R)]
    convs = [torch.nn.Conv2d(Ci, Co, 1, bias=False).to(dev) for _ in range(R)]
    def cudnn_bn():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return [GF.bn_act(c(x), bn, relu=True) for c, x, bn in zip(convs, xs, bns)]
    def ours_bn():
        GF.USE_CONV1X1_TC = True
        with torch.autocast("cuda", dtype=torch.bfloat16):
            r = [GF.conv1x1_bn_act(x, c, bn, relu=True) for c, x, bn in zip(convs, xs, bns)]
        GF.USE_CONV1X1_TC = False
        return r
    with torch.no_grad():
        a, b = cudnn_bn()[0], ours_bn()[0]
        berr = float((a.float() - b.float()).abs().max())
        t_cb, t_ob = graph_time(cudnn_bn), graph_time(ours_bn)
    mb = (P * (Ci + Co) + Ci * Co) * 2 / 1e6
    print(f"{N}x{H}x{H} {Ci:5d}->{Co:<5d} {err:9.4f} {serr:9.2e} | {t_c:7.1f} {t_o:7.1f} {mb / t_o * 1e3:6.0f} | {t_cb:9.1f} {t_ob:8.1f}   bn-out diff {berr:.3f}")
