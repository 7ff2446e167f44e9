"""One call of every custom kernel family at its config-2 shape between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off` (DRAM traffic / pipe utilisation per launch).  Round 2: the current kernels
(segmented BatchNorm with the 1-bit ReLU mask, tcgen05 1x1-conv GEMM, fused loss, TGCN recurrence, ...)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graphecho_b200 import functional as GF
dev = torch.device("cuda:0")
torch.manual_seed(0)
N, C, H = 256, 256, 28
cl = torch.channels_last
x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=cl)
res = torch.randn_like(x)
bn = torch.nn.BatchNorm2d(C).to(dev)
conv = torch.nn.Conv2d(C, 64, 1, bias=False).to(dev)
bn64 = torch.nn.BatchNorm2d(64).to(dev)
gamma, beta, pre_bias = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.randn(C, device=dev) * 0.1
xn = torch.randn(N, H * H, C, device=dev).bfloat16()
A, B = torch.randn(1, 252, 512, device=dev), torch.randn(1, 252, 512, device=dev)
w2, b2 = torch.randn(512, device=dev), torch.randn(1, device=dev)
M = torch.randn(1, 252, 252, device=dev)
M512 = torch.randn(512, 252, 252, device=dev)
s = [torch.randn(N, 128, H, H, device=dev).bfloat16().contiguous(memory_format=cl) for _ in range(4)]
W3, b3 = torch.randn(2, 128, 1, 1, device=dev), torch.randn(2, device=dev)
stem = torch.randn(N, 64, 56, 56, device=dev).bfloat16().contiguous(memory_format=cl)
logits = torch.randn(128, 2, 112, 112, device=dev)
masks = (torch.rand(128, 2, 112, 112, device=dev) > 0.5).float()
emb = torch.randn(8, 8, 256, 64, device=dev)
Wg, bg = torch.randn(256, 128, 1, 1, device=dev) * 0.05, torch.zeros(256, device=dev)
top = torch.randn(N, C, 14, 14, device=dev).bfloat16().contiguous(memory_format=cl)


def run():
    with GF.domain_split(N // 2):
        xr = x.clone().requires_grad_()
        o = GF.bn_act(xr, bn, residual=res, relu=True)
        o.backward(torch.ones_like(o))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            xc = x.clone().requires_grad_()
            o = GF.conv1x1_bn_act(xc, conv, bn64, relu=True)
        o.backward(torch.ones_like(o))
    xg = x.clone().requires_grad_()
    o = GF.gn_relu(xg, gamma, beta, 32, pre_bias=pre_bias)
    o.backward(torch.ones_like(o))
    xh = x.clone().requires_grad_()
    o = GF.gn_relu_upsample(xh, gamma, beta, (H, H))
    o.backward(torch.ones_like(o))
    tp = top.clone().requires_grad_()
    o = GF.upsample_add(tp, x)
    o.backward(torch.ones_like(o))
    st = stem.clone().requires_grad_()
    o = GF.maxpool3s2(st)
    o.backward(torch.ones_like(o))
    e = GF.knn_graph_nmajor(xn, None, 9, 1)[0]
    xr2 = xn.clone().requires_grad_()
    f = GF.mr_gather_nmajor(xr2, e)
    f.backward(torch.ones_like(f))
    Ar, Br = A.clone().requires_grad_(), B.clone().requires_grad_()
    Mx = GF.affinity_pairwise(Ar, Br, w2, b2)
    Mx.sum().backward()
    Mr = M.clone().requires_grad_()
    P = GF.sinkhorn_rpm_exp(Mr, 20, True)
    P.sum().backward()
    GF.sinkhorn_rpm_exp(M512, 20, True)
    sr = [t.clone().requires_grad_() for t in s]
    lg = GF.seg_tail(*sr, W3, b3, 4)
    lg.sum().backward()
    lr = logits.clone().requires_grad_()
    GF.seg_loss(lr, masks).backward()
    GF.mask_boxes(GF.LogitMap(logits))
    er = emb.clone().requires_grad_()
    h, _ = GF.tgcn_recurrence(er, Wg, bg, 9)
    h.sum().backward()


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
